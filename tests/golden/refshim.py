"""Kept for the fixture generator's `import refshim`: the shim lives in oracle/refshim.py."""
from oracle.refshim import *  # noqa: F401,F403
from oracle.refshim import _install_stubs, _stub  # noqa: F401
