"""intrinsicnerf_b200 - B200 (sm_100a) implementation of IntrinsicNeRF's volumetric
ray-marching hot path behind the reference's Python API.  See DESIGN.md / INTEGRATION.md."""
from . import _lib, ops  # noqa: F401
from .nerf import Embedder, NeRF, Semantic_NeRF, get_embedder  # noqa: F401
from .ops import set_default_precision  # noqa: F401

__version__ = "0.1.0"
