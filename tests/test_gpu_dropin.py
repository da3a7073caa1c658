"""Drop-in proof (SURVEY section 4 tier T4, section 8b): the reference's own callers run UNCHANGED on top of
intrinsicnerf_b200.dropin - `run_nerf.train()` (object_level/run_nerf.py:664) with its render call (:942), loss.backward,
checkpoint write and render_path(update_cluster=True); `SSRTrainer.step` (SSR/training/trainer.py:851) and
`SSRTrainer.render_path` (:1221).  The reference sources come from /root/reference or from the byte-identical staged copy
under oracle/_ref (oracle/build_ref.py).  Each scenario runs in its own process: the object-level entry point switches the
default tensor type to CUDA and both rebind module globals."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _reference_present():
    from oracle import refshim
    return refshim.available()


def _run(script, marker, *args):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "tools", script), *args], cwd=ROOT, capture_output=True,
                         text=True, timeout=900)
    assert f"{marker} PASS" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]


@pytest.mark.skipif(not _reference_present(), reason="reference sources not staged (python oracle/build_ref.py)")
def test_reference_object_level_training_loop_on_dropin():
    _run("dropin_object.py", "DROPIN_OBJECT", "24")


@pytest.mark.skipif(not _reference_present(), reason="reference sources not staged (python oracle/build_ref.py)")
def test_reference_ssr_trainer_step_and_render_path_on_dropin():
    _run("dropin_ssr.py", "DROPIN_SSR", "24")
