"""Deferred device status (include/inrf.h: inrf_poll_status): conditions only a running kernel can see are
reported at the next call boundary instead of passing garbage on.

  * a stuck barrier in the tensor-core kernel -> INRF_ECUDA from the following call, later launches healthy;
  * weights / hidden activations outside the fp16 range of INRF_PREC_TC -> INRF_ERANGE, outputs saturated (finite),
    never inf/NaN; the strict-fp32 path renders the same network correctly.
The reference computes in fp32 (run_nerf_helpers.py:284-325), so these are the conditions under which the
fp16-operand path must refuse instead of silently disagreeing with it."""
import os
import subprocess
import sys

import pytest
import torch

from oracle import nerf_oracle as orc
from tests.util import build_nets, rec_get, rel_err

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def _sample_inputs(n=640, seed=0):
    g = torch.Generator().manual_seed(seed)
    pts = torch.rand(n, 3, generator=g) * 4 - 2
    vd = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1)
    return pts, vd


def test_healthy_run_reports_nothing(dev):
    from intrinsicnerf_b200 import ops
    coarse, fine, pc, pf = build_nets("object")
    rays = orc.blender_rays(8, 8).to(dev)
    ops.render_chunk(rays, coarse.packed(), fine.packed(), white_bkgd=True)
    torch.cuda.synchronize()
    ops.poll_status()            # raises if anything was recorded


def test_activation_overflow_is_reported_not_inf(dev):
    """A trunk scaled so that hidden activations pass 65504: tc launch saturates (finite raw), the next call raises
    InrfRangeError, the record is cleared, the fp32 path still matches the oracle on the same weights."""
    from intrinsicnerf_b200 import ops
    from intrinsicnerf_b200._lib import InrfRangeError
    coarse, fine, pc, pf = build_nets("object")
    with torch.no_grad():
        for i in (1, 2, 3):
            fine.pts_linears[i].weight *= 100.0          # activations grow ~1e6 x
            pf[f"pts_linears.{i}.weight"] *= 100.0
    pts, vd = _sample_inputs()
    emb = torch.cat([orc.posenc(pts, 10), orc.posenc(vd, 4)], -1)
    want = orc.mlp_forward(pf, emb)
    raw = ops.mlp_forward(fine.packed(), 0, 0, pts.to(dev), vd.to(dev), precision="tc")
    torch.cuda.synchronize()
    assert torch.isfinite(raw).all(), "tensor-core path produced inf/NaN instead of saturating"
    with pytest.raises(InrfRangeError, match="fp16 limit"):
        ops.poll_status()
    ops.poll_status()                                     # cleared
    raw32 = ops.mlp_forward(fine.packed(), 0, 0, pts.to(dev), vd.to(dev), precision="fp32")
    torch.cuda.synchronize()
    ops.poll_status()
    assert rel_err(raw32.cpu(), want, floor=1e-2 * float(want.abs().max())) < 1e-3


def test_activation_overflow_surfaces_at_next_call_without_sync(dev):
    """The hot entry points poll on entry: after the offending launch has finished, the NEXT library call returns
    INRF_ERANGE without the caller ever calling poll_status."""
    from intrinsicnerf_b200 import ops
    from intrinsicnerf_b200._lib import InrfRangeError
    coarse, fine, _, _ = build_nets("object")
    with torch.no_grad():
        for i in (1, 2, 3):
            fine.pts_linears[i].weight *= 100.0
    pts, vd = _sample_inputs(256)
    ops.mlp_forward(fine.packed(), 0, 0, pts.to(dev), vd.to(dev), precision="tc")
    torch.cuda.synchronize()                              # the kernel has run; nobody polled
    with pytest.raises(InrfRangeError):
        ops.mlp_forward(coarse.packed(), 0, 0, pts.to(dev), vd.to(dev), precision="tc")
    out = ops.mlp_forward(coarse.packed(), 0, 0, pts.to(dev), vd.to(dev), precision="tc")   # healthy network: fine again
    torch.cuda.synchronize()
    ops.poll_status()
    assert torch.isfinite(out).all()


def test_weight_overflow_and_underflow_at_pack(dev):
    from intrinsicnerf_b200 import ops
    from intrinsicnerf_b200._lib import InrfRangeError
    coarse, fine, _, _ = build_nets("object")
    with torch.no_grad():
        fine.pts_linears[2].weight[5, 7] = 1.0e6          # > 65504
    fine.packed()
    torch.cuda.synchronize()
    with pytest.raises(InrfRangeError, match="exceeds the fp16 limit"):
        ops.poll_status()
    with torch.no_grad():
        coarse.pts_linears[3].weight *= 1e-7              # |w| ~ 6e-9: below fp16's subnormal range
    coarse.packed()
    torch.cuda.synchronize()
    with pytest.raises(InrfRangeError, match="below 2\\^-17"):
        ops.poll_status()


@pytest.mark.parametrize("scale", [1.0 / 16.0, 16.0])
def test_moderately_scaled_weights_stay_in_range(dev, scale):
    """Weights 16x larger / smaller than the default init on two layers (compensated on the next, so the function is
    unchanged up to fp32 rounding): no status record, tc raw within the usual bound of the oracle."""
    from intrinsicnerf_b200 import ops
    coarse, fine, pc, pf = build_nets("object")
    with torch.no_grad():
        for i, s in ((2, scale), (3, 1.0 / scale)):
            fine.pts_linears[i].weight *= s
            pf[f"pts_linears.{i}.weight"] *= s
        fine.pts_linears[2].bias *= scale
        pf["pts_linears.2.bias"] *= scale
    pts, vd = _sample_inputs()
    emb = torch.cat([orc.posenc(pts, 10), orc.posenc(vd, 4)], -1)
    want = orc.mlp_forward(pf, emb)
    raw = ops.mlp_forward(fine.packed(), 0, 0, pts.to(dev), vd.to(dev), precision="tc")
    torch.cuda.synchronize()
    ops.poll_status()
    assert float((raw.cpu() - want).abs().max()) < 2e-3 * max(1.0, float(want.abs().max()))


_WATCHDOG_SCRIPT = r"""
import sys, torch
sys.path.insert(0, %r)
from tests.util import build_nets
from intrinsicnerf_b200 import ops
from intrinsicnerf_b200._lib import InrfError, InrfRangeError
dev = torch.device("cuda:0")
coarse, fine, _, _ = build_nets("object")
pts = torch.rand(1024, 3, device=dev); vd = torch.nn.functional.normalize(torch.randn(1024, 3, device=dev), dim=-1)
p = fine.packed()
ops.mlp_forward(p, 0, 0, pts, vd, precision="tc")      # launch 1: the producer starves the ring (INRF_TC_FAULT=1)
torch.cuda.synchronize()
try:
    ops.mlp_forward(p, 0, 0, pts, vd, precision="tc")  # call 2: must report the tripped watchdog, launches nothing
    print("NO_ERROR")
except InrfRangeError as e:
    print("WRONG_CLASS", e)
except InrfError as e:
    print("GOT_ECUDA" if "watchdog" in str(e) else "OTHER", e)
ops.poll_status()
print("CLEARED")
good = ops.mlp_forward(p, 0, 0, pts, vd, precision="tc")   # launch 3, same process: healthy again (claim words are per launch)
ref = ops.mlp_forward(p, 0, 0, pts, vd, precision="fp32")
torch.cuda.synchronize()
ops.poll_status()
print("HEALTHY_AFTER" if float((good - ref).abs().max()) < 2e-3 else "POISONED", float((good - ref).abs().max()))
"""


def test_stuck_barrier_is_reported_and_does_not_poison_later_launches():
    """INRF_TC_FAULT=1 makes the weight producer of the FIRST tensor-core launch of the process stop after three ring
    fills; the MMA issuer's wait exceeds the (shortened) watchdog, every role runs out, the launch returns, and the
    next call gets INRF_ECUDA.  The launch after that runs in the same process and must be correct: the device-side
    abort word is cleared before every launch (round 1 left it set, so one trip poisoned the process).  The env
    switches are read once per process, hence the subprocess."""
    env = dict(os.environ, INRF_TC_FAULT="1", INRF_TC_WATCHDOG_CYCLES="200000000")
    out = subprocess.run([sys.executable, "-c", _WATCHDOG_SCRIPT % ROOT], env=env, cwd=ROOT, capture_output=True, text=True,
                         timeout=300)
    assert "GOT_ECUDA" in out.stdout and "CLEARED" in out.stdout and "HEALTHY_AFTER" in out.stdout, \
        out.stdout[-2000:] + out.stderr[-2000:]
