"""Training converges the same way on the tensor-core path and on the strict-fp32 path.

A fixed batch of rays is fitted to a smooth synthetic target (photometric loss on the fine and coarse maps plus the
intrinsic terms, Adam, lr 5e-4 - the loop of object_level/run_nerf.py:942-1027 on a fixed batch) for 150 steps, once
with ops.MlpTcFn (fp16 operands in the forward AND backward GEMMs, default) and once with ops.MlpFn (fp32 everywhere),
from identical initial weights.  Both must reduce the loss substantially and end within a few percent of each other:
the operand rounding of the tensor-core path acts as small unbiased noise on the gradients, not as a bias."""
import pytest
import torch

from oracle import nerf_oracle as orc
from tests.util import build_nets

pytestmark = pytest.mark.gpu


def _fit(precision, steps=150, n=384):
    from intrinsicnerf_b200 import object_level as ol, ops
    ops.set_default_precision(precision)
    try:
        coarse, fine, _, _ = build_nets("object")
        e, _ = ol.get_embedder(10, 0)
        ed, _ = ol.get_embedder(4, 0)
        kw = dict(network_fn=coarse, network_fine=fine, network_query_fn=ol._FusedQuery(e, ed, 65536), N_samples=64, N_importance=128,
                  perturb=0., white_bkgd=True, raw_noise_std=0.)
        rays = orc.blender_rays(32, 32)[torch.randperm(1024, generator=torch.Generator().manual_seed(3))[:n]].contiguous().cuda()
        d = torch.nn.functional.normalize(rays[:, 3:6], dim=-1)
        target = torch.stack([0.5 + 0.4 * d[:, 0], 0.4 + 0.3 * d[:, 1], 0.6 + 0.3 * d[:, 0] * d[:, 1]], -1).clamp(0, 1)
        mask = torch.ones(n, device="cuda")
        opt = torch.optim.Adam(list(coarse.parameters()) + list(fine.parameters()), lr=5e-4)
        losses = []
        for _ in range(steps):
            opt.zero_grad(set_to_none=True)
            out = ol.render_rays(rays, **kw)
            loss = ((out["rgb_map"] - target) ** 2).mean() + ((out["rgb0"] - target) ** 2).mean()
            terms = ops.intrinsic_losses(None, out["albedo_map"], out["shading_map"], out["residual_map"], target, mask, None, "object")
            loss = loss + terms[1] + 0.02 * terms[3] + terms[4] + 0.1 * terms[6] + terms[2]
            loss.backward()
            opt.step()
            losses.append(float(loss))
        torch.cuda.synchronize()
        ops.poll_status()
        return losses
    finally:
        ops.set_default_precision("tc")


def test_tc_and_fp32_training_converge_alike():
    tc = _fit("tc")
    fp32 = _fit("fp32")
    print("tc   loss: start %.4f  step 50 %.4f  end %.4f" % (tc[0], tc[50], tc[-1]))
    print("fp32 loss: start %.4f  step 50 %.4f  end %.4f" % (fp32[0], fp32[50], fp32[-1]))
    assert abs(tc[0] - fp32[0]) < 1e-3 * fp32[0]                       # same starting point
    end_tc, end_32 = sum(tc[-10:]) / 10, sum(fp32[-10:]) / 10
    assert end_tc < 0.5 * tc[0] and end_32 < 0.5 * fp32[0]             # both learn
    assert abs(end_tc - end_32) < 0.08 * end_32, (end_tc, end_32)      # and arrive at the same place
