"""GPU parity of the tensor-core training path (inrf_mlp_fwd_train_tc / inrf_mlp_bwd_tc through ops.MlpTcFn).

Forward: raw rows against the fp32 oracle; every stashed activation tile against the oracle evaluated in the kernels'
arithmetic (oracle.mlp_forward_tc_arith: fp16-rounded GEMM operands, exact sums).
Backward: parameter gradients against float64 autograd through that oracle taking the SAME ReLU branches as the
kernel (the 0/1 masks are read back from the kernel's own stash).  The network is piecewise linear, so this is the exact
gradient for those branch decisions; without pinning them, fp32-vs-fp64 accumulation flips a handful of near-zero units
per thousand samples and each flip is a full-size difference in one sample's gradient.
Bars: raw values 1e-3 relative with a 0.25 floor (the rendered maps' 1e-4 bar is test_gpu_render.py); stashed activations 2e-3
absolute (one fp16 ulp at 2.0); gradients 2.5e-3 of each parameter's largest gradient - the operand rounding of the
backward GEMMs (2^-11 per operand, averaged over the batch; one-element biases have no larger entry to be measured against)."""
import pytest
import torch

from oracle import nerf_oracle as orc
from tests.util import build_nets, rel_err, stash_activations, stash_mask_bits

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _inputs(variant, C, endpoint, M, seed, g_scale):
    coarse, fine, pc, pf = build_nets(variant, C)
    gen = torch.Generator().manual_seed(seed)
    pts = torch.rand(M, 3, generator=gen) * 6 - 3
    vd = torch.nn.functional.normalize(torch.randn(M, 3, generator=gen), dim=-1)
    scale = 1.0 if variant == "object" else 10.0
    ch = 11 + C + (128 if endpoint else 0)
    g_raw = torch.randn(M, ch, generator=gen) * g_scale
    emb = torch.cat([orc.posenc(pts.double(), 10, scale), orc.posenc(vd.double(), 4)], -1)
    return fine, pf, pts, vd, scale, g_raw, emb


@pytest.mark.parametrize("variant,C,endpoint,M,g_scale", [("object", 0, False, 1000, 1.0), ("ssr", 28, True, 2390, 1e-4),
                                                           ("ssr", 28, False, 128, 3e3), ("ssr", 5, False, 150, 1.0),
                                                           ("ssr", 0, False, 77, 1e-7), ("ssr", 112, False, 300, 1.0)])
def test_tc_training_forward_stash_and_gradients(variant, C, endpoint, M, g_scale):
    """g_scale spans 1e-7 .. 3e3: the device-side power-of-two gradient scaling keeps fp16 in range either way."""
    from intrinsicnerf_b200 import ops
    assert ops.default_precision() == ops.PREC_TC
    fine, pf, pts, vd, scale, g_raw, emb = _inputs(variant, C, endpoint, M, 21, g_scale)
    fine.zero_grad()
    out = fine.evaluate("pts", pts.to(DEV), vd.to(DEV), endpoint, scale)
    assert out.grad_fn is not None and type(out.grad_fn).__name__.startswith("MlpTcFn")
    with torch.no_grad():
        want32 = orc.mlp_forward(pf, emb.float(), variant, C, endpoint)
    assert rel_err(out.detach(), want32, floor=0.25) < 1e-3         # |error| < 2.5e-4 on O(1) raw rows (fp16 operands)
    # ---- stash: every activation tile of the forward, as the backward will read it ----------------------------------
    act = stash_activations(out.grad_fn.saved_tensors[3], M, C)
    names = orc.OBJECT_HEADS if variant == "object" else orc.SSR_HEADS
    # the bit words the dX epilogue reads are exactly "stored activation > 0" of the images the dW GEMM reads
    mb = stash_mask_bits(out.grad_fn.saved_tensors[3], M)
    imgs = torch.cat([act[f"h{l}"] for l in range(8)] + [act["v"], act["as"]] + ([act["s1"]] if C > 0 else []), 1)
    assert torch.equal(mb[:, :imgs.shape[1]], imgs > 0)
    masks = [act[f"h{l}"] > 0 for l in range(8)] + [act["as"][:, :128] > 0, act["as"][:, 128:] > 0, act["v"] > 0]
    if C > 0:
        masks.append(act["s1"] > 0)
    p64 = {k: v.double().requires_grad_(True) for k, v in pf.items()}
    out64, _ = orc.mlp_forward_tc_arith(p64, emb, variant, C, endpoint, masks=masks)
    with torch.no_grad():
        free, _ = orc.mlp_forward_tc_arith(p64, emb, variant, C, endpoint)          # its own branch decisions
        assert float((free - out64).abs().max()) < 1e-3                                  # the kernel's masks are (nearly) the oracle's
        h = emb[:, :63]
        assert float((act["pe"][:, :63].double() - h).abs().max()) < 2e-3 and float(act["pe"][:, 63].abs().max()) == 0.0
        rnd = lambda x: x.to(torch.float16).double()  # noqa: E731
        n_flip = n_all = 0
        for l, name in enumerate(orc.TRUNK):
            zl = rnd(h) @ rnd(p64[name + ".weight"]).t() + p64[name + ".bias"]
            # independence of the branch decisions: where the kernel's mask (read from its stash) differs from the
            # oracle's own sign(z), the oracle's pre-activation sits on the kink - within the fp32-vs-fp64 accumulation
            # error of zero - so the masked-oracle gradient below is the exact gradient of a function that agrees
            # with the free oracle everywhere except on a measure-zero-like set of (sample, unit) pairs
            differ = (zl > 0) != masks[l]
            n_flip, n_all = n_flip + int(differ.sum()), n_all + differ.numel()
            if bool(differ.any()):
                assert float(zl[differ].abs().max()) < 1e-4 * max(1.0, float(zl.abs().max())), (l, float(zl[differ].abs().max()))
            h = torch.relu(zl)
            assert float((act[f"h{l}"].double() - h).abs().max()) < 2e-3, l
            if l == 4:
                h = torch.cat([emb[:, :63], h], -1)
        a1 = torch.relu(rnd(h) @ rnd(p64[names["albedo1"] + ".weight"]).t() + p64[names["albedo1"] + ".bias"])
        assert float((act["as"][:, :128].double() - a1).abs().max()) < 2e-3
        assert n_flip <= 1e-4 * n_all, (n_flip, n_all)        # < 0.01 % of the trunk's (sample, unit) pairs
    # ---- backward ------------------------------------------------------------------------------------------------------
    (out64 * g_raw.double()).sum().backward()
    (out * g_raw.to(DEV)).sum().backward()
    errs = {}
    for name, p in fine.named_parameters():
        a, b = p64[name].grad.float(), p.grad.cpu()
        assert torch.isfinite(b).all(), name
        errs[name] = float((a - b).abs().max()) / (float(a.abs().max()) + 1e-30)
    assert max(errs.values()) < 2.5e-3, sorted(errs.items(), key=lambda kv: -kv[1])[:6]
    # a second backward accumulates into .grad (zero_grad is the caller's business, as in PyTorch)
    out2 = fine.evaluate("pts", pts.to(DEV), vd.to(DEV), endpoint, scale)
    (out2 * g_raw.to(DEV)).sum().backward()
    name, p = next(iter(fine.named_parameters()))
    ref = 2 * p64[name].grad.float()
    assert float((p.grad.cpu() - ref).abs().max()) < 3e-3 * float(ref.abs().max())


def test_tc_training_zero_and_nonfinite_gradients():
    """All-zero upstream gradient -> exactly zero parameter gradients; an inf in grad_raw must not hang or poison the
    scale selection (the affected sample's gradients are non-finite, as autograd's would be)."""
    fine, pf, pts, vd, scale, g_raw, emb = _inputs("object", 0, False, 300, 5, 1.0)
    fine.zero_grad()
    out = fine.evaluate("pts", pts.to(DEV), vd.to(DEV), False, scale)
    (out * 0.0).sum().backward()
    assert all(float(p.grad.abs().max()) == 0.0 for p in fine.parameters())
    g = g_raw.clone()
    g[7, 2] = float("inf")
    out = fine.evaluate("pts", pts.to(DEV), vd.to(DEV), False, scale)
    (out * g.to(DEV)).sum().backward()
    torch.cuda.synchronize()
    # ... and the non-finite weight gradient is REPORTED (deferred status, INRF_ERANGE) instead of silently reaching Adam
    from intrinsicnerf_b200 import ops
    from intrinsicnerf_b200._lib import InrfRangeError
    with pytest.raises(InrfRangeError, match="non-finite weight gradient"):
        ops.poll_status()
    ops.poll_status()


def test_rays_mode_and_embedded_mode_agree():
    """The three addressing modes of the training forward feed the same backward."""
    from intrinsicnerf_b200 import ops
    coarse, fine, pc, pf = build_nets("object")
    rays = orc.blender_rays(4, 4).to(DEV)
    z = torch.linspace(2.0, 6.0, 64, device=DEV).expand(16, 64).contiguous()
    g = torch.randn(16 * 64, 11, generator=torch.Generator().manual_seed(0)).to(DEV)
    grads = []
    for mode in ("rays", "pts"):
        fine.zero_grad()
        if mode == "rays":
            out = fine.evaluate("rays", rays, z)
        else:
            pts = rays[:, None, 0:3] + rays[:, None, 3:6] * z[:, :, None]
            out = fine.evaluate("pts", pts.reshape(-1, 3), rays[:, None, 8:11].expand(16, 64, 3).reshape(-1, 3))
        (out * g).sum().backward()
        grads.append(torch.cat([p.grad.reshape(-1) for p in fine.parameters()]).clone())
    assert float((grads[0] - grads[1]).abs().max()) < 2e-3 * float(grads[1].abs().max())
    assert ops.default_precision() == ops.PREC_TC


def test_tc_training_gradients_are_bit_reproducible():
    """No atomics anywhere in the tensor-core backward (split-K partial tiles are added in a fixed order): two runs on the
    same inputs give identical parameter gradients, bit for bit."""
    fine, pf, pts, vd, scale, g_raw, emb = _inputs("ssr", 28, False, 3000, 9, 1.0)
    runs = []
    for _ in range(2):
        fine.zero_grad()
        out = fine.evaluate("pts", pts.to(DEV), vd.to(DEV), False, scale)
        (out * g_raw.to(DEV)).sum().backward()
        runs.append(torch.cat([p.grad.reshape(-1) for p in fine.parameters()]).clone())
    assert torch.equal(runs[0], runs[1])
