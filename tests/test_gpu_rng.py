"""Training-mode random draws generated inside the stage kernels (include/inrf.h: inrf_*_rng).

The reference draws t_rand / u / sigma noise with torch.rand / torch.randn (run_nerf.py:478, run_nerf_helpers.py:414,
run_nerf.py:387).  Exact streams cannot be matched (and the reference does not pin them either - its tests inject the
numbers), so parity here is distributional plus the properties the training step relies on: a draw is a pure function of
(seed, tensor, element) - reproducible, and the backward sees the forward's noise - and the three tensors of one step are
independent streams."""
import math

import pytest
import torch

from oracle import nerf_oracle as orc
from tests.util import build_nets

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_stratified_jitter_in_kernel():
    from intrinsicnerf_b200 import ops
    rays = orc.blender_rays(64, 64).to(DEV)                       # 4096 rays
    z0 = ops.coarse_z(rays, 64)
    za, zb, zc = ops.coarse_z(rays, 64, seed=11), ops.coarse_z(rays, 64, seed=11), ops.coarse_z(rays, 64, seed=12)
    assert torch.equal(za, zb) and not torch.equal(za, zc)
    mid = 0.5 * (z0[:, 1:] + z0[:, :-1])
    lower, upper = torch.cat([z0[:, :1], mid], -1), torch.cat([mid, z0[:, -1:]], -1)
    assert bool((za >= lower).all()) and bool((za <= upper).all())       # every sample stays in its stratum (run_nerf.py:474-486)
    t = ((za - lower) / (upper - lower))[:, 1:-1].double()                # the implied U[0,1) draws (inner strata)
    n = t.numel()
    assert abs(float(t.mean()) - 0.5) < 4 / math.sqrt(12 * n) and abs(float(t.var()) - 1 / 12) < 2e-3
    assert abs(float(torch.corrcoef(torch.stack([t[:, :-1].reshape(-1), t[:, 1:].reshape(-1)]))[0, 1])) < 0.01   # neighbours independent
    # same values as the injected-tensor path when fed the draws it implies (exactly representable check on one stratum)
    zi = ops.coarse_z(rays, 64, t_rand=((za - lower) / (upper - lower).clamp_min(1e-30)).clamp(0, 1))
    assert float((zi - za).abs().max()) < 1e-5


def test_sample_pdf_uniform_draws_in_kernel():
    from intrinsicnerf_b200 import ops
    N = 2048
    bins = torch.linspace(2.0, 6.0, 63, device=DEV).expand(N, 63).contiguous()
    w = torch.ones(N, 62, device=DEV)
    a, _, _ = ops.sample_pdf(bins, w, 128, seed=5)
    b, _, _ = ops.sample_pdf(bins, w, 128, seed=5)
    c, _, _ = ops.sample_pdf(bins, w, 128, seed=6)
    assert torch.equal(a, b) and not torch.equal(a, c)
    u = ((a - 2.0) / 4.0).double()                                          # uniform pdf: samples = 2 + 4 u
    assert bool((u >= 0).all()) and bool((u < 1).all())
    assert abs(float(u.mean()) - 0.5) < 4 / math.sqrt(12 * u.numel()) and abs(float(u.var()) - 1 / 12) < 2e-3
    # the jitter stream and the u stream of the same seed are different streams
    rays = orc.blender_rays(46, 46)[:N].contiguous().to(DEV)
    z = ops.coarse_z(rays, 128, seed=5)
    z0 = ops.coarse_z(rays, 128)
    assert abs(float(torch.corrcoef(torch.stack([(z - z0).reshape(-1).double(), (u - 0.5).reshape(-1)]))[0, 1])) < 0.01


def test_sigma_noise_in_kernel_is_gaussian_and_shared_by_forward_and_backward():
    from intrinsicnerf_b200 import ops
    N, S = 4096, 64
    raw = torch.zeros(N, S, 11, device=DEV)
    raw[..., :3] = 0.5
    z = torch.arange(S, device=DEV, dtype=torch.float32).expand(N, S).contiguous() + 2.0     # dist = 1 with |d| = 1
    d = torch.tensor([0.0, 0.0, 1.0], device=DEV).expand(N, 3).contiguous()
    rec, w = ops.composite(raw, z, d, rng=(1.0, 99, False))
    rec2, w2 = ops.composite(raw, z, d, rng=(1.0, 99, False))
    rec3, w3 = ops.composite(raw, z, d, rng=(1.0, 99, True))
    assert torch.equal(w, w2) and not torch.equal(w, w3)                                     # coarse / fine noise: different streams
    relu_n = -torch.log1p(-w[:, 0].double())                                                # first sample: w = 1 - exp(-relu(noise))
    assert abs(float((relu_n > 0).double().mean()) - 0.5) < 0.03                             # P(noise > 0) = 1/2
    assert abs(float(relu_n.mean()) - 1 / math.sqrt(2 * math.pi)) < 0.02                     # E relu(N(0,1)) = 0.3989
    assert abs(float((relu_n ** 2).mean()) - 0.5) < 0.03                                     # E relu(N(0,1))^2 = 1/2
    # backward regenerates the same noise: analytic gradient == finite difference of the forward with the same seed
    raw = torch.randn(8, S, 11, device=DEV) * 0.3
    raw[..., 3] = raw[..., 3] + 0.5
    zz, dd = z[:8].contiguous(), d[:8].contiguous()
    g_rec = torch.randn(8, 13, device=DEV)
    g_rec[:, 3] = 0
    x = raw.clone().requires_grad_(True)
    out, _ = ops.composite(x, zz, dd, rng=(0.7, 1234, False))
    (out * g_rec).sum().backward()
    for (r, s_, ch) in ((0, 5, 3), (3, 20, 3), (5, 0, 1), (7, 63, 3)):
        eps = 1e-2
        p, m = raw.clone(), raw.clone()
        p[r, s_, ch] += eps
        m[r, s_, ch] -= eps
        fp = (ops.composite(p, zz, dd, rng=(0.7, 1234, False))[0] * g_rec).sum()
        fm = (ops.composite(m, zz, dd, rng=(0.7, 1234, False))[0] * g_rec).sum()
        fd = float((fp - fm) / (2 * eps))
        assert abs(fd - float(x.grad[r, s_, ch])) < 2e-2 * max(1.0, abs(fd)), (r, s_, ch, fd, float(x.grad[r, s_, ch]))


def test_ssr_training_step_launches_no_generator_kernels_and_is_reproducible():
    """SSRRenderer in training mode (perturb = 1, raw_noise_std = 1: SSR_room0_config.yaml:29,34) under autograd: two
    steps from the same torch seed give identical losses; the seed sequence advances between steps."""
    from intrinsicnerf_b200 import ssr
    C = 5
    coarse, fine, _, _ = build_nets("ssr", C)

    class T(ssr.SSRRenderer):
        pass
    t = T()
    t.N_samples, t.N_importance, t.perturb, t.raw_noise_std, t.training = 64, 128, 1.0, 1.0, True
    t.white_bkgd, t.enable_semantic, t.num_valid_semantic_class, t.endpoint_feat = False, True, C, False
    t.netchunk = t.chunk = 32768
    t.ssr_net_coarse, t.ssr_net_fine = coarse, fine
    t.embed_fn, _ = ssr.get_embedder(10, 0, scalar_factor=10)
    t.embeddirs_fn, _ = ssr.get_embedder(4, 0, scalar_factor=1)
    rays = orc.replica_rays(12, 16).to(DEV)
    vals = []
    for seed in (3, 3, 4):
        torch.manual_seed(seed)
        a = float(t.render_rays(rays)["rgb_fine"].sum())
        b = float(t.render_rays(rays)["rgb_fine"].sum())
        vals.append((a, b))
    assert vals[0] == vals[1] and vals[0] != vals[2] and vals[0][0] != vals[0][1]


def test_ssr_eval_render_is_fused_and_raw_is_lazy():
    """Eval-mode volumetric_rendering returns raw_coarse / raw_fine as the reference does (trainer.py:777-802), but
    produces them on first access: the render itself is the two fused launches."""
    from intrinsicnerf_b200 import ops, ssr
    C = 28
    coarse, fine, _, _ = build_nets("ssr", C)

    class T(ssr.SSRRenderer):
        pass
    t = T()
    t.N_samples, t.N_importance, t.perturb, t.raw_noise_std, t.training = 64, 128, 1.0, 1.0, False
    t.white_bkgd, t.enable_semantic, t.num_valid_semantic_class, t.endpoint_feat = False, True, C, False
    t.netchunk = t.chunk = 32768
    t.ssr_net_coarse, t.ssr_net_fine = coarse, fine
    t.embed_fn, _ = ssr.get_embedder(10, 0, scalar_factor=10)
    t.embeddirs_fn, _ = ssr.get_embedder(4, 0, scalar_factor=1)
    rays = orc.replica_rays(24, 32).to(DEV)
    coarse.packed(), fine.packed()
    with torch.no_grad():
        n0 = ops.launch_count()
        out = t.render_rays(rays)
        assert ops.launch_count() - n0 == 2
        assert "raw_fine" in out and "raw_coarse" in out and len(out) == 19
        rgb = out["rgb_fine"]
        assert ops.launch_count() - n0 == 2
        raw = out["raw_fine"]                                       # first access: one stage-path render
        assert raw.shape == (rays.shape[0], 192, 11 + C) and out["raw_coarse"].shape == (rays.shape[0], 64, 11 + C)
        want = ops.render_chunk(rays, coarse.packed(), fine.packed(), variant=1, n_classes=C, pe_scalar_factor=10.0, want_raw=True)
    torch.cuda.synchronize()
    assert torch.equal(raw, want["raw_fine"]) and torch.equal(rgb, want["rec_fine"][:, 0:3])


@pytest.mark.gpu
def test_cuda_graph_replays_draw_fresh_numbers():
    """A captured step bakes its seeds into the graph; the device-resident epoch (inrf_rng_epoch_bump, captured as the
    first node) makes every replay draw new jitter / noise, the backward of a replay regenerates that replay's forward
    noise, and resetting the epoch reproduces the sequence."""
    from intrinsicnerf_b200 import ops
    dev = torch.device("cuda:0")
    N, S, C = 256, 64, 0
    rays = orc.blender_rays(32, 32)[:N].contiguous().to(dev)
    raw = torch.randn(N, S, 11, device=dev)
    g_rec = torch.randn(N, 13, device=dev)
    seed = 1234567
    ops.rng_epoch_reset()
    z_eager = ops.coarse_z(rays, S, False, None, seed)                    # epoch 0: key = seed
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    out = {}

    def step():
        ops.rng_epoch_bump()
        out["z"] = ops.coarse_z(rays, S, False, None, seed)
        rec, w = ops.raw2outputs_rec(raw, out["z"], rays[:, 3:6], None, False, C, False, True, (1.0, seed, False))
        out["rec"] = rec
        out["graw"] = ops.raw2outputs_bwd(raw, out["z"], rays[:, 3:6], None, g_rec, None, False, C, False, (1.0, seed, False))
    with torch.cuda.stream(side):
        step()                                                            # warm-up outside the capture (epoch 1)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        step()
    runs = []
    for _ in range(3):                                                    # epochs 2, 3, 4
        graph.replay()
        torch.cuda.synchronize()
        runs.append({k: v.clone() for k, v in out.items()})
    assert not torch.equal(runs[0]["z"], runs[1]["z"]) and not torch.equal(runs[1]["z"], runs[2]["z"])
    assert not torch.equal(runs[0]["z"], z_eager)
    assert not torch.equal(runs[0]["rec"], runs[1]["rec"])
    # reproducible: the same epochs give the same numbers
    ops.rng_epoch_reset()
    ops.rng_epoch_bump()                                                  # epoch 1 (the warm-up's)
    for i in range(3):
        graph.replay()
        torch.cuda.synchronize()
        for k in out:
            assert torch.equal(out[k], runs[i][k]), (i, k)
    # the backward of a replay used that replay's forward noise: an eager backward with explicit noise differs per epoch,
    # and two replays' gradients differ
    assert not torch.equal(runs[0]["graw"], runs[1]["graw"])
    ops.rng_epoch_reset()
    assert torch.equal(ops.coarse_z(rays, S, False, None, seed), z_eager)
    ops.poll_status()
