"""GPU parity: every stage kernel against the CPU oracle (and the reference's golden vectors)."""
import numpy as np
import pytest
import torch

from oracle import nerf_oracle as orc
from intrinsicnerf_b200 import ops as ops_mod
from tests.util import build_nets, load_golden, rel_err, relu_margin

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def test_embed_matches_oracle_and_golden(dev, golden_dir):
    from intrinsicnerf_b200 import ops
    g = load_golden(golden_dir, "object_stages.npz")
    x = torch.from_numpy(g["x"])
    out = ops.embed(x.to(dev), 10).cpu()
    # |x*2^9| reaches ~2e3 rad: CUDA sinf/cosf (<=2 ulp) vs ATen's vectorised sin (<=1 ulp)
    assert np.abs(out.numpy() - g["emb_pts"]).max() < 5e-7
    d = x / x.norm(dim=-1, keepdim=True)
    assert np.abs(ops.embed(d.to(dev), 4).cpu().numpy() - g["emb_dir"]).max() < 5e-7
    out10 = ops.embed(x.to(dev), 10, 10.0).cpu()
    assert (out10 - orc.posenc(x, 10, 10.0)).abs().max() < 5e-7
    assert ops.embed(torch.zeros(0, 3, device=dev), 10).shape == (0, 63)      # empty input


def test_coarse_z_bitexact(dev):
    from intrinsicnerf_b200 import ops
    rays = orc.blender_rays(5, 7)
    rays[:, 6] = torch.linspace(0.5, 2.0, rays.shape[0])
    g = torch.Generator().manual_seed(0)
    t_rand = torch.rand(rays.shape[0], 64, generator=g)
    for lindisp in (False, True):
        for tr in (None, t_rand):
            want = orc.coarse_z(rays[:, 6:7], rays[:, 7:8], 64, lindisp, tr)
            got = ops.coarse_z(rays.to(dev), 64, lindisp, None if tr is None else tr.to(dev)).cpu()
            assert torch.equal(got, want), (lindisp, tr is None, (got - want).abs().max())


@pytest.mark.parametrize("variant,C", [("object", 0), ("ssr", 28), ("ssr", 0)])
def test_mlp_fp32_matches_oracle(dev, variant, C):
    from intrinsicnerf_b200 import ops
    coarse, fine, pc, pf = build_nets(variant, C)
    g = torch.Generator().manual_seed(5)
    M = 200                                            # not a multiple of the 64-row tile
    pts = torch.rand(M, 3, generator=g) * 8 - 4
    vd = torch.randn(M, 3, generator=g)
    vd = vd / vd.norm(dim=-1, keepdim=True)
    scale = 1.0 if variant == "object" else 10.0
    emb = torch.cat([orc.posenc(pts, 10, scale), orc.posenc(vd, 4)], -1)
    for endpoint in ([False] if variant == "object" else [False, True]):
        want = orc.mlp_forward(pf, emb, variant, C, endpoint)
        got = ops.mlp_forward(fine.packed(), fine.variant, C, pts.to(dev), vd.to(dev), endpoint, scale, "fp32").cpu()
        assert got.shape == want.shape
        # semantic logits are unbounded and cross zero: error relative to max(|ref|, 1e-2)
        assert rel_err(got, want, floor=1e-2) < 2e-5, rel_err(got, want, floor=1e-2)
        got_e = ops.mlp_forward_embedded(fine.packed(), fine.variant, C, emb.to(dev), endpoint, "fp32").cpu()
        assert rel_err(got_e, want, floor=1e-2) < 2e-5


def test_module_forward_and_run_network(dev):
    """NeRF.forward on embedded rows and run_network on points give the same rows."""
    from intrinsicnerf_b200 import object_level as ol, ops
    ops.set_default_precision("fp32")
    try:
        coarse, fine, pc, pf = build_nets("object")
        embed_fn, _ = ol.get_embedder(10, 0)
        embeddirs_fn, _ = ol.get_embedder(4, 0)
        g = torch.Generator().manual_seed(9)
        pts = (torch.rand(7, 64, 3, generator=g) * 6 - 3).to(dev)
        vd = torch.nn.functional.normalize(torch.randn(7, 3, generator=g), dim=-1).to(dev)
        with torch.no_grad():
            a = ol.run_network(pts, vd, coarse, embed_fn, embeddirs_fn)
            emb = torch.cat([embed_fn(pts.reshape(-1, 3)), embeddirs_fn(vd[:, None].expand(pts.shape).reshape(-1, 3))], -1)
            b = coarse(emb).reshape(7, 64, 11)
        assert rel_err(a, b) < 1e-5
        want = orc.query_field(pts.cpu(), vd.cpu(), pc)
        assert rel_err(a, want) < 2e-5
        out = coarse(emb)                                 # autograd mode -> fp32 training kernels (MlpFn)
        assert out.requires_grad and rel_err(out.detach().reshape(7, 64, 11), want) < 2e-5
    finally:
        ops.set_default_precision("tc")


def test_raw2outputs_golden_and_variants(dev, golden_dir):
    from intrinsicnerf_b200 import object_level as ol, ops, ssr
    g = load_golden(golden_dir, "object_stages.npz")
    raw, z, rd = (torch.from_numpy(g[k]).to(dev) for k in ("raw", "z", "rays_d"))
    for wb in (0, 1):
        out = ol.raw2outputs(raw, z, rd, 0, bool(wb))
        for name, t in zip(("rgb", "disp", "acc", "weights", "depth", "albedo", "shading", "residual"), out):
            assert rel_err(t, g[f"wb{wb}_{name}"], floor=1e-2) < 2e-5, name
    assert torch.isnan(ol.raw2outputs(raw, z, rd)[1][5])                 # NaN disparity is preserved (A8)
    # semantic + endpoint channels, white background, noise
    gen = torch.Generator().manual_seed(2)
    C = 28
    raw2 = torch.randn(9, 192, 11 + C + 128, generator=gen)
    z2 = torch.sort(torch.rand(9, 192, generator=gen) * 9 + 0.1, dim=-1)[0]
    rd2 = torch.randn(9, 3, generator=gen)
    noise = torch.randn(9, 192, generator=gen)
    want = orc.composite(raw2, z2, rd2, noise, True, C, True)
    rec, w = ops.raw2outputs_rec(raw2.to(dev), z2.to(dev), rd2.to(dev), noise.to(dev), True, C, True)
    assert rel_err(w, want["weights"], floor=1e-2) < 1e-4        # weights live in [0,1]
    assert rel_err(rec[:, 13:13 + C], want["sem"], floor=1e-2) < 5e-5
    assert rel_err(rec[:, 13 + C:], want["feat"], floor=1e-2) < 5e-5
    assert rel_err(rec[:, 0:3], want["rgb"], floor=1e-2) < 2e-5
    assert rel_err(rec[:, 9:12], want["residual"], floor=1e-2) < 2e-5
    tup = ssr.raw2outputs(raw2[..., :11 + C].to(dev), z2.to(dev), rd2.to(dev), 0, False, True, C, False)
    want2 = orc.composite(raw2[..., :11 + C], z2, rd2, None, False, C, False)
    assert len(tup) == 10 and rel_err(tup[5], want2["sem"], floor=1e-2) < 5e-5
    assert rel_err(tup[4], want2["depth"], floor=1e-2) < 2e-5


def test_sample_pdf_indices_exact_given_cdf(dev, golden_dir):
    """for identical cdf and u the indices equal searchsorted(right=True) exactly,
    including u=0, u=1, all-zero weights and single-spike weights (appendix A7)."""
    from intrinsicnerf_b200 import ops
    g = load_golden(golden_dir, "object_stages.npz")
    bins, w, u = (torch.from_numpy(g[k]) for k in ("bins", "w", "u"))
    cdf = orc.build_cdf(w)
    for uu in (u, torch.linspace(0, 1, 128).expand(7, 128).contiguous()):
        want_s, want_i = orc.invert_cdf(bins, cdf, uu)
        got_s, got_i = ops.invert_cdf(bins.to(dev), cdf.to(dev), uu.to(dev))
        assert torch.equal(got_i.cpu(), want_i)
        assert torch.equal(got_s.cpu(), want_s)
    # the library's own cdf: golden samples reproduced to rounding, indices self-consistent
    s, inds, cdf_k = ops.sample_pdf(bins.to(dev), w.to(dev), 128, None, True, True)
    assert np.abs(s.cpu().numpy() - g["s_det"]).max() < 2e-6
    assert (cdf_k.cpu() - cdf).abs().max() < 2e-7
    assert torch.equal(inds.cpu(), torch.searchsorted(cdf_k.cpu(), torch.linspace(0, 1, 128).expand(7, 128).contiguous(), right=True))
    assert int(inds.min()) >= 1 and int(inds.max()) <= 63
    s2 = ops.sample_pdf(bins.to(dev), w.to(dev), 128, u.to(dev))[0]
    assert np.abs(s2.cpu().numpy() - g["s_rnd"]).max() < 5e-6


def test_merge_sorted_exact(dev):
    from intrinsicnerf_b200 import ops
    gen = torch.Generator().manual_seed(4)
    for Sa, Sb in ((64, 128), (64, 0), (5, 3), (1, 1), (300, 724)):
        za = torch.sort(torch.rand(11, Sa, generator=gen) * 4 + 2, dim=-1)[0]
        zb = torch.rand(11, Sb, generator=gen) * 4 + 2
        if Sb > 2:
            zb[:, 1] = zb[:, 0]                         # duplicates
        out, std = ops.merge_sorted(za.to(dev), zb.to(dev), Sb > 0)
        want = torch.sort(torch.cat([za, zb], -1), -1)[0]
        assert torch.equal(out.cpu(), want)
        if Sb > 0:
            assert rel_err(std, torch.std(zb, dim=-1, unbiased=False), floor=1e-3) < 1e-5
    # both lists ascending (what the renderer passes): the rank-merge fast path, incl. ties within and across lists
    for Sa, Sb, N in ((64, 128, 4099), (64, 0, 7), (5, 3, 9), (1, 1, 3), (300, 724, 5), (1, 128, 2)):
        za = torch.sort(torch.rand(N, Sa, generator=gen) * 4 + 2, dim=-1)[0]
        zb = torch.sort(torch.rand(N, Sb, generator=gen) * 4 + 2, dim=-1)[0]
        if Sb > 2:
            zb[:, 1] = zb[:, 0]
            zb[:, 2] = za[:, min(2, Sa - 1)]
            zb = torch.sort(zb, dim=-1)[0]
        out, std = ops.merge_sorted(za.to(dev), zb.to(dev), Sb > 0)
        assert torch.equal(out.cpu(), torch.sort(torch.cat([za, zb], -1), -1)[0]), (Sa, Sb)
    from intrinsicnerf_b200._lib import InrfError
    with pytest.raises(InrfError):
        ops.merge_sorted(torch.zeros(2, 1000, device=dev), torch.zeros(2, 100, device=dev))


def test_get_rays_bitexact(dev):
    """inrf_get_rays == get_rays + render()'s packing (run_nerf_helpers.py:359-368, run_nerf.py:100-128)."""
    from intrinsicnerf_b200 import ops
    for (H, W, theta) in ((7, 5, -180.0), (33, 40, 57.0)):
        K = orc.blender_intrinsics(H, W)
        c2w = orc.pose_spherical(theta, -30.0, 4.0)[:3, :4]
        want = orc.blender_rays(H, W, theta=theta)
        got = ops.get_rays_packed(H, W, K, c2w, 2.0, 6.0, dev).cpu()
        assert got.shape == want.shape
        assert torch.equal(got[:, :8], want[:, :8])
        assert (got[:, 8:] - want[:, 8:]).abs().max() <= 2e-7          # viewdir = d/|d|: 1 ulp (norm summation order)


def test_full_frame_properties(dev):
    """Size-independent properties at BASELINE's full size (800x800, 64+128): merged depths are
    sorted and inside [near, far], accumulated opacity in [0,1], compositing weights sum to acc,
    chunking does not change a single bit, NaNs only where acc == 0."""
    from intrinsicnerf_b200 import ops
    from tests.util import build_nets, rec_get
    coarse, fine, _, _ = build_nets("object")
    K = orc.blender_intrinsics(800, 800)
    rays = ops.get_rays_packed(800, 800, K, orc.pose_spherical(30.0, -30.0, 4.0)[:3, :4], 2.0, 6.0, dev)
    assert rays.shape == (640000, 11)
    a = ops.render_chunk(rays[:200000], coarse.packed(), fine.packed(), white_bkgd=True, want_z=True, want_weights=True)
    z = a["z_fine"]
    assert bool((z[:, 1:] >= z[:, :-1]).all()) and float(z.min()) >= 2.0 - 1e-5 and float(z.max()) <= 6.0 + 1e-5
    acc = rec_get(a["rec_fine"], "acc")
    assert float(acc.min()) >= 0.0 and float(acc.max()) <= 1.0 + 1e-5
    assert float((a["weights_fine"].sum(-1) - acc).abs().max()) < 1e-5
    disp = rec_get(a["rec_fine"], "disp")
    assert bool((torch.isnan(disp) == (acc == 0)).all())
    b = ops.render_chunk(rays[:200000][77777:123457], coarse.packed(), fine.packed(), white_bkgd=True)
    assert torch.equal(b["rec_fine"], a["rec_fine"][77777:123457])     # bit-identical under re-chunking
    assert torch.equal(b["rec_coarse"], a["rec_coarse"][77777:123457])


@pytest.mark.parametrize("white,C,endpoint,with_noise", [(True, 0, False, False), (False, 28, True, True), (True, 5, False, True)])
def test_raw2outputs_backward_matches_autograd(dev, white, C, endpoint, with_noise):
    """CUDA backward of the compositing stage vs PyTorch autograd through the oracle (fp64 reference)."""
    from intrinsicnerf_b200 import ops
    gen = torch.Generator().manual_seed(12)
    N, S = 7, 192
    ch = 11 + C + (128 if endpoint else 0)
    raw = torch.randn(N, S, ch, generator=gen)
    raw[..., 3] = raw[..., 3] * 2.0 + 0.5
    z = torch.sort(torch.rand(N, S, generator=gen) * 4 + 2, dim=-1)[0]
    rd = torch.randn(N, 3, generator=gen)
    noise = torch.randn(N, S, generator=gen) if with_noise else None
    rc = 13 + C + (128 if endpoint else 0)
    g_rec = torch.randn(N, rc, generator=gen)
    g_w = torch.randn(N, S, generator=gen) * 0.1
    # reference gradient in double precision
    r64 = raw.double().requires_grad_(True)
    o = orc.composite(r64, z.double(), rd.double(), None if noise is None else noise.double(), white, C, endpoint)
    parts = [o["rgb"], o["disp"][:, None], o["acc"][:, None], o["albedo"], o["shading"][:, None], o["residual"], o["depth"][:, None]]
    if C > 0:
        parts.append(o["sem"])
    if endpoint:
        parts.append(o["feat"])
    rec64 = torch.cat(parts, -1)
    loss = (rec64 * g_rec.double()).sum() + (o["weights"] * g_w.double()).sum()
    loss.backward()
    want = r64.grad.float()
    rg = raw.to(dev).requires_grad_(True)
    rec, w = ops.composite(rg, z.to(dev), rd.to(dev), None if noise is None else noise.to(dev), white, C, endpoint)
    ((rec * g_rec.to(dev)).sum() + (w * g_w.to(dev)).sum()).backward()
    got = rg.grad.cpu()
    scale = float(want.abs().max())
    assert float((got - want).abs().max()) < 2e-4 * scale, float((got - want).abs().max()) / scale


@pytest.mark.parametrize("variant,C,endpoint", [("object", 0, False), ("ssr", 28, True), ("ssr", 28, False), ("ssr", 5, False), ("ssr", 0, False)])
def test_mlp_backward_matches_autograd(dev, variant, C, endpoint, monkeypatch):
    """k_mlp_bwd_fp32 (through ops.MlpFn and the module's torch.cat of parameters) vs PyTorch autograd
    through the oracle's functional MLP in float64."""
    coarse, fine, pc, pf = build_nets(variant, C)
    gen = torch.Generator().manual_seed(21)
    M = 2390 if endpoint else 150          # full 64-row tiles + a ragged one
    pts = torch.rand(M, 3, generator=gen) * 6 - 3
    vd = torch.nn.functional.normalize(torch.randn(M, 3, generator=gen), dim=-1)
    scale = 1.0 if variant == "object" else 10.0
    ch = 11 + C + (128 if endpoint else 0)
    g_raw = torch.randn(M, ch, generator=gen)
    p64 = {k: v.double().requires_grad_(True) for k, v in pf.items()}
    emb = torch.cat([orc.posenc(pts.double(), 10, scale), orc.posenc(vd.double(), 4)], -1)
    out64, margin = relu_margin(lambda: orc.mlp_forward(p64, emb, variant, C, endpoint))
    # rows with a hidden unit within fp32 rounding of the ReLU kink get no upstream gradient (util.relu_margin):
    # the oracle's own fp32-vs-fp64 gradients differ by 3-5 % on this weight set without the filter, 3e-6 with it
    keep = margin > 5e-6
    assert int(keep.sum()) > 0.8 * M
    g_raw = g_raw * keep[:, None]
    (out64 * g_raw.double()).sum().backward()
    fine.zero_grad()
    monkeypatch.setattr(ops_mod, "_DEFAULT_PRECISION", ops_mod.PREC_FP32)      # the strict-fp32 training mode
    out = fine.evaluate("pts", pts.to(dev), vd.to(dev), endpoint, scale)
    assert rel_err(out.detach(), out64.detach(), floor=1e-2) < 2e-5
    (out * g_raw.to(dev)).sum().backward()
    errs = {}
    for name, p in fine.named_parameters():
        a, b = p64[name].grad.float(), p.grad.cpu()
        errs[name] = float((a - b).abs().max()) / (float(a.abs().max()) + 1e-12)
    assert max(errs.values()) < 5e-4, sorted(errs.items(), key=lambda kv: -kv[1])[:6]
    # second backward accumulates (zero_grad semantics are the caller's)
    out2 = fine.evaluate("pts", pts.to(dev), vd.to(dev), endpoint, scale)
    (out2 * g_raw.to(dev)).sum().backward()
    name, p = next(iter(fine.named_parameters()))
    assert rel_err(p.grad.cpu(), 2 * p64[name].grad.float(), floor=float(p64[name].grad.abs().max())) < 1e-3


# ---- ray generation for whole images and for sampled pixels ------------------------
@pytest.mark.parametrize("conv", ["opencv", "opengl"])
@pytest.mark.parametrize("dt", ["z", "euclidean"])
def test_rays_from_pixels_matches_reference(dev, golden_dir, conv, dt):
    """inrf_rays_from_pixels vs the reference's create_rays (golden, both conventions and depth types): origins,
    near/far exact; directions and view directions to 2 ulp of the largest component (the reference rotates with a
    batched matmul whose summation order is not specified)."""
    import intrinsicnerf_b200.ssr as ssr
    g = load_golden(golden_dir, "aux.npz")
    H, W, K, poses = int(g["ray_H"]), int(g["ray_W"]), g["ray_K"], torch.tensor(g["ray_poses"])
    want = torch.tensor(g[f"ssr_rays_{conv}_{dt}"])
    got = ssr.create_rays(2, poses.to(dev), H, W, float(K[0, 0]), float(K[1, 1]), float(K[0, 2]), float(K[1, 2]), 0.1, 10.0,
                          depth_type=dt, convention=conv).cpu()
    assert got.shape == want.shape
    assert torch.equal(got[..., 0:3], want[..., 0:3]) and torch.equal(got[..., 6:8], want[..., 6:8])
    assert float((got - want).abs().max()) <= 2.5e-7 * float(want[..., 3:6].abs().max())
    # sampled pixels (sampling_index layout: n random pixels followed by their clamped neighbours)
    gen = torch.Generator().manual_seed(5)
    idx = torch.randint(0, H * W, (1, 50), generator=gen)
    sub = ssr.rays_for_batch(idx.to(dev), poses[1], H, W, float(K[0, 0]), float(K[1, 1]), float(K[0, 2]), float(K[1, 2]), 0.1, 10.0,
                             depth_type=dt, convention=conv).cpu()
    assert torch.equal(sub, got[1][idx.reshape(-1)])


def test_object_rays_for_pixels(dev, golden_dir):
    """get_rays + gather + packing of run_nerf.py:913-932 in one launch: bit-exact against the reference's get_rays."""
    import intrinsicnerf_b200.object_level as ol
    g = load_golden(golden_dir, "aux.npz")
    H, W, K, pose = int(g["ray_H"]), int(g["ray_W"]), g["ray_K"], g["ray_poses"][0][:3, :4]
    ro, rd = torch.tensor(g["obj_rays_o"]), torch.tensor(g["obj_rays_d"])
    gen = torch.Generator().manual_seed(6)
    coords = torch.stack([torch.randint(0, H, (40,), generator=gen), torch.randint(0, W, (40,), generator=gen)], -1)
    rays = ol.rays_for_pixels(H, W, K, pose, coords, 2.0, 6.0).cpu()
    assert torch.equal(rays[:, 0:3], ro[coords[:, 0], coords[:, 1]])
    assert torch.equal(rays[:, 3:6], rd[coords[:, 0], coords[:, 1]])
    d = rd[coords[:, 0], coords[:, 1]]
    assert float((rays[:, 8:11] - d / torch.norm(d, dim=-1, keepdim=True)).abs().max()) <= 2e-7
    assert torch.equal(rays[:, 6:8], torch.tensor([2.0, 6.0]).expand(40, 2))


# ---- fused training losses ---------------------------------------------------------------
@pytest.mark.parametrize("tag", ["even", "odd"])
@pytest.mark.parametrize("fork", ["obj", "ssr"])
def test_intrinsic_losses_match_reference(dev, golden_dir, fork, tag):
    """inrf_intrinsic_loss_fwd/_bwd vs the reference's img2mse + compute_intrinsic_loss + cluster term and their
    autograd gradients (float64 golden; the kernels take fp32 inputs: 2e-5 relative on the terms, 1e-4 of the
    largest gradient entry per map)."""
    from intrinsicnerf_b200 import ops
    g = load_golden(golden_dir, "aux.npz")
    f = {k: torch.tensor(g[f"loss_{tag}_{k}"]).float().to(dev) for k in ("albedo", "shading", "residual", "rgb", "gt", "mask", "label", "target")}
    for k in ("albedo", "shading", "residual", "rgb"):
        f[k].requires_grad_(True)
    lab = f["mask"] if fork == "obj" else f["label"]
    terms = ops.intrinsic_losses(f["rgb"], f["albedo"], f["shading"], f["residual"], f["gt"], lab, f["target"], "object" if fork == "obj" else "ssr")
    want = torch.tensor(g[f"loss_{fork}_{tag}_terms"])
    assert rel_err(terms.detach(), want, floor=1e-6) < 2e-5, (terms.detach().cpu(), want)
    (terms * torch.tensor(g["loss_weights"]).float().to(dev)).sum().backward()
    for k in ("albedo", "shading", "residual", "rgb"):
        wg = torch.tensor(g[f"loss_{fork}_{tag}_g_{k}"]).float()
        assert float((f[k].grad.cpu() - wg).abs().max()) < 1e-4 * float(wg.abs().max()), k


def test_compute_intrinsic_loss_mirrors(dev, golden_dir):
    """The reference-named entry points (6 terms, reference order) composed by the caller exactly as in
    run_nerf.py:975-981, gradients flowing back through views of one render record."""
    import intrinsicnerf_b200.object_level as ol
    import intrinsicnerf_b200.ssr as ssr
    g = load_golden(golden_dir, "aux.npz")
    wts = torch.tensor(g["loss_weights"]).float()
    for fork, fn, key in (("obj", ol.compute_intrinsic_loss, "mask"), ("ssr", ssr.compute_intrinsic_loss, "label")):
        rec = torch.zeros(64, 13, device=dev)
        rec[:, 5:8] = torch.tensor(g["loss_even_albedo"]).float()
        rec[:, 8] = torch.tensor(g["loss_even_shading"]).float()
        rec[:, 9:12] = torch.tensor(g["loss_even_residual"]).float()
        rec.requires_grad_(True)
        gt, lab = torch.tensor(g["loss_even_gt"]).float().to(dev), torch.tensor(g[f"loss_even_{key}"]).float().to(dev)
        terms = fn(rec[:, 5:8], rec[:, 8], rec[:, 9:12], gt, rec[:, 3], rec[:, 4], lab)
        want = torch.tensor(g[f"loss_{fork}_even_terms"])[1:7]
        assert rel_err(torch.stack(terms).detach(), want, floor=1e-6) < 2e-5
        sum(w * t for w, t in zip(wts[1:7].tolist(), terms)).backward()
        wg = torch.tensor(g[f"loss_{fork}_even_g_albedo"]).float()
        # the golden albedo gradient also contains the cluster term (weight 0.4): remove it
        wg = wg - 0.4 * 2 * (torch.tensor(g["loss_even_albedo"]) - torch.tensor(g["loss_even_target"])).float() / (3 * 64)
        assert float((rec.grad[:, 5:8].cpu() - wg).abs().max()) < 1e-4 * float(wg.abs().max())
        assert float(rec.grad[:, 0:5].abs().max()) == 0.0
