"""Pin the CPU oracle against golden vectors produced by the unmodified reference
(tests/golden/make_golden.py).  Same torch build => bit-exact for everything that goes
through identical ATen calls; a small tolerance is kept for robustness across BLAS
thread counts (MKL sgemm blocking depends on the thread count)."""
import os

import numpy as np
import pytest
import torch

from oracle import nerf_oracle as orc

TOL = dict(rtol=2e-5, atol=2e-6)


def _load(golden_dir, name):
    return {k: v for k, v in np.load(os.path.join(golden_dir, name), allow_pickle=False).items()}


def t(a):
    return torch.from_numpy(np.asarray(a))


def close(a, b, **kw):
    kw = {**TOL, **kw}
    np.testing.assert_allclose(np.asarray(a), np.asarray(b), equal_nan=True, **kw)


def wsum(p):
    return float(sum(v.double().abs().sum() for v in p.values()))


@pytest.fixture(scope="module")
def obj(golden_dir):
    g = _load(golden_dir, "object_render.npz")
    torch.manual_seed(int(g["seed"]))
    c, f = orc.init_params("object"), orc.init_params("object")
    np.testing.assert_allclose([wsum(c), wsum(f)], g["weight_abs_sums"], rtol=1e-12)
    return g, orc.make_opaque(c), orc.make_opaque(f)


def test_stage_vectors(golden_dir):
    g = _load(golden_dir, "object_stages.npz")
    torch.manual_seed(int(g["seed"]))
    c = orc.make_opaque(orc.init_params("object"))
    x = t(g["x"])
    e_pts = orc.posenc(x, 10)
    e_dir = orc.posenc(x / x.norm(dim=-1, keepdim=True), 4)
    assert np.array_equal(e_pts.numpy(), g["emb_pts"])
    assert np.array_equal(e_dir.numpy(), g["emb_dir"])
    close(orc.mlp_forward(c, torch.cat([e_pts, e_dir], -1)), g["mlp_out"])
    for wb in (0, 1):
        o = orc.composite(t(g["raw"]), t(g["z"]), t(g["rays_d"]), None, bool(wb))
        for k in ("rgb", "disp", "acc", "weights", "depth", "albedo", "shading", "residual"):
            assert np.array_equal(o[k].numpy(), g[f"wb{wb}_{k}"], equal_nan=True), k
    assert np.isnan(g["wb0_disp"][5])                       # appendix A8 case is in the fixture
    s_det, inds, _ = orc.sample_pdf(t(g["bins"]), t(g["w"]), 128)
    assert np.array_equal(s_det.numpy(), g["s_det"])
    assert inds.min() >= 1 and inds.max() <= 63             # appendix A7
    assert (inds[:, 0] == 1).all() and (inds[:, -1] == 63).all()
    s_rnd, _, _ = orc.sample_pdf(t(g["bins"]), t(g["w"]), 128, t(g["u"]))
    assert np.array_equal(s_rnd.numpy(), g["s_rnd"])


KEYMAP = dict(rgb_map="rgb", disp_map="disp", acc_map="acc", albedo_map="albedo", shading_map="shading",
              residual_map="residual")
KEYMAP0 = dict(rgb0="rgb", disp0="disp", acc0="acc", albedo0="albedo", shading0="shading", residual0="residual")


def _check(g, prefix, res, tol):
    for k, kk in KEYMAP.items():
        src = res["fine"] if "fine" in res else res["coarse"]
        close(src[kk], g[prefix + k], **tol)
    if "fine" in res:
        for k, kk in KEYMAP0.items():
            close(res["coarse"][kk], g[prefix + k], **tol)
        close(res["z_std"], g[prefix + "z_std"], **tol)


def test_render_rays_det(obj):
    g, c, f = obj
    res = orc.render_rays(t(g["rays"]), c, f, white_bkgd=True)
    _check(g, "det_", res, {})
    close(res["raw_fine"], g["det_raw"])


def test_render_rays_stochastic_pytest_hooks(obj):
    """Replays object_level/run_nerf.py's pytest=True hooks: every draw is
    np.random.seed(0); np.random.rand(shape) (lines 389-393, 480-484; helpers 416-425)."""
    g, c, f = obj
    N = g["rays"].shape[0]

    def draw(*shape):
        np.random.seed(0)
        return torch.Tensor(np.random.rand(*shape))
    res = orc.render_rays(t(g["rays"]), c, f, white_bkgd=True, t_rand=draw(N, 64), u=draw(N, 128),
                          noise_coarse=draw(N, 64) * 1.0, noise_fine=draw(N, 192) * 1.0)
    _check(g, "sto_", res, dict(rtol=1e-4, atol=1e-5))


def test_render_rays_lindisp_black(obj):
    g, c, f = obj
    res = orc.render_rays(t(g["rays"]), c, f, white_bkgd=False, lindisp=True)
    _check(g, "lin_", res, {})


def test_render_rays_coarse_only(obj):
    g, c, f = obj
    res = orc.render_rays(t(g["rays"]), c, None, white_bkgd=True, n_importance=0)
    _check(g, "co_", res, {})


def test_render_image(obj):
    g, c, f = obj
    K = orc.blender_intrinsics(6, 6)
    img = orc.render_image(6, 6, K, orc.pose_spherical(-180.0, -30.0, 4.0)[:3, :4], 2.0, 6.0, c, f, white_bkgd=True)
    for k in ("rgb", "disp", "acc", "albedo", "shading", "residual"):
        close(img[k + "_map"], g["img_" + k])
        close(img[k + "0"], g["img_" + k + "0"])
    close(img["z_std"], g["img_z_std"])


@pytest.fixture(scope="module")
def ssr(golden_dir):
    g = _load(golden_dir, "ssr_render.npz")
    C = int(g["C"])
    torch.manual_seed(int(g["seed"]))
    c, f = orc.init_params("ssr", C), orc.init_params("ssr", C)
    np.testing.assert_allclose([wsum(c), wsum(f)], g["weight_abs_sums"], rtol=1e-12)
    return g, C, orc.make_opaque(c), orc.make_opaque(f)


SSR_KEYS = dict(rgb="rgb", disp="disp", acc="acc", depth="depth", albedo="albedo", shading="shading",
                residual="residual", sem_logits="sem")


def test_ssr_eval(ssr):
    g, C, c, f = ssr
    res = orc.render_rays(t(g["rays"]), c, f, "ssr", C, pe_scale_pts=10.0)
    for k, kk in SSR_KEYS.items():
        close(res["coarse"][kk], g[f"eval_{k}_coarse"])
        close(res["fine"][kk], g[f"eval_{k}_fine"])
    close(res["z_std"], g["eval_z_std"])
    close(res["raw_fine"][:2], g["eval_raw_fine_head"])


def test_ssr_train_mode_replay(ssr):
    g, C, c, f = ssr
    res = orc.render_rays(t(g["rays"]), c, f, "ssr", C, pe_scale_pts=10.0, t_rand=t(g["train_t_rand"]),
                          u=t(g["train_u"]), noise_coarse=t(g["train_noise_coarse"]) * 1.0,
                          noise_fine=t(g["train_noise_fine"]) * 1.0)
    for k, kk in SSR_KEYS.items():
        close(res["coarse"][kk], g[f"train_{k}_coarse"], rtol=1e-4, atol=1e-5)
        close(res["fine"][kk], g[f"train_{k}_fine"], rtol=1e-4, atol=1e-5)


def test_ssr_endpoint_and_mlp(ssr):
    g, C, c, f = ssr
    res = orc.render_rays(t(g["rays"]), c, f, "ssr", C, pe_scale_pts=10.0, endpoint=True)
    close(res["fine"]["feat"], g["ep_feat_map_fine"])
    close(res["fine"]["rgb"], g["ep_rgb_fine"])
    emb = torch.cat([orc.posenc(t(g["mlp_x"]), 10, 10.0), orc.posenc(t(g["mlp_d"]), 4)], -1)
    close(orc.mlp_forward(f, emb, "ssr", C), g["mlp_out"])
    close(orc.mlp_forward(f, emb, "ssr", C, endpoint=True), g["mlp_out_endpoint"])


# ---- ray generation and training losses ------------------------------------
def test_oracle_rays_match_reference(golden_dir):
    g = _load(golden_dir, "aux.npz")
    H, W, K, poses = int(g["ray_H"]), int(g["ray_W"]), g["ray_K"], torch.tensor(g["ray_poses"])
    ro, rd = orc.get_rays(H, W, K, poses[0][:3, :4])
    assert torch.equal(ro, torch.tensor(g["obj_rays_o"])) and torch.equal(rd, torch.tensor(g["obj_rays_d"]))
    for conv in ("opencv", "opengl"):
        for dt in ("z", "euclidean"):
            r = orc.create_rays(poses, H, W, float(K[0, 0]), float(K[1, 1]), float(K[0, 2]), float(K[1, 2]), 0.1, 10.0, dt, conv)
            assert torch.equal(r, torch.tensor(g[f"ssr_rays_{conv}_{dt}"])), (conv, dt)


@pytest.mark.parametrize("tag", ["even", "odd"])
@pytest.mark.parametrize("fork", ["obj", "ssr"])
def test_oracle_losses_match_reference(golden_dir, fork, tag):
    """oracle.intrinsic_losses == img2mse + compute_intrinsic_loss + cluster term of the reference, values and
    autograd gradients (float64 fixtures: 1e-12)."""
    g = _load(golden_dir, "aux.npz")
    tt = {k: torch.tensor(g[f"loss_{tag}_{k}"]) for k in ("albedo", "shading", "residual", "rgb", "gt", "mask", "label", "target")}
    for k in ("albedo", "shading", "residual", "rgb"):
        tt[k].requires_grad_(True)
    lab = tt["mask"] if fork == "obj" else tt["label"]
    terms = orc.intrinsic_losses(tt["rgb"], tt["albedo"], tt["shading"], tt["residual"], tt["gt"], lab, tt["target"], "object" if fork == "obj" else "ssr")
    assert torch.allclose(terms, torch.tensor(g[f"loss_{fork}_{tag}_terms"]), rtol=1e-12, atol=1e-15)
    (terms * torch.tensor(g["loss_weights"])).sum().backward()
    for k in ("albedo", "shading", "residual", "rgb"):
        assert torch.allclose(tt[k].grad, torch.tensor(g[f"loss_{fork}_{tag}_g_{k}"]), rtol=1e-10, atol=1e-14), k


def test_tc_arithmetic_oracle_is_the_same_network():
    """oracle.mlp_forward_tc_arith (fp16-rounded GEMM operands, the arithmetic of the tensor-core kernels) evaluates the
    same network as mlp_forward: values agree to the operand-rounding level, pinning its own ReLU branch decisions
    changes nothing, and its straight-through gradient is close to the exact one on samples away from the ReLU kinks."""
    import torch
    from oracle import nerf_oracle as orc
    for variant, C, endpoint in (("object", 0, False), ("ssr", 7, True)):
        pc, _ = orc.seeded_nets(variant, C)
        g = torch.Generator().manual_seed(0)
        pts = torch.rand(64, 3, generator=g, dtype=torch.float64) * 4 - 2
        vd = torch.nn.functional.normalize(torch.randn(64, 3, generator=g, dtype=torch.float64), dim=-1)
        emb = torch.cat([orc.posenc(pts, 10, 10.0 if variant == "ssr" else 1.0), orc.posenc(vd, 4)], -1)
        p64 = {k: v.double().requires_grad_(True) for k, v in pc.items()}
        exact = orc.mlp_forward(p64, emb, variant, C, endpoint)
        arith, margin = orc.mlp_forward_tc_arith(p64, emb, variant, C, endpoint)
        assert arith.shape == exact.shape and margin.shape == (64,) and bool((margin >= 0).all())
        assert float((arith - exact).detach().abs().max()) < 2e-3
        # masks recorded from the function's own activations reproduce it exactly
        names = orc.OBJECT_HEADS if variant == "object" else orc.SSR_HEADS
        with torch.no_grad():
            rnd = lambda x: x.to(torch.float16).double()  # noqa: E731
            masks, h = [], emb[:, :63]
            for i, name in enumerate(orc.TRUNK):
                z = rnd(h) @ rnd(p64[name + ".weight"]).t() + p64[name + ".bias"]
                masks.append(z > 0)
                h = torch.relu(z)
                if i == 4:
                    h = torch.cat([emb[:, :63], h], -1)
            lin = lambda n, x: rnd(x) @ rnd(p64[n + ".weight"]).t() + p64[n + ".bias"]  # noqa: E731
            masks += [lin(names["albedo1"], h) > 0, lin(names["shading1"], h) > 0]
            wv, wf = p64[names["views"] + ".weight"], p64[names["feature"] + ".weight"]
            zv = rnd(h) @ rnd(wv[:, :256] @ wf).t() + rnd(emb[:, 63:]) @ rnd(wv[:, 256:]).t() \
                + wv[:, :256] @ p64[names["feature"] + ".bias"] + p64[names["views"] + ".bias"]
            masks.append(zv > 0)
            if C > 0:
                masks.append(lin(names["sem1"], h) > 0)
        pinned, _ = orc.mlp_forward_tc_arith(p64, emb, variant, C, endpoint, masks=masks)
        assert float((pinned - arith).detach().abs().max()) < 1e-12
        keep = (margin > 5e-3).double()[:, None]
        if float(keep.sum()) >= 4:
            w = torch.randn(exact.shape, generator=g, dtype=torch.float64) * keep
            ga = torch.autograd.grad((arith * w).sum(), p64["pts_linears.3.weight"], retain_graph=True)[0]
            ge = torch.autograd.grad((exact * w).sum(), p64["pts_linears.3.weight"])[0]
            assert float((ga - ge).abs().max()) < 2e-2 * float(ge.abs().max())


# ---- wide fixtures: 2048 rays per fork through the unmodified reference, three weight regimes ----------------------------
@pytest.mark.parametrize("regime", ["opaque", "default", "trained_like"])
def test_oracle_matches_reference_on_2048_rays(golden_dir, regime):
    """The oracle's render_rays against tests/golden/wide.npz (reference fp32 outputs).  Same ATen primitives on the same
    torch build: agreement is at rounding level on the well-conditioned set, and within the reference's own fp32-vs-fp64
    disagreement (stored beside the outputs) elsewhere."""
    import numpy as np
    from tests.util import load_golden
    g = load_golden(golden_dir, "wide.npz")
    torch.manual_seed(20220414)
    coarse, fine = orc.seeded_nets("object", opaque=False)
    for p in (coarse, fine):
        if regime == "opaque":
            p["alpha_linear.bias"] += 1.0
            p["pts_linears.7.weight"] *= 3.0
        elif regime == "trained_like":
            p["pts_linears.7.weight"] *= 6.0
            p["alpha_linear.weight"] *= 30.0
    rays = orc.blender_rays(64, 64)[::2].contiguous()
    torch.set_num_threads(8)
    with torch.no_grad():
        r = orc.render_rays(rays, coarse, fine, white_bkgd=True)
    for k in ("rgb", "disp", "acc", "albedo", "shading", "residual"):
        for ours, key in ((r["fine"][k], f"{k}_map"), (r["coarse"][k], f"{k}0")):
            want = torch.from_numpy(g[f"object_{regime}_{key}"])
            floor = float(g[f"object_{regime}_floor_{key}"])
            a, b = ours.double().reshape(-1), want.double().reshape(-1)
            ok = ~(torch.isnan(a) | torch.isnan(b))
            err = float(((a[ok] - b[ok]).abs() / b[ok].abs().clamp_min(1e-3)).max())
            assert err <= max(2e-6, 10.0 * floor), (regime, key, err, floor)
