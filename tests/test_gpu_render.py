"""GPU parity of the fused renderer (render_rays / volumetric_rendering / render) against the
CPU oracle and the reference's golden vectors.  Tolerance: 1e-4 relative (floor 1e-3 on the
denominator), the bound BASELINE.json's north_star states, on the well-conditioned 'opaque'
weight set."""
import numpy as np
import pytest
import torch

from oracle import nerf_oracle as orc
from tests.util import build_nets, load_golden, rec_get, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-4
KEYS = ("rgb", "disp", "acc", "albedo", "shading", "residual")


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def obj_nets():
    return build_nets("object")


def _object_kwargs(coarse, fine, **over):
    from intrinsicnerf_b200 import object_level as ol
    e, _ = ol.get_embedder(10, 0)
    ed, _ = ol.get_embedder(4, 0)
    kw = dict(network_fn=coarse, network_fine=fine, network_query_fn=ol._FusedQuery(e, ed, 65536), N_samples=64,
              N_importance=128, perturb=0., white_bkgd=True, raw_noise_std=0.)
    kw.update(over)
    return kw


def _check_param_grads(pairs, precision):
    """fp32 training mode: 2e-3 of each parameter's largest gradient vs the exact oracle.  Tensor-core mode: the same
    oracle takes the other ReLU branch for ~0.05 % of the (sample, unit) pairs (fp16 operand rounding), each a
    full-size difference in one sample's contribution - on these few-thousand-sample batches that is percent-level on
    single entries, so the end-to-end bar is direction + magnitude (cosine > 0.999, norm within 2 %); the tight
    per-entry bar of the tensor-core backward lives in test_gpu_train_tc.py against the kernel-arithmetic oracle."""
    for ref, net in pairs:
        for name, p in net.named_parameters():
            a, b = ref[name].grad.double().reshape(-1), p.grad.cpu().double().reshape(-1)
            if precision == "fp32":
                assert float((a - b).abs().max()) < 2e-3 * float(a.abs().max()) + 1e-9, name
            elif float(a.norm()) == 0.0:                     # a head the loss does not reach (e.g. the coarse colour heads)
                assert float(b.abs().max()) < 1e-9, name
            else:
                cos = float((a @ b) / (a.norm() * b.norm() + 1e-300))
                assert cos > 0.999 and abs(float(b.norm() / (a.norm() + 1e-300)) - 1.0) < 0.02, (name, cos)


@pytest.fixture
def train_precision(request, monkeypatch):
    from intrinsicnerf_b200 import ops
    monkeypatch.setattr(ops, "_DEFAULT_PRECISION", ops.PREC_FP32 if request.param == "fp32" else ops.PREC_TC)
    return request.param


@pytest.mark.parametrize("precision", ["fp32", "tc"])
def test_render_rays_golden(dev, golden_dir, obj_nets, precision):
    """The reference's own outputs (golden) for the deterministic, stochastic (pytest hooks),
    lindisp/black-background and coarse-only configurations."""
    from intrinsicnerf_b200 import object_level as ol, ops
    coarse, fine, pc, pf = obj_nets
    g = load_golden(golden_dir, "object_render.npz")
    rays = torch.from_numpy(g["rays"]).to(dev)
    ops.set_default_precision(precision)
    try:
        with torch.no_grad():
            det = ol.render_rays(rays, retraw=True, **_object_kwargs(coarse, fine))
            sto = ol.render_rays(rays, pytest=True, **_object_kwargs(coarse, fine, perturb=1.0, raw_noise_std=1.0))
            lin = ol.render_rays(rays, **_object_kwargs(coarse, fine, lindisp=True, white_bkgd=False))
            co = ol.render_rays(rays, **_object_kwargs(coarse, None, N_importance=0))
    finally:
        ops.set_default_precision("tc")
    for prefix, res in (("det_", det), ("sto_", sto), ("lin_", lin)):
        for k in KEYS:
            assert rel_err(res[k + "_map"], g[prefix + k + "_map"]) < TOL, (prefix, k, rel_err(res[k + "_map"], g[prefix + k + "_map"]))
            assert rel_err(res[k + "0"], g[prefix + k + "0"]) < TOL, (prefix, k)
        assert rel_err(res["z_std"], g[prefix + "z_std"]) < TOL
    assert set(co.keys()) == {k + "_map" for k in KEYS}
    for k in KEYS:
        assert rel_err(co[k + "_map"], g["co_" + k + "_map"]) < TOL
    assert det["raw"].shape == (rays.shape[0], 192, 11)
    assert rel_err(det["raw"], g["det_raw"], floor=1e-2) < 5e-4       # per-sample raw: fine z differs by rounding


@pytest.mark.parametrize("precision", ["fp32", "tc"])
def test_render_image_api_and_chunking(dev, golden_dir, obj_nets, precision):
    """render(): ray generation, packing, chunk loop, reshape and the 7-entry return list."""
    from intrinsicnerf_b200 import object_level as ol, ops
    coarse, fine, pc, pf = obj_nets
    g = load_golden(golden_dir, "object_render.npz")
    K = np.array(orc.blender_intrinsics(6, 6))
    c2w = orc.pose_spherical(-180.0, -30.0, 4.0)[:3, :4].to(dev)
    kw = _object_kwargs(coarse, fine)
    ops.set_default_precision(precision)
    try:
        with torch.no_grad():
            out = ol.render(6, 6, K, chunk=32768, c2w=c2w, ndc=False, near=2., far=6., use_viewdirs=True, **kw)
            out_small = ol.render(6, 6, K, chunk=7, c2w=c2w, ndc=False, near=2., far=6., use_viewdirs=True, **kw)
    finally:
        ops.set_default_precision("tc")
    assert len(out) == 7 and out[0].shape == (6, 6, 3) and out[1].shape == (6, 6)
    for i, k in enumerate(KEYS):
        assert rel_err(out[i], g["img_" + k]) < TOL, k
        assert torch.equal(out[i], out_small[i]), "results must not depend on the chunk size"
    for k in ("rgb0", "disp0", "acc0", "albedo0", "shading0", "residual0", "z_std"):
        assert rel_err(out[6][k], g["img_" + k]) < TOL, k


@pytest.mark.parametrize("precision", ["fp32", "tc"])
def test_ssr_renderer_golden(dev, golden_dir, precision):
    from intrinsicnerf_b200 import ops, ssr
    g = load_golden(golden_dir, "ssr_render.npz")
    C = int(g["C"])
    coarse, fine, pc, pf = build_nets("ssr", C)

    class T(ssr.SSRRenderer):
        pass
    t = T()
    t.N_samples, t.N_importance, t.perturb, t.raw_noise_std = 64, 128, 1, 1.0
    t.white_bkgd, t.enable_semantic, t.num_valid_semantic_class, t.endpoint_feat = False, True, C, False
    t.netchunk = t.chunk = 32768
    t.ssr_net_coarse, t.ssr_net_fine = coarse, fine
    t.embed_fn, _ = ssr.get_embedder(10, 0, scalar_factor=10)
    t.embeddirs_fn, _ = ssr.get_embedder(4, 0, scalar_factor=1)
    rays = torch.from_numpy(g["rays"]).to(dev)
    ops.set_default_precision(precision)
    try:
        t.training = False
        with torch.no_grad():
            ev = t.render_rays(rays)
            t.endpoint_feat = True
            ep = t.render_rays(rays)
    finally:
        ops.set_default_precision("tc")
    names = ("rgb", "disp", "acc", "depth", "albedo", "shading", "residual", "sem_logits")
    for k in names:
        for lvl in ("coarse", "fine"):
            ref = g[f"eval_{k}_{lvl}"]
            # semantic logits are unbounded and cross zero (sums of +/- terms, O(0.05) for these
            # weights); they feed a softmax, so their error is bounded on the absolute scale
            # max(1, |logit|) instead of the 1e-3 relative floor
            floor = 1.0 if k == "sem_logits" else 1e-3
            e = rel_err(ev[f"{k}_{lvl}"], ref, floor=floor)
            assert e < TOL, (k, lvl, e)
    assert rel_err(ev["z_std"], g["eval_z_std"]) < TOL
    assert ev["raw_coarse"].shape == (rays.shape[0], 64, 11 + C) and ev["raw_fine"].shape == (rays.shape[0], 192, 11 + C)
    # endpoint features are raw ReLU hidden activations (many exactly or nearly zero), not rendered
    # quantities: their error is taken relative to the map's scale; fp16-operand rounding of a
    # 256-term dot product leaves ~2e-4 of that scale on the tensor-core path
    feat = g["ep_feat_map_fine"]
    assert rel_err(ep["feat_map_fine"], feat, floor=float(np.abs(feat).max())) < (TOL if precision == "fp32" else 1e-3)
    assert ep["raw_fine"].shape[-1] == 11 + C + 128


def test_ssr_training_mode_replay_against_oracle(dev, golden_dir):
    """Training-mode randomness (jitter, sigma noise, random u) injected identically into the
    oracle and the kernels (fp32 mode)."""
    from intrinsicnerf_b200 import ops
    g = load_golden(golden_dir, "ssr_render.npz")
    C = int(g["C"])
    coarse, fine, pc, pf = build_nets("ssr", C)
    rays = torch.from_numpy(g["rays"])
    rnd = {k: torch.from_numpy(g["train_" + k]) for k in ("t_rand", "u", "noise_coarse", "noise_fine")}
    o = ops.render_chunk(rays.to(dev), coarse.packed(), fine.packed(), variant=1, n_classes=C, pe_scalar_factor=10.0,
                         precision="fp32", **{k: v.to(dev) for k, v in rnd.items()})
    names = dict(rgb="rgb", disp="disp", acc="acc", depth="depth", albedo="albedo", shading="shading", residual="residual")
    for k in names:
        assert rel_err(rec_get(o["rec_fine"], k), g[f"train_{k}_fine"]) < TOL, k
        assert rel_err(rec_get(o["rec_coarse"], k), g[f"train_{k}_coarse"]) < TOL, k
    sem = g["train_sem_logits_fine"]
    assert rel_err(o["rec_fine"][:, 13:13 + C], sem, floor=1.0) < TOL


def test_tc_matches_fp32_on_larger_batch(dev, obj_nets):
    """Tensor-core path vs the strict fp32 path on 4096 rays (sizes the CPU oracle would need
    minutes for): every map within 1e-4, merged depths within 2e-6."""
    from intrinsicnerf_b200 import ops
    coarse, fine, pc, pf = obj_nets
    rays = orc.blender_rays(64, 64).to(dev)
    a = ops.render_chunk(rays, coarse.packed(), fine.packed(), white_bkgd=True, precision="fp32", want_z=True)
    b = ops.render_chunk(rays, coarse.packed(), fine.packed(), white_bkgd=True, precision="tc", want_z=True)
    for k in KEYS:
        assert rel_err(rec_get(b["rec_fine"], k), rec_get(a["rec_fine"], k)) < TOL, k
        assert rel_err(rec_get(b["rec_coarse"], k), rec_get(a["rec_coarse"], k)) < TOL, k
    assert (a["z_fine"] - b["z_fine"]).abs().max() < 1e-3
    assert bool((b["z_fine"][:, 1:] >= b["z_fine"][:, :-1]).all())      # sortedness at full size


def test_oracle_end_to_end_medium(dev, obj_nets):
    """256 rays through the full CPU oracle vs the kernels (both precisions)."""
    from intrinsicnerf_b200 import ops
    coarse, fine, pc, pf = obj_nets
    rays = orc.blender_rays(16, 16)
    want = orc.render_rays(rays, pc, pf, white_bkgd=True)
    for prec in ("fp32", "tc"):
        o = ops.render_chunk(rays.to(dev), coarse.packed(), fine.packed(), white_bkgd=True, precision=prec)
        for k in KEYS:
            assert rel_err(rec_get(o["rec_fine"], k), want["fine"][k]) < TOL, (prec, k)
            assert rel_err(rec_get(o["rec_coarse"], k), want["coarse"][k]) < TOL, (prec, k)
        assert rel_err(o["z_std"], want["z_std"]) < TOL


def test_training_step_through_stage_kernels_with_foreign_network(dev):
    """Training path available in round 1: a foreign PyTorch network (here the oracle's functional
    MLP on the GPU) trains through our sampling / compositing kernels - forward by the CUDA stages,
    backward by k_raw2outputs_bwd - and the parameter gradients equal full PyTorch autograd through
    the CPU oracle.  (The fused tensor-core path is forward-only and says so.)"""
    from intrinsicnerf_b200 import object_level as ol
    torch.manual_seed(20220414)
    pc = orc.make_opaque(orc.init_params("object"))
    pf = orc.make_opaque(orc.init_params("object"))
    rays = orc.blender_rays(6, 6)[::3].contiguous()                   # 12 rays
    target = torch.rand(rays.shape[0], 3, generator=torch.Generator().manual_seed(1))

    def loss_of(res_rgb, res_rgb0, res_alb):
        return ((res_rgb - target) ** 2).mean() + ((res_rgb0 - target) ** 2).mean() + 0.1 * res_alb.mean()

    # reference gradients: full autograd through the oracle on the CPU
    cc = {k: v.clone().requires_grad_(True) for k, v in pc.items()}
    cf = {k: v.clone().requires_grad_(True) for k, v in pf.items()}
    r = orc.render_rays(rays, cc, cf, white_bkgd=True)
    loss_of(r["fine"]["rgb"], r["coarse"]["rgb"], r["fine"]["albedo"]).backward()

    gc = {k: v.clone().to(dev).requires_grad_(True) for k, v in pc.items()}
    gf = {k: v.clone().to(dev).requires_grad_(True) for k, v in pf.items()}

    def query(pts, viewdirs, params):                                  # foreign network_query_fn
        return orc.query_field(pts, viewdirs, params)
    out = ol.render_rays(rays.to(dev), gc, query, N_samples=64, N_importance=128, network_fine=gf, perturb=0.,
                         white_bkgd=True, raw_noise_std=0.)
    tgt = target.to(dev)
    loss = ((out["rgb_map"] - tgt) ** 2).mean() + ((out["rgb0"] - tgt) ** 2).mean() + 0.1 * out["albedo_map"].mean()
    loss.backward()
    for name in ("pts_linears.0.weight", "pts_linears.7.weight", "alpha_linear.weight", "albedo_linear2.weight", "shading_linear.bias"):
        for ref, got in ((cc, gc), (cf, gf)):
            a, b = ref[name].grad, got[name].grad.cpu()
            assert float((a - b).abs().max()) < 2e-3 * float(a.abs().max()) + 1e-9, name


@pytest.mark.parametrize("train_precision", ["fp32", "tc"], indirect=True)
def test_training_step_with_our_modules(dev, train_precision):
    """run_nerf.py:942-1019 shape: render(..., retraw=True) under autograd with OUR NeRF modules, a loss
    on fine and coarse maps, loss.backward(): parameter gradients equal full autograd through the CPU
    oracle (perturb=0 so that both sides see identical samples)."""
    from intrinsicnerf_b200 import object_level as ol
    coarse, fine, pc, pf = build_nets("object")
    rays = orc.blender_rays(6, 6)[::3].contiguous()
    target = torch.rand(rays.shape[0], 3, generator=torch.Generator().manual_seed(1))
    cc = {k: v.clone().requires_grad_(True) for k, v in pc.items()}
    cf = {k: v.clone().requires_grad_(True) for k, v in pf.items()}
    r = orc.render_rays(rays, cc, cf, white_bkgd=True)
    (((r["fine"]["rgb"] - target) ** 2).mean() + ((r["coarse"]["rgb"] - target) ** 2).mean()
     + 0.1 * r["fine"]["albedo"].mean() + 0.05 * r["fine"]["shading"].mean() + 0.05 * r["coarse"]["residual"].mean()).backward()
    kw = _object_kwargs(coarse, fine)
    out = ol.render_rays(rays.to(dev), retraw=True, **kw)
    tgt = target.to(dev)
    loss = ((out["rgb_map"] - tgt) ** 2).mean() + ((out["rgb0"] - tgt) ** 2).mean() + 0.1 * out["albedo_map"].mean() \
        + 0.05 * out["shading_map"].mean() + 0.05 * out["residual0"].mean()
    loss.backward()
    _check_param_grads(((cc, coarse), (cf, fine)), train_precision)
    # and an optimizer step runs, after which the inference path sees the new weights
    opt = torch.optim.Adam(list(coarse.parameters()) + list(fine.parameters()), lr=5e-4)
    opt.step()
    with torch.no_grad():
        after = ol.render_rays(rays.to(dev), **kw)
    assert torch.isfinite(after["rgb_map"]).all()


@pytest.mark.parametrize("train_precision", ["fp32", "tc"], indirect=True)
def test_ssr_training_step(dev, golden_dir, train_precision):
    """SSRTrainer.step shape (trainer.py:882-990): volumetric_rendering in training mode, cross-entropy on the
    semantic logits + photometric loss, backward through our kernels; gradients vs the oracle."""
    from intrinsicnerf_b200 import ssr
    g = load_golden(golden_dir, "ssr_render.npz")
    C = int(g["C"])
    coarse, fine, pc, pf = build_nets("ssr", C)

    class T(ssr.SSRRenderer):
        pass
    t = T()
    t.N_samples, t.N_importance, t.perturb, t.raw_noise_std = 64, 128, 0, 0.0
    t.white_bkgd, t.enable_semantic, t.num_valid_semantic_class, t.endpoint_feat = False, True, C, False
    t.netchunk = t.chunk = 32768
    t.ssr_net_coarse, t.ssr_net_fine = coarse, fine
    t.embed_fn, _ = ssr.get_embedder(10, 0, scalar_factor=10)
    t.embeddirs_fn, _ = ssr.get_embedder(4, 0, scalar_factor=1)
    t.training = True
    rays = torch.from_numpy(g["rays"])
    labels = torch.arange(rays.shape[0]) % C
    cc = {k: v.clone().requires_grad_(True) for k, v in pc.items()}
    cf = {k: v.clone().requires_grad_(True) for k, v in pf.items()}
    r = orc.render_rays(rays, cc, cf, "ssr", C, pe_scale_pts=10.0)
    ce = torch.nn.functional.cross_entropy
    (ce(r["fine"]["sem"], labels) + ce(r["coarse"]["sem"], labels) + (r["fine"]["rgb"] ** 2).mean() + r["fine"]["depth"].mean()).backward()
    out = t.render_rays(rays.to(dev))
    lab = labels.to(dev)
    (ce(out["sem_logits_fine"], lab) + ce(out["sem_logits_coarse"], lab) + (out["rgb_fine"] ** 2).mean() + out["depth_fine"].mean()).backward()
    _check_param_grads(((cc, coarse), (cf, fine)), train_precision)


def test_dense_grid_query_zero_viewdirs(dev):
    """SURVEY section 8f row 4 (SSR/extract_colour_mesh.py:149-166): grid points through the fine network with
    zero view directions - the field query of the meshing script as one chunked device call."""
    from intrinsicnerf_b200 import ops
    coarse, fine, pc, pf = build_nets("ssr", 28)
    n = 21
    ax = torch.linspace(-2.5, 2.5, n)
    pts = torch.stack(torch.meshgrid(ax, ax, ax, indexing="ij"), -1).reshape(-1, 3)
    emb = torch.cat([orc.posenc(pts, 10, 10.0), orc.posenc(torch.zeros_like(pts), 4)], -1)
    want = orc.mlp_forward(pf, emb, "ssr", 28, False)
    got = fine.query_points(pts.to(dev), pe_scalar_factor=10.0, chunk=4000)
    assert got.shape == want.shape
    assert rel_err(got[:, :11], want[:, :11], floor=1e-2) < 2e-3          # tensor-core path, raw (pre-compositing) values
    assert float((got[:, 11:].cpu() - want[:, 11:]).abs().max()) < 2e-3 * float(want[:, 11:].abs().max())
    ops.set_default_precision("fp32")
    try:
        got32 = fine.query_points(pts.to(dev), pe_scalar_factor=10.0)
    finally:
        ops.set_default_precision("tc")
    assert rel_err(got32[:, :11], want[:, :11], floor=1e-2) < 2e-5
