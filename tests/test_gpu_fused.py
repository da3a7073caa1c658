"""The fused renderer (inrf_render_fwd, two launches per chunk: k_mlp_tc with the ray back end) against the
stage-by-stage path (k_coarse_z, k_mlp_tc, k_raw2outputs, k_zmid, k_sample_pdf, k_merge_sorted, ...).

The back-end warp restates raw2outputs / sample_pdf / the merge (run_nerf.py:359-412, 499-503, 519) with the SAME
operations in the SAME order as the stage kernels, so the two paths must agree BIT FOR BIT on every output - records,
merged depths (i.e. the resampled positions and their sorted order) and z_std - for both network variants, ragged ray
counts, jitter, sigma noise, lindisp and both backgrounds.  The stage path is selected in-process by asking for an
output only it produces (weights_fine)."""
import pytest
import torch

from oracle import nerf_oracle as orc
from tests.util import build_nets

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def _same(a, b):
    return a.shape == b.shape and torch.equal(torch.nan_to_num(a, nan=-7.0), torch.nan_to_num(b, nan=-7.0)) \
        and torch.equal(torch.isnan(a), torch.isnan(b))


def _both(rays, pc, pf, **kw):
    from intrinsicnerf_b200 import ops
    torch.cuda.synchronize()
    n0 = ops.launch_count()
    fused = ops.render_chunk(rays, pc, pf, want_z=True, **kw)
    n_fused = ops.launch_count() - n0
    staged = ops.render_chunk(rays, pc, pf, want_z=True, want_weights=True, **kw)
    torch.cuda.synchronize()
    ops.poll_status()
    return fused, staged, n_fused


@pytest.mark.parametrize("n_rays", [1, 2, 3, 127, 129, 1000, 4099])
def test_object_fused_equals_staged(dev, n_rays):
    coarse, fine, _, _ = build_nets("object")
    rays = orc.blender_rays(72, 72)[:n_rays].contiguous().to(dev)
    f, s, launches = _both(rays, coarse.packed(), fine.packed(), white_bkgd=True)
    for k in ("rec_coarse", "rec_fine", "z_std", "z_fine"):
        assert _same(f[k], s[k]), k
    assert launches == 2, launches                        # coarse launch + fine launch; no stage kernels


@pytest.mark.parametrize("opts", [dict(white_bkgd=False), dict(white_bkgd=True, lindisp=True), dict(white_bkgd=False, jitter=True, noise=True)])
def test_object_fused_options(dev, opts):
    coarse, fine, _, _ = build_nets("object")
    opts = dict(opts)
    n = 777
    rays = orc.blender_rays(40, 40)[:n].contiguous().to(dev)
    g = torch.Generator().manual_seed(5)
    kw = dict(white_bkgd=opts["white_bkgd"], lindisp=opts.get("lindisp", False))
    if opts.get("jitter"):
        kw["t_rand"] = torch.rand(n, 64, generator=g).to(dev)
    if opts.get("noise"):
        kw["noise_coarse"] = torch.randn(n, 64, generator=g).to(dev)
        kw["noise_fine"] = torch.randn(n, 192, generator=g).to(dev)
    f, s, _ = _both(rays, coarse.packed(), fine.packed(), **kw)
    for k in ("rec_coarse", "rec_fine", "z_std", "z_fine"):
        assert _same(f[k], s[k]), k


@pytest.mark.parametrize("C,wb", [(5, False), (28, False), (28, True), (112, False)])
def test_ssr_fused_equals_staged(dev, C, wb):
    """Semantic logits are composited by the back-end warp as well (model_utils.py:90-94, white background :113-114)."""
    coarse, fine, _, _ = build_nets("ssr", C)
    rays = orc.replica_rays(24, 32)[:700].contiguous().to(dev)
    f, s, launches = _both(rays, coarse.packed(), fine.packed(), variant=1, n_classes=C, white_bkgd=wb, pe_scalar_factor=10.0)
    for k in ("rec_coarse", "rec_fine", "z_std", "z_fine"):
        assert _same(f[k], s[k]), k
    assert launches == 2


def test_coarse_only_and_unusual_sample_counts(dev):
    """N_importance = 0 (BASELINE config 1): one launch.  Sample counts other than 64 + 128 that are multiples of 32 still
    composite in-kernel when there is no fine pass; anything else takes the stage path and must give the same records."""
    from intrinsicnerf_b200 import ops
    coarse, fine, _, _ = build_nets("object")
    rays = orc.blender_rays(32, 32)[:500].contiguous().to(dev)
    pc = coarse.packed()                                  # (packs on first use: 4 launches that are not part of a render)
    for S in (64, 32, 96, 128):
        n0 = ops.launch_count()
        f = ops.render_chunk(rays, pc, None, white_bkgd=True, n_samples=S, n_importance=0)
        assert ops.launch_count() - n0 == 1, S
        raw = ops.mlp_forward_rays(coarse.packed(), 0, 0, rays, ops.coarse_z(rays, S))
        rec, _ = ops.raw2outputs_rec(raw, ops.coarse_z(rays, S), rays[:, 3:6].contiguous(), None, True)
        torch.cuda.synchronize()
        assert _same(f["rec_coarse"], rec), S
    f = ops.render_chunk(rays, coarse.packed(), fine.packed(), white_bkgd=True, n_samples=48, n_importance=80, want_z=True)   # stage path
    assert f["z_fine"].shape == (500, 128) and torch.isfinite(f["rec_fine"][:, :3]).all()
    ops.poll_status()


def test_fused_chunking_and_sharding_are_bitwise_neutral(dev):
    """A ray's result must not depend on where it sits in a chunk (the reference's chunk argument 'does not affect final
    results', run_nerf.py:83): contiguous per-CTA runs of whole ray groups keep that true for the fused kernel."""
    from intrinsicnerf_b200 import ops
    coarse, fine, _, _ = build_nets("object")
    rays = orc.blender_rays(64, 64).contiguous().to(dev)
    whole = ops.render_chunk(rays, coarse.packed(), fine.packed(), white_bkgd=True)
    parts = [ops.render_chunk(rays[a:b], coarse.packed(), fine.packed(), white_bkgd=True) for a, b in ((0, 1), (1, 1000), (1000, 1003), (1003, 4096))]
    torch.cuda.synchronize()
    for k in ("rec_coarse", "rec_fine", "z_std"):
        assert _same(whole[k], torch.cat([p[k] for p in parts], 0)), k


@pytest.mark.parametrize("conv,dt", [("opengl", "z"), ("opencv", "z"), ("opencv", "euclidean")])
def test_in_kernel_ray_generation_equals_ray_table(dev, conv, dt):
    """inrf_render_fwd_camera (rays generated per pixel inside the kernels) against the same frame
    rendered from the [H*W, 11] table that inrf_rays_from_pixels writes: bit-identical records, any pixel sub-range."""
    from intrinsicnerf_b200 import ops
    coarse, fine, _, _ = build_nets("object")
    H, W = 37, 53
    K = orc.blender_intrinsics(H, W)
    c2w = torch.as_tensor(orc.pose_spherical(40.0, -30.0, 4.0))[:3, :4]
    rays = ops.rays_from_pixels(None, H, W, K[0][0], K[1][1], K[0][2], K[1][2], c2w, 2.0, 6.0, conv, dt, dev)
    want = ops.render_chunk(rays, coarse.packed(), fine.packed(), white_bkgd=True, want_z=True)
    n0 = ops.launch_count()
    got = ops.render_frame_camera(H, W, K, c2w, 2.0, 6.0, coarse.packed(), fine.packed(), dev, white_bkgd=True, convention=conv,
                                  depth_type=dt, want_z=True)
    assert ops.launch_count() - n0 == 2                  # no k_get_rays
    part = ops.render_frame_camera(H, W, K, c2w, 2.0, 6.0, coarse.packed(), fine.packed(), dev, pix0=700, n=333, white_bkgd=True,
                                   convention=conv, depth_type=dt)
    torch.cuda.synchronize()
    ops.poll_status()
    for k in ("rec_coarse", "rec_fine", "z_std", "z_fine"):
        assert _same(got[k], want[k]), k
    assert _same(part["rec_fine"], want["rec_fine"][700:1033])


def test_render_with_c2w_uses_the_camera_path_and_matches_rays_path(dev):
    """object_level.render(c2w=...) (run_nerf.py:100-103): same maps as render(rays=get_rays(...)) - which goes through the
    reference's own torch ray generation - to the ray-generation rounding."""
    from intrinsicnerf_b200 import object_level as ol, ops
    coarse, fine, _, _ = build_nets("object")
    e, _ = ol.get_embedder(10, 0)
    ed, _ = ol.get_embedder(4, 0)
    kw = dict(network_fn=coarse, network_fine=fine, network_query_fn=ol._FusedQuery(e, ed, 65536), N_samples=64, N_importance=128,
              perturb=0., white_bkgd=True, raw_noise_std=0.)
    H = W = 24
    K = orc.blender_intrinsics(H, W)
    c2w = torch.as_tensor(orc.pose_spherical(-120.0, -30.0, 4.0))[:3, :4].to(dev)
    coarse.packed(), fine.packed()
    with torch.no_grad():                                # as render_path does (run_nerf.py:167); under autograd render() trains
        n0 = ops.launch_count()
        a = ol.render(H, W, K, chunk=200, c2w=c2w, ndc=False, near=2., far=6., use_viewdirs=True, **kw)
        assert ops.launch_count() - n0 == 2 * 3          # three chunks, two launches each, no ray-table kernel
        b = ol.render(H, W, K, chunk=4096, rays=ol.get_rays(H, W, K, c2w), ndc=False, near=2., far=6., use_viewdirs=True, **kw)
    torch.cuda.synchronize()
    for x, y in zip(a[:6], b[:6]):
        assert x.shape == y.shape
        assert float((x - y).abs().max()) < 2e-5
    assert set(a[6]) == set(b[6])
