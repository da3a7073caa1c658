"""GPU parity of the full-image driver kernels through the C ABI: inrf_frame_finish /
inrf_edit_recompose against the reference's own render_path outputs (tests/golden/frame.npz) and against the numpy
oracle on full-size frames; render_path of both forks end to end.  Bars: bit-exact for every 8/16-bit plane, label
and cluster sample; 2e-6 absolute for the entropy map (expf/logf vs numpy)."""
import os

import numpy as np
import pytest
import torch

from oracle import frame_oracle as fo
from tests.util import build_nets, load_golden

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _rec(rgb, disp, acc, albedo, shading, residual, depth=None, logits=None, pad=0):
    f = lambda a, w: torch.as_tensor(np.asarray(a), dtype=torch.float32).reshape(-1, w)  # noqa: E731
    cols = [f(rgb, 3), f(disp, 1), f(acc, 1), f(albedo, 3), f(shading, 1), f(residual, 3),
            f(depth, 1) if depth is not None else torch.zeros(f(acc, 1).shape)]
    if logits is not None:
        cols.append(f(logits, np.asarray(logits).shape[-1]))
    if pad:
        cols.append(torch.full((cols[0].shape[0], pad), float("nan")))
    return torch.cat(cols, 1).contiguous().to(DEV)


def test_object_frame_planes_equal_reference(golden_dir):
    from intrinsicnerf_b200 import ops
    g = load_golden(golden_dir, "frame.npz")
    px, lb = [], []
    for i in range(2):
        m = {k: g[f"obj{i}_{k}"] for k in ("rgb", "disp", "acc", "albedo", "shading", "residual")}
        H, W = m["acc"].shape
        rec = _rec(**m)
        o = ops.frame_finish(rec, H, W, 0, ("rgb8", "albedo8", "shading8", "residual8", "label8", "labels64"), sub_step=2)
        for k in ("rgb8", "albedo8", "shading8", "residual8", "label8"):
            assert np.array_equal(o[k].cpu().numpy(), g[f"obj{i}_{k}"]), (i, k)
        assert np.array_equal(o["labels64"].cpu().numpy().reshape(-1, 1), g[f"obj{i}_dc_label"])
        c8, e8 = ops.edit_recompose(torch.from_numpy(g[f"obj{i}_dc_result"]).to(DEV), rec)
        assert np.array_equal(c8.cpu().numpy().reshape(H, W, 3), g[f"obj{i}_c8"])
        assert np.array_equal(e8.cpu().numpy().reshape(H, W, 3), g[f"obj{i}_edit8"])
        px.append(o["sample_pixels"].cpu().numpy()), lb.append(o["sample_labels"].cpu().numpy())
    assert np.array_equal(np.concatenate(px, 0), g["obj_uc_pixels"], equal_nan=True)
    assert np.array_equal(np.concatenate(lb, 0), g["obj_uc_labels"])


def test_ssr_frame_planes_equal_reference(golden_dir):
    from intrinsicnerf_b200 import ops
    g = load_golden(golden_dir, "frame.npz")
    H, W, C = int(g["ssr_H"]), int(g["ssr_W"]), int(g["ssr_C"])
    planes = ("rgb8", "albedo8", "shading8", "residual8", "disp16", "depth_mm16", "label8", "vis_label8", "entropy8",
              "entropy", "labels64")
    px, lb = [], []
    for i in range(2):
        r = lambda k: g[f"ssr{i}_{k}_fine"]  # noqa: E731
        rec = _rec(r("rgb"), r("disp"), np.zeros(H * W), r("albedo"), r("shading"), r("residual"), r("depth"), r("sem_logits"),
                   pad=128 * (i == 1))                                   # a record with the endpoint feature behind the logits
        o = ops.frame_finish(rec, H, W, C, planes, colour_map=g["ssr_colour_map"], sub_step=2)
        for k in planes[:8]:
            assert np.array_equal(o[k].cpu().numpy(), g[f"ssr{i}_{k}"]), (i, k)
        ent = o["entropy"].cpu().numpy()
        np.testing.assert_allclose(ent, g["ssr_entropys"][i], rtol=0, atol=2e-6)
        e8 = o["entropy8"].cpu().numpy().astype(int)
        assert np.abs(e8 - g[f"ssr{i}_entropy8"].astype(int)).max() <= 1   # a 1e-6 entropy difference may cross a 1/255 step
        assert np.array_equal(e8, fo.to8b(ent))
        c8, ed8 = ops.edit_recompose(torch.from_numpy(g[f"ssr{i}_dc_result"]).to(DEV), rec)
        assert np.array_equal(c8.cpu().numpy().reshape(H, W, 3), g[f"ssr{i}_c8"])
        assert np.array_equal(ed8.cpu().numpy().reshape(H, W, 3), g[f"ssr{i}_edit8"])
        px.append(o["sample_pixels"].cpu().numpy()), lb.append(o["sample_labels"].cpu().numpy())
    assert np.array_equal(np.stack(px, 0), g["ssr_uc_pixels"], equal_nan=True)
    assert np.array_equal(np.stack(lb, 0), g["ssr_uc_labels"])


@pytest.mark.parametrize("H,W", [(800, 800), (1, 1), (3, 5), (121, 67)])
def test_full_size_frame_matches_oracle(H, W):
    """BASELINE-size frame (and odd / tiny sizes for the [::2, ::2] sub-sampling) against the numpy oracle."""
    from intrinsicnerf_b200 import ops
    gen = torch.Generator().manual_seed(H * 1000 + W)
    r = lambda *s: (torch.rand(*s, generator=gen) * 1.3 - 0.15)  # noqa: E731
    m = dict(rgb=r(H, W, 3), disp=r(H, W) * 50, acc=r(H, W) * 14, albedo=r(H, W, 3), shading=r(H, W), residual=r(H, W, 3) * 0.2)
    rec = _rec(**m)
    o = ops.frame_finish(rec, H, W, 0, ("rgb8", "albedo8", "shading8", "residual8", "label8", "labels64"), sub_step=2)
    want = fo.object_frame(**{k: v.numpy() for k, v in m.items()})
    for k, v in want.items():
        assert np.array_equal(o[k].cpu().numpy().reshape(v.shape), v), k
    res = r(H * W, 3)
    c8, e8 = ops.edit_recompose(res.to(DEV), rec)
    wc, we = fo.edit_recompose(res.numpy().reshape(H, W, 3), m["shading"].numpy(), m["residual"].numpy())
    assert np.array_equal(c8.cpu().numpy().reshape(H, W, 3), wc) and np.array_equal(e8.cpu().numpy().reshape(H, W, 3), we)


def test_frame_argument_errors():
    from intrinsicnerf_b200 import ops
    from intrinsicnerf_b200._lib import InrfError
    rec = torch.zeros(12, 13, device=DEV)
    assert ops.frame_finish(torch.zeros(0, 13, device=DEV), 0, 0, 0, ("rgb8",))["rgb8"].numel() == 0
    with pytest.raises(InrfError):
        ops.frame_finish(rec, 3, 4, 5, ("label8",))                      # stride 13 cannot hold 5 logits
    with pytest.raises(InrfError):
        ops.frame_finish(rec, 3, 4, 0, ("entropy",))                     # semantic plane without classes
    with pytest.raises(RuntimeError):
        ops.frame_finish(rec.cpu(), 3, 4, 0, ("rgb8",))                  # no CPU path


def test_object_render_path_end_to_end(tmp_path):
    """render_path (run_nerf.py:142-272 signature) on two 40x40 views with update_cluster: returned maps equal
    render(), the PNG planes equal the oracle's conversion of those maps, a cluster manager comes back and the
    c###/edit### images equal the oracle's recomposition of its dest_color."""
    import cv2
    from intrinsicnerf_b200 import object_level as ol
    from oracle import nerf_oracle as orc
    coarse, fine, _, _ = build_nets("object", device=DEV)
    e10, _ = ol.get_embedder(10, 0)
    e4, _ = ol.get_embedder(4, 0)
    kw = dict(network_fn=coarse, network_fine=fine, network_query_fn=ol._FusedQuery(e10, e4, 65536), N_samples=64, N_importance=128,
              perturb=False, white_bkgd=True, raw_noise_std=0., use_viewdirs=True, ndc=False, lindisp=False, near=2., far=6.)
    H = W = 40
    K = np.array([[55.0, 0, 20.0], [0, 55.0, 20.0], [0, 0, 1]], dtype=np.float32)
    poses = [torch.as_tensor(orc.pose_spherical(a, -30.0, 4.0)).float() for a in (20.0, 140.0)]
    rgbs, disps, cm = ol.render_path(poses, (H, W, 55.0), K, 1024, kw, savedir=str(tmp_path), update_cluster=True)
    assert rgbs.shape == (2, H, W, 3) and disps.shape == (2, H, W) and cm is not None and cm.clusters[0] is not None
    for i, pose in enumerate(poses):
        with torch.no_grad():
            rgb, disp, acc, albedo, shading, residual, _ = ol.render(H, W, K, chunk=1024, c2w=pose[:3, :4], **kw)
        assert np.array_equal(rgbs[i], rgb.cpu().numpy()) and np.array_equal(disps[i], disp.cpu().numpy(), equal_nan=True)
        want = fo.object_frame(*(t.cpu().numpy() for t in (rgb, disp, acc, albedo, shading, residual)))
        for prefix, key in (("", "rgb8"), ("a", "albedo8"), ("s", "shading8"), ("res", "residual8"), ("acc", "label8")):
            img = cv2.imread(os.path.join(str(tmp_path), "{}{:03d}.png".format(prefix, i)), cv2.IMREAD_UNCHANGED)
            img = img[..., ::-1] if img.ndim == 3 else img
            assert np.array_equal(img, want[key]), (i, key)
        result = cm.dest_color(albedo.reshape(-1, 3), torch.zeros(H * W, 1, dtype=torch.long, device=DEV))
        wc, we = fo.edit_recompose(result.cpu().numpy().reshape(H, W, 3), shading.cpu().numpy(), residual.cpu().numpy())
        for name, w8 in (("c", wc), ("edit", we)):
            img = cv2.imread(os.path.join(str(tmp_path), "{}{:03d}.png".format(name, i)), cv2.IMREAD_UNCHANGED)[..., ::-1]
            assert np.array_equal(img, w8), (i, name)


def test_ssr_render_path_end_to_end(tmp_path):
    """SSRRenderer.render_path (trainer.py:1221 signature): 12-tuple layout, maps equal render_rays' fine outputs,
    label / colour / entropy planes equal the oracle's conversion of the rendered logits."""
    from intrinsicnerf_b200 import ssr
    from oracle import nerf_oracle as orc
    C, H, W = 5, 24, 32
    coarse, fine, _, _ = build_nets("ssr", n_classes=C, device=DEV)

    class T(ssr.SSRRenderer):
        pass
    t = T()
    t.N_samples, t.N_importance, t.perturb, t.training, t.raw_noise_std, t.white_bkgd = 64, 128, 1.0, True, 1.0, False
    t.enable_semantic, t.num_valid_semantic_class, t.endpoint_feat, t.chunk, t.netchunk = True, C, False, 512, 65536
    t.ssr_net_coarse, t.ssr_net_fine = coarse, fine
    t.embed_fn, _ = ssr.get_embedder(10, 0, scalar_factor=10)
    t.embeddirs_fn, _ = ssr.get_embedder(4, 0, scalar_factor=1)
    t.H_scaled, t.W_scaled, t.near, t.far, t.no_semantic_tree = H, W, 0.1, 10.0, False
    t.valid_colour_map = (torch.rand(C, 3, generator=torch.Generator().manual_seed(2)) * 255).to(torch.uint8).to(DEV)
    poses = torch.tensor(np.stack([orc.pose_spherical(a, -20.0, 3.0) for a in (10.0, 200.0)])).float()
    rays = ssr.create_rays(2, poses, H, W, 30.0, 30.0, W / 2, H / 2, 0.1, 10.0, convention="opengl")
    out = t.render_path(rays, save_dir=str(tmp_path), update_cluster=True)
    assert len(out) == 12 and t.training is True
    rgbs, disps, deps, vis_deps, sems, vis_sems, ents, vis_ents, albedos, shadings, residuals, cm = out
    assert cm.class_num == C and len(cm.clusters) == C
    t.training = False
    with torch.no_grad():
        d = t.render_rays(rays[1])
    t.training = True
    assert np.array_equal(rgbs[1].reshape(-1, 3), d["rgb_fine"].cpu().numpy())
    assert np.array_equal(deps[1].reshape(-1), d["depth_fine"].cpu().numpy())
    assert np.array_equal(albedos[1].reshape(-1, 3), d["albedo_fine"].cpu().numpy())
    g = lambda k, *s: d[k].cpu().numpy().reshape(H, W, *s)  # noqa: E731
    want = fo.ssr_frame(g("rgb_fine", 3), g("disp_fine"), g("depth_fine"), g("albedo_fine", 3), g("shading_fine"),
                        g("residual_fine", 3), g("sem_logits_fine", C), t.valid_colour_map.cpu().numpy())
    assert np.array_equal(sems[1], want["label8"]) and np.array_equal(vis_sems[1], want["vis_label8"])
    np.testing.assert_allclose(ents[1], want["entropy"], rtol=0, atol=2e-6)
    for f in ("rgb_001.png", "disp_001.png", "depth_001.png", "label_001.png", "vis_label_001.png", "entropy_001.png",
              "c001.png", "edit001.png"):
        assert os.path.exists(os.path.join(str(tmp_path), f)), f
