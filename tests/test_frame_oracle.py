"""The frame oracle (oracle/frame_oracle.py) against what the unmodified reference render_path handed to
imageio.imwrite / Cluster_Manager (tests/golden/frame.npz).  Bit-exact: byte planes."""
import numpy as np

from oracle import frame_oracle as fo
from tests.util import load_golden

OBJ_PLANES = ("rgb8", "albedo8", "shading8", "residual8", "label8")
SSR_PLANES = ("rgb8", "albedo8", "shading8", "residual8", "disp16", "depth_mm16", "label8", "vis_label8", "entropy8")


def test_object_frame_planes_equal_reference(golden_dir):
    g = load_golden(golden_dir, "frame.npz")
    px, lb = [], []
    for i in range(2):
        m = {k: g[f"obj{i}_{k}"] for k in ("rgb", "disp", "acc", "albedo", "shading", "residual")}
        o = fo.object_frame(**m)
        for k in OBJ_PLANES:
            assert np.array_equal(o[k], g[f"obj{i}_{k}"]), (i, k)
        assert np.array_equal(o["labels64"].reshape(-1, 1), g[f"obj{i}_dc_label"])
        c8, e8 = fo.edit_recompose(g[f"obj{i}_dc_result"].reshape(m["albedo"].shape), m["shading"], m["residual"])
        assert np.array_equal(c8, g[f"obj{i}_c8"]) and np.array_equal(e8, g[f"obj{i}_edit8"])
        px.append(o["sample_pixels"]), lb.append(o["sample_labels"])
    assert np.array_equal(np.concatenate(px, 0), g["obj_uc_pixels"], equal_nan=True)
    assert np.array_equal(np.concatenate(lb, 0), g["obj_uc_labels"])
    assert g["obj1_label8"].max() == 255 and g["obj0_label8"].max() == 0      # the edge frame has acc > 10, a real frame never


def test_ssr_frame_planes_equal_reference(golden_dir):
    g = load_golden(golden_dir, "frame.npz")
    H, W, C = int(g["ssr_H"]), int(g["ssr_W"]), int(g["ssr_C"])
    px, lb = [], []
    for i in range(2):
        r = lambda k, *s: g[f"ssr{i}_{k}_fine"].reshape(H, W, *s)  # noqa: E731
        o = fo.ssr_frame(r("rgb", 3), r("disp"), r("depth"), r("albedo", 3), r("shading"), r("residual", 3), r("sem_logits", C),
                         g["ssr_colour_map"])
        for k in SSR_PLANES:
            assert np.array_equal(o[k], g[f"ssr{i}_{k}"]), (i, k)
        assert np.array_equal(o["label8"], g["ssr_sems"][i])
        np.testing.assert_allclose(o["entropy"], g["ssr_entropys"][i], rtol=0, atol=2e-6)
        c8, e8 = fo.edit_recompose(g[f"ssr{i}_dc_result"].reshape(H, W, 3), r("shading"), r("residual", 3))
        assert np.array_equal(c8, g[f"ssr{i}_c8"]) and np.array_equal(e8, g[f"ssr{i}_edit8"])
        px.append(o["sample_pixels"]), lb.append(o["sample_labels"])
    assert np.array_equal(np.stack(px, 0), g["ssr_uc_pixels"], equal_nan=True)
    assert np.array_equal(np.stack(lb, 0), g["ssr_uc_labels"])
    assert g["ssr1_label8"][0, 0] == 0 and g["ssr1_label8"][0, 1] == 2        # exact ties: first maximum
