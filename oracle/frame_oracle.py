"""CPU oracle for the full-image driver's per-frame conversions.  TEST INFRASTRUCTURE ONLY (see nerf_oracle.py):
only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may import this module.

numpy restatement of what render_path does to every rendered frame on the host:
  object fork  object_level/run_nerf.py:164-240   (to8b run_nerf_helpers.py:13)
  SSR fork     SSR/training/trainer.py:1241-1441
Pinned by tests/test_frame_oracle.py against tests/golden/frame.npz, which holds the arrays the UNMODIFIED
reference render_path handed to imageio.imwrite / Cluster_Manager (tests/golden/make_golden.py --frame).
"""
import numpy as np


def to8b(x):
    """run_nerf_helpers.py:13 / trainer.py:1241: (255*np.clip(x,0,1)).astype(np.uint8)."""
    with np.errstate(invalid="ignore"):
        return (255 * np.clip(np.asarray(x, dtype=np.float32), 0, 1)).astype(np.uint8)


def object_frame(rgb, disp, acc, albedo, shading, residual, update_cluster=True, acc_threshold=10):
    """One iteration of the loop at run_nerf.py:164-215: the arrays written as ###/a###/s###/res###/acc###.png
    and the cluster samples (albedo[::2, ::2], label[::2, ::2])."""
    label = (np.asarray(acc) > acc_threshold).astype(int)                 # :174
    out = {"rgb8": to8b(rgb), "albedo8": to8b(albedo), "shading8": to8b(shading), "residual8": to8b(residual),
           "label8": to8b(label.astype(np.float32)), "labels64": label.astype(np.int64)}
    if update_cluster:
        out["sample_pixels"] = np.asarray(albedo)[::2, ::2, :].reshape(-1, 3)   # :183-184
        out["sample_labels"] = label[::2, ::2].reshape(-1, 1)                   # :185-186
    return out


def edit_recompose(result, shading, residual):
    """run_nerf.py:228-240 / trainer.py:1428-1441: c### = to8b(result); edit### = to8b(result*shading + residual)."""
    shape = np.asarray(result).shape
    edit = np.asarray(result).reshape(-1, 3) * np.asarray(shading).reshape(-1, 1) + np.asarray(residual).reshape(-1, 3)
    return to8b(result), to8b(edit.reshape(shape))


def ssr_frame(rgb, disp, depth, albedo, shading, residual, sem_logits, colour_map, update_cluster=True):
    """One iteration of the loop at trainer.py:1248-1389 for the maps of the last pass.  sem_logits [H,W,C] float32,
    colour_map [C,3] uint8 (valid_colour_map)."""
    x = np.asarray(sem_logits, dtype=np.float32)
    m = x.max(-1, keepdims=True)
    e = np.exp(x - m)
    z = e.sum(-1, keepdims=True)
    soft = e / z
    log_soft = (x - m) - np.log(z)
    label = np.argmax(soft, -1)                                            # logits_2_label, :1243
    entropy = np.sum(-log_soft * soft, -1)                                 # logits_2_uncertainty, :1244
    with np.errstate(invalid="ignore"):
        out = {"rgb8": to8b(rgb), "albedo8": to8b(albedo), "shading8": to8b(shading), "residual8": to8b(residual),
               "disp16": np.asarray(disp).astype(np.uint16), "depth_mm16": (np.asarray(depth) * 1000).astype(np.uint16),   # :1351-1352
               "label8": label.astype(np.uint8), "vis_label8": np.asarray(colour_map)[label].astype(np.uint8),
               "entropy": entropy.astype(np.float32), "entropy8": to8b(entropy), "labels64": label.astype(np.int64)}
    if update_cluster:
        out["sample_pixels"] = np.asarray(albedo)[::2, ::2, :].reshape(-1, 3)   # :1330-1332
        out["sample_labels"] = out["label8"][::2, ::2].reshape(-1, 1)           # :1333-1334
    return out
