"""Timing of the clustering kernels on the synthetic albedo mixture of SURVEY section 8d
(P = 1 000 000 px, 12 modes) and of the per-image / per-step dest_color lookups."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from intrinsicnerf_b200 import cluster as cl  # noqa: E402
from oracle import cluster_oracle as co  # noqa: E402

dev = torch.device("cuda:0")
P = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
px, which = co.synthetic_albedo(P, n_modes=12, seed=0)


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps, out


c = cl.Cluster(device=dev)
t_fit, _ = timed(lambda: c.update_center(px.numpy(), band_factor=0.5), reps=1)
A, K = c.anchors.shape[0], c.rgb_centers.shape[0]
q_img = px[:160000].to(dev)
q_step = px[:2048].to(dev)
t_img, _ = timed(lambda: c.dest_color(q_img))
t_step, _ = timed(lambda: c.dest_color(q_step), reps=20)
print(f"CLUSTER_BENCH P={P} anchors={A} clusters={K} update_center={t_fit:.3f}s "
      f"dest_color_160k={t_img * 1e3:.3f}ms ({160000 / t_img / 1e6:.1f} Mpx/s) dest_color_2048={t_step * 1e6:.1f}us")
# CPU comparison on a bounded sample with the oracle (reference algorithm, torch CPU + numpy)
t0 = time.perf_counter()
idx, _ = co.nearest_anchor(c.anchors.cpu(), co.map_color(px[:20000]))
t_cpu = time.perf_counter() - t0
print(f"CLUSTER_BENCH cpu_oracle nearest_anchor 20000 px x {A} anchors: {t_cpu:.3f}s ({20000 / t_cpu / 1e3:.1f} kpx/s, {torch.get_num_threads()} threads)")
# the reference's update_center core (sklearn on the host, cluster.py:136-141) on a bounded sample of the same mixture
try:
    import numpy as np
    from sklearn.cluster import MeanShift, estimate_bandwidth
    Ps = min(P, 200000)
    X = co.map_color(px[:Ps]).numpy()
    t0 = time.perf_counter()
    bw = max(estimate_bandwidth(X, quantile=0.3, n_samples=5000) * 0.5, 0.01)
    ms = MeanShift(bandwidth=bw, bin_seeding=True).fit(X)
    t_sk = time.perf_counter() - t0
    t0 = time.perf_counter()
    c2 = cl.Cluster(device=dev)
    c2.update_center(px[:Ps].numpy(), band_factor=0.5)
    torch.cuda.synchronize()
    t_us = time.perf_counter() - t0
    print(f"CLUSTER_BENCH update_center on {Ps} px: sklearn estimate_bandwidth + MeanShift(bin_seeding) on the host {t_sk:.2f}s "
          f"({len(ms.cluster_centers_)} clusters) vs CUDA path {t_us:.3f}s ({c2.rgb_centers.shape[0]} clusters)")
except ImportError:
    pass
