"""Chunk-level timing of the fused renderer (2 launches) against the stage path (8 launches), same rays and weights:
full 64+128 chunk, coarse-only chunk, and the two fused launches separately (the coarse launch carries the resampler)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from intrinsicnerf_b200 import ops  # noqa: E402
from oracle import nerf_oracle as orc  # noqa: E402
from tests.util import build_nets  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 160000
dev = torch.device("cuda:0")
coarse, fine, _, _ = build_nets("object")
rays = orc.blender_rays(400, 400)[:n].contiguous().to(dev)
pc, pf = coarse.packed(), fine.packed()


def t(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / reps


full_f = t(lambda: ops.render_chunk(rays, pc, pf, white_bkgd=True))
full_s = t(lambda: ops.render_chunk(rays, pc, pf, white_bkgd=True, want_weights=True))
co_f = t(lambda: ops.render_chunk(rays, pc, None, white_bkgd=True, n_importance=0))
z = ops.coarse_z(rays, 64)
co_mlp = t(lambda: ops.mlp_forward_rays(pc, 0, 0, rays, z))
zf = ops.render_chunk(rays, pc, pf, white_bkgd=True, want_z=True)["z_fine"]
fi_mlp = t(lambda: ops.mlp_forward_rays(pf, 0, 0, rays, zf))
ops.poll_status()
print(f"FUSED_PERF rays={n}: full chunk fused {full_f:.2f} ms vs staged {full_s:.2f} ms ({n / full_f / 1e3:.3f} vs {n / full_s / 1e3:.3f} Mrays/s); "
      f"coarse-only fused (composite, no resampling) {co_f:.2f} ms vs bare coarse MLP {co_mlp:.2f} ms; bare fine MLP {fi_mlp:.2f} ms; "
      f"=> fused coarse+resample launch ~ {full_f - fi_mlp:.2f} ms if the fine launch cost the bare MLP")
