"""Short single-GPU target for ncu: a few fine-pass MLP launches and one full render chunk."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from intrinsicnerf_b200 import ops  # noqa: E402
from oracle import nerf_oracle as orc  # noqa: E402
from tests.util import build_nets  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
prec = sys.argv[2] if len(sys.argv) > 2 else "tc"
dev = torch.device("cuda:0")
coarse, fine, _, _ = build_nets("object", opaque=False)
rays = orc.blender_rays(400, 400)[:n].contiguous().to(dev)
for _ in range(3):
    o = ops.render_chunk(rays, coarse.packed(), fine.packed(), white_bkgd=True, precision=prec, want_z=True)
for _ in range(3):
    ops.mlp_forward_rays(fine.packed(), 0, 0, rays, o["z_fine"], precision=prec)
torch.cuda.synchronize()
print("done", n, prec)
