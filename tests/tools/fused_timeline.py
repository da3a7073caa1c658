"""Development: one fused 64+128 chunk through a library built with -DINRF_TC_TIMELINE (tools/gpu_timeline.sh); the
library prints the clock64 stamps of the issuer, two epilogue warps and back-end warp 12 of CTA 0 for both launches."""
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
for _ in range(2):
    ops.render_chunk(rays, pc, pf, white_bkgd=True)
    torch.cuda.synchronize()
    sys.stderr.write("TCTL ---- chunk done\n")
ops.poll_status()
