"""Kernel-only timing of the MLP launch (fine-pass shape) + a quick correctness check vs the fp32 kernel.
Env: INRF_TC_CLUSTER, INRF_TC_BIASMMA select the variant (read once per process)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
os.environ.setdefault("INRF_TC_CHECK", "0")
import torch  # noqa: E402

from intrinsicnerf_b200 import ops  # noqa: E402
from oracle import nerf_oracle as orc  # noqa: E402
from tests.util import build_nets  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 160000
variant = sys.argv[2] if len(sys.argv) > 2 else "object"
C = 28 if variant == "ssr" else 0
dev = torch.device("cuda:0")
coarse, fine, _, _ = build_nets(variant, C)
rays = orc.blender_rays(400, 400)[:n].contiguous().to(dev)
z = torch.sort(torch.rand(n, 192, device=dev) * 4 + 2, dim=-1)[0]
scale = 1.0 if variant == "object" else 10.0
# correctness on the first 4096 rays
a = ops.mlp_forward_rays(fine.packed(), fine.variant, C, rays[:4096], z[:4096], False, scale, "fp32")
b = ops.mlp_forward_rays(fine.packed(), fine.variant, C, rays[:4096], z[:4096], False, scale, "tc")
torch.cuda.synchronize()
err = (a - b).abs().max().item()
chk = int(b.contiguous().view(torch.int32).to(torch.int64).sum().item())     # bit-level checksum: variants must agree exactly
for _ in range(2):
    ops.mlp_forward_rays(fine.packed(), fine.variant, C, rays, z, False, scale, "tc")
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 5
s.record()
for _ in range(reps):
    ops.mlp_forward_rays(fine.packed(), fine.variant, C, rays, z, False, scale, "tc")
e.record()
torch.cuda.synchronize()
ms = s.elapsed_time(e) / reps
flop = n * 192 * (1318912 if variant == "object" else 2 * (692224 + 128 * C))
print(f"TC_PERF variant={variant} cluster={os.environ.get('INRF_TC_CLUSTER', '2')} biasmma={os.environ.get('INRF_TC_BIASMMA', '1')} "
      f"rows={n * 192} ms={ms:.3f} us_per_tile={ms * 1e3 / (n * 192 / 128 / 148):.2f} TFLOPs={flop / ms / 1e9:.1f} "
      f"rays_per_s_equiv={n / ms * 1e3 * 192 / 256:.0f} max_abs_err_vs_fp32={err:.3e} checksum={chk} "
      f"split={os.environ.get('INRF_TC_SPLIT', 'default')}")
