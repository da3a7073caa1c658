"""Short single-GPU target for ncu: the HBM-bound stage kernels at the headline shape (one 160 000-ray chunk: compositing,
resampling, merge) and the frame kernel on an 800x800 record."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from intrinsicnerf_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
N = 160000
g = torch.Generator().manual_seed(0)
z = torch.sort(torch.rand(N, 64, generator=g) * 4 + 2, dim=-1)[0].to(dev)
raw = torch.rand(N, 192, 11, generator=g).to(dev)
zf = torch.sort(torch.rand(N, 192, generator=g) * 4 + 2, dim=-1)[0].to(dev)
d = torch.randn(N, 3, generator=g).to(dev)
for _ in range(3):
    rec, w = ops.raw2outputs_rec(raw, zf, d, None, True, 0, False, True)
    zs, = ops.sample_pdf(0.5 * (z[:, 1:] + z[:, :-1]), torch.rand(N, 62, device=dev), 128, None)[:1]
    ops.merge_sorted(z, zs)
    ops.frame_finish(torch.rand(640000, 13, device=dev), 800, 800, 0, ("rgb8", "albedo8", "shading8", "residual8", "label8"), sub_step=2)
torch.cuda.synchronize()
print("done")
