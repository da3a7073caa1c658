"""Timing of the section-8f rows next to the hot path: ray generation and the fused training losses, with the
oracle's PyTorch composition (same ops as the reference, run eagerly on the GPU) as the baseline."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from intrinsicnerf_b200 import ops  # noqa: E402
from oracle import nerf_oracle as orc  # noqa: E402

dev = torch.device("cuda:0")


def timeit(fn, reps=50, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / reps * 1e3       # microseconds


# ---- rays: 800x800 frame -----------------------------------------------------------------------------
H = W = 800
K = orc.blender_intrinsics(H, W)
c2w = orc.pose_spherical(30.0, -30.0, 4.0)[:3, :4]
t_k = timeit(lambda: ops.get_rays_packed(H, W, K, c2w, 2.0, 6.0, dev))


def torch_rays():
    c = torch.as_tensor(c2w, device=dev)
    jj, ii = torch.meshgrid(torch.linspace(0, H - 1, H, device=dev), torch.linspace(0, W - 1, W, device=dev), indexing="ij")
    dirs = torch.stack([(ii - K[0][2]) / K[0][0], -(jj - K[1][2]) / K[1][1], -torch.ones_like(ii)], -1)
    rd = torch.sum(dirs[..., None, :] * c[:3, :3], -1).reshape(-1, 3)
    ro = c[:3, -1].expand(rd.shape)
    vd = rd / torch.norm(rd, dim=-1, keepdim=True)
    return torch.cat([ro, rd, 2.0 * torch.ones_like(rd[:, :1]), 6.0 * torch.ones_like(rd[:, :1]), vd], -1)


t_t = timeit(torch_rays)
print(f"AUX rays 800x800: kernel {t_k:.1f} us ({H * W * 44 / t_k / 1e3:.0f} GB/s written) vs eager PyTorch composition {t_t:.1f} us")
pix = torch.randint(0, H * W, (4096,), device=dev)
t_p = timeit(lambda: ops.rays_from_pixels(pix, H, W, K[0][0], K[1][1], K[0][2], K[1][2], c2w, 2.0, 6.0))
print(f"AUX rays for 4096 sampled pixels: {t_p:.1f} us per call (replaces a gather from a precomputed table)")

# ---- losses: one training step's worth (N_rand = 1024 pixels + 1024 neighbours) ---------------------------
N = 2048
g = torch.Generator().manual_seed(0)
mk = lambda *s: torch.rand(*s, generator=g).to(dev)   # noqa: E731
rgb, alb, sh, res = mk(N, 3).requires_grad_(True), mk(N, 3).requires_grad_(True), mk(N).requires_grad_(True), (mk(N, 3) * 0.2).requires_grad_(True)
gt, mask, tgt = mk(N, 3), (mk(N) > 0.3).float(), mk(N, 3)
w = torch.tensor([1.0, 1.0, 0.7, 0.01, 0.02, 0.03, 0.5, 0.4], device=dev)


def fused():
    for t in (rgb, alb, sh, res):
        t.grad = None
    (ops.intrinsic_losses(rgb, alb, sh, res, gt, mask, tgt, "object") * w).sum().backward()


def eager():
    for t in (rgb, alb, sh, res):
        t.grad = None
    (orc.intrinsic_losses(rgb, alb, sh, res, gt, mask, tgt, "object") * w).sum().backward()


t_f, t_e = timeit(fused), timeit(eager)
fused(); ga = alb.grad.clone(); eager()
err = float((ga - alb.grad).abs().max() / alb.grad.abs().max())
print(f"AUX losses fwd+bwd, N={N}: fused kernels {t_f:.1f} us vs eager PyTorch composition {t_e:.1f} us ({t_e / t_f:.1f}x), "
      f"albedo-gradient difference {err:.1e}")

# ---- frame: render_path's per-frame conversions on an 800x800 record (section 8f row 3) -----------------------
import time  # noqa: E402

import numpy as np  # noqa: E402

from oracle import frame_oracle as fo  # noqa: E402

rec = torch.rand(H * W, 13, generator=g).to(dev)
planes = ("rgb8", "albedo8", "shading8", "residual8", "label8", "labels64")
t_fr = timeit(lambda: ops.frame_finish(rec, H, W, 0, planes, sub_step=2))
nbytes = H * W * (52 + 11 + 8) + (H // 2) * (W // 2) * 20
host8 = [torch.empty(H, W, c, dtype=torch.uint8).pin_memory() for c in (3, 3, 1, 3, 1)]


def ours_e2e():
    f = ops.frame_finish(rec, H, W, 0, planes, sub_step=2)
    for dst, k in zip(host8, planes[:5]):
        dst.copy_(f[k].reshape(dst.shape), non_blocking=True)
    torch.cuda.synchronize()


def reference_style():
    m = [rec[:, 0:3], rec[:, 5:8], rec[:, 8], rec[:, 9:12], rec[:, 4]]
    host = [x.reshape(H, W, -1).cpu().numpy() for x in m]           # .cpu().numpy() per map (run_nerf.py:168-174)
    out = [fo.to8b(x) for x in host[:4]]
    label = (host[4] > 10).astype(int)
    out.append(fo.to8b(label.astype(np.float32)))
    return out, host[1][::2, ::2, :].reshape(-1, 3), label[::2, ::2].reshape(-1, 1)


def wall(fn, reps=10):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps * 1e3


print(f"AUX frame 800x800: inrf_frame_finish {t_fr:.1f} us ({nbytes / t_fr / 1e3:.0f} GB/s moved); per-frame host hand-over "
      f"{wall(ours_e2e):.2f} ms (kernel + D2H of 11 B/pixel) vs reference-style D2H of float maps + numpy to8b {wall(reference_style):.2f} ms")
