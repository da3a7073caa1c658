"""Kernel-only timing of the TRAINING forward (inrf_mlp_fwd_train_tc: k_mlp_tc<.,STASH>) next to the inference launch on
the same rows.  Args: rays [variant].  Env INRF_TC_STASH_ABL (1 no mask words | 2 no bulk copies | 4 no copy waits)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from intrinsicnerf_b200 import _lib, ops  # noqa: E402
from oracle import nerf_oracle as orc  # noqa: E402
from tests.util import build_nets  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
variant = sys.argv[2] if len(sys.argv) > 2 else "ssr"
C = 28 if variant == "ssr" else 0
dev = torch.device("cuda:0")
coarse, fine, _, _ = build_nets(variant, C)
rays = orc.blender_rays(400, 400)[:n].contiguous().to(dev)
z = torch.sort(torch.rand(n, 192, device=dev) * 4 + 2, dim=-1)[0]
scale = 1.0 if variant == "object" else 10.0
M = n * 192
L = _lib.lib()
packed = fine.packed()
raw = torch.empty(M, ops.RAW_BASE + C, device=dev)
stash = torch.empty(int(L.inrf_mlp_stash_img_bytes(M)), dtype=torch.uint8, device=dev)


def train_fwd():
    ops.check(L.inrf_mlp_fwd_train_tc(ops._ptr(packed), fine.variant, C, 0, float(scale), None, None, ops._ptr(rays), ops._ptr(z), 192, None,
                                      M, ops._ptr(raw), ops._ptr(stash), ops._stream()))


def infer():
    ops.mlp_forward_rays(packed, fine.variant, C, rays, z, False, scale, "tc")


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / reps * 1e3


tiles = M / 128
it = -(-tiles // 148)
a, b = timeit(train_fwd), timeit(infer)
print(f"STASH_PERF variant={variant} abl={os.environ.get('INRF_TC_STASH_ABL', '0')} rays={n} tiles={tiles:.0f} iters={it:.0f} "
      f"train_fwd={a:.1f} us ({a / it:.2f} us/iter) inference={b:.1f} us ({b / it:.2f} us/iter) stash={stash.numel() / 1e6:.0f} MB "
      f"-> {stash.numel() / a / 1e6:.2f} TB/s")
