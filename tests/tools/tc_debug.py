"""GPU bring-up diagnostics for the tcgen05 MLP kernel: compares k_mlp_tc with k_mlp_fp32 channel by
channel on a few hundred rows and prints the watchdog record if the kernel stalled."""
import os
import sys

os.environ.setdefault("INRF_TC_CHECK", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from intrinsicnerf_b200 import ops  # noqa: E402
from tests.util import build_nets  # noqa: E402


def run(variant, C, M, endpoint=False):
    dev = torch.device("cuda:0")
    coarse, fine, pc, pf = build_nets(variant, C)
    g = torch.Generator().manual_seed(1)
    pts = (torch.rand(M, 3, generator=g) * 6 - 3).to(dev)
    vd = torch.nn.functional.normalize(torch.randn(M, 3, generator=g), dim=-1).to(dev)
    scale = 1.0 if variant == "object" else 10.0
    a = ops.mlp_forward(fine.packed(), fine.variant, C, pts, vd, endpoint, scale, "fp32")
    try:
        b = ops.mlp_forward(fine.packed(), fine.variant, C, pts, vd, endpoint, scale, "tc")
    except Exception as e:  # noqa: BLE001
        print(f"[{variant} C={C} M={M}] TC FAILED: {e}")
        return False
    torch.cuda.synchronize()
    err = (a - b).abs()
    rel = err / a.abs().clamp_min(1e-2)
    print(f"[{variant} C={C} M={M} ep={endpoint}] max abs {err.max().item():.3e} max rel {rel.max().item():.3e} "
          f"nan {int(torch.isnan(b).sum())}")
    names = ["r", "g", "b", "sigma", "alb_r", "alb_g", "alb_b", "shading", "res_r", "res_g", "res_b"]
    for c in range(min(a.shape[1], 11)):
        print(f"   ch {c:2d} {names[c]:8s} abs {err[:, c].max().item():.3e}  fp32[0]={a[0, c].item():+.5f} tc[0]={b[0, c].item():+.5f}"
              f"  fp32[-1]={a[-1, c].item():+.5f} tc[-1]={b[-1, c].item():+.5f}")
    if a.shape[1] > 11:
        print(f"   extra channels max abs {err[:, 11:].max().item():.3e}")
    rows = err.max(dim=1)[0]
    bad = (rows > 1e-2).nonzero().flatten()
    print(f"   rows with abs err > 1e-2: {bad.numel()} / {M}   first: {bad[:16].tolist()}")
    return bool(err.max() < 1e-3)


if __name__ == "__main__":
    ok = run("object", 0, 128)
    ok &= run("object", 0, 300)
    ok &= run("object", 0, 128 * 148 * 3 + 17)
    ok &= run("ssr", 28, 512)
    ok &= run("ssr", 28, 512, True)
    ok &= run("ssr", 0, 200)
    print("TC_DEBUG", "PASS" if ok else "FAIL")
