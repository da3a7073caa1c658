"""Diagnose the MLP backward: isolate the semantic branch by zeroing parts of the upstream gradient."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from oracle import nerf_oracle as orc  # noqa: E402
from tests.util import build_nets  # noqa: E402

dev = torch.device("cuda:0")
C = 5
coarse, fine, pc, pf = build_nets("ssr", C)
gen = torch.Generator().manual_seed(21)
M = 150
pts = torch.rand(M, 3, generator=gen) * 6 - 3
vd = torch.nn.functional.normalize(torch.randn(M, 3, generator=gen), dim=-1)
g_full = torch.randn(M, 11 + C, generator=gen)


def run(mask_name, g_raw):
    p64 = {k: v.double().to(dev).requires_grad_(True) for k, v in pf.items()}
    emb = torch.cat([orc.posenc(pts.double().to(dev), 10, 10.0), orc.posenc(vd.double().to(dev), 4)], -1)
    out64 = orc.mlp_forward(p64, emb, "ssr", C, False)
    (out64 * g_raw.double().to(dev)).sum().backward()
    fine.zero_grad()
    out = fine.evaluate("pts", pts.to(dev), vd.to(dev), False, 10.0)
    (out * g_raw.to(dev)).sum().backward()
    errs = {}
    for name, p in fine.named_parameters():
        a, b = p64[name].grad.float(), p.grad
        errs[name] = float((a - b).abs().max()) / (float(a.abs().max()) + 1e-30)
    top = sorted(errs.items(), key=lambda kv: -kv[1])[:5]
    print(mask_name, " | ".join(f"{k}:{v:.2e}" for k, v in top))


g = g_full.clone(); g[:, :11] = 0
run("sem-only   ", g)
g = g_full.clone(); g[:, 11:] = 0
run("no-sem     ", g)
g = torch.zeros_like(g_full); g[:, 11] = g_full[:, 11]
run("sem-chan0  ", g)
run("all        ", g_full)
# stash check: relu(sem1(h)) from the oracle vs what the forward stashed
from intrinsicnerf_b200 import ops, _lib  # noqa: E402
import ctypes as Cc  # noqa: E402
flat = fine.flat_params()
packed = ops.pack_weights(flat, 1, C)
L = _lib.lib()
raw = torch.empty(M, 11 + C, device=dev)
stash = torch.zeros(M, int(L.inrf_stash_floats_per_row()), device=dev)
p = lambda t: None if t is None else Cc.c_void_p(t.data_ptr())
ptsd, vdd = pts.to(dev).contiguous(), vd.to(dev).contiguous()
L.inrf_mlp_fwd_train(p(packed), 1, C, 0, 10.0, p(ptsd), p(vdd), None, None, 1, None, M, p(raw), p(stash), None)
torch.cuda.synchronize()
p32 = {k: v.to(dev) for k, v in pf.items()}
emb = torch.cat([orc.posenc(ptsd, 10, 10.0), orc.posenc(vdd, 4)], -1)
h = emb[:, :63]
hs = []
for i in range(8):
    h = torch.relu(torch.nn.functional.linear(h, p32[f"pts_linears.{i}.weight"], p32[f"pts_linears.{i}.bias"]))
    hs.append(h)
    if i == 4:
        h = torch.cat([emb[:, :63], h], -1)
sem1 = torch.relu(torch.nn.functional.linear(h, p32["semantic_linear.0.0.weight"], p32["semantic_linear.0.0.bias"]))
print("stash H7 err", float((stash[:, 7 * 256:8 * 256] - hs[7]).abs().max()), "stash SEM1 err", float((stash[:, 2688:2816] - sem1).abs().max()),
      "SEM1 max", float(sem1.abs().max()))
