"""Bring-up check of the tensor-core training path: stash images vs the oracle's activations, then parameter
gradients vs fp64 autograd through the oracle.  Usage: python tests/tools/tcbwd_debug.py [object|ssr] [M]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
os.environ.setdefault("INRF_TC_CHECK", "1")
import torch  # noqa: E402

from intrinsicnerf_b200 import ops  # noqa: E402
from oracle import nerf_oracle as orc  # noqa: E402
from tests.util import build_nets  # noqa: E402

variant = sys.argv[1] if len(sys.argv) > 1 else "object"
M = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
C = 28 if variant == "ssr" else 0
endpoint = variant == "ssr" and len(sys.argv) > 3
dev = torch.device("cuda:0")
coarse, fine, pc, pf = build_nets(variant, C)


def decode(buf, n_tiles, slots):
    """[n_tiles*slots*16384] uint8 images -> [n_tiles, slots, 128, 64] float32 (undo the 128B swizzle)."""
    x = buf.view(torch.float16).reshape(n_tiles, slots, 16, 8, 8, 8)
    r = torch.arange(8, device=buf.device).view(8, 1)
    u = torch.arange(8, device=buf.device).view(1, 8)
    idx = (u ^ r).view(1, 1, 1, 8, 8, 1).expand(n_tiles, slots, 16, 8, 8, 8)
    return torch.gather(x, 4, idx).reshape(n_tiles, slots, 128, 64).float()


g = torch.Generator().manual_seed(3)
pts = (torch.rand(M, 3, generator=g) * 4 - 2)
vd = torch.nn.functional.normalize(torch.randn(M, 3, generator=g), dim=-1)
scale = 10.0 if variant == "ssr" else 1.0
out_ch = 11 + C + (128 if endpoint else 0)
g_raw = torch.randn(M, out_ch, generator=g) * 1e-3

# ---- oracle in fp64 with intermediates -------------------------------------------------------------------------
p64 = {k: v.double().clone().requires_grad_(True) for k, v in pc.items()}
emb = torch.cat([orc.posenc(pts.double(), 10, scale), orc.posenc(vd.double(), 4, 1.0)], -1)
want = orc.mlp_forward(p64, emb, variant, C, endpoint)
(want * g_raw.double()).sum().backward()

# explicit forward with retained pre-activation gradients, EMULATING the tensor-core arithmetic (operands rounded to
# fp16 with a straight-through gradient, sums exact) so that every ReLU takes the same branch as in the kernel: an
# exact-arithmetic oracle flips ~0.05 % of the masks, and one flipped (row, unit) is a full-size error in dZ
q64 = {k: v.detach().clone().requires_grad_(True) for k, v in p64.items()}
hn = orc.OBJECT_HEADS if variant == "object" else orc.SSR_HEADS
rnd = lambda x: x + (x.to(torch.float16).to(x.dtype) - x).detach()  # noqa: E731


def lin(name, x):
    return rnd(x) @ rnd(q64[name + ".weight"]).t() + q64[name + ".bias"]


zs = []
pe_r = emb[:, :63]
h = pe_r
for i, name in enumerate(orc.TRUNK):
    zl = lin(name, h)
    zl.retain_grad()
    zs.append(zl)
    h = torch.relu(zl)
    if i == 4:
        h = torch.cat([pe_r, h], -1)
z_a1 = lin(hn["albedo1"], h); z_a1.retain_grad()
z_s1 = lin(hn["shading1"], h); z_s1.retain_grad()
wv, wf = q64[hn["views"] + ".weight"], q64[hn["feature"] + ".weight"]
wc = wv[:, :256] @ wf
bc = wv[:, :256] @ q64[hn["feature"] + ".bias"] + q64[hn["views"] + ".bias"]
z_v = rnd(h) @ rnd(wc).t() + rnd(emb[:, 63:]) @ rnd(wv[:, 256:]).t() + bc
z_v.retain_grad()
sigma = h @ q64[hn["alpha"] + ".weight"].t() + q64[hn["alpha"] + ".bias"]
alb = torch.sigmoid(lin(hn["albedo2"], torch.relu(z_a1)))
shd = torch.sigmoid(lin(hn["shading2"], torch.relu(z_s1)))
h2 = torch.relu(z_v)
res = torch.sigmoid(lin(hn["residual"], h2))
cols = [alb * shd + res, sigma, alb, shd, res]
if C > 0:
    z_m = lin(hn["sem1"], h); z_m.retain_grad()
    cols.append(lin(hn["sem2"], torch.relu(z_m)))
if endpoint:
    cols.append(h2)
(torch.cat(cols, -1) * g_raw.double()).sum().backward()

# ---- ours ------------------------------------------------------------------------------------------------------
flat = coarse.flat_params_diff()
raw = ops.MlpTcFn.apply(flat, coarse.variant, C, endpoint, scale, "pts", pts.to(dev), vd.to(dev))
torch.cuda.synchronize()
err = float((raw.detach().cpu().double() - want.detach()).abs().max())
print(f"[fwd] raw max abs err {err:.3e}")
T = (M + 127) // 128
stash = raw.grad_fn.saved_tensors[3]
img = decode(stash, T, stash.numel() // (T * 16384))
names = orc.TRUNK
h = emb[:, :63]
for i, name in enumerate(names):
    h = torch.relu(orc._lin({k: v.detach() for k, v in p64.items()}, name, h))
    got = img[:, 2 + 4 * i:6 + 4 * i].permute(0, 2, 1, 3).reshape(T * 128, 256)[:M].cpu().double()
    print(f"[stash] H{i} max abs err {float((got - h).abs().max()):.3e} (max {float(h.abs().max()):.2f})")
    if i == 4:
        h = torch.cat([emb[:, :63], h], -1)
pe = img[:, 0].reshape(T * 128, 64)[:M, :63].cpu().double()
print(f"[stash] PE max abs err {float((pe - emb[:, :63]).abs().max()):.3e}")

(raw * g_raw.to(dev)).sum().backward()
torch.cuda.synchronize()
# ---- gradient images ----------------------------------------------------------------------------------------------
ws = ops._Workspace.bufs[str(dev) + "bwd"]
head = (4096 + (128 * 256 + 128) * 4 + 1023) // 1024 * 1024 + 48 * 32768
amax = float(g_raw.abs().max())
import math
S = 2.0 ** (7 - math.frexp(amax)[1])
bimg = decode(ws[head:head + T * 42 * 16384], T, 42) / S


def cmp(tag, slot0, n_chunks, ref):
    got = bimg[:, slot0:slot0 + n_chunks].permute(0, 2, 1, 3).reshape(T * 128, 64 * n_chunks)[:M, :ref.shape[1]].cpu().double()
    e = float((got - ref).abs().max() / (ref.abs().max() + 1e-30))
    print(f"[dz] {tag:8s} rel-to-max err {e:.3e}  (max |ref| {float(ref.abs().max()):.3e}, max |got| {float(got.abs().max()):.3e})")
    bad = ((got - ref).abs() > 0.02 * ref.abs().max()).nonzero()
    zg = int(((got == 0) & (ref != 0)).sum()), int(((got != 0) & (ref == 0)).sum())
    print(f"[dzd] {tag}: {bad.shape[0]} bad elements of {got.numel()}; got==0&ref!=0: {zg[0]}, got!=0&ref==0: {zg[1]}; "
          f"bad rows%128 {sorted(set((bad[:, 0] % 128).tolist()))[:20]} cols {sorted(set(bad[:, 1].tolist()))[:20]}")
    for r, c in bad[:6].tolist():
        print(f"[dzd]    ({r},{c}) got {float(got[r, c]):.4e} ref {float(ref[r, c]):.4e}")


cmp("DAS", 2, 4, torch.cat([z_a1.grad, z_s1.grad], -1))
cmp("DV", 6, 2, z_v.grad)
if C > 0:
    cmp("DS1", 8, 2, z_m.grad)
for i in range(7, -1, -1):
    cmp(f"DZ{i}", 10 + 4 * i, 4, zs[i].grad)
worst = 0.0
for name, p in coarse.named_parameters():
    a, b = q64[name].grad, p.grad.cpu().double()
    e = float((a - b).abs().max() / (a.abs().max() + 1e-30))
    worst = max(worst, e)
    print(f"[grad] {name:28s} rel-to-max err {e:.3e}   (max |g| {float(a.abs().max()):.3e})")
print(f"TCBWD {variant} M={M} endpoint={endpoint}: worst gradient error {worst:.3e}")
