"""Timing of one SSR training step's render + backward (BASELINE config 5 shape on one GPU: 1024 rays, 64+128
samples, Semantic_NeRF with C=28) through our training path (fp32 kernels behind autograd) next to the same
algorithm composed from eager PyTorch ops on the same GPU (the oracle's code with torch's default device set to
CUDA - what the reference's own GPU path executes: cuBLAS fp32 GEMMs + elementwise kernels)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from oracle import nerf_oracle as orc  # noqa: E402
from tests.util import build_nets  # noqa: E402

dev = torch.device("cuda:0")
C, N = 28, int(sys.argv[1]) if len(sys.argv) > 1 else 1024


def timeit(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / reps


from intrinsicnerf_b200 import ssr  # noqa: E402

coarse, fine, pc, pf = build_nets("ssr", C)


class T(ssr.SSRRenderer):
    pass


t = T()
t.N_samples, t.N_importance, t.perturb, t.raw_noise_std = 64, 128, 1.0, 1.0
t.white_bkgd, t.enable_semantic, t.num_valid_semantic_class, t.endpoint_feat = False, True, C, False
t.netchunk = t.chunk = 32768
t.ssr_net_coarse, t.ssr_net_fine = coarse, fine
t.embed_fn, _ = ssr.get_embedder(10, 0, scalar_factor=10)
t.embeddirs_fn, _ = ssr.get_embedder(4, 0, scalar_factor=1)
t.training = True
rays = orc.replica_rays(120, 160)[:N].contiguous().to(dev)
labels = (torch.arange(N) % C).to(dev)
ce = torch.nn.functional.cross_entropy
params = list(coarse.parameters()) + list(fine.parameters())


def ours():
    for p in params:
        p.grad = None
    out = t.render_rays(rays)
    (ce(out["sem_logits_fine"], labels) + ce(out["sem_logits_coarse"], labels) + (out["rgb_fine"] ** 2).mean()
     + (out["rgb_coarse"] ** 2).mean() + out["albedo_fine"].mean() + out["shading_fine"].mean()).backward()


def ours_fwd_only():
    with torch.no_grad():
        t.render_rays(rays)


from intrinsicnerf_b200 import ops  # noqa: E402

ms_ours = timeit(ours)                       # default training mode: tensor cores (ops.MlpTcFn)
ops.set_default_precision("fp32")
ms_fp32 = timeit(ours)                       # strict fp32 CUDA-core mode (ops.MlpFn)
ops.set_default_precision("tc")
ms_inf = timeit(ours_fwd_only)

torch.set_default_device("cuda")
cc = {k: v.clone().to(dev).requires_grad_(True) for k, v in pc.items()}
cf = {k: v.clone().to(dev).requires_grad_(True) for k, v in pf.items()}


def eager():
    for p in list(cc.values()) + list(cf.values()):
        p.grad = None
    tr = torch.rand(N, 64)
    u = torch.rand(N, 128)
    r = orc.render_rays(rays, cc, cf, "ssr", C, pe_scale_pts=10.0, t_rand=tr, u=u, noise_coarse=torch.randn(N, 64),
                        noise_fine=torch.randn(N, 192))
    (ce(r["fine"]["sem"], labels) + ce(r["coarse"]["sem"], labels) + (r["fine"]["rgb"] ** 2).mean()
     + (r["coarse"]["rgb"] ** 2).mean() + r["fine"]["albedo"].mean() + r["fine"]["shading"].mean()).backward()


ms_eager = timeit(eager)
flop = N * 256 * 2 * (692224 + 128 * C) * 3
print(f"TRAIN step (render+backward) N={N} rays, 64+128, C={C}: tensor-core mode {ms_ours:.2f} ms ({flop / ms_ours / 1e9:.1f} TFLOP/s), "
      f"strict-fp32 mode {ms_fp32:.2f} ms, eager PyTorch composition on the same GPU {ms_eager:.2f} ms "
      f"({ms_eager / ms_ours:.2f}x vs tensor-core mode); eval render of the same rays {ms_inf:.2f} ms")
