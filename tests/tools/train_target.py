"""Short single-GPU target for ncu: three SSR training steps (render + backward, 1024 rays, 64+128, C=28) on the
tensor-core training path."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from intrinsicnerf_b200 import ssr  # noqa: E402
from oracle import nerf_oracle as orc  # noqa: E402
from tests.util import build_nets  # noqa: E402

dev = torch.device("cuda:0")
C, N = 28, 1024
coarse, fine, _, _ = build_nets("ssr", C)


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
for _ in range(3):
    for p in list(coarse.parameters()) + list(fine.parameters()):
        p.grad = None
    out = t.render_rays(rays)
    (ce(out["sem_logits_fine"], labels) + ce(out["sem_logits_coarse"], labels) + (out["rgb_fine"] ** 2).mean()
     + (out["rgb_coarse"] ** 2).mean() + out["albedo_fine"].mean() + out["shading_fine"].mean()).backward()
torch.cuda.synchronize()
print("done")
