"""NCCL check of the sharded paths on real GPUs (launched with torchrun, one rank per GPU):
  1. render_frame_pipelined: a 160x160 frame sharded over the ranks, gather overlapped chunk by chunk, must equal
     the single-GPU frame bit for bit on every rank;
  2. data-parallel training step: ray_shard + gather_maps_for_loss + allreduce_gradients must give the gradient of
     the single-process step on the whole batch (fp32 training mode: deterministic up to atomics order).
Prints one MULTI line per check from rank 0."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from intrinsicnerf_b200 import object_level as ol, ops, parallel  # noqa: E402
from oracle import nerf_oracle as orc  # noqa: E402
from tests.util import build_nets  # noqa: E402

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
coarse, fine, _, _ = build_nets("object", device=dev)
for m in (coarse, fine):
    parallel.broadcast_weights(m, 0)

# ---- 1. sharded frame ------------------------------------------------------------------------------------------------
H = W = 160
rays = orc.blender_rays(H, W).to(dev)
pc, pf = coarse.packed(), fine.packed()
rec_fn = lambda r: ops.render_chunk(r, pc, pf, white_bkgd=True)["rec_fine"]  # noqa: E731
with torch.no_grad():
    full = rec_fn(rays)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    got = parallel.render_frame_pipelined(rays, rec_fn, 2048)
    ev0.record()
    got = parallel.render_frame_pipelined(rays, rec_fn, 2048)
    ev1.record()
    torch.cuda.synchronize()
same = torch.equal(got, full)
flag = torch.tensor([int(same)], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"MULTI frame {H}x{W} over {world} GPUs: pipelined gather == single-GPU frame on every rank: {bool(flag.item())}; "
          f"{ev0.elapsed_time(ev1):.2f} ms per frame")

# ---- 1b. image-sharded render_path (config 4 shape: views round-robin over the ranks) -----------------------------------
import contextlib  # noqa: E402
import io  # noqa: E402
import tempfile  # noqa: E402

import numpy as np  # noqa: E402

e10, _ = ol.get_embedder(10, 0)
e4, _ = ol.get_embedder(4, 0)
kwp = dict(network_fn=coarse, network_fine=fine, network_query_fn=ol._FusedQuery(e10, e4, 65536), N_samples=64, N_importance=128,
           perturb=False, white_bkgd=True, raw_noise_std=0., use_viewdirs=True, ndc=False, lindisp=False, near=2., far=6.)
Hs = Ws = 48
Ks = orc.blender_intrinsics(Hs, Ws)
poses = [torch.tensor(orc.pose_spherical(a, -30.0, 4.0)).float() for a in (0.0, 50.0, 100.0, 150.0, 200.0)]
with tempfile.TemporaryDirectory() as d1, tempfile.TemporaryDirectory() as d2, contextlib.redirect_stdout(io.StringIO()):
    r1, dsp1, cm1 = ol.render_path(poses, (Hs, Ws, Ks[0][0]), Ks, 4096, kwp, savedir=d1, update_cluster=True)
    r2, dsp2, cm2 = ol.render_path(poses, (Hs, Ws, Ks[0][0]), Ks, 4096, kwp, savedir=d2, update_cluster=True, sharded=True)
    mine = parallel.image_shard(len(poses), rank, world)
    files_ok = all(open(os.path.join(d1, f"{p}{i:03d}.png"), "rb").read() == open(os.path.join(d2, f"{p}{i:03d}.png"), "rb").read()
                   for i in mine for p in ("", "a", "s", "res", "acc", "c", "edit"))
    n_files = len(os.listdir(d2))
ok = np.array_equal(r1, r2) and np.array_equal(dsp1, dsp2, equal_nan=True) and files_ok and n_files == 7 * len(mine) \
    and torch.equal(cm1.clusters[0].anchors, cm2.clusters[0].anchors) and torch.equal(cm1.clusters[0].rgb_centers, cm2.clusters[0].rgb_centers)
flag = torch.tensor([int(ok)], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"MULTI render_path over {world} GPUs ({len(poses)} views {Hs}x{Ws}, update_cluster): maps, PNG bytes and cluster state equal the "
          f"single-process run on every rank: {bool(flag.item())}")

# ---- 2. data-parallel training step ----------------------------------------------------------------------------------------
ops.set_default_precision("fp32")
N = 512
kw = dict(network_fn=coarse, network_fine=fine, network_query_fn=ol._FusedQuery(e10, e4, 65536), N_samples=64, N_importance=128,
          perturb=0., white_bkgd=True, raw_noise_std=0.)
batch = rays[torch.randperm(H * W, generator=torch.Generator().manual_seed(0))[:N].to(dev)]
target = torch.rand(N, 3, generator=torch.Generator().manual_seed(1)).to(dev)


def loss_fn(m):
    h = N // 2                                       # pixel i is paired with pixel i + N/2, as compute_intrinsic_loss does
    return ((m["rgb_map"] - target) ** 2).mean() + ((m["rgb0"] - target) ** 2).mean() \
        + ((m["albedo_map"][:h] - m["albedo_map"][h:]) ** 2).mean() + m["shading_map"].mean()


nets = [coarse, fine]
for n_ in nets:
    n_.zero_grad()
loss_fn(ol.render_rays(batch, **kw)).backward()
ref = torch.cat([p.grad.reshape(-1) for n_ in nets for p in n_.parameters()]).clone()
for n_ in nets:
    n_.zero_grad()
a, b = parallel.ray_shard(N, rank, world)
loc = ol.render_rays(batch[a:b], **kw)
maps = parallel.gather_maps_for_loss({k: loc[k] for k in ("rgb_map", "rgb0", "albedo_map", "shading_map")}, N)
loss_fn(maps).backward()
parallel.allreduce_gradients(nets)
got = torch.cat([p.grad.reshape(-1) for n_ in nets for p in n_.parameters()])
err = float((got - ref).abs().max() / ref.abs().max())
t = torch.tensor([err], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"MULTI data-parallel step over {world} GPUs ({N} rays): gradient vs single-process step, max error / max |g| = {float(t.item()):.2e}")
dist.destroy_process_group()
