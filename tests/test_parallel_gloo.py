"""World-size-2 (and 3) gloo tests of the ray-sharding host logic on CPU."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from intrinsicnerf_b200 import parallel


def test_ray_shard_partitions_exactly():
    for n in (0, 1, 127, 128, 129, 640000, 76800, 1024, 2048):
        for world in (1, 2, 3, 4, 8):
            spans = [parallel.ray_shard(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a0, b0), (a1, b1) in zip(spans, spans[1:]):
                assert b0 == a1 and a0 <= b0
            assert all(a % parallel.TILE == 0 for a, _ in spans if a < n)
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= parallel.TILE or n < parallel.TILE * world
    assert parallel.image_shard(100, 3, 8) == list(range(3, 100, 8))
    assert sorted(sum((parallel.image_shard(100, r, 8) for r in range(8)), [])) == list(range(100))


def _fake_render(rays):
    # deterministic per-ray function standing in for the renderer (rays are independent)
    return {"rgb_map": torch.sin(rays[:, :3]) + rays[:, 3:6], "acc_map": rays[:, 6] * 2 + rays[:, 7], "z_std": rays.sum(-1)}


def _worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)
    rays = torch.randn(n, 11, generator=g)
    full = parallel.render_rays_sharded(rays, _fake_render)
    ref = _fake_render(rays)
    ok = all(torch.equal(full[k], ref[k]) for k in ref)
    lin = torch.nn.Linear(4, 4)
    if rank != 0:
        with torch.no_grad():
            lin.weight.zero_()
    parallel.broadcast_weights(lin, 0)
    w = [torch.zeros_like(lin.weight) for _ in range(world)]
    dist.all_gather(w, lin.weight.data)
    ok &= all(torch.equal(w[0], x) for x in w) and bool(w[0].abs().sum() > 0)
    # chunk-pipelined frame gather: every rank ends with the single-process frame
    rec_fn = lambda r: torch.cat([torch.sin(r[:, :3]) + r[:, 3:6], r[:, 6:8] * 2], 1)  # noqa: E731
    for chunk in (128, 256, 4096):
        ok &= torch.equal(parallel.render_frame_pipelined(rays, rec_fn, chunk), rec_fn(rays))
    # data-parallel training step: gathered maps + summed gradients == single-process gradients
    torch.manual_seed(1)
    net = torch.nn.Sequential(torch.nn.Linear(11, 16), torch.nn.ReLU(), torch.nn.Linear(16, 4))
    ref_net = torch.nn.Sequential(torch.nn.Linear(11, 16), torch.nn.ReLU(), torch.nn.Linear(16, 4))
    ref_net.load_state_dict(net.state_dict())

    def loss_fn(m):          # pairs ray i with ray i + N/2 like compute_intrinsic_loss
        h = m["rgb"].shape[0] // 2
        return ((m["rgb"][:h] - m["rgb"][h:2 * h]) ** 2).mean() + m["acc"].mean()
    full_out = ref_net(rays)
    loss_fn({"rgb": full_out[:, :3], "acc": full_out[:, 3]}).backward()
    a, b = parallel.ray_shard(n, rank, world)
    loc = net(rays[a:b])
    maps = parallel.gather_maps_for_loss({"rgb": loc[:, :3], "acc": loc[:, 3]}, n)
    loss_fn(maps).backward()
    parallel.allreduce_gradients([net])
    for p, q_ in zip(net.parameters(), ref_net.parameters()):
        ok &= bool(torch.allclose(p.grad, q_.grad, rtol=1e-5, atol=1e-6))
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


# (8, 1024) is BASELINE config 5 at 8 ranks: 128 rays = one tile per rank; (4, 200): ranks 2 and 3 get an empty / short shard
@pytest.mark.parametrize("world,n", [(2, 1000), (3, 129), (2, 5), (4, 200), (8, 1024)])
def test_sharded_equals_single_process(world, n):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res
