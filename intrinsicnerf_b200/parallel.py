"""Ray-sharded data parallelism (one process per GPU, torch.distributed).

Rays are independent (every reduction in render_rays is along the sample axis), so the path
shards with NO data-path collective: rank r renders a contiguous, tile-aligned slice of the flat
ray list (or whole images when there are at least as many views as ranks).  The only collective
is the gather of finished per-ray records (13+C floats per ray) when one rank needs the whole
frame, and a one-time weight broadcast.  Per-ray results do not depend on the world size (same
kernel, same per-tile order), so sharded == single-GPU bit for bit.
"""
import torch
import torch.distributed as dist

TILE = 128   # rows per CTA tile of the MLP kernels; shard boundaries are multiples of it


def ray_shard(n_rays, rank, world, align=TILE):
    """Contiguous [start, end) slice of rank `rank`; boundaries are multiples of `align`."""
    tiles = (n_rays + align - 1) // align
    per, extra = divmod(tiles, world)
    t0 = rank * per + min(rank, extra)
    t1 = t0 + per + (1 if rank < extra else 0)
    return min(t0 * align, n_rays), min(t1 * align, n_rays)


def image_shard(n_views, rank, world):
    """Views rendered by `rank` when sharding by image (config 4: 100 views over 8 GPUs)."""
    return list(range(rank, n_views, world))


def gather_records(local, n_total, group=None):
    """All-gather per-ray records [n_local, R] of ray_shard() slices into [n_total, R] on every rank."""
    world = dist.get_world_size(group)
    if world == 1:
        return local
    sizes = [ray_shard(n_total, r, world)[1] - ray_shard(n_total, r, world)[0] for r in range(world)]
    pad = max(sizes)
    buf = local.new_zeros((pad,) + tuple(local.shape[1:]))
    buf[: local.shape[0]] = local
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    return torch.cat([o[:s] for o, s in zip(out, sizes)], 0)


def broadcast_weights(module, src=0, group=None):
    """One-time replication of the (5.3 MB) network weights."""
    for p in module.parameters():
        dist.broadcast(p.data, src, group=group)


def render_rays_sharded(rays, render_fn, group=None, gather=True):
    """rays [N,11] (identical on every rank) -> dict of [N,...] tensors (or the local slice when
    gather=False).  `render_fn(rays_slice) -> dict` is e.g. functools.partial(batchify_rays, ...)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    a, b = ray_shard(rays.shape[0], rank, world)
    local = render_fn(rays[a:b])
    if not gather or world == 1:
        return local
    out = {}
    for k in sorted(local):
        v = local[k]
        width = 1
        for d in v.shape[1:]:
            width *= int(d)
        flat = v.reshape(v.shape[0], width)          # explicit width: an empty shard has 0 rows
        full = gather_records(flat.contiguous(), rays.shape[0], group)
        out[k] = full.reshape((rays.shape[0],) + tuple(v.shape[1:]))
    return out
