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


def render_frame_pipelined(rays, record_fn, chunk, group=None):
    """Full frame on every rank with the gather overlapped: rank r renders its ray_shard() slice chunk by chunk
    (`record_fn(rays_chunk) -> [n, R]` packed per-ray records) and the all-gather of chunk k (NCCL, its own stream,
    async_op) runs while chunk k+1 is computed - SURVEY 8e: the collective moves 52 B/ray and is latency-bound, so
    it is hidden behind the 340 MFLOP/ray of compute instead of being paid at the end of the frame.
    Returns [N, R] on every rank, bit-identical to the single-process frame (rays are independent)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n = rays.shape[0]
    spans = [ray_shard(n, r, world) for r in range(world)]
    a, b = spans[rank]
    if world == 1:
        parts = [record_fn(rays[i:min(i + chunk, n)]) for i in range(0, n, chunk)]
        return parts[0] if len(parts) == 1 else torch.cat(parts, 0)
    steps = max((hi - lo + chunk - 1) // chunk for lo, hi in spans)
    out, pending = None, None

    def drain(p):
        work, buf, k = p
        work.wait()
        for r, (lo, hi) in enumerate(spans):
            s0 = lo + k * chunk
            s1 = min(s0 + chunk, hi)
            if s1 > s0:
                out[s0:s1] = buf[r, : s1 - s0]
    for k in range(steps):
        s0 = a + k * chunk
        s1 = min(s0 + chunk, b)
        rec = record_fn(rays[s0:s1]) if s1 > s0 else None
        if out is None:
            # record width: known from the first non-empty local chunk; ranks with an empty shard learn it from rank 0
            width = torch.tensor([rec.shape[1] if rec is not None else 0], device=rays.device)
            dist.all_reduce(width, op=dist.ReduceOp.MAX, group=group)
            R = int(width.item())
            out = rays.new_empty(n, R)
        send = rays.new_zeros(chunk, R)
        if rec is not None:
            send[: rec.shape[0]] = rec
        buf = rays.new_empty(world, chunk, R)
        work = dist.all_gather_into_tensor(buf.view(world * chunk, R), send, group=group, async_op=True)
        if pending is not None:
            drain(pending)               # chunk k-1 landed while chunk k was being rendered
        pending = (work, buf, k)
    if pending is not None:
        drain(pending)
    return out


class _GatherRows(torch.autograd.Function):
    """all-gather of ray_shard() row slices with a gradient: backward hands each rank the rows it contributed."""

    @staticmethod
    def forward(ctx, local, n_total, group):
        ctx.span = ray_shard(n_total, dist.get_rank(group), dist.get_world_size(group))
        return gather_records(local.contiguous(), n_total, group)

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.span
        return g[a:b].contiguous(), None, None


def gather_maps_for_loss(local_maps, n_total, group=None):
    """Training (BASELINE config 5): every rank renders ray_shard(n_total) of the step's rays; the losses pair ray i
    with ray i + N/2 and with rays of another quarter (compute_intrinsic_loss, training_utils.py:201-205), so the
    rendered maps (N x ~(14+C) floats, ~170 KB) are all-gathered and every rank evaluates the same full-batch loss.
    The backward pass returns each rank the gradient rows of its own rays; allreduce_gradients(SUM) then yields
    exactly the single-process gradient."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return dict(local_maps)
    # ONE collective for all maps: their columns are concatenated into a [n_local, sum(widths)] record, gathered, and
    # split again (ten separate all-gathers cost ten NCCL latencies per step; the gradient needs no collective at all)
    keys = sorted(local_maps)
    shapes, widths, flats = [], [], []
    for k in keys:
        v = local_maps[k]
        width = 1
        for d in v.shape[1:]:
            width *= int(d)
        shapes.append(tuple(v.shape[1:]))
        widths.append(width)
        flats.append(v.reshape(v.shape[0], width))   # explicit width: an empty shard has 0 rows
    full = _GatherRows.apply(torch.cat(flats, 1), n_total, group)
    out, off = {}, 0
    for k, shp, w in zip(keys, shapes, widths):
        out[k] = full[:, off:off + w].reshape((n_total,) + shp)
        off += w
    return out


def allreduce_gradients(modules, group=None, average=False):
    """All parameter gradients of the coarse + fine networks (1.32 M fp32 = 5.3 MB) in one all-reduce per
    network - in place on the flat gradient buffers when the .grad tensors tile one - instead of one per parameter tensor."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    params = [p for m in modules for p in m.parameters() if p.requires_grad]
    for p in params:
        if p.grad is None:
            p.grad = torch.zeros_like(p)
    world = dist.get_world_size(group)
    # Our modules' backward returns ONE flat gradient per network and autograd hands the parameters views of it, so the
    # .grad tensors of a network tile one storage: reduce those ranges in place (no concatenation, no copies back).
    by_store = {}
    for p in params:
        by_store.setdefault(p.grad.untyped_storage().data_ptr(), []).append(p)
    ranges, loose = [], []
    for ps in by_store.values():
        ps = sorted(ps, key=lambda q: q.grad.storage_offset())
        tiled = len(ps) > 1 and all(q.grad.is_contiguous() for q in ps) and all(
            a.grad.storage_offset() + a.grad.numel() == b.grad.storage_offset() for a, b in zip(ps, ps[1:]))
        if tiled:
            g0 = ps[0].grad
            total = sum(q.grad.numel() for q in ps)
            ranges.append(g0.new_empty(0).set_(g0.untyped_storage(), g0.storage_offset(), (total,)))
        else:
            loose += ps
    works = [dist.all_reduce(r, op=dist.ReduceOp.SUM, group=group, async_op=True) for r in ranges]
    if loose:
        flat = torch.cat([p.grad.reshape(-1) for p in loose])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        if average:
            flat /= world
        off = 0
        for p in loose:
            k = p.numel()
            p.grad.copy_(flat[off:off + k].view_as(p.grad))
            off += k
    for w, r in zip(works, ranges):
        w.wait()
        if average:
            r /= world
