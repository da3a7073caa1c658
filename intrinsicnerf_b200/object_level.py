"""Drop-in for the renderer half of ``object_level/run_nerf.py`` and the model/sampling half
of ``object_level/run_nerf_helpers.py``: same names, same signatures, same return layout.

  render           run_nerf.py:74-139       batchify_rays   run_nerf.py:59-71
  render_rays      run_nerf.py:415-528      run_network     run_nerf.py:42-56
  raw2outputs      run_nerf.py:359-412      batchify        run_nerf.py:32-39
  sample_pdf       run_nerf_helpers.py:402  create_nerf     run_nerf.py:275-356
  get_rays/ndc_rays  run_nerf_helpers.py:359-398 (ray generation stays PyTorch: SURVEY 8f row 1)

When ``network_fn``/``network_fine`` are :class:`intrinsicnerf_b200.nerf.NeRF` modules the
whole chunk goes through one ``inrf_render_fwd`` call (PE + MLP + compositing + resampling
on the GPU, nothing materialised per layer).  With foreign callables the stage kernels
(sampling, compositing, merge) still run, around the caller's network.
"""
import os

import numpy as np
import torch

from . import ops
from .nerf import NeRF, Embedder, get_embedder  # noqa: F401  (re-exported like `from run_nerf_helpers import *`)

DEBUG = False

# keys of the 13(+C)-wide per-ray record (include/inrf.h)
_REC = dict(rgb=(0, 3), disp=(3, 4), acc=(4, 5), albedo=(5, 8), shading=(8, 9), residual=(9, 12), depth=(12, 13))


def _split_rec(rec, key):
    a, b = _REC[key]
    v = rec[:, a:b]
    return v if b - a > 1 else v[:, 0]


def batchify(fn, chunk):
    """Apply ``fn`` in slices of ``chunk`` rows (kept for API parity; the fused kernels need no
    netchunk, so ``chunk=None`` is the efficient setting)."""
    if chunk is None:
        return fn

    def run(inputs):
        return torch.cat([fn(inputs[i:i + chunk]) for i in range(0, inputs.shape[0], chunk)], 0)
    return run


class _FusedQuery:
    """The ``network_query_fn`` create_nerf hands out: callable like the reference's lambda,
    and recognisable by render_rays so that it can fuse the whole chunk."""

    def __init__(self, embed_fn, embeddirs_fn, netchunk):
        self.embed_fn, self.embeddirs_fn, self.netchunk = embed_fn, embeddirs_fn, netchunk

    def __call__(self, inputs, viewdirs, network_fn):
        return run_network(inputs, viewdirs, network_fn, self.embed_fn, self.embeddirs_fn, self.netchunk)

    def fusable_with(self, *nets):
        ok_emb = isinstance(self.embed_fn, Embedder) and isinstance(self.embeddirs_fn, Embedder) \
            and self.embed_fn.n_freqs == 10 and self.embeddirs_fn.n_freqs == 4 and self.embeddirs_fn.scalar_factor == 1.0
        return ok_emb and all(n is None or isinstance(n, NeRF) for n in nets)


def run_network(inputs, viewdirs, fn, embed_fn, embeddirs_fn, netchunk=1024 * 64):
    """inputs [N,S,3], viewdirs [N,3] -> [N,S,out].  With our modules the embedding and the
    network are one kernel launch; otherwise embed -> fn like the reference."""
    if isinstance(fn, NeRF) and isinstance(embed_fn, Embedder) and isinstance(embeddirs_fn, Embedder) \
            and viewdirs is not None and embed_fn.n_freqs == 10 and embeddirs_fn.n_freqs == 4:
        dirs = viewdirs[:, None].expand(inputs.shape)
        out = fn.evaluate("pts", inputs.reshape(-1, 3), dirs.reshape(-1, 3), False, embed_fn.scalar_factor)
        return out.reshape(list(inputs.shape[:-1]) + [out.shape[-1]])
    flat = torch.reshape(inputs, [-1, inputs.shape[-1]])
    embedded = embed_fn(flat)
    if viewdirs is not None:
        dirs = viewdirs[:, None].expand(inputs.shape)
        embedded = torch.cat([embedded, embeddirs_fn(torch.reshape(dirs, [-1, dirs.shape[-1]]))], -1)
    out = batchify(fn, netchunk)(embedded)
    return torch.reshape(out, list(inputs.shape[:-1]) + [out.shape[-1]])


def raw2outputs(raw, z_vals, rays_d, raw_noise_std=0, white_bkgd=False, pytest=False):
    """-> (rgb_map, disp_map, acc_map, weights, depth_map, albedo_map, shading_map, residual_map)."""
    noise = None
    if raw_noise_std > 0.:
        if pytest:
            np.random.seed(0)
            noise = torch.Tensor(np.random.rand(*list(raw[..., 3].shape))).to(raw.device) * raw_noise_std
        else:
            noise = torch.randn(raw[..., 3].shape, device=raw.device) * raw_noise_std
    rec, w = ops.composite(raw[..., :11], z_vals, rays_d, noise, white_bkgd)
    g = lambda k: _split_rec(rec, k)  # noqa: E731
    return g("rgb"), g("disp"), g("acc"), w, g("depth"), g("albedo"), g("shading"), g("residual")


def sample_pdf(bins, weights, N_samples, det=False, pytest=False):
    u = None
    if pytest:
        np.random.seed(0)
        shape = list(bins.shape[:-1]) + [N_samples]
        u = np.broadcast_to(np.linspace(0., 1., N_samples), shape) if det else np.random.rand(*shape)
        u = torch.Tensor(np.ascontiguousarray(u)).to(bins.device)
    elif not det:
        u = torch.rand(list(bins.shape[:-1]) + [N_samples], device=bins.device)
    return ops.sample_pdf(bins, weights, N_samples, u)[0]


def render_rays(ray_batch, network_fn, network_query_fn, N_samples, retraw=False, lindisp=False, perturb=0.,
                N_importance=0, network_fine=None, white_bkgd=False, raw_noise_std=0., verbose=False, pytest=False):
    N = ray_batch.shape[0]
    dev = ray_batch.device
    if ray_batch.shape[-1] != 11:
        raise NotImplementedError("the CUDA renderer needs use_viewdirs=True ray records [N,11]")
    St = N_samples + N_importance

    def draw_uniform(*shape):
        if pytest:
            np.random.seed(0)
            return torch.Tensor(np.random.rand(*shape)).to(dev)
        return torch.rand(*shape, device=dev)

    def draw_noise(*shape):
        if raw_noise_std <= 0.:
            return None
        if pytest:
            return draw_uniform(*shape) * raw_noise_std
        return torch.randn(*shape, device=dev) * raw_noise_std

    fused = isinstance(network_query_fn, _FusedQuery) and network_query_fn.fusable_with(network_fn, network_fine)
    if fused and (network_fn.needs_grad() or (network_fine is not None and network_fine.needs_grad())):
        fused = False          # training: stage kernels + differentiable MLP / compositing (MlpFn, CompositeFn)
    # stage path without the pytest hooks: the reference's three random tensors per pass (run_nerf.py:478, 387,
    # run_nerf_helpers.py:414) are generated inside the stage kernels from one seed per call (inrf_*_rng)
    in_kernel_rng = (not fused) and (not pytest)
    seed = ops.next_seed() if in_kernel_rng and (perturb > 0. or raw_noise_std > 0.) else None
    t_rand = draw_uniform(N, N_samples) if (perturb > 0. and not in_kernel_rng) else None
    u = draw_uniform(N, N_importance) if (N_importance > 0 and perturb != 0. and not in_kernel_rng) else None
    ret = {}
    if fused:
        fine = network_fine if network_fine is not None else network_fn
        o = ops.render_chunk(ray_batch, network_fn.packed(), fine.packed() if N_importance > 0 else None,
                             variant=network_fn.variant, n_samples=N_samples, n_importance=N_importance,
                             lindisp=lindisp, white_bkgd=white_bkgd,
                             pe_scalar_factor=network_query_fn.embed_fn.scalar_factor, t_rand=t_rand, u=u,
                             noise_coarse=draw_noise(N, N_samples),
                             noise_fine=draw_noise(N, St) if N_importance > 0 else None, want_raw=retraw)
        rec = o["rec_fine"] if N_importance > 0 else o["rec_coarse"]
        for k in ("rgb", "disp", "acc", "albedo", "shading", "residual"):
            ret[k + "_map"] = _split_rec(rec, k)
        if retraw:
            ret["raw"] = o["raw_fine"] if N_importance > 0 else o["raw_coarse"]
        if N_importance > 0:
            for k in ("rgb", "disp", "acc", "albedo", "shading", "residual"):
                ret[k + "0"] = _split_rec(o["rec_coarse"], k)
            ret["z_std"] = o["z_std"]
    else:
        rays_d, viewdirs = ray_batch[:, 3:6], ray_batch[:, -3:]
        rays_o = ray_batch[:, 0:3]
        z_vals = ops.coarse_z(ray_batch, N_samples, lindisp, t_rand, seed if (in_kernel_rng and perturb > 0.) else None)
        # our modules take (rays, z) and form o + d z in the kernel's front end (same two roundings as the line below):
        # no [N,S,3] tensors, four launches fewer per pass
        by_rays = isinstance(network_query_fn, _FusedQuery) and network_query_fn.fusable_with(network_fn, network_fine) \
            and ray_batch.shape[1] == 11

        def query(z, net):
            if by_rays:
                out = net.evaluate("rays", ray_batch, z, False, network_query_fn.embed_fn.scalar_factor)
                return out.reshape(z.shape[0], z.shape[1], out.shape[-1])
            pts = rays_o[..., None, :] + rays_d[..., None, :] * z[..., :, None]
            return network_query_fn(pts, viewdirs, net)
        raw = query(z_vals, network_fn)
        rng_c = (raw_noise_std, seed, False) if (in_kernel_rng and raw_noise_std > 0.) else None
        rng_f = (raw_noise_std, seed, True) if (in_kernel_rng and raw_noise_std > 0.) else None
        rec, weights = ops.composite(raw[..., :11], z_vals, rays_d, None if in_kernel_rng else draw_noise(N, N_samples), white_bkgd, rng=rng_c)
        rec0 = rec
        if N_importance > 0:
            z_mid = .5 * (z_vals[..., 1:] + z_vals[..., :-1])
            z_samples = ops.sample_pdf(z_mid, weights[..., 1:-1].detach(), N_importance, u,
                                       seed=seed if (in_kernel_rng and perturb != 0.) else None)[0]   # detached (run_nerf.py:501)
            z_vals, z_std = ops.merge_sorted(z_vals, z_samples)
            raw = query(z_vals, network_fn if network_fine is None else network_fine)
            rec, weights = ops.composite(raw[..., :11], z_vals, rays_d, None if in_kernel_rng else draw_noise(N, St), white_bkgd, rng=rng_f)
        maps = ops.split_rec(rec)                       # under autograd: one backward kernel for all maps (ops.SplitRecFn)
        for k in ("rgb", "disp", "acc", "albedo", "shading", "residual"):
            ret[k + "_map"] = maps[k]
        if retraw:
            ret["raw"] = raw
        if N_importance > 0:
            maps0 = ops.split_rec(rec0)
            for k in ("rgb", "disp", "acc", "albedo", "shading", "residual"):
                ret[k + "0"] = maps0[k]
            ret["z_std"] = z_std
    if DEBUG:
        for k in ret:
            if torch.isnan(ret[k]).any() or torch.isinf(ret[k]).any():
                print(f"! [Numerical Error] {k} contains nan or inf.")
    return ret


def batchify_rays(rays_flat, chunk=1024 * 32, **kwargs):
    """Chunked render_rays.  ``chunk`` only bounds the scratch memory (raw tensors); results
    do not depend on it."""
    pieces = {}
    for i in range(0, rays_flat.shape[0], chunk):
        r = render_rays(rays_flat[i:i + chunk], **kwargs)
        for k, v in r.items():
            pieces.setdefault(k, []).append(v)
    return {k: (v[0] if len(v) == 1 else torch.cat(v, 0)) for k, v in pieces.items()}


def get_rays(H, W, K, c2w):
    c2w = torch.as_tensor(c2w, dtype=torch.float32)
    dev = c2w.device
    jj, ii = torch.meshgrid(torch.linspace(0, H - 1, H, device=dev), torch.linspace(0, W - 1, W, device=dev), indexing="ij")
    dirs = torch.stack([(ii - K[0][2]) / K[0][0], -(jj - K[1][2]) / K[1][1], -torch.ones_like(ii)], -1)
    rays_d = torch.sum(dirs[..., None, :] * c2w[:3, :3], -1)
    rays_o = c2w[:3, -1].expand(rays_d.shape)
    return rays_o, rays_d


img2mse = lambda x, y: torch.mean((x - y) ** 2)  # noqa: E731  (run_nerf_helpers.py:11)


def compute_intrinsic_loss(albedo, shading, residual, gt_rgb, disp, acc, obj_mask):
    """Mirror of run_nerf_helpers.py:59-86: (chroma, residual, reflect_sparsity, shading_smooth, far_reflect,
    intensity) - one fused forward launch and one fused backward launch instead of ~40 small kernels each way.
    disp / acc only feed compute_depth_weight, whose result the reference discards (it passes w_depth = 1)."""
    t = ops.intrinsic_losses(None, albedo, shading, residual, gt_rgb, obj_mask, None, "object")
    return t[1], t[2], t[3], t[4], t[5], t[6]


def rays_for_pixels(H, W, K, c2w, select_coords, near, far):
    """Packed [N, 11] ray records for select_coords[N, 2] = (row, column) - the gather of run_nerf.py:913-932
    (random pixels + their neighbours) fused with get_rays and render()'s packing."""
    sc = torch.as_tensor(select_coords)
    pix = (sc[:, 0].long() * W + sc[:, 1].long()).cuda()
    return ops.rays_from_pixels(pix, H, W, K[0][0], K[1][1], K[0][2], K[1][2], c2w, near, far, "opengl", "z")


def get_rays_np(H, W, K, c2w):
    i, j = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32), indexing="xy")
    dirs = np.stack([(i - K[0][2]) / K[0][0], -(j - K[1][2]) / K[1][1], -np.ones_like(i)], -1)
    rays_d = np.sum(dirs[..., np.newaxis, :] * c2w[:3, :3], -1)
    rays_o = np.broadcast_to(c2w[:3, -1], np.shape(rays_d))
    return rays_o, rays_d


def ndc_rays(H, W, focal, near, rays_o, rays_d):
    """Forward-facing (LLFF) NDC warp, run_nerf_helpers.py:381-398."""
    t = -(near + rays_o[..., 2]) / rays_d[..., 2]
    rays_o = rays_o + t[..., None] * rays_d
    sx, sy = -1. / (W / (2. * focal)), -1. / (H / (2. * focal))
    o = torch.stack([sx * rays_o[..., 0] / rays_o[..., 2], sy * rays_o[..., 1] / rays_o[..., 2],
                     1. + 2. * near / rays_o[..., 2]], -1)
    d = torch.stack([sx * (rays_d[..., 0] / rays_d[..., 2] - rays_o[..., 0] / rays_o[..., 2]),
                     sy * (rays_d[..., 1] / rays_d[..., 2] - rays_o[..., 1] / rays_o[..., 2]),
                     -2. * near / rays_o[..., 2]], -1)
    return o, d


def _render_frame_from_camera(H, W, K, c2w, near, far, chunk, dev, network_fn=None, network_query_fn=None, N_samples=64, retraw=False,
                              lindisp=False, perturb=0., N_importance=0, network_fine=None, white_bkgd=False, raw_noise_std=0., **_):
    """render()'s full-image case with the rays generated inside the fused kernels (inrf_render_fwd_camera): the dict
    batchify_rays would return, or None when this call needs the general path (training, random draws, raw output,
    foreign networks, sample counts the fused kernel does not cover)."""
    q = network_query_fn
    if retraw or perturb or raw_noise_std or not (isinstance(q, _FusedQuery) and q.fusable_with(network_fn, network_fine)) \
            or not ops.frame_camera_ok(N_samples, N_importance) or network_fn.needs_grad() \
            or (network_fine is not None and network_fine.needs_grad()):
        return None
    fine = network_fine if network_fine is not None else network_fn
    pc, pf = network_fn.packed(), (fine.packed() if N_importance > 0 else None)
    parts = []
    for i in range(0, H * W, chunk):
        parts.append(ops.render_frame_camera(H, W, K, c2w, near, far, pc, pf, dev, pix0=i, n=min(chunk, H * W - i), variant=network_fn.variant,
                                             n_samples=N_samples, n_importance=N_importance, lindisp=lindisp, white_bkgd=white_bkgd,
                                             pe_scalar_factor=q.embed_fn.scalar_factor))
    cat = lambda k: parts[0][k] if len(parts) == 1 else torch.cat([p[k] for p in parts], 0)  # noqa: E731
    rec = cat("rec_fine" if N_importance > 0 else "rec_coarse")
    ret = {k + "_map": _split_rec(rec, k) for k in ("rgb", "disp", "acc", "albedo", "shading", "residual")}
    if N_importance > 0:
        rec0 = cat("rec_coarse")
        ret.update({k + "0": _split_rec(rec0, k) for k in ("rgb", "disp", "acc", "albedo", "shading", "residual")})
        ret["z_std"] = cat("z_std")
    return ret


def render(H, W, K, chunk=1024 * 32, rays=None, c2w=None, ndc=True, near=0., far=1., use_viewdirs=False,
           c2w_staticcam=None, **kwargs):
    """-> [rgb_map, disp_map, acc_map, albedo_map, shading_map, residual_map, extras_dict]."""
    if c2w is not None and use_viewdirs and not ndc and c2w_staticcam is None and not torch.is_tensor(near) \
            and not torch.is_tensor(far):
        # full-image fast path: rays are generated and packed on the device by one kernel
        net = kwargs.get("network_fn")
        dev = next(net.parameters()).device if isinstance(net, torch.nn.Module) else torch.device("cuda")
        all_ret = _render_frame_from_camera(H, W, K, c2w, near, far, chunk, dev, **kwargs)
        if all_ret is None:
            packed = ops.get_rays_packed(H, W, K, c2w, near, far, dev)
            all_ret = batchify_rays(packed, chunk, **kwargs)
        for k in all_ret:
            all_ret[k] = torch.reshape(all_ret[k], [H, W] + list(all_ret[k].shape[1:]))
        main = ["rgb_map", "disp_map", "acc_map", "albedo_map", "shading_map", "residual_map"]
        return [all_ret[k] for k in main] + [{k: v for k, v in all_ret.items() if k not in main}]
    if c2w is not None:
        rays_o, rays_d = get_rays(H, W, K, c2w)
    else:
        rays_o, rays_d = rays
    if not use_viewdirs:
        raise NotImplementedError("the CUDA renderer implements use_viewdirs=True (what every reference config sets)")
    viewdirs = rays_d
    if c2w_staticcam is not None:
        rays_o, rays_d = get_rays(H, W, K, c2w_staticcam)
    viewdirs = viewdirs / torch.norm(viewdirs, dim=-1, keepdim=True)
    viewdirs = torch.reshape(viewdirs, [-1, 3]).float()
    sh = rays_d.shape
    if ndc:
        rays_o, rays_d = ndc_rays(H, W, K[0][0], 1., rays_o, rays_d)
    rays_o = torch.reshape(rays_o, [-1, 3]).float()
    rays_d = torch.reshape(rays_d, [-1, 3]).float()
    near, far = near * torch.ones_like(rays_d[..., :1]), far * torch.ones_like(rays_d[..., :1])
    packed = torch.cat([rays_o, rays_d, near, far, viewdirs], -1)
    all_ret = batchify_rays(packed, chunk, **kwargs)
    for k in all_ret:
        all_ret[k] = torch.reshape(all_ret[k], list(sh[:-1]) + list(all_ret[k].shape[1:]))
    main = ["rgb_map", "disp_map", "acc_map", "albedo_map", "shading_map", "residual_map"]
    return [all_ret[k] for k in main] + [{k: v for k, v in all_ret.items() if k not in main}]


to8b = lambda x: (255 * np.clip(x, 0, 1)).astype(np.uint8)  # noqa: E731  (run_nerf_helpers.py:13; host twin of the kernel's)


def imwrite(filename, array):
    """PNG writer of render_path: imageio like the reference when it is installed, else OpenCV (RGB -> BGR)."""
    array = array.cpu().numpy() if torch.is_tensor(array) else np.asarray(array)
    try:
        import imageio
        if getattr(imageio, "__version__", None) is not None:        # a real install (test harnesses stub the module)
            imageio.imwrite(filename, array)
            return
    except ImportError:
        pass
    import cv2
    cv2.imwrite(filename, array[..., ::-1] if array.ndim == 3 and array.shape[-1] == 3 else array)


class _FrameWriter:
    """Host hand-over of finished frames, one frame behind the renderer: the planes of frame k are copied to pinned
    host memory on a side stream (ordered after frame k's kernels only) and written as PNGs while the GPU renders
    frame k+1 - the reference's loop instead blocks on `.cpu()` and encodes with the GPU idle (run_nerf.py:168-215)."""

    def __init__(self, device):
        self.stream = torch.cuda.Stream(device=device)
        self.pending = None
        self._pinned = {}          # (plane name, shape, dtype) -> two pinned host buffers used alternately

    def _host_buffer(self, key, t):
        """Pinned staging buffer for plane `key`, reused across frames.  Two per plane: one is being written to disk
        (previous frame) while the other receives this frame's copy.  device='cpu' is explicit because the reference's
        entry point sets a CUDA default tensor type (run_nerf.py:1129) and pinning needs a CPU tensor."""
        k = (key, tuple(t.shape), t.dtype)
        slot = self._pinned.setdefault(k, [None, None, 0])
        i = slot[2]
        if slot[i] is None:
            slot[i] = torch.empty(t.shape, dtype=t.dtype, device="cpu", pin_memory=True)
        slot[2] = 1 - i
        return slot[i]

    def submit(self, planes, names):
        """planes: {name: device uint8/uint16 tensor}; names: {name: file path}.  Returns immediately."""
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream())
        host = {}
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ready)
            for k, path in names.items():
                t = planes[k]
                t.record_stream(self.stream)
                h = self._host_buffer(k, t)
                h.copy_(t, non_blocking=True)
                host[path] = h
            done = torch.cuda.Event()
            done.record(self.stream)
        self.flush()                                   # write the PREVIOUS frame while this one's copies are in flight
        self.pending = (done, host)

    def flush(self):
        if self.pending is not None:
            done, host = self.pending
            done.synchronize()
            for path, h in host.items():
                imwrite(path, h.numpy())
            self.pending = None


def render_record(H, W, K, chunk, c2w, near=0., far=1., **kwargs):
    """One full frame as the packed per-ray record [H*W, 13] (include/inrf.h) - what render() slices its six
    maps out of, kept whole so that the frame kernels read every pixel once."""
    fn, fine, q = kwargs.get("network_fn"), kwargs.get("network_fine"), kwargs.get("network_query_fn")
    n_imp = kwargs.get("N_importance", 0)
    if isinstance(q, _FusedQuery) and q.fusable_with(fn, fine) and not kwargs.get("perturb", 0.) \
            and not kwargs.get("raw_noise_std", 0.) and kwargs.get("use_viewdirs", False) and not kwargs.get("ndc", True) \
            and kwargs.get("c2w_staticcam") is None:
        dev = next(fn.parameters()).device
        pc, pf = fn.packed(), ((fine if fine is not None else fn).packed() if n_imp > 0 else None)
        recs = []
        if ops.frame_camera_ok(kwargs["N_samples"], n_imp):
            # rays generated inside the kernels from (K, c2w, pixel index): no [H*W, 11] table is written or read
            for i in range(0, H * W, chunk):
                o = ops.render_frame_camera(H, W, K, c2w, near, far, pc, pf, dev, pix0=i, n=min(chunk, H * W - i), variant=fn.variant,
                                            n_samples=kwargs["N_samples"], n_importance=n_imp, lindisp=kwargs.get("lindisp", False),
                                            white_bkgd=kwargs.get("white_bkgd", False), pe_scalar_factor=q.embed_fn.scalar_factor)
                recs.append(o["rec_fine"] if n_imp > 0 else o["rec_coarse"])
            return recs[0] if len(recs) == 1 else torch.cat(recs, 0)
        rays = ops.get_rays_packed(H, W, K, c2w, near, far, dev)
        for i in range(0, rays.shape[0], chunk):
            o = ops.render_chunk(rays[i:i + chunk], pc, pf, variant=fn.variant, n_samples=kwargs["N_samples"], n_importance=n_imp,
                                 lindisp=kwargs.get("lindisp", False), white_bkgd=kwargs.get("white_bkgd", False),
                                 pe_scalar_factor=q.embed_fn.scalar_factor)
            recs.append(o["rec_fine"] if n_imp > 0 else o["rec_coarse"])
        return recs[0] if len(recs) == 1 else torch.cat(recs, 0)
    rgb, disp, acc, albedo, shading, residual, _ = render(H, W, K, chunk=chunk, c2w=c2w, near=near, far=far, **kwargs)
    cols = [rgb, disp[..., None], acc[..., None], albedo, shading[..., None], residual, torch.zeros_like(acc)[..., None]]
    return torch.cat([c.reshape(H * W, -1) for c in cols], -1).contiguous()


def render_path(render_poses, hwf, K, chunk, render_kwargs, gt_imgs=None, savedir=None, render_factor=0,
                update_cluster=False, b_f=0.5, sharded=False, group=None):
    """Mirror of run_nerf.py:142-272 -> (rgbs [n,H,W,3], disps [n,H,W], cluster_manager).  Per frame the
    reference copies six float maps to the host (48 B/pixel) and converts them with numpy; here the frame's record
    stays in HBM, inrf_frame_finish writes the 8-bit planes (11 B/pixel cross PCIe, plus the rgb/disp floats this
    function returns) and the albedo[::2, ::2] cluster samples, and the c###/edit### pass (dest_color +
    inrf_edit_recompose) runs on the resident records instead of re-uploading every albedo map.

    ``sharded=True`` under torch.distributed (BASELINE config 4: 100 views over 8 GPUs): rank r renders views
    r, r+world, ... (parallel.image_shard) and writes their PNGs under the global view index; the cluster samples are
    all-gathered so that every rank fits the same (deterministic) cluster manager, and rgbs / disps come back complete
    and in view order on every rank (one NCCL all-gather of 16 B/pixel)."""
    from . import parallel
    from .cluster import Cluster_Manager
    import torch.distributed as dist
    H, W, focal = hwf
    if render_factor != 0:
        H, W, focal = H // render_factor, W // render_factor, focal / render_factor
    H, W = int(H), int(W)
    kw = dict(render_kwargs)
    near, far = kw.pop("near", 0.), kw.pop("far", 1.)
    n_views = len(render_poses)
    world = dist.get_world_size(group) if (sharded and dist.is_initialized()) else 1
    rank = dist.get_rank(group) if world > 1 else 0
    mine = parallel.image_shard(n_views, rank, world)
    rgbd, recs, labels, sample_pixels, sample_labels = [], [], [], [], []
    planes = ("rgb8", "albedo8", "shading8", "residual8", "label8") + (("labels64",) if update_cluster else ())
    dev = None
    writer = None
    for i in mine:
        c2w = torch.as_tensor(render_poses[i])[:3, :4]
        with torch.no_grad():
            rec = render_record(H, W, K, chunk, c2w, near=near, far=far, **kw)
            f = ops.frame_finish(rec, H, W, 0, planes, acc_threshold=10.0, sub_step=2 if update_cluster else 0)
        dev = rec.device
        rgbd.append(rec[:, 0:4].reshape(H, W, 4))
        if i == mine[0]:
            print((H, W, 3), (H, W))
        if update_cluster:
            recs.append(rec)
            labels.append(f["labels64"])
            sample_pixels.append(f["sample_pixels"])
            sample_labels.append(f["sample_labels"])
        if savedir is not None:
            writer = writer or _FrameWriter(dev)
            writer.submit(f, {name: os.path.join(savedir, "{}{:03d}.png".format(prefix, i)) for prefix, name in
                              (("", "rgb8"), ("a", "albedo8"), ("s", "shading8"), ("res", "residual8"), ("acc", "label8"))})
    if writer is not None:
        writer.flush()
    if world > 1:
        dev = dev or torch.device("cuda", torch.cuda.current_device())
        per = (n_views + world - 1) // world                                     # views per rank, padded

        def gather_views(parts, tail, dtype):                                    # [n_mine, *tail] -> [n_views, *tail] in view order
            buf = torch.zeros((per,) + tail, dtype=dtype, device=dev)
            if parts:
                buf[: len(parts)] = torch.stack(parts, 0)
            out = torch.empty((world, per) + tail, dtype=dtype, device=dev)
            dist.all_gather_into_tensor(out.view((world * per,) + tail), buf, group=group)
            return out.transpose(0, 1).reshape((world * per,) + tail)[:n_views]   # view v sits at [v % world, v // world]
        all_rgbd = gather_views(rgbd, (H, W, 4), torch.float32)
        if update_cluster:
            ns = ((H + 1) // 2) * ((W + 1) // 2)
            sample_pixels = list(gather_views(sample_pixels, (ns, 3), torch.float32))
            sample_labels = list(gather_views(sample_labels, (ns, 1), torch.int64))
    else:
        all_rgbd = torch.stack(rgbd, 0) if rgbd else torch.zeros(0, H, W, 4)
    cluster_manager = None
    if update_cluster:
        sample_pixels, sample_labels = torch.cat(sample_pixels, 0), torch.cat(sample_labels, 0)
        cluster_manager = Cluster_Manager(class_num=1, device=sample_pixels.device)
        print(sample_pixels.shape, sample_labels.shape)
        cluster_manager.update_center(sample_labels, sample_pixels, band_factor=b_f)
        print("cluster albedo...")
        for i, rec, lab in zip(mine, recs, labels):
            result = cluster_manager.dest_color(rec[:, 5:8].contiguous(), lab.reshape(-1, 1))
            c8, e8 = ops.edit_recompose(result, rec)
            if savedir is not None:
                writer = writer or _FrameWriter(rec.device)
                writer.submit({"c8": c8.reshape(H, W, 3), "edit8": e8.reshape(H, W, 3)},
                              {"c8": os.path.join(savedir, "c{:03d}.png".format(i)), "edit8": os.path.join(savedir, "edit{:03d}.png".format(i))})
        if writer is not None:
            writer.flush()
    host = all_rgbd.cpu().numpy()
    return np.ascontiguousarray(host[..., 0:3]), np.ascontiguousarray(host[..., 3]), cluster_manager


def create_nerf(args, device=None):
    """-> (render_kwargs_train, render_kwargs_test, start, grad_vars, optimizer), with the
    reference's kwargs keys and checkpoint format ('network_fn_state_dict', ...)."""
    device = device or torch.device("cuda")
    embed_fn, input_ch = get_embedder(args.multires, args.i_embed)
    if not args.use_viewdirs:
        raise NotImplementedError("use_viewdirs=False is not implemented by the CUDA path")
    embeddirs_fn, input_ch_views = get_embedder(args.multires_views, args.i_embed)
    output_ch = 5 if args.N_importance > 0 else 4
    model = NeRF(D=args.netdepth, W=args.netwidth, input_ch=input_ch, output_ch=output_ch, skips=[4],
                 input_ch_views=input_ch_views, use_viewdirs=True).to(device)
    grad_vars = list(model.parameters())
    model_fine = None
    if args.N_importance > 0:
        model_fine = NeRF(D=args.netdepth_fine, W=args.netwidth_fine, input_ch=input_ch, output_ch=output_ch,
                          skips=[4], input_ch_views=input_ch_views, use_viewdirs=True).to(device)
        grad_vars += list(model_fine.parameters())
    query = _FusedQuery(embed_fn, embeddirs_fn, args.netchunk)
    optimizer = torch.optim.Adam(params=grad_vars, lr=args.lrate, betas=(0.9, 0.999))
    start = 0
    if getattr(args, "ft_path", None) not in (None, "None"):
        ckpts = [args.ft_path]
    else:
        d = os.path.join(args.basedir, args.expname)
        ckpts = [os.path.join(d, f) for f in sorted(os.listdir(d)) if "tar" in f] if os.path.isdir(d) else []
    print("Found ckpts", ckpts)
    if ckpts and not args.no_reload:
        print("Reloading from", ckpts[-1])
        ck = torch.load(ckpts[-1], map_location=device)
        start = ck["global_step"]
        optimizer.load_state_dict(ck["optimizer_state_dict"])
        model.load_state_dict(ck["network_fn_state_dict"])
        if model_fine is not None:
            model_fine.load_state_dict(ck["network_fine_state_dict"])
    train = dict(network_query_fn=query, perturb=args.perturb, N_importance=args.N_importance, network_fine=model_fine,
                 N_samples=args.N_samples, network_fn=model, use_viewdirs=args.use_viewdirs,
                 white_bkgd=args.white_bkgd, raw_noise_std=args.raw_noise_std)
    if args.dataset_type != "llff" or args.no_ndc:
        print("Not ndc!")
        train["ndc"] = False
        train["lindisp"] = args.lindisp
    test = dict(train)
    test["perturb"] = False
    test["raw_noise_std"] = 0.
    return train, test, start, grad_vars, optimizer
