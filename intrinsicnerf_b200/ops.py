"""Tensor-level wrappers over the C ABI.  Inputs must be CUDA float32 tensors; everything
is enqueued on torch's current stream.  No CPU path exists."""
import ctypes as C

import torch

from . import _lib
from ._lib import PREC_FP32, PREC_TC, RAW_BASE, REC_BASE, RenderCfg, check  # noqa: F401

_DEFAULT_PRECISION = PREC_TC


def set_default_precision(p):
    """'tc' (tcgen05 fp16-operand / fp32-accumulate, default) or 'fp32' (CUDA-core FFMA)."""
    global _DEFAULT_PRECISION
    _DEFAULT_PRECISION = {"tc": PREC_TC, "fp32": PREC_FP32, PREC_TC: PREC_TC, PREC_FP32: PREC_FP32}[p]


def default_precision():
    return _DEFAULT_PRECISION


def poll_status():
    """Report (raise) what kernels enqueued so far have recorded: a tripped barrier watchdog (InrfError) or a value
    outside the fp16 range of the tensor-core path (InrfRangeError).  The hot entry points poll on entry without
    synchronising; call this after a stream/device synchronise to check everything launched before it."""
    check(_lib.lib().inrf_poll_status())


_SEED_STATE = [None, 0]


def next_seed(device=None):
    """Seed for one step's in-kernel random tensors (inrf_*_rng).  Derived on the host from torch's CUDA generator -
    its seed and Philox offset, which torch.manual_seed resets - and the offset is advanced, exactly as a torch.rand
    call would consume it: runs are reproducible from torch.manual_seed, and no generator kernel is launched.
    (While a CUDA graph is being captured the generator cannot be read; a host-side counter is used instead.  The seed
    is then a constant of the graph: call rng_epoch_bump() at the start of the captured step and every replay draws
    fresh numbers.)"""
    try:
        idx = torch.cuda.current_device() if device is None else torch.device(device).index or 0
        gen = torch.cuda.default_generators[idx]
        off = int(gen.get_offset())
        gen.set_offset(off + 4)
        base, ctr = int(gen.initial_seed()), off // 4 + 1
    except Exception:
        base = torch.initial_seed()
        if _SEED_STATE[0] != base:
            _SEED_STATE[0], _SEED_STATE[1] = base, 0
        _SEED_STATE[1] += 1
        ctr = _SEED_STATE[1]
    return (base * 0x9E3779B97F4A7C15 + ctr * 0xD1B54A32D192ED03) & 0x7FFFFFFFFFFFFFFF


def rng_epoch_bump():
    """inrf_rng_epoch_bump on the current stream: increments the device-resident epoch that is mixed into the key of every
    in-kernel draw.  Capture it as the first call of a CUDA-graphed training step (the seeds are baked into the graph);
    eager code does not need it (next_seed already changes per step)."""
    check(_lib.lib().inrf_rng_epoch_bump(_stream()))


def rng_epoch_reset():
    """Epoch back to 0 (key = seed), e.g. to reproduce a run from its first step."""
    check(_lib.lib().inrf_rng_epoch_reset(_stream()))


def launch_count():
    """Kernels launched by libinrf.so in this process so far (inrf_launch_count)."""
    return int(_lib.lib().inrf_launch_count())


def _prec(p):
    if p is None:
        return _DEFAULT_PRECISION
    return {"tc": PREC_TC, "fp32": PREC_FP32, PREC_TC: PREC_TC, PREC_FP32: PREC_FP32}[p]


def _f32(t, name):
    if not torch.is_tensor(t):
        raise TypeError(f"{name}: expected a tensor")
    if not t.is_cuda:
        raise RuntimeError(f"{name}: intrinsicnerf_b200 runs on CUDA tensors only (got {t.device}); there is no CPU fallback")
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream():
    """Current CUDA stream of the current device as a void* (the raw-handle query: torch.cuda.current_stream() builds a
    Stream object and costs ~20 us of host time per call, 17 calls per training step)."""
    if _raw_stream is not None:
        return C.c_void_p(_raw_stream(torch.cuda.current_device()))
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class _dev:
    """`with _dev(d)` without the cost when d already is the current device (the usual case)."""
    __slots__ = ("idx", "prev")

    def __init__(self, device):
        self.idx = device.index if getattr(device, "index", None) is not None else -1

    def __enter__(self):
        self.prev = torch.cuda.current_device()
        if self.idx >= 0 and self.idx != self.prev:
            torch.cuda.set_device(self.idx)
        return self

    def __exit__(self, *exc):
        if self.idx >= 0 and self.idx != self.prev:
            torch.cuda.set_device(self.prev)
        return False


_LINSPACE = {}


def linspace01(n, device):
    """torch.linspace(0,1,n) on `device`, cached (t_vals / deterministic u)."""
    key = (n, str(device))
    t = _LINSPACE.get(key)
    if t is None:
        t = torch.linspace(0.0, 1.0, steps=n).to(device)   # CPU values, bit-identical to the reference's
        _LINSPACE[key] = t
    return t


def flat_param_count(variant, n_classes):
    return check(_lib.lib().inrf_flat_param_count(variant, n_classes))


def pack_weights(flat, variant, n_classes):
    flat = _f32(flat, "flat_params")
    n = flat_param_count(variant, n_classes)
    if flat.numel() != n:
        raise ValueError(f"flat parameter vector has {flat.numel()} values, expected {n}")
    nbytes = check(_lib.lib().inrf_packed_bytes(variant, n_classes))
    packed = torch.empty(nbytes, dtype=torch.uint8, device=flat.device)
    with _dev(flat.device):
        check(_lib.lib().inrf_pack_weights(_ptr(flat), variant, n_classes, _ptr(packed), nbytes, _stream()))
    return packed


def embed(x, n_freqs, scalar_factor=1.0):
    x = _f32(x, "x")
    lead = x.shape[:-1]
    x2 = x.reshape(-1, 3)
    out = torch.empty(x2.shape[0], 3 + 6 * n_freqs, dtype=torch.float32, device=x.device)
    with _dev(x.device):
        check(_lib.lib().inrf_embed(_ptr(x2), x2.shape[0], n_freqs, float(scalar_factor), _ptr(out), _stream()))
    return out.reshape(*lead, out.shape[-1])


def mlp_forward(packed, variant, n_classes, pts, viewdirs, endpoint=False, pe_scalar_factor=1.0, precision=None):
    pts, viewdirs = _f32(pts, "pts").reshape(-1, 3), _f32(viewdirs, "viewdirs").reshape(-1, 3)
    M = pts.shape[0]
    if viewdirs.shape[0] != M:
        raise ValueError("pts and viewdirs must have one row per sample")
    raw = torch.empty(M, RAW_BASE + n_classes + (128 if endpoint else 0), dtype=torch.float32, device=pts.device)
    with _dev(pts.device):
        check(_lib.lib().inrf_mlp_fwd(_ptr(packed), variant, n_classes, int(endpoint), float(pe_scalar_factor),
                                      _ptr(pts), _ptr(viewdirs), M, _ptr(raw), _prec(precision), _stream()))
    return raw


def mlp_forward_embedded(packed, variant, n_classes, emb, endpoint=False, precision=None):
    emb = _f32(emb, "embedded")
    if emb.shape[-1] != 90:
        raise ValueError("embedded input must be [...,90] = gamma(x)[63] | gamma(d)[27]")
    lead = emb.shape[:-1]
    e2 = emb.reshape(-1, 90)
    raw = torch.empty(e2.shape[0], RAW_BASE + n_classes + (128 if endpoint else 0), dtype=torch.float32, device=emb.device)
    with _dev(emb.device):
        check(_lib.lib().inrf_mlp_fwd_embedded(_ptr(packed), variant, n_classes, int(endpoint), _ptr(e2), e2.shape[0],
                                               _ptr(raw), _prec(precision), _stream()))
    return raw.reshape(*lead, raw.shape[-1])


def mlp_forward_rays(packed, variant, n_classes, rays, z, endpoint=False, pe_scalar_factor=1.0, precision=None):
    rays, z = _f32(rays, "rays"), _f32(z, "z")
    N, S = z.shape
    raw = torch.empty(N, S, RAW_BASE + n_classes + (128 if endpoint else 0), dtype=torch.float32, device=z.device)
    with _dev(z.device):
        check(_lib.lib().inrf_mlp_fwd_rays(_ptr(packed), variant, n_classes, int(endpoint), float(pe_scalar_factor),
                                           _ptr(rays), _ptr(z), N, S, _ptr(raw), _prec(precision), _stream()))
    return raw


def raw2outputs_rec(raw, z_vals, rays_d, noise=None, white_bkgd=False, n_classes=0, endpoint=False, want_weights=True, rng=None):
    """rng = (noise_std, seed, fine_pass): sigma noise generated inside the kernel (inrf_raw2outputs_rng) instead of `noise`."""
    raw, z_vals, rays_d = _f32(raw, "raw"), _f32(z_vals, "z_vals"), _f32(rays_d, "rays_d")
    N, S, ch = raw.shape
    if ch != RAW_BASE + n_classes + (128 if endpoint else 0):
        raise ValueError(f"raw has {ch} channels, expected {RAW_BASE + n_classes + (128 if endpoint else 0)}")
    rec = torch.empty(N, REC_BASE + n_classes + (128 if endpoint else 0), dtype=torch.float32, device=raw.device)
    weights = torch.empty(N, S, dtype=torch.float32, device=raw.device) if want_weights else None
    noise = None if noise is None else _f32(noise, "noise")
    with _dev(raw.device):
        if rng is not None and noise is None:
            check(_lib.lib().inrf_raw2outputs_rng(_ptr(raw), _ptr(z_vals), _ptr(rays_d), 3, float(rng[0]), int(rng[1]), int(bool(rng[2])), N, S,
                                                  n_classes, int(endpoint), int(bool(white_bkgd)), _ptr(rec), _ptr(weights), _stream()))
        else:
            check(_lib.lib().inrf_raw2outputs(_ptr(raw), _ptr(z_vals), _ptr(rays_d), 3, _ptr(noise), N, S, n_classes,
                                              int(endpoint), int(bool(white_bkgd)), _ptr(rec), _ptr(weights), _stream()))
    return rec, weights


def raw2outputs_bwd(raw, z_vals, rays_d, noise, grad_rec, grad_weights, white_bkgd=False, n_classes=0, endpoint=False, rng=None):
    raw, z_vals, rays_d, grad_rec = _f32(raw, "raw"), _f32(z_vals, "z_vals"), _f32(rays_d, "rays_d"), _f32(grad_rec, "grad_rec")
    N, S, ch = raw.shape
    noise = None if noise is None else _f32(noise, "noise")
    grad_weights = None if grad_weights is None else _f32(grad_weights, "grad_weights")
    grad_raw = torch.empty_like(raw)
    with _dev(raw.device):
        if rng is not None and noise is None:
            check(_lib.lib().inrf_raw2outputs_bwd_rng(_ptr(raw), _ptr(z_vals), _ptr(rays_d), 3, float(rng[0]), int(rng[1]), int(bool(rng[2])), N, S,
                                                      n_classes, int(endpoint), int(bool(white_bkgd)), _ptr(grad_rec), _ptr(grad_weights),
                                                      _ptr(grad_raw), _stream()))
        else:
            check(_lib.lib().inrf_raw2outputs_bwd(_ptr(raw), _ptr(z_vals), _ptr(rays_d), 3, _ptr(noise), N, S, n_classes,
                                                  int(endpoint), int(bool(white_bkgd)), _ptr(grad_rec), _ptr(grad_weights),
                                                  _ptr(grad_raw), _stream()))
    return grad_raw


class MlpFn(torch.autograd.Function):
    """Differentiable field-network evaluation (training path, fp32 CUDA cores).

    forward : inrf_mlp_fwd_train (k_mlp_fp32 writing the activation stash)
    backward: inrf_mlp_bwd       (k_mlp_bwd_fp32) -> dL/d(flat parameters)
    The flat parameter vector is built with torch.cat from the module's parameters, so autograd
    scatters the flat gradient back to every nn.Parameter.  Inputs (points / rays / depths) get no
    gradient - the reference never differentiates them (z_samples is detached, run_nerf.py:501)."""

    @staticmethod
    def forward(ctx, flat, variant, n_classes, endpoint, pe_scalar_factor, mode, a, b):
        flat_c = _f32(flat.detach(), "flat_params")
        packed = pack_weights(flat_c, variant, n_classes)
        L = _lib.lib()
        ch = RAW_BASE + n_classes + (128 if endpoint else 0)
        pts = vd = rays = z = emb = None
        S = 1
        if mode == "pts":
            pts, vd = _f32(a, "pts").reshape(-1, 3), _f32(b, "viewdirs").reshape(-1, 3)
            M = pts.shape[0]
        elif mode == "rays":
            rays, z = _f32(a, "rays"), _f32(b, "z")
            S = z.shape[1]
            M = z.numel()
        else:
            emb = _f32(a, "embedded").reshape(-1, 90)
            M = emb.shape[0]
        dev = flat_c.device
        raw = torch.empty(M, ch, dtype=torch.float32, device=dev)
        stash = torch.empty(M, int(L.inrf_stash_floats_per_row()), dtype=torch.float32, device=dev)
        with _dev(dev):
            check(L.inrf_mlp_fwd_train(_ptr(packed), variant, n_classes, int(endpoint), float(pe_scalar_factor), _ptr(pts), _ptr(vd),
                                       _ptr(rays), _ptr(z), S, _ptr(emb), M, _ptr(raw), _ptr(stash), _stream()))
        ctx.save_for_backward(flat_c, raw, stash, *[t for t in (pts, vd, rays, z, emb) if t is not None])
        ctx.cfg = (variant, n_classes, bool(endpoint), float(pe_scalar_factor), mode, S, M)
        return raw

    @staticmethod
    def backward(ctx, g_raw):
        variant, n_classes, endpoint, pe, mode, S, M = ctx.cfg
        flat_c, raw, stash, *addr = ctx.saved_tensors
        pts = vd = rays = z = emb = None
        if mode == "pts":
            pts, vd = addr
        elif mode == "rays":
            rays, z = addr
        else:
            (emb,) = addr
        g_raw = _f32(g_raw, "grad_raw").reshape(M, -1)
        g_flat = torch.zeros_like(flat_c)
        with _dev(flat_c.device):
            check(_lib.lib().inrf_mlp_bwd(_ptr(flat_c), variant, n_classes, int(endpoint), pe, _ptr(pts), _ptr(vd), _ptr(rays),
                                          _ptr(z), S, _ptr(emb), M, _ptr(raw), _ptr(stash), _ptr(g_raw), _ptr(g_flat), _stream()))
        return g_flat, None, None, None, None, None, None, None


class MlpTcFn(torch.autograd.Function):
    """Differentiable field-network evaluation on the tensor cores (training path, default).

    forward : inrf_mlp_fwd_train_tc - the tcgen05 inference kernel, additionally writing every activation tile to
              the stash as 16 KB operand images (5.25 KB per sample)
    backward: inrf_mlp_bwd_tc       - tcgen05 dX chain + one dW launch over those images -> dL/d(flat parameters)
    fp16 operands / fp32 accumulation in both directions; `set_default_precision("fp32")` selects MlpFn instead."""

    @staticmethod
    def forward(ctx, flat, variant, n_classes, endpoint, pe_scalar_factor, mode, a, b):
        flat_c = _f32(flat.detach(), "flat_params")
        packed = pack_weights(flat_c, variant, n_classes)
        L = _lib.lib()
        ch = RAW_BASE + n_classes + (128 if endpoint else 0)
        pts = vd = rays = z = emb = None
        S = 1
        if mode == "pts":
            pts, vd = _f32(a, "pts").reshape(-1, 3), _f32(b, "viewdirs").reshape(-1, 3)
            M = pts.shape[0]
        elif mode == "rays":
            rays, z = _f32(a, "rays"), _f32(b, "z")
            S = z.shape[1]
            M = z.numel()
        else:
            emb = _f32(a, "embedded").reshape(-1, 90)
            M = emb.shape[0]
        dev = flat_c.device
        raw = torch.empty(M, ch, dtype=torch.float32, device=dev)
        stash = torch.empty(int(L.inrf_mlp_stash_img_bytes(M)), dtype=torch.uint8, device=dev)
        with _dev(dev):
            check(L.inrf_mlp_fwd_train_tc(_ptr(packed), variant, n_classes, int(endpoint), float(pe_scalar_factor), _ptr(pts), _ptr(vd),
                                          _ptr(rays), _ptr(z), S, _ptr(emb), M, _ptr(raw), _ptr(stash), _stream()))
        ctx.save_for_backward(flat_c, packed, raw, stash)
        ctx.cfg = (variant, n_classes, bool(endpoint), M)
        return raw

    @staticmethod
    def backward(ctx, g_raw):
        variant, n_classes, endpoint, M = ctx.cfg
        flat_c, packed, raw, stash = ctx.saved_tensors
        g_raw = _f32(g_raw, "grad_raw").reshape(M, -1)
        g_flat = torch.zeros_like(flat_c)
        L = _lib.lib()
        nbytes = int(L.inrf_mlp_bwd_tc_workspace_bytes(variant, n_classes, M))
        ws = _Workspace.get(flat_c.device, nbytes, "bwd")
        with _dev(flat_c.device):
            check(L.inrf_mlp_bwd_tc(_ptr(packed), _ptr(flat_c), variant, n_classes, int(endpoint), M, _ptr(raw), _ptr(stash),
                                    _ptr(g_raw), _ptr(ws), ws.numel(), _ptr(g_flat), _stream()))
        return g_flat, None, None, None, None, None, None, None


class CompositeFn(torch.autograd.Function):
    """raw2outputs with a CUDA backward for `raw` (z_vals, rays_d and the noise are constants in the
    reference: z_samples is detached, run_nerf.py:501).  Lets a foreign PyTorch network train
    through the compositing kernel."""

    @staticmethod
    def forward(ctx, raw, z_vals, rays_d, noise, white_bkgd, n_classes, endpoint, rng=None):
        ctx.set_materialize_grads(False)            # an unused `weights` output costs no zero-fill and no extra read in the backward
        rec, w = raw2outputs_rec(raw.detach(), z_vals, rays_d, noise, white_bkgd, n_classes, endpoint, True, rng)
        ctx.save_for_backward(raw.detach(), z_vals, rays_d, noise if noise is not None else torch.empty(0))
        ctx.cfg = (bool(white_bkgd), int(n_classes), bool(endpoint), noise is not None, rng)
        return rec, w

    @staticmethod
    def backward(ctx, g_rec, g_w):
        raw, z_vals, rays_d, noise = ctx.saved_tensors
        wb, C, ep, has_noise, rng = ctx.cfg
        g_rec = torch.zeros(raw.shape[0], REC_BASE + C + (128 if ep else 0), device=raw.device) if g_rec is None else g_rec
        g = raw2outputs_bwd(raw, z_vals, rays_d, noise if has_noise else None, g_rec.contiguous(),
                            None if g_w is None else g_w.contiguous(), wb, C, ep, rng)   # the same (seed, element) -> the same noise
        return g, None, None, None, None, None, None, None


def composite(raw, z_vals, rays_d, noise=None, white_bkgd=False, n_classes=0, endpoint=False, rng=None):
    """rec, weights = raw2outputs; differentiable w.r.t. raw when autograd is recording.  rng = (noise_std, seed,
    fine_pass) generates the sigma noise inside the kernels (forward and backward) instead of taking a `noise` tensor."""
    if rng is not None and not rng[0] > 0:
        rng = None
    if torch.is_grad_enabled() and raw.requires_grad:
        return CompositeFn.apply(raw, z_vals, rays_d, noise, white_bkgd, n_classes, endpoint, rng)
    return raw2outputs_rec(raw, z_vals, rays_d, noise, white_bkgd, n_classes, endpoint, True, rng)


REC_LAYOUT = (("rgb", 0, 3), ("disp", 3, 4), ("acc", 4, 5), ("albedo", 5, 8), ("shading", 8, 9), ("residual", 9, 12), ("depth", 12, 13))
_ZERO_BLOCKS = {}


def _zero_block(n, width, device):
    """Read-only zeros [n, width] (cached): the gradient of a record column block nobody differentiated."""
    key = (n, width, str(device))
    z = _ZERO_BLOCKS.get(key)
    if z is None:
        if len(_ZERO_BLOCKS) > 64:
            _ZERO_BLOCKS.clear()
        z = torch.zeros(n, width, dtype=torch.float32, device=device)
        _ZERO_BLOCKS[key] = z
    return z


def _rec_blocks(width, n_classes, endpoint):
    blocks = list(REC_LAYOUT)
    if n_classes > 0:
        blocks.append(("sem", REC_BASE, REC_BASE + n_classes))
    if endpoint:
        blocks.append(("feat", REC_BASE + n_classes, REC_BASE + n_classes + 128))
    if blocks[-1][2] != width:
        raise ValueError(f"record has {width} columns, expected {blocks[-1][2]}")
    return blocks


class SplitRecFn(torch.autograd.Function):
    """The packed per-ray record -> its maps, with ONE kernel in the backward.  Plain slicing under autograd costs a
    zero-filled [N, 13+C] tensor, a slice copy and an accumulation per map that receives a gradient - 40 small kernels per
    training step for the 2 x 8 maps of a coarse + fine render; here the incoming map gradients are concatenated in
    column order (cached zero blocks for the maps without a gradient)."""

    @staticmethod
    def forward(ctx, rec, n_classes, endpoint):
        ctx.set_materialize_grads(False)
        blocks = _rec_blocks(rec.shape[1], n_classes, endpoint)
        ctx.blocks, ctx.n, ctx.dev = blocks, rec.shape[0], rec.device
        r = rec.detach()
        return tuple(r[:, a:b] if b - a > 1 else r[:, a] for _, a, b in blocks)

    @staticmethod
    def backward(ctx, *grads):
        parts = [_zero_block(ctx.n, b - a, ctx.dev) if g is None else g.reshape(ctx.n, b - a) for (_, a, b), g in zip(ctx.blocks, grads)]
        return torch.cat(parts, 1), None, None


def split_rec(rec, n_classes=0, endpoint=False):
    """dict of the record's maps: rgb [N,3], disp [N], acc [N], albedo [N,3], shading [N], residual [N,3], depth [N],
    (sem [N,C]), (feat [N,128]) - views of `rec`; under autograd through SplitRecFn."""
    blocks = _rec_blocks(rec.shape[1], n_classes, endpoint)
    if torch.is_grad_enabled() and rec.requires_grad:
        outs = SplitRecFn.apply(rec, n_classes, endpoint)
    else:
        outs = tuple(rec[:, a:b] if b - a > 1 else rec[:, a] for _, a, b in blocks)
    return {name: o for (name, _, _), o in zip(blocks, outs)}


LOSS_TERMS = ("img", "chroma", "residual", "reflect_sparsity", "shading_smooth", "far_reflect", "intensity", "cluster")


class IntrinsicLossFn(torch.autograd.Function):
    """losses[8] = (img, chroma, residual, reflect_sparsity, shading_smooth, far_reflect, intensity, cluster) in one
    launch, and the gradient of any weighted sum of them in one more (inrf_intrinsic_loss_fwd/_bwd)."""

    @staticmethod
    def forward(ctx, rgb, albedo, shading, residual, gt_rgb, label, target, mode):
        albedo, residual, gt_rgb = _f32(albedo.detach(), "albedo"), _f32(residual.detach(), "residual"), _f32(gt_rgb, "gt_rgb")
        shading = _f32(shading.detach(), "shading").reshape(-1)
        label = _f32(label.detach().float() if torch.is_tensor(label) else label, "label").reshape(-1)
        rgb = None if rgb is None else _f32(rgb.detach(), "rgb")
        target = None if target is None else _f32(target.detach(), "target_albedo")
        N = albedo.shape[0]
        if albedo.shape != (N, 3) or residual.shape != (N, 3) or gt_rgb.shape != (N, 3) or shading.numel() != N or label.numel() != N:
            raise ValueError("intrinsic loss: albedo/residual/gt_rgb must be [N,3], shading/label [N]")
        losses = torch.empty(8, dtype=torch.float32, device=albedo.device)
        with _dev(albedo.device):
            check(_lib.lib().inrf_intrinsic_loss_fwd(_ptr(rgb), 3, _ptr(albedo), 3, _ptr(shading), 1, _ptr(residual), 3, _ptr(gt_rgb),
                                                     _ptr(label), _ptr(target), N, int(mode), _ptr(losses), _stream()))
        ctx.save_for_backward(*[t if t is not None else torch.empty(0) for t in (rgb, albedo, shading, residual, gt_rgb, label, target)])
        ctx.cfg = (rgb is not None, target is not None, int(mode))
        return losses

    @staticmethod
    def backward(ctx, g_losses):
        rgb, albedo, shading, residual, gt_rgb, label, target = ctx.saved_tensors
        has_rgb, has_target, mode = ctx.cfg
        N = albedo.shape[0]
        w = _f32(g_losses, "grad")
        g_alb, g_res, g_sh = torch.empty_like(albedo), torch.empty_like(residual), torch.empty_like(shading)
        g_rgb = torch.empty_like(rgb) if has_rgb else None
        with _dev(albedo.device):
            check(_lib.lib().inrf_intrinsic_loss_bwd(_ptr(rgb if has_rgb else None), 3, _ptr(albedo), 3, _ptr(shading), 1, _ptr(residual), 3,
                                                     _ptr(gt_rgb), _ptr(label), _ptr(target if has_target else None), N, mode, _ptr(w),
                                                     _ptr(g_rgb), _ptr(g_alb), _ptr(g_sh), _ptr(g_res), _stream()))
        return g_rgb, g_alb, g_sh, g_res, None, None, None, None


def intrinsic_losses(rgb, albedo, shading, residual, gt_rgb, label, target_albedo=None, mode="object"):
    """All loss terms of one training step as a differentiable tensor [8] (order: LOSS_TERMS).  mode "object":
    label is the object mask (run_nerf_helpers.py:25-37); mode "ssr": label is the semantic class
    (training_utils.py:141-152).  rgb / target_albedo may be None (their terms are 0)."""
    if mode not in ("object", "ssr"):
        raise ValueError("mode must be 'object' or 'ssr'")
    return IntrinsicLossFn.apply(rgb, albedo, shading.reshape(-1), residual, gt_rgb, label, target_albedo, 0 if mode == "object" else 1)


def sample_pdf(bins, weights, n_samples, u=None, want_inds=False, want_cdf=False, seed=None):
    """seed: draw u ~ U[0,1) inside the kernel (inrf_sample_pdf_rng) instead of taking a `u` tensor / the deterministic u."""
    bins, weights = _f32(bins, "bins"), _f32(weights, "weights")
    N, B = bins.shape
    if weights.shape != (N, B - 1):
        raise ValueError("weights must be [N, bins-1]")
    if seed is not None and u is None and not want_inds and not want_cdf:
        samples = torch.empty(N, n_samples, dtype=torch.float32, device=bins.device)
        with _dev(bins.device):
            check(_lib.lib().inrf_sample_pdf_rng(_ptr(bins), _ptr(weights), B - 1, int(seed), N, B, n_samples, _ptr(samples), _stream()))
        return samples, None, None
    u = None if u is None else _f32(u, "u")
    u_det = linspace01(n_samples, bins.device) if u is None else None
    samples = torch.empty(N, n_samples, dtype=torch.float32, device=bins.device)
    inds = torch.empty(N, n_samples, dtype=torch.int64, device=bins.device) if want_inds else None
    cdf = torch.empty(N, B, dtype=torch.float32, device=bins.device) if want_cdf else None
    with _dev(bins.device):
        check(_lib.lib().inrf_sample_pdf(_ptr(bins), _ptr(weights), B - 1, _ptr(u), _ptr(u_det), N, B, n_samples,
                                         _ptr(samples), _ptr(inds), _ptr(cdf), _stream()))
    return samples, inds, cdf


def invert_cdf(bins, cdf, u):
    bins, cdf, u = _f32(bins, "bins"), _f32(cdf, "cdf"), _f32(u, "u")
    N, B = bins.shape
    n = u.shape[1]
    samples = torch.empty(N, n, dtype=torch.float32, device=bins.device)
    inds = torch.empty(N, n, dtype=torch.int64, device=bins.device)
    with _dev(bins.device):
        check(_lib.lib().inrf_invert_cdf(_ptr(bins), _ptr(cdf), _ptr(u), N, B, n, _ptr(samples), _ptr(inds), _stream()))
    return samples, inds


def merge_sorted(z_a, z_b, want_std=True):
    z_a, z_b = _f32(z_a, "z_a"), _f32(z_b, "z_b")
    N, Sa = z_a.shape
    Sb = z_b.shape[1]
    out = torch.empty(N, Sa + Sb, dtype=torch.float32, device=z_a.device)
    std = torch.empty(N, dtype=torch.float32, device=z_a.device) if want_std else None
    with _dev(z_a.device):
        check(_lib.lib().inrf_merge_sorted(_ptr(z_a), _ptr(z_b), N, Sa, Sb, _ptr(out), _ptr(std), _stream()))
    return out, std


def coarse_z(rays, n_samples, lindisp=False, t_rand=None, seed=None):
    """seed: stratified jitter drawn inside the kernel (inrf_coarse_z_rng) instead of a `t_rand` tensor."""
    rays = _f32(rays, "rays")
    N = rays.shape[0]
    z = torch.empty(N, n_samples, dtype=torch.float32, device=rays.device)
    if seed is not None and t_rand is None:
        with _dev(rays.device):
            check(_lib.lib().inrf_coarse_z_rng(_ptr(rays), _ptr(linspace01(n_samples, rays.device)), int(seed), N, n_samples,
                                               int(bool(lindisp)), _ptr(z), _stream()))
        return z
    t_rand = None if t_rand is None else _f32(t_rand, "t_rand")
    with _dev(rays.device):
        check(_lib.lib().inrf_coarse_z(_ptr(rays), _ptr(linspace01(n_samples, rays.device)), _ptr(t_rand), N, n_samples,
                                       int(bool(lindisp)), _ptr(z), _stream()))
    return z


def get_rays_packed(H, W, K, c2w, near, far, device):
    """Full-image ray records [H*W, 11] (get_rays + render()'s packing, use_viewdirs=True, ndc=False)."""
    m = torch.as_tensor(c2w, dtype=torch.float32).detach().cpu()[:3, :4].contiguous()
    arr = (C.c_float * 12)(*[float(v) for v in m.reshape(-1).tolist()])
    rays = torch.empty(H * W, 11, dtype=torch.float32, device=device)
    with _dev(device):
        check(_lib.lib().inrf_get_rays(int(H), int(W), float(K[0][0]), float(K[1][1]), float(K[0][2]), float(K[1][2]), arr,
                                       float(near), float(far), _ptr(rays), _stream()))
    return rays


def rays_from_pixels(pix, H, W, fx, fy, cx, cy, c2w, near, far, convention="opengl", depth_type="z", device=None):
    """Ray records [N, 11] for flat pixel indices pix[N] = row*W + column (None: the whole image), either
    camera convention (get_rays, run_nerf_helpers.py:359-368 / create_rays, SSR/models/rays.py:223-256)."""
    if convention not in ("opengl", "opencv") or depth_type not in ("z", "euclidean"):
        raise ValueError("convention must be 'opengl' or 'opencv', depth_type 'z' or 'euclidean'")
    m = torch.as_tensor(c2w, dtype=torch.float32).detach().cpu()[:3, :4].contiguous()
    arr = (C.c_float * 12)(*[float(v) for v in m.reshape(-1).tolist()])
    if pix is None:
        if device is None:
            raise ValueError("device is required for full-image generation")
        n = H * W
    else:
        if not torch.is_tensor(pix) or not pix.is_cuda:
            raise RuntimeError("pix: intrinsicnerf_b200 runs on CUDA tensors only; there is no CPU fallback")
        pix = pix.detach().to(torch.int64).reshape(-1).contiguous()
        device, n = pix.device, pix.numel()
    rays = torch.empty(n, 11, dtype=torch.float32, device=device)
    with _dev(device):
        check(_lib.lib().inrf_rays_from_pixels(_ptr(pix), n, int(H), int(W), float(fx), float(fy), float(cx), float(cy), arr,
                                               1 if convention == "opencv" else 0, 1 if depth_type == "euclidean" else 0,
                                               float(near), float(far), _ptr(rays), _stream()))
    return rays


class _Workspace:
    """Grow-only scratch buffer per device (the library allocates nothing itself)."""
    bufs = {}

    @classmethod
    def get(cls, device, nbytes, tag=""):
        key = str(device) + tag
        b = cls.bufs.get(key)
        if b is None or b.numel() < nbytes:
            b = torch.empty(int(nbytes * 1.1) + 1024, dtype=torch.uint8, device=device)
            cls.bufs[key] = b
        return b


def render_chunk(rays, packed_coarse, packed_fine, variant=0, n_classes=0, n_samples=64, n_importance=128,
                 lindisp=False, white_bkgd=False, endpoint=False, pe_scalar_factor=1.0, precision=None,
                 t_rand=None, u=None, noise_coarse=None, noise_fine=None, want_raw=False, want_z=False,
                 want_weights=False):
    """One inrf_render_fwd call.  Returns a dict of packed outputs:
    rec_coarse [N,13+C], rec_fine [N,13+C(+128)], z_std [N] (+ raw_coarse/raw_fine/z_fine/weights_fine)."""
    rays = _f32(rays, "rays")
    if rays.ndim != 2 or rays.shape[1] != 11:
        raise ValueError("rays must be [N,11] = o3 d3 near far viewdir3 (use_viewdirs=True)")
    dev, N = rays.device, rays.shape[0]
    St = n_samples + n_importance
    cfg = RenderCfg(variant=variant, n_classes=n_classes, n_samples=n_samples, n_importance=n_importance,
                    lindisp=int(bool(lindisp)), white_bkgd=int(bool(white_bkgd)), endpoint_feat=int(bool(endpoint)),
                    precision=_prec(precision), pe_scalar_factor=float(pe_scalar_factor))
    L = _lib.lib()
    ws_bytes = check(L.inrf_render_workspace_bytes(C.byref(cfg), N))
    ws = _Workspace.get(dev, ws_bytes)
    new = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)  # noqa: E731
    out = {"rec_coarse": new(N, REC_BASE + n_classes)}
    if n_importance > 0:
        out["rec_fine"] = new(N, REC_BASE + n_classes + (128 if endpoint else 0))
        out["z_std"] = new(N)
    if want_raw:
        out["raw_coarse"] = new(N, n_samples, RAW_BASE + n_classes)
        if n_importance > 0:
            out["raw_fine"] = new(N, St, RAW_BASE + n_classes + (128 if endpoint else 0))
    if want_z and n_importance > 0:
        out["z_fine"] = new(N, St)
    if want_weights and n_importance > 0:
        out["weights_fine"] = new(N, St)
    opt = lambda t, name: None if t is None else _f32(t, name)  # noqa: E731
    t_rand, u = opt(t_rand, "t_rand"), opt(u, "u")
    noise_coarse, noise_fine = opt(noise_coarse, "noise_coarse"), opt(noise_fine, "noise_fine")
    u_det = linspace01(n_importance, dev) if (n_importance > 0 and u is None) else None
    with _dev(dev):
        check(L.inrf_render_fwd(_ptr(rays), N, _ptr(packed_coarse), _ptr(packed_fine), C.byref(cfg),
                                _ptr(linspace01(n_samples, dev)), _ptr(u_det), _ptr(t_rand), _ptr(u),
                                _ptr(noise_coarse), _ptr(noise_fine), _ptr(out["rec_coarse"]), _ptr(out.get("rec_fine")),
                                _ptr(out.get("z_std")), _ptr(out.get("raw_coarse")), _ptr(out.get("raw_fine")),
                                _ptr(out.get("z_fine")), _ptr(out.get("weights_fine")), _ptr(ws), ws.numel(), _stream()))
    return out


def frame_camera_ok(n_samples, n_importance, precision=None):
    """True when inrf_render_fwd_camera covers this configuration (see include/inrf.h)."""
    import os
    return _prec(precision) == PREC_TC and os.environ.get("INRF_NO_FUSE", "0")[:1] != "1" and n_samples % 32 == 0 \
        and (n_importance == 0 or (n_samples == 64 and n_importance == 128))


def render_frame_camera(H, W, K, c2w, near, far, packed_coarse, packed_fine, device, pix0=0, n=None, variant=0, n_classes=0,
                        n_samples=64, n_importance=128, lindisp=False, white_bkgd=False, pe_scalar_factor=1.0,
                        convention="opengl", depth_type="z", want_z=False):
    """inrf_render_fwd_camera: pixels [pix0, pix0 + n) of the frame seen by (K, c2w), rays generated inside the kernels
    (no [N,11] table).  Returns the same dict as render_chunk.  Raises InrfError (INRF_EUNSUPPORTED) for configurations
    the fused kernel does not cover - callers then build the rays and use render_chunk."""
    n = H * W - pix0 if n is None else int(n)
    m = torch.as_tensor(c2w, dtype=torch.float32).detach().cpu()[:3, :4].contiguous()
    cam = _lib.Camera(H=int(H), W=int(W), fx=float(K[0][0]), fy=float(K[1][1]), cx=float(K[0][2]), cy=float(K[1][2]),
                      near_=float(near), far_=float(far), convention=1 if convention == "opencv" else 0,
                      euclidean=1 if depth_type == "euclidean" else 0)
    for i, v in enumerate(m.reshape(-1).tolist()):
        cam.c2w[i] = float(v)
    cfg = RenderCfg(variant=variant, n_classes=n_classes, n_samples=n_samples, n_importance=n_importance, lindisp=int(bool(lindisp)),
                    white_bkgd=int(bool(white_bkgd)), endpoint_feat=0, precision=_DEFAULT_PRECISION, pe_scalar_factor=float(pe_scalar_factor))
    L = _lib.lib()
    ws = _Workspace.get(device, check(L.inrf_render_workspace_bytes(C.byref(cfg), n)))
    new = lambda *s: torch.empty(*s, dtype=torch.float32, device=device)  # noqa: E731
    out = {"rec_coarse": new(n, REC_BASE + n_classes)}
    if n_importance > 0:
        out["rec_fine"], out["z_std"] = new(n, REC_BASE + n_classes), new(n)
        if want_z:
            out["z_fine"] = new(n, n_samples + n_importance)
    with _dev(device):
        check(L.inrf_render_fwd_camera(C.byref(cam), int(pix0), n, _ptr(packed_coarse), _ptr(packed_fine), C.byref(cfg),
                                       _ptr(linspace01(n_samples, device)), _ptr(linspace01(n_importance, device)) if n_importance > 0 else None,
                                       _ptr(out["rec_coarse"]), _ptr(out.get("rec_fine")), _ptr(out.get("z_std")), _ptr(out.get("z_fine")),
                                       _ptr(ws), ws.numel(), _stream()))
    return out


_PLANE_SPEC = {  # name -> (dtype, trailing width)
    "rgb8": (torch.uint8, 3), "albedo8": (torch.uint8, 3), "shading8": (torch.uint8, 1), "residual8": (torch.uint8, 3),
    "label8": (torch.uint8, 1), "vis_label8": (torch.uint8, 3), "entropy8": (torch.uint8, 1), "entropy": (torch.float32, 1),
    "disp16": (torch.uint16, 1), "depth_mm16": (torch.uint16, 1), "labels64": (torch.int64, 1),
}


def frame_finish(rec, H, W, n_classes=0, planes=("rgb8", "albedo8", "shading8", "residual8", "label8"), acc_threshold=10.0,
                 colour_map=None, sub_step=0):
    """inrf_frame_finish: the per-ray record of one rendered frame, rec[H*W, >=13+C], -> dict of device planes
    shaped [H, W(, 3)] (render_path's to8b / uint16 / label / entropy conversions, run_nerf.py:164-215,
    trainer.py:1241-1389), plus sample_pixels [ceil(H/s)*ceil(W/s), 3] and sample_labels [.., 1] (the
    albedo[::s, ::s] / label[::s, ::s] sub-sampling that feeds the cluster refresh) when sub_step=s > 0."""
    rec = _f32(rec, "rec")
    if rec.ndim != 2 or rec.shape[0] != H * W:
        raise ValueError("rec must be [H*W, 13 + n_classes (+128)]")
    dev = rec.device
    out, tab = {}, _lib.FramePlanes()
    for name in planes:
        dt, w = _PLANE_SPEC[name]
        t = torch.empty((H, W, 3) if w == 3 else (H, W), dtype=dt, device=dev)
        out[name] = t
        setattr(tab, name, t.data_ptr())
    if sub_step > 0:
        hs, ws = (H + sub_step - 1) // sub_step, (W + sub_step - 1) // sub_step
        out["sample_pixels"] = torch.empty(hs * ws, 3, dtype=torch.float32, device=dev)
        out["sample_labels"] = torch.empty(hs * ws, 1, dtype=torch.int64, device=dev)
        tab.sample_pixels, tab.sample_labels = out["sample_pixels"].data_ptr(), out["sample_labels"].data_ptr()
    cmap = None
    if colour_map is not None:
        cmap = torch.as_tensor(colour_map).to(device=dev, dtype=torch.uint8).contiguous()
        if cmap.ndim != 2 or cmap.shape[1] != 3 or cmap.shape[0] < n_classes:
            raise ValueError("colour_map must be uint8 [>= n_classes, 3]")
    with _dev(dev):
        check(_lib.lib().inrf_frame_finish(_ptr(rec), int(H), int(W), rec.shape[1], int(n_classes), float(acc_threshold),
                                           _ptr(cmap), int(sub_step), C.byref(tab), _stream()))
    return out


def edit_recompose(cluster_rgb, rec, want_c8=True, want_edit8=True):
    """inrf_edit_recompose: c8 = to8b(cluster_rgb), edit8 = to8b(cluster_rgb * shading + residual), both [P, 3] uint8
    (run_nerf.py:228-240, trainer.py:1425-1441)."""
    cluster_rgb, rec = _f32(cluster_rgb, "cluster_rgb").reshape(-1, 3), _f32(rec, "rec")
    P = cluster_rgb.shape[0]
    if rec.ndim != 2 or rec.shape[0] != P:
        raise ValueError("rec must have one record per pixel")
    c8 = torch.empty(P, 3, dtype=torch.uint8, device=rec.device) if want_c8 else None
    e8 = torch.empty(P, 3, dtype=torch.uint8, device=rec.device) if want_edit8 else None
    with _dev(rec.device):
        check(_lib.lib().inrf_edit_recompose(_ptr(cluster_rgb), _ptr(rec), P, rec.shape[1], _ptr(c8), _ptr(e8), _stream()))
    return c8, e8
