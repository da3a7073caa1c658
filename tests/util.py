"""Shared helpers for the parity tests."""
import numpy as np
import torch

from oracle import nerf_oracle as orc

SEED = 20220414


def rel_err(a, b, floor=1e-3):
    """max |a-b| / max(|b|, floor), NaNs must coincide."""
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    na, nb = torch.isnan(a), torch.isnan(b)
    assert torch.equal(na, nb), "NaN pattern differs"
    a, b = a[~na], b[~nb]
    if a.numel() == 0:
        return 0.0
    return float(((a - b).abs() / b.abs().clamp_min(floor)).max())


def apply_regime(net, regime):
    """Weight regimes of SURVEY section 7-1 (same as tests/golden/make_golden.py:apply_regime)."""
    with torch.no_grad():
        if regime == "opaque":
            net.alpha_linear.bias += 1.0
            net.pts_linears[7].weight *= 3.0
        elif regime == "trained_like":
            net.pts_linears[7].weight *= 6.0
            net.alpha_linear.weight *= 30.0
        elif regime != "default":
            raise ValueError(regime)


def build_nets(variant="object", n_classes=0, opaque=True, device="cuda", regime=None):
    """Our modules with the reference's seeded init (+ optional 'opaque' tweak / a named weight regime), and the same
    weights as oracle parameter dicts on the CPU."""
    import intrinsicnerf_b200 as inrf
    torch.manual_seed(SEED)
    if variant == "object":
        mk = lambda: inrf.NeRF(D=8, W=256, input_ch=63, output_ch=5, skips=[4], input_ch_views=27, use_viewdirs=True)  # noqa: E731
    else:
        mk = lambda: inrf.Semantic_NeRF(n_classes > 0, n_classes, D=8, W=256, input_ch=63, output_ch=5, skips=[4],  # noqa: E731
                                        input_ch_views=27, use_viewdirs=True)
    coarse, fine = mk(), mk()
    if regime is None:
        regime = "opaque" if opaque else "default"
    for net in (coarse, fine):
        apply_regime(net, regime)
    pc = {k: v.detach().clone() for k, v in coarse.state_dict().items()}
    pf = {k: v.detach().clone() for k, v in fine.state_dict().items()}
    return coarse.to(device), fine.to(device), pc, pf


REC = dict(rgb=(0, 3), disp=(3, 4), acc=(4, 5), albedo=(5, 8), shading=(8, 9), residual=(9, 12), depth=(12, 13))


def rec_get(rec, key):
    a, b = REC[key]
    v = rec[:, a:b]
    return v if b - a > 1 else v[:, 0]


def load_golden(golden_dir, name):
    import os
    return {k: v for k, v in np.load(os.path.join(golden_dir, name), allow_pickle=False).items()}


def relu_margin(fn):
    """Run `fn()` and return (result, per-row min |pre-activation| over every torch.relu it applied to a
    2-D [rows, units] input).  A hidden unit whose pre-activation lies within fp32 rounding of zero
    takes a different ReLU branch in an fp32 implementation than in the fp64 oracle, and that single
    branch moves a weight gradient by O(1/rows); gradient parity tests use this margin to leave such
    rows out (zero upstream gradient on both sides) instead of loosening the tolerance."""
    margins = []
    real = torch.relu

    def recording(x):
        if x.dim() == 2:
            margins.append(x.detach().abs().amin(dim=1))
        return real(x)

    torch.relu = recording
    try:
        out = fn()
    finally:
        torch.relu = real
    return out, torch.stack(margins, 0).amin(0)


def decode_images(buf, n_tiles, slots):
    """Tensor-core operand images (include/inrf.h: inrf_mlp_fwd_train_tc) -> [n_tiles, slots, 128, 64] float32:
    each image is 128 rows x 64 fp16 in 8-row atoms of 1024 B, the 16-byte unit index XORed with (row & 7)."""
    x = buf[: n_tiles * slots * 16384].view(torch.float16).reshape(n_tiles, slots, 16, 8, 8, 8)
    r = torch.arange(8, device=buf.device).view(8, 1)
    u = torch.arange(8, device=buf.device).view(1, 8)
    idx = (u ^ r).view(1, 1, 1, 8, 8, 1).expand(n_tiles, slots, 16, 8, 8, 8)
    return torch.gather(x, 4, idx).reshape(n_tiles, slots, 128, 64).float()


def stash_mask_bits(stash, M):
    """The ReLU-mask bit words of the forward stash (csrc/common.cuh IS_MASK) -> bool [M, 40 * 64]: column (s * 64 + c) is
    the decision for column c of image slot 2 + s (trunk h0..h7, views', albedo1|shading1, sem1)."""
    T = (M + 127) // 128
    slots = stash.numel() // (T * 16384)
    w = stash.view(T, slots * 16384)[:, 42 * 16384: 42 * 16384 + 40 * 1024].contiguous().view(torch.int32).reshape(T, 40, 2, 128)
    i = torch.arange(32, device=stash.device, dtype=torch.int32)
    bits = (w.unsqueeze(-1) >> ((i >> 1) + 16 * (i & 1))) & 1            # [T, 40, 2, 128, 32]: even columns in bits 0..15, odd in 16..31
    return bits.permute(0, 3, 1, 2, 4).reshape(T * 128, 40 * 64)[:M].bool().cpu()


def stash_activations(stash, M, n_classes):
    """Decoded forward stash -> dict of [M, width] activations: 'pe' 64, 'dir' 32, 'h0'..'h7' 256, 'v' 128, 'as' 256, ['s1' 128]."""
    T = (M + 127) // 128
    img = decode_images(stash, T, stash.numel() // (T * 16384))      # 42 image slots + the ReLU-mask bit words
    rows = lambda s0, n: img[:, s0:s0 + n].permute(0, 2, 1, 3).reshape(T * 128, 64 * n)[:M].cpu()  # noqa: E731
    out = {"pe": rows(0, 1), "dir": rows(1, 1)[:, :32], "v": rows(34, 2), "as": rows(36, 4)}
    for l in range(8):
        out[f"h{l}"] = rows(2 + 4 * l, 4)
    if n_classes > 0:
        out["s1"] = rows(40, 2)
    return out


def encode_images(x):
    """Inverse of decode_images: [n_tiles, slots, 128, 64] values -> flat uint8 buffer of fp16 operand images."""
    n_tiles, slots = x.shape[:2]
    v = x.to(torch.float16).reshape(n_tiles, slots, 16, 8, 8, 8)
    r = torch.arange(8, device=x.device).view(8, 1)
    u = torch.arange(8, device=x.device).view(1, 8)
    idx = (u ^ r).view(1, 1, 1, 8, 8, 1).expand(n_tiles, slots, 16, 8, 8, 8)
    return torch.gather(v, 4, idx).contiguous().view(torch.uint8).reshape(-1)    # XOR is an involution: same gather
