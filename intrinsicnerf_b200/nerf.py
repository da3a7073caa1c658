"""Field-network modules with the reference's constructor signatures, parameter names and
state_dict layout, backed by the sm_100a kernels.

  Embedder / get_embedder   object_level/run_nerf_helpers.py:195-243, SSR/models/semantic_nerf.py:14-65
  NeRF                      object_level/run_nerf_helpers.py:247-325
  Semantic_NeRF             SSR/models/semantic_nerf.py:74-181

Checkpoints written by the reference load unchanged (same module attribute names, same
construction order, so ``torch.manual_seed`` + construction gives the same initial weights).
Only the architecture the reference's configs use is implemented in CUDA (D=8, W=256,
skips=[4], use_viewdirs=True, multires 10 / 4); anything else raises - there is no fallback.
"""
import torch
import torch.nn as nn

from . import ops
from ._lib import NET_OBJECT, NET_SSR


class Embedder:
    """gamma(x) = [x, sin(2^k x), cos(2^k x)]_k.  Same kwargs as the reference class."""

    def __init__(self, **kwargs):
        self.kwargs = kwargs
        if kwargs.get("input_dims", 3) != 3 or not kwargs.get("include_input", True) \
                or not kwargs.get("log_sampling", True):
            raise NotImplementedError("only include_input=True, input_dims=3, log_sampling=True is implemented")
        self.n_freqs = int(kwargs["num_freqs"])
        if int(kwargs["max_freq_log2"]) != self.n_freqs - 1:
            raise NotImplementedError("frequency bands must be 2^0..2^(L-1)")
        self.scalar_factor = float(kwargs.get("scalar_factor", 1.0))
        self.out_dim = 3 + 6 * self.n_freqs

    def embed(self, inputs):
        return ops.embed(inputs, self.n_freqs, self.scalar_factor)

    __call__ = embed


def get_embedder(multires, i=0, scalar_factor=1):
    """Returns (embed_fn, out_dim).  ``scalar_factor`` is the SSR fork's extra argument
    (semantic_nerf.py:50, the input is divided by it)."""
    if i == -1:
        raise NotImplementedError("i_embed=-1 (no positional encoding) is not implemented by the CUDA path")
    eo = Embedder(include_input=True, input_dims=3, max_freq_log2=multires - 1, num_freqs=multires,
                  log_sampling=True, periodic_fns=[torch.sin, torch.cos], scalar_factor=scalar_factor)
    return eo, eo.out_dim


class _FieldNet(nn.Module):
    variant = NET_OBJECT

    def _check_arch(self, D, W, input_ch, input_ch_views, skips, use_viewdirs):
        if (D, W, input_ch, input_ch_views, list(skips), bool(use_viewdirs)) != (8, 256, 63, 27, [4], True):
            raise NotImplementedError(
                "the CUDA path implements D=8, W=256, input_ch=63, input_ch_views=27, skips=[4], use_viewdirs=True "
                f"(got D={D}, W={W}, input_ch={input_ch}, input_ch_views={input_ch_views}, skips={skips}, "
                f"use_viewdirs={use_viewdirs})")

    # canonical layer order of include/inrf.h: inrf_flat_param_count
    def _ordered_layers(self):
        raise NotImplementedError

    @property
    def n_classes(self):
        return 0

    def flat_params(self):
        parts = []
        for lin in self._ordered_layers():
            parts.append(lin.weight.reshape(-1))
            parts.append(lin.bias.reshape(-1))
        return torch.cat(parts).detach().float()

    def packed(self):
        """Device blob for the kernels, re-packed only when a parameter changed."""
        ps = list(self.parameters())
        key = (ps[0].device, tuple(p._version for p in ps), tuple(p.data_ptr() for p in ps))
        if getattr(self, "_packed_key", None) != key:
            if not ps[0].is_cuda:
                raise RuntimeError("intrinsicnerf_b200 networks run on CUDA only: call .cuda() first (no CPU fallback)")
            self._packed_blob = ops.pack_weights(self.flat_params(), self.variant, self.n_classes)
            self._packed_key = key
        return self._packed_blob

    def needs_grad(self):
        """True when the call must be recorded by autograd (training): the training kernels are used (forward
        with activation stash + backward; tensor cores by default, strict fp32 CUDA cores under
        ops.set_default_precision("fp32")); otherwise the tensor-core inference path."""
        return torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())

    def flat_params_diff(self):
        """Canonical flat parameter vector as a differentiable function of the module parameters."""
        parts = []
        for lin in self._ordered_layers():
            parts.append(lin.weight.reshape(-1))
            parts.append(lin.bias.reshape(-1))
        return torch.cat(parts).float()

    def evaluate(self, mode, a, b=None, endpoint=False, pe_scalar_factor=1.0):
        """raw rows for (pts, viewdirs) | (rays, z) | embedded rows; differentiable w.r.t. the parameters
        in training mode."""
        C = self.n_classes
        if self.needs_grad():
            if not next(self.parameters()).is_cuda:
                raise RuntimeError("intrinsicnerf_b200 networks run on CUDA only: call .cuda() first (no CPU fallback)")
            fn = ops.MlpFn if ops.default_precision() == ops.PREC_FP32 else ops.MlpTcFn
            return fn.apply(self.flat_params_diff(), self.variant, C, bool(endpoint), float(pe_scalar_factor), mode, a, b)
        if mode == "pts":
            return ops.mlp_forward(self.packed(), self.variant, C, a, b, endpoint, pe_scalar_factor)
        if mode == "rays":
            out = ops.mlp_forward_rays(self.packed(), self.variant, C, a, b, endpoint, pe_scalar_factor)
            return out.reshape(-1, out.shape[-1])
        out = ops.mlp_forward_embedded(self.packed(), self.variant, C, a, endpoint)
        return out.reshape(-1, out.shape[-1])


    @torch.no_grad()
    def query_points(self, pts, viewdirs=None, endpoint=False, pe_scalar_factor=1.0, chunk=1 << 22):
        """Dense field query (SSR/extract_colour_mesh.py:149-166: 256^3 grid points through the fine network with
        zero view directions, then marching cubes on the density): raw rows [M, 11 + C (+128)] for pts[M,3] in
        `chunk`-row launches, everything staying on the device.  viewdirs None = zeros, as the reference passes."""
        pts = pts.reshape(-1, 3)
        out = torch.empty(pts.shape[0], ops.RAW_BASE + self.n_classes + (128 if endpoint else 0), dtype=torch.float32, device=pts.device)
        for i in range(0, pts.shape[0], chunk):
            p = pts[i:i + chunk]
            d = torch.zeros_like(p) if viewdirs is None else viewdirs.reshape(-1, 3)[i:i + chunk]
            out[i:i + chunk] = ops.mlp_forward(self.packed(), self.variant, self.n_classes, p, d, endpoint, pe_scalar_factor)
        return out


class NeRF(_FieldNet):
    variant = NET_OBJECT

    def __init__(self, D=8, W=256, input_ch=3, input_ch_views=3, output_ch=4, skips=[4], use_viewdirs=False):
        super().__init__()
        self._check_arch(D, W, input_ch, input_ch_views, skips, use_viewdirs)
        self.D, self.W, self.input_ch, self.input_ch_views = D, W, input_ch, input_ch_views
        self.skips, self.use_viewdirs = skips, use_viewdirs
        # construction order = reference order (RNG stream and state_dict key order)
        self.pts_linears = nn.ModuleList(
            [nn.Linear(input_ch, W)] + [nn.Linear(W + input_ch, W) if i in skips else nn.Linear(W, W) for i in range(D - 1)])
        self.views_linears = nn.ModuleList([nn.Linear(input_ch_views + W, W // 2)])
        self.feature_linear = nn.Linear(W, W)
        self.alpha_linear = nn.Linear(W, 1)
        self.shading_linear = nn.Linear(W // 2, 3)      # residual head (sic, appendix A10)
        self.albedo_linear1 = nn.Linear(W, W // 2)
        self.albedo_linear2 = nn.Linear(W // 2, 3)
        self.test_linear1 = nn.Linear(W, W // 2)        # shading head (sic)
        self.test_linear2 = nn.Linear(W // 2, 1)

    def _ordered_layers(self):
        return list(self.pts_linears) + [self.alpha_linear, self.feature_linear, self.views_linears[0],
                                         self.albedo_linear1, self.albedo_linear2, self.test_linear1,
                                         self.test_linear2, self.shading_linear]

    def forward(self, x):
        """x: [..., 90] embedded rows (gamma(x) | gamma(d)) -> [..., 11]."""
        out = self.evaluate("emb", x)
        return out.reshape(*x.shape[:-1], out.shape[-1])


class Semantic_NeRF(_FieldNet):
    variant = NET_SSR

    def __init__(self, enable_semantic, num_semantic_classes, D=8, W=256, input_ch=3, input_ch_views=3, output_ch=4,
                 skips=[4], use_viewdirs=False):
        super().__init__()
        self._check_arch(D, W, input_ch, input_ch_views, skips, use_viewdirs)
        self.D, self.W, self.input_ch, self.input_ch_views = D, W, input_ch, input_ch_views
        self.skips, self.use_viewdirs, self.enable_semantic = skips, use_viewdirs, enable_semantic
        self.num_semantic_classes = int(num_semantic_classes) if enable_semantic else 0
        self.pts_linears = nn.ModuleList(
            [nn.Linear(input_ch, W)] + [nn.Linear(W + input_ch, W) if i in skips else nn.Linear(W, W) for i in range(D - 1)])
        self.views_linears = nn.ModuleList([nn.Linear(input_ch_views + W, W // 2)])
        self.feature_linear = nn.Linear(W, W)
        self.alpha_linear = nn.Linear(W, 1)
        if enable_semantic:
            self.semantic_linear = nn.Sequential(nn.Sequential(nn.Linear(W, W // 2), nn.ReLU(True)),
                                                 nn.Linear(W // 2, num_semantic_classes))
        self.residual_linear = nn.Linear(W // 2, 3)
        self.albedo_linear1 = nn.Linear(W, W // 2)
        self.albedo_linear2 = nn.Linear(W // 2, 3)
        self.shading_linear1 = nn.Linear(W, W // 2)
        self.shading_linear2 = nn.Linear(W // 2, 1)

    @property
    def n_classes(self):
        return self.num_semantic_classes

    def _ordered_layers(self):
        layers = list(self.pts_linears) + [self.alpha_linear, self.feature_linear, self.views_linears[0],
                                           self.albedo_linear1, self.albedo_linear2, self.shading_linear1,
                                           self.shading_linear2, self.residual_linear]
        if self.enable_semantic:
            layers += [self.semantic_linear[0][0], self.semantic_linear[1]]
        return layers

    def forward(self, x, show_endpoint=False):
        out = self.evaluate("emb", x, endpoint=bool(show_endpoint))
        return out.reshape(*x.shape[:-1], out.shape[-1])
