"""CPU oracle for the IntrinsicNeRF ray-marching hot path.  TEST INFRASTRUCTURE ONLY.

This file is a plain PyTorch-fp32 (CPU) restatement of the reference algorithm for the
path SURVEY.md section 8 scopes: positional encoding, the 8x256 intrinsic MLP (object
fork and SSR fork), alpha compositing, inverse-CDF resampling, merge-sort and the
render_rays driver.  It exists so that the CUDA path can be checked against it; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import
it.  The product (``intrinsicnerf_b200``) never does.

Pinning: ``tests/golden/make_golden.py`` runs the unmodified reference (imported from
/root/reference in the build container) on seeded inputs and stores inputs+outputs in
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks this oracle against
those files (and against the live reference when it is present).  The third-party
arithmetic underneath is PyTorch ATen (searchsorted/cumsum/cumprod/sort); the oracle
calls the same ATen primitives, so on the same torch build it reproduces the reference
to the last bit on CPU.

Reference citations are ``path:line`` relative to the reference checkout.
"""
import math

import numpy as np
import torch

# ----------------------------------------------------------------------------------
# Parameter naming (state_dict keys) of the two forks
# ----------------------------------------------------------------------------------
TRUNK = [f"pts_linears.{i}" for i in range(8)]
# object fork: object_level/run_nerf_helpers.py:259-279 (shading_linear is the residual
# head, test_linear1/2 the shading head - SURVEY appendix A10)
OBJECT_HEADS = dict(alpha="alpha_linear", feature="feature_linear", views="views_linears.0",
                    albedo1="albedo_linear1", albedo2="albedo_linear2",
                    shading1="test_linear1", shading2="test_linear2",
                    residual="shading_linear")
# SSR fork: SSR/models/semantic_nerf.py:100-118
SSR_HEADS = dict(alpha="alpha_linear", feature="feature_linear", views="views_linears.0",
                 albedo1="albedo_linear1", albedo2="albedo_linear2",
                 shading1="shading_linear1", shading2="shading_linear2",
                 residual="residual_linear",
                 sem1="semantic_linear.0.0", sem2="semantic_linear.1")


def _lin(p, name, x):
    return torch.nn.functional.linear(x, p[name + ".weight"], p[name + ".bias"])


# ----------------------------------------------------------------------------------
# E1  positional encoding
# ----------------------------------------------------------------------------------
def posenc(x, n_freqs, scale=1.0):
    """gamma(x) = [x, sin(2^0 x), cos(2^0 x), ..., sin(2^(L-1) x), cos(2^(L-1) x)].

    object_level/run_nerf_helpers.py:195-225 (Embedder) and SSR/models/semantic_nerf.py:14-65
    (same, with the input divided by ``scalar_factor`` first, line 64).
    Output is [..., 3 + 6L]; each sin/cos block is 3 wide.
    """
    if scale != 1.0:
        x = x / scale
    parts = [x]
    for k in range(n_freqs):
        f = float(2.0 ** k)          # 2.**linspace(0, L-1, L) is exact powers of two
        parts.append(torch.sin(x * f))
        parts.append(torch.cos(x * f))
    return torch.cat(parts, dim=-1)


# ----------------------------------------------------------------------------------
# M1 / M2  field network
# ----------------------------------------------------------------------------------
def mlp_forward(params, emb, variant="object", n_classes=0, endpoint=False):
    """Forward of NeRF (object_level/run_nerf_helpers.py:284-325) or Semantic_NeRF
    (SSR/models/semantic_nerf.py:123-181) on an embedded batch ``emb`` [M, 63+27].

    ``params`` is a state_dict-style mapping.  Returns [M, 11 (+C) (+128)]:
    rgb3 | sigma1 | albedo3 | shading1 | residual3 | sem_logits C | endpoint feature 128.
    """
    names = OBJECT_HEADS if variant == "object" else SSR_HEADS
    pe_pts, pe_dir = emb[:, :63], emb[:, 63:]
    h = pe_pts
    for i, name in enumerate(TRUNK):
        h = torch.relu(_lin(params, name, h))
        if i == 4:                       # skips=[4]: concat AFTER layer 4's ReLU
            h = torch.cat([pe_pts, h], dim=-1)
    sigma = _lin(params, names["alpha"], h)                        # no activation here
    albedo = torch.sigmoid(_lin(params, names["albedo2"], torch.relu(_lin(params, names["albedo1"], h))))
    shading = torch.sigmoid(_lin(params, names["shading2"], torch.relu(_lin(params, names["shading1"], h))))
    feat = _lin(params, names["feature"], h)
    h2 = torch.relu(_lin(params, names["views"], torch.cat([feat, pe_dir], dim=-1)))
    residual = torch.sigmoid(_lin(params, names["residual"], h2))
    rgb = albedo * shading + residual
    cols = [rgb, sigma, albedo, shading, residual]
    if variant == "ssr" and n_classes > 0:
        sem = _lin(params, names["sem2"], torch.relu(_lin(params, names["sem1"], h)))
        cols.append(sem)
    if endpoint:
        cols.append(h2)
    return torch.cat(cols, dim=-1)


def mlp_forward_tc_arith(params, emb, variant="object", n_classes=0, endpoint=False, masks=None):
    """The same network in the ARITHMETIC of the tensor-core kernels: every GEMM operand
    (weights and incoming activations) rounded to fp16 (round-to-nearest, straight-through gradient), products
    and sums exact, biases / sigma head / sigmoids unrounded, views_linears.0 o feature_linear composed into one
    matrix before rounding.  Not a different algorithm - it exists because a ReLU unit whose pre-activation lies
    within the operand rounding of zero takes the other branch in exact arithmetic, and one flipped (sample, unit)
    pair is a full-size difference in that sample's gradient: gradient parity of the tensor-core training path is
    measured against THIS function (float64 inputs), value parity against mlp_forward.
    ``masks`` (optional): one 0/1 tensor per ReLU in call order (trunk 0..7, albedo1, shading1, views, [sem1]);
    when given, relu(z) is evaluated as z * mask, i.e. with the branch decisions of another run of the same network
    (the kernel's, read back from its activation stash) - the network is piecewise linear, so autograd then yields the
    exact gradient for precisely those decisions.
    Returns (out, margin): margin[m] = min |pre-activation| over every ReLU of sample m."""
    names = OBJECT_HEADS if variant == "object" else SSR_HEADS
    rnd = lambda x: x + (x.to(torch.float16).to(x.dtype) - x).detach()  # noqa: E731
    margins = []

    def lin(name, x):
        return rnd(x) @ rnd(params[name + ".weight"]).t() + params[name + ".bias"]

    def relu(z):
        margins.append(z.detach().abs().amin(dim=1))
        if masks is not None:
            return z * masks[len(margins) - 1].to(z.dtype)
        return torch.relu(z)

    pe_pts, pe_dir = emb[:, :63], emb[:, 63:]
    h = pe_pts
    for i, name in enumerate(TRUNK):
        h = relu(lin(name, h))
        if i == 4:
            h = torch.cat([pe_pts, h], dim=-1)
    sigma = _lin(params, names["alpha"], h)
    albedo = torch.sigmoid(lin(names["albedo2"], relu(lin(names["albedo1"], h))))
    shading = torch.sigmoid(lin(names["shading2"], relu(lin(names["shading1"], h))))
    wv, wf = params[names["views"] + ".weight"], params[names["feature"] + ".weight"]
    wc = wv[:, :256] @ wf
    bc = wv[:, :256] @ params[names["feature"] + ".bias"] + params[names["views"] + ".bias"]
    h2 = relu(rnd(h) @ rnd(wc).t() + rnd(pe_dir) @ rnd(wv[:, 256:]).t() + bc)
    residual = torch.sigmoid(lin(names["residual"], h2))
    cols = [albedo * shading + residual, sigma, albedo, shading, residual]
    if variant == "ssr" and n_classes > 0:
        cols.append(lin(names["sem2"], relu(lin(names["sem1"], h))))
    if endpoint:
        cols.append(h2)
    return torch.cat(cols, dim=-1), torch.stack(margins, 0).amin(0)


def query_field(pts, viewdirs, params, variant="object", n_classes=0, endpoint=False,
                pe_scale_pts=1.0, netchunk=65536):
    """run_network: object_level/run_nerf.py:42-56, SSR/models/model_utils.py:19-35.

    pts [N,S,3], viewdirs [N,3] -> [N,S,out].  View directions are expanded per sample
    and embedded per sample exactly like the reference (lines 49-51).
    """
    N, S, _ = pts.shape
    flat = pts.reshape(-1, 3)
    dirs = viewdirs[:, None, :].expand(N, S, 3).reshape(-1, 3)
    emb = torch.cat([posenc(flat, 10, pe_scale_pts), posenc(dirs, 4, 1.0)], dim=-1)
    outs = [mlp_forward(params, emb[i:i + netchunk], variant, n_classes, endpoint)
            for i in range(0, emb.shape[0], netchunk)]
    out = torch.cat(outs, dim=0)
    return out.reshape(N, S, out.shape[-1])


# ----------------------------------------------------------------------------------
# O1 / O2  alpha compositing
# ----------------------------------------------------------------------------------
def composite(raw, z_vals, rays_d, noise=None, white_bkgd=False, n_classes=0, endpoint=False):
    """raw2outputs: object_level/run_nerf.py:359-412 and SSR/models/model_utils.py:39-116.

    ``noise`` (already scaled by raw_noise_std) is added to sigma before the ReLU
    (run_nerf.py:385-395).  Returns a dict; ``disp`` is NaN where sum(weights)==0
    (run_nerf.py:404).  White background is added to rgb, albedo,
    shading and the semantic map, NOT to the residual (run_nerf.py:407-410,
    model_utils.py:109-114).
    """
    delta = z_vals[:, 1:] - z_vals[:, :-1]
    delta = torch.cat([delta, torch.full_like(delta[:, :1], 1e10)], dim=-1)
    delta = delta * torch.norm(rays_d[:, None, :], dim=-1)
    sigma = raw[..., 3] if noise is None else raw[..., 3] + noise
    alpha = 1.0 - torch.exp(-torch.relu(sigma) * delta)
    trans = torch.cumprod(torch.cat([torch.ones_like(alpha[:, :1]), 1.0 - alpha + 1e-10], dim=-1), dim=-1)[:, :-1]
    w = alpha * trans
    out = {"weights": w}
    out["rgb"] = torch.sum(w[..., None] * raw[..., 0:3], dim=-2)
    out["albedo"] = torch.sum(w[..., None] * raw[..., 4:7], dim=-2)
    out["shading"] = torch.sum(w * raw[..., 7], dim=-1)
    out["residual"] = torch.sum(w[..., None] * raw[..., 8:11], dim=-2)
    if n_classes > 0:
        out["sem"] = torch.sum(w[..., None] * raw[..., 11:11 + n_classes], dim=-2)
    if endpoint:
        out["feat"] = torch.sum(w[..., None] * raw[..., -128:], dim=-2)
    out["depth"] = torch.sum(w * z_vals, dim=-1)
    out["acc"] = torch.sum(w, dim=-1)
    out["disp"] = 1.0 / torch.max(1e-10 * torch.ones_like(out["depth"]), out["depth"] / torch.sum(w, dim=-1))
    if white_bkgd:
        bg = 1.0 - out["acc"]
        out["rgb"] = out["rgb"] + bg[:, None]
        out["albedo"] = out["albedo"] + bg[:, None]
        out["shading"] = out["shading"] + bg
        if n_classes > 0:
            out["sem"] = out["sem"] + bg[:, None]
    return out


# ----------------------------------------------------------------------------------
# S1  hierarchical sampling
# ----------------------------------------------------------------------------------
def build_cdf(weights):
    """pdf/cdf of run_nerf_helpers.py:404-407: w+1e-5, normalise, cumsum, prepend 0."""
    w = weights + 1e-5
    pdf = w / torch.sum(w, dim=-1, keepdim=True)
    cdf = torch.cumsum(pdf, dim=-1)
    return torch.cat([torch.zeros_like(cdf[:, :1]), cdf], dim=-1)


def invert_cdf(bins, cdf, u):
    """Inverse-CDF lookup of run_nerf_helpers.py:428-443 (SSR/models/rays.py:201-218).

    Returns (samples [N,Sf], inds int64 [N,Sf]) with inds = searchsorted(cdf, u, right=True).
    """
    u = u.contiguous()
    inds = torch.searchsorted(cdf.detach(), u, right=True)
    lo = torch.clamp(inds - 1, min=0)
    hi = torch.clamp(inds, max=cdf.shape[-1] - 1)
    cdf_lo, cdf_hi = torch.gather(cdf, 1, lo), torch.gather(cdf, 1, hi)
    bin_lo, bin_hi = torch.gather(bins, 1, lo), torch.gather(bins, 1, hi)
    denom = cdf_hi - cdf_lo
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)
    t = (u - cdf_lo) / denom
    return bin_lo + t * (bin_hi - bin_lo), inds


def sample_pdf(bins, weights, n_samples, u=None):
    """sample_pdf (run_nerf_helpers.py:402-445): u=None means det (linspace(0,1,Sf))."""
    cdf = build_cdf(weights)
    if u is None:
        u = torch.linspace(0.0, 1.0, n_samples).expand(cdf.shape[0], n_samples)
    samples, inds = invert_cdf(bins, cdf, u)
    return samples, inds, cdf


# ----------------------------------------------------------------------------------
# R3  render_rays
# ----------------------------------------------------------------------------------
def coarse_z(near, far, n_samples, lindisp=False, t_rand=None):
    """z_vals of run_nerf.py:464-486 / trainer.py:730-746 (stratified jitter iff t_rand)."""
    t = torch.linspace(0.0, 1.0, n_samples)
    if not lindisp:
        z = near * (1.0 - t) + far * t
    else:
        z = 1.0 / (1.0 / near * (1.0 - t) + 1.0 / far * t)
    z = z.expand(near.shape[0], n_samples)
    if t_rand is not None:
        mid = 0.5 * (z[:, 1:] + z[:, :-1])
        upper = torch.cat([mid, z[:, -1:]], dim=-1)
        lower = torch.cat([z[:, :1], mid], dim=-1)
        z = lower + (upper - lower) * t_rand
    return z


def render_rays(rays, coarse, fine, variant="object", n_classes=0, n_samples=64,
                n_importance=128, lindisp=False, white_bkgd=False, pe_scale_pts=1.0,
                t_rand=None, u=None, noise_coarse=None, noise_fine=None, endpoint=False,
                netchunk=65536):
    """Volumetric rendering of one ray chunk.

    object fork: object_level/run_nerf.py:415-528; SSR fork:
    SSR/training/trainer.py:717-808.  ``rays`` is [N,11] = o3 d3 near far viewdir3.
    Random draws are injected (t_rand [N,Sc], u [N,Sf], noise_* [N,S] pre-scaled); None
    selects the deterministic branch (perturb=0 / raw_noise_std=0).
    Returns dict with 'coarse' and 'fine' composite dicts, 'z_coarse', 'z_fine',
    'z_samples', 'z_std', 'raw_coarse', 'raw_fine', 'inds'.
    """
    o, d, vd = rays[:, 0:3], rays[:, 3:6], rays[:, 8:11]
    near, far = rays[:, 6:7], rays[:, 7:8]
    z = coarse_z(near, far, n_samples, lindisp, t_rand)
    pts = o[:, None, :] + d[:, None, :] * z[:, :, None]
    raw_c = query_field(pts, vd, coarse, variant, n_classes, False, pe_scale_pts, netchunk)
    comp_c = composite(raw_c, z, d, noise_coarse, white_bkgd, n_classes, False)
    res = {"coarse": comp_c, "z_coarse": z, "raw_coarse": raw_c}
    if n_importance > 0:
        z_mid = 0.5 * (z[:, 1:] + z[:, :-1])
        z_samples, inds, cdf = sample_pdf(z_mid, comp_c["weights"][:, 1:-1], n_importance, u)
        z_samples = z_samples.detach()
        z_all, _ = torch.sort(torch.cat([z, z_samples], dim=-1), dim=-1)
        pts = o[:, None, :] + d[:, None, :] * z_all[:, :, None]
        raw_f = query_field(pts, vd, fine if fine is not None else coarse, variant, n_classes,
                            endpoint, pe_scale_pts, netchunk)
        comp_f = composite(raw_f, z_all, d, noise_fine, white_bkgd, n_classes, endpoint)
        res.update(fine=comp_f, z_fine=z_all, z_samples=z_samples, inds=inds, cdf=cdf,
                   raw_fine=raw_f, z_std=torch.std(z_samples, dim=-1, unbiased=False))
    return res


def render_image(H, W, K, c2w, near, far, coarse, fine, chunk=32768, **kw):
    """render(): object_level/run_nerf.py:74-139 with c2w given (full image), ndc=False,
    use_viewdirs=True.  Returns dict of [H,W,...] maps (fine + '*0' coarse + z_std)."""
    rays_o, rays_d = get_rays(H, W, K, c2w)
    vd = rays_d / torch.norm(rays_d, dim=-1, keepdim=True)
    o, d, vd = rays_o.reshape(-1, 3).float(), rays_d.reshape(-1, 3).float(), vd.reshape(-1, 3).float()
    rays = torch.cat([o, d, near * torch.ones_like(d[:, :1]), far * torch.ones_like(d[:, :1]), vd], dim=-1)
    acc = {}
    for i in range(0, rays.shape[0], chunk):
        r = render_rays(rays[i:i + chunk], coarse, fine, **kw)
        flat = {}
        for k in ("rgb", "disp", "acc", "albedo", "shading", "residual", "depth"):
            flat[k + "_map"] = r["fine"][k]
            flat[k + "0"] = r["coarse"][k]
        flat["z_std"] = r["z_std"]
        for k, v in flat.items():
            acc.setdefault(k, []).append(v)
    return {k: torch.cat(v, 0).reshape([H, W] + list(v[0].shape[1:])) for k, v in acc.items()}


# ----------------------------------------------------------------------------------
# Ray generation / synthetic cameras (callers of the path; used to build inputs)
# ----------------------------------------------------------------------------------
def get_rays(H, W, K, c2w):
    """object_level/run_nerf_helpers.py:359-368 (pinhole, OpenGL convention)."""
    c2w = torch.as_tensor(c2w, dtype=torch.float32)
    jj, ii = torch.meshgrid(torch.linspace(0, H - 1, H), torch.linspace(0, W - 1, W), indexing="ij")
    dirs = torch.stack([(ii - K[0][2]) / K[0][0], -(jj - K[1][2]) / K[1][1], -torch.ones_like(ii)], dim=-1)
    rays_d = torch.sum(dirs[..., None, :] * c2w[:3, :3], dim=-1)
    rays_o = c2w[:3, -1].expand(rays_d.shape)
    return rays_o, rays_d


def pose_spherical(theta, phi, radius):
    """object_level/load_blender.py:9-34."""
    def t_(t):
        return np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, t], [0, 0, 0, 1]], dtype=np.float32)

    def rphi(a):
        return np.array([[1, 0, 0, 0], [0, np.cos(a), -np.sin(a), 0], [0, np.sin(a), np.cos(a), 0], [0, 0, 0, 1]], dtype=np.float32)

    def rth(a):
        return np.array([[np.cos(a), 0, -np.sin(a), 0], [0, 1, 0, 0], [np.sin(a), 0, np.cos(a), 0], [0, 0, 0, 1]], dtype=np.float32)

    m = t_(radius)
    m = rphi(phi / 180.0 * np.pi) @ m
    m = rth(theta / 180.0 * np.pi) @ m
    m = np.array([[-1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=np.float32) @ m
    return torch.from_numpy(m.astype(np.float32))


def blender_intrinsics(H, W):
    """K of run_nerf.py:762-767 with camera_angle_x of the Blender 'chair' scene
    (load_blender.py:72-73)."""
    f = 0.5 * W / math.tan(0.5 * 0.6911112070083618)
    return [[f, 0.0, 0.5 * W], [0.0, f, 0.5 * H], [0.0, 0.0, 1.0]]


def blender_rays(H, W, theta=-180.0, phi=-30.0, radius=4.0, near=2.0, far=6.0):
    """Packed [H*W, 11] ray records for a synthetic Blender view."""
    K = blender_intrinsics(H, W)
    c2w = pose_spherical(theta, phi, radius)[:3, :4]
    ro, rd = get_rays(H, W, K, c2w)
    vd = rd / torch.norm(rd, dim=-1, keepdim=True)
    o, d, vd = ro.reshape(-1, 3).float(), rd.reshape(-1, 3).float(), vd.reshape(-1, 3).float()
    return torch.cat([o, d, near * torch.ones_like(d[:, :1]), far * torch.ones_like(d[:, :1]), vd], dim=-1).contiguous()


def replica_rays(H=240, W=320, near=0.1, far=10.0, yaw_deg=0.0):
    """Synthetic Replica-like view (SSR/training/trainer.py:61-74: hfov 90, fx=fy=W/2,
    cx=(W-1)/2, cy=(H-1)/2; OpenCV convention, SSR/models/rays.py:27-67)."""
    fx = fy = W / 2.0
    cx, cy = (W - 1.0) / 2.0, (H - 1.0) / 2.0
    jj, ii = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    dirs = torch.stack([(ii - cx) / fx, (jj - cy) / fy, torch.ones_like(ii)], dim=-1).reshape(-1, 3)
    a = yaw_deg / 180.0 * math.pi
    R = torch.tensor([[math.cos(a), 0.0, math.sin(a)], [0.0, 1.0, 0.0], [-math.sin(a), 0.0, math.cos(a)]])
    d = dirs @ R.t()
    o = torch.zeros_like(d)
    vd = d / torch.norm(d, dim=-1, keepdim=True)
    return torch.cat([o, d, near * torch.ones_like(d[:, :1]), far * torch.ones_like(d[:, :1]), vd], dim=-1).contiguous()


def create_rays(Ts_c2w, H, W, fx, fy, cx, cy, near, far, depth_type="z", convention="opencv"):
    """SSR/models/rays.py:48-76 get_rays_camera + :79-84 get_rays_world + :223-256 create_rays
    (use_viewdirs=True, no static camera): [B, H*W, 11]."""
    Ts = torch.as_tensor(Ts_c2w, dtype=torch.float32)
    B = Ts.shape[0]
    jj, ii = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    x = (ii - cx) / fx
    if convention == "opencv":
        y, z = (jj - cy) / fy, torch.ones_like(ii)
    else:
        y, z = -(jj - cy) / fy, -torch.ones_like(ii)
    dirs = torch.stack((x, y, z), dim=-1)
    if depth_type == "euclidean":
        dirs = dirs * (1.0 / torch.norm(dirs, dim=-1, keepdim=True))
    dirs_C = dirs.reshape(1, -1, 3).expand(B, -1, -1)
    dirs_W = torch.matmul(Ts[:, None, :3, :3], dirs_C[..., None]).squeeze(-1)
    origins = Ts[:, None, :3, -1].expand_as(dirs_W)
    viewdirs = dirs_W / torch.norm(dirs_W, dim=-1, keepdim=True)
    nr, fr = near * torch.ones_like(dirs_W[..., :1]), far * torch.ones_like(dirs_W[..., :1])
    return torch.cat([origins, dirs_W, nr, fr, viewdirs], dim=-1)


# ----------------------------------------------------------------------------------
# Training losses (object_level/run_nerf_helpers.py:11-86; SSR/training/training_utils.py:124-207)
# ----------------------------------------------------------------------------------
def _chroma(c):
    s = torch.sum(c, dim=-1) + 1e-5
    return c[:, 0] / s, c[:, 1] / s


def chroma_weight(color1, color2, l1, l2, mode):
    """compute_chroma_weight: object fork multiplies both weights by the two object masks
    (run_nerf_helpers.py:25-35); the SSR fork masks only the first by label equality (training_utils.py:141-152)."""
    r1, g1 = _chroma(color1)
    r2, g2 = _chroma(color2)
    dc = (r1 - r2) ** 2 + (g1 - g2) ** 2
    if mode == "object":
        return torch.exp(-60 * dc) * l1 * l2, dc * l1 * l2
    return torch.exp(-60 * dc) * (l1 == l2).to(dc.dtype), dc


def intrinsic_losses(rgb, albedo, shading, residual, gt_rgb, label, target_albedo=None, mode="object"):
    """(img, chroma, residual, reflect_sparsity, shading_smooth, far_reflect, intensity, cluster): img2mse
    (run_nerf_helpers.py:11), compute_intrinsic_loss (:59-86 / training_utils.py:179-207; the depth weight
    it computes is replaced by the constant 1 in the calls at :78-79, :83) and the cluster term
    (run_nerf.py:987-991)."""
    zero = albedo.new_zeros(())
    img = torch.mean((rgb - gt_rgb) ** 2) if rgb is not None else zero
    r1, g1 = _chroma(albedo)
    r2, g2 = _chroma(gt_rgb)
    chroma = torch.mean((r1 - r2) ** 2) + torch.mean((g1 - g2) ** 2)
    resid = torch.mean(residual ** 2)
    split = albedo.shape[0] // 2
    a1, a2 = albedo[:split], albedo[-1 * split:]
    s1, s2 = shading[:split], shading[-1 * split:]
    c1, c2 = gt_rgb[:split], gt_rgb[-1 * split:]
    l1, l2 = label[:split], label[-1 * split:]
    w, w_inv = chroma_weight(c1, c2, l1, l2, mode)
    reflect = torch.mean(w * torch.sum((a1 - a2) ** 2, dim=-1))
    shade = torch.mean(w_inv * (s1 - s2) ** 2)
    split2 = a1.shape[0] // 2
    wf, _ = chroma_weight(c1[:split2], c1[-1 * split2:], l1[:split2], l1[-1 * split2:], mode)
    far = torch.mean(wf * torch.sum((a1[:split2] - a1[-1 * split2:]) ** 2, dim=-1))
    intens = (torch.mean(gt_rgb) - torch.mean(albedo)) ** 2
    cluster = torch.mean((albedo - target_albedo) ** 2) if target_albedo is not None else zero
    return torch.stack([img, chroma, resid, reflect, shade, far, intens, cluster])


# ----------------------------------------------------------------------------------
# Synthetic weights
# ----------------------------------------------------------------------------------
def param_shapes(variant="object", n_classes=0):
    """Ordered (name, shape) list in the reference's construction order
    (run_nerf_helpers.py:259-279 / semantic_nerf.py:96-118) - the order matters because
    the seeded default init consumes the RNG stream layer by layer."""
    shapes = [("pts_linears.0", (256, 63))]
    for i in range(1, 8):
        shapes.append((f"pts_linears.{i}", (256, 319 if i == 5 else 256)))
    shapes.append(("views_linears.0", (128, 283)))
    if variant == "object":
        shapes += [("feature_linear", (256, 256)), ("alpha_linear", (1, 256)), ("shading_linear", (3, 128)),
                   ("albedo_linear1", (128, 256)), ("albedo_linear2", (3, 128)),
                   ("test_linear1", (128, 256)), ("test_linear2", (1, 128))]
    else:
        shapes += [("feature_linear", (256, 256)), ("alpha_linear", (1, 256))]
        if n_classes > 0:
            shapes += [("semantic_linear.0.0", (128, 256)), ("semantic_linear.1", (n_classes, 128))]
        shapes += [("residual_linear", (3, 128)), ("albedo_linear1", (128, 256)), ("albedo_linear2", (3, 128)),
                   ("shading_linear1", (128, 256)), ("shading_linear2", (1, 128))]
    return shapes


def init_params(variant="object", n_classes=0, generator=None):
    """Default nn.Linear init (kaiming_uniform(a=sqrt(5)) weight, U(-1/sqrt(fan_in),..) bias)
    drawn layer by layer in construction order, i.e. what ``NeRF(...)`` /
    ``Semantic_NeRF(...)`` produce after ``torch.manual_seed(seed)``."""
    p = {}
    for name, (n_out, n_in) in param_shapes(variant, n_classes):
        lin = torch.nn.Linear(n_in, n_out)   # consumes the global RNG like the reference ctor
        p[name + ".weight"] = lin.weight.detach().clone()
        p[name + ".bias"] = lin.bias.detach().clone()
    return p


def make_opaque(params):
    """'Opaque' well-conditioned regime of SURVEY section 8d: alpha bias +1, last trunk x3."""
    q = {k: v.clone() for k, v in params.items()}
    q["alpha_linear.bias"] = q["alpha_linear.bias"] + 1.0
    q["pts_linears.7.weight"] = q["pts_linears.7.weight"] * 3.0
    return q


def seeded_nets(variant="object", n_classes=0, seed=20220414, opaque=True):
    torch.manual_seed(seed)                       # run_nerf.py:1130
    coarse = init_params(variant, n_classes)      # coarse first, then fine (run_nerf.py:286-295)
    fine = init_params(variant, n_classes)
    if opaque:
        coarse, fine = make_opaque(coarse), make_opaque(fine)
    return coarse, fine
