// extern "C" surface of libinrf.so (see include/inrf.h) and the render_rays orchestration.
#include "common.cuh"

namespace inrf {
const char* last_error();
long long launch_count();
int pack_weights(const float* flat, int variant, int n_classes, void* packed, int64_t packed_bytes, cudaStream_t st);
int launch_embed(const float* x, int64_t M, int L, float scale, float* out, cudaStream_t st);
int launch_coarse_z(const float* rays, const float* t_vals, const float* t_rand, int64_t N, int S, int lindisp, float* z, cudaStream_t st,
                    Rng rng = Rng{});
int launch_raw2outputs(const float* raw, const float* z, const float* rays_d, int ld_d, const float* noise, int64_t N,
                       int S, int n_classes, int endpoint, int white_bkgd, float* rec, float* weights, cudaStream_t st, Rng rng = Rng{});
int launch_raw2outputs_bwd(const float* raw, const float* z, const float* rays_d, int ld_d, const float* noise, int64_t N,
                           int S, int n_classes, int endpoint, int white_bkgd, const float* grec, const float* gweights,
                           float* graw, cudaStream_t st, Rng rng = Rng{});
int launch_sample_pdf(const float* bins, const float* weights, int ld_w, const float* cdf_in, const float* u,
                      const float* u_det, int64_t N, int B, int n_samples, float* samples, int64_t* inds, float* cdf_out, cudaStream_t st,
                      Rng rng = Rng{});
int launch_merge_sorted(const float* za, const float* zb, int64_t N, int Sa, int Sb, float* zout, float* zstd, cudaStream_t st);
int launch_zmid(const float* z, int64_t N, int S, float* zmid, cudaStream_t st);
int launch_get_rays(int H, int W, float fx, float fy, float cx, float cy, const float* c2w, int opencv, int euclidean,
                    float nearv, float farv, const int64_t* pix, int64_t n, float* rays, cudaStream_t st);

static int run_mlp(const MlpArgs& a, int precision, cudaStream_t st) {
  if (precision == INRF_PREC_FP32) return launch_mlp_fp32(a, st);
  if (precision == INRF_PREC_TC) return launch_mlp_tc(a, st);
  set_error("unknown precision %d", precision);
  return INRF_EINVAL;
}

static inline int64_t align256(int64_t x) { return (x + 255) / 256 * 256; }

struct WsPlan { int64_t z_c, w_c, zmid, zs, z_f, raw_c, raw_f, ring, total; };

// Fused renderer (two launches per chunk, no raw tensor): tensor-core precision, 32 | S for both passes, no endpoint
// feature, and - when there is a fine pass - the reference's 64 + 128 samples with deterministic u.  Everything else
// (caller wants raw / weights, random u, strict fp32) takes the stage-by-stage path.  INRF_NO_FUSE=1 forces that path.
static bool can_fuse(const InrfRenderCfg& c, bool wants_raw_or_weights, bool random_u) {
  static const bool off = getenv("INRF_NO_FUSE") != nullptr && getenv("INRF_NO_FUSE")[0] == '1';
  if (off || c.precision != INRF_PREC_TC || c.endpoint_feat || wants_raw_or_weights) return false;
  if ((c.n_samples & 31) || ((c.n_samples + c.n_importance) & 31)) return false;
  if (c.n_importance > 0 && (random_u || c.n_samples != 64 || c.n_importance != 128)) return false;
  return true;
}

static WsPlan plan_ws_fused(const InrfRenderCfg& c, int64_t N, bool own_z_f) {
  WsPlan p{};
  int64_t o = 0;
  p.ring = o; o = align256(o + mlp_tc_ring_bytes(c.n_classes));
  p.z_f = o; if (own_z_f && c.n_importance > 0) o = align256(o + N * (c.n_samples + c.n_importance) * 4);
  p.total = o;
  return p;
}

static WsPlan plan_ws(const InrfRenderCfg& c, int64_t N, bool own_raw_c, bool own_raw_f, bool own_z_f) {
  WsPlan p{};
  int64_t o = 0;
  const int Sc = c.n_samples, Sf = c.n_importance, St = Sc + Sf;
  p.z_c = o; o = align256(o + N * Sc * 4);
  p.w_c = o; o = align256(o + N * Sc * 4);
  p.zmid = o; o = align256(o + N * (Sc - 1) * 4);
  p.zs = o; o = align256(o + N * (Sf > 0 ? Sf : 1) * 4);
  p.z_f = o; if (own_z_f) o = align256(o + N * St * 4);
  p.raw_c = o; if (own_raw_c) o = align256(o + N * Sc * (int64_t)raw_channels(c.n_classes, 0) * 4);
  p.raw_f = o; if (own_raw_f && Sf > 0) o = align256(o + N * St * (int64_t)raw_channels(c.n_classes, c.endpoint_feat) * 4);
  p.total = o;
  return p;
}

static int check_cfg(const InrfRenderCfg* c) {
  if (c == nullptr) { set_error("render cfg is null"); return INRF_EINVAL; }
  if (c->variant != INRF_NET_OBJECT && c->variant != INRF_NET_SSR) { set_error("unknown variant %d", c->variant); return INRF_EINVAL; }
  if (c->n_samples < 2 || c->n_samples > 256) { set_error("n_samples %d outside [2,256]", c->n_samples); return INRF_EUNSUPPORTED; }
  if (c->n_importance < 0 || c->n_samples + c->n_importance > 1024) { set_error("n_samples+n_importance > 1024"); return INRF_EUNSUPPORTED; }
  if (c->variant == INRF_NET_OBJECT && (c->n_classes != 0 || c->endpoint_feat)) { set_error("object network has no semantic/endpoint outputs"); return INRF_EINVAL; }
  if (c->n_classes < 0 || c->n_classes > MAX_CLASSES) { set_error("n_classes %d outside [0,%d]", c->n_classes, MAX_CLASSES); return INRF_EUNSUPPORTED; }
  if (c->variant == INRF_NET_SSR && c->lindisp) { set_error("SSR renderer samples linearly in depth only (trainer.py:732)"); return INRF_EUNSUPPORTED; }
  if (!(c->pe_scalar_factor > 0.f)) { set_error("pe_scalar_factor must be positive"); return INRF_EINVAL; }
  return INRF_OK;
}

// The fused renderer: coarse launch (depths generated in the front end, rows composited by the back-end warps, which
// also resample and write the merged depths of the fine pass), then the fine launch, composited the same way.  No raw
// tensor exists.  `a` carries the ray addressing (table or camera).
static int render_fused(MlpArgs a, int64_t N, const void* packed_coarse, const void* packed_fine, const InrfRenderCfg& c,
                        const float* t_vals, const float* u_det, const float* t_rand, const float* noise_coarse,
                        const float* noise_fine, float* rec_coarse, float* rec_fine, float* z_std, float* z_f, float* ring,
                        cudaStream_t st) {
  const int Sc = c.n_samples, Sf = c.n_importance, St = Sc + Sf;
  a.packed = packed_coarse; a.variant = c.variant; a.n_classes = c.n_classes; a.endpoint = 0;
  a.pe_scale = c.pe_scalar_factor; a.z = nullptr; a.S = Sc; a.M = N * Sc; a.raw = nullptr;
  FuseArgs f{};
  f.white_bkgd = c.white_bkgd; f.lindisp = c.lindisp; f.t_vals = t_vals; f.t_rand = t_rand; f.noise = noise_coarse;
  f.rec = rec_coarse; f.n_importance = Sf; f.u_det = u_det; f.z_out = Sf > 0 ? z_f : nullptr; f.z_std = z_std; f.ring = ring;
  int rc = launch_mlp_tc(a, st, &f);
  if (rc || Sf == 0) return rc;
  a.packed = packed_fine ? packed_fine : packed_coarse;
  a.z = z_f; a.S = St; a.M = N * St;
  f.t_vals = nullptr; f.t_rand = nullptr; f.noise = noise_fine; f.rec = rec_fine; f.n_importance = 0; f.z_out = nullptr; f.z_std = nullptr;
  return launch_mlp_tc(a, st, &f);
}

}  // namespace inrf

using namespace inrf;

extern "C" {

const char* inrf_last_error_string(void) { return last_error(); }
int inrf_version(void) { return 200; }
int inrf_poll_status(void) { return status_poll(); }
int64_t inrf_launch_count(void) { return launch_count(); }
int inrf_rng_epoch_bump(void* stream) { return rng_epoch_set(1, (cudaStream_t)stream); }
int inrf_rng_epoch_reset(void* stream) { return rng_epoch_set(0, (cudaStream_t)stream); }
#define INRF_POLL() do { int rc__ = status_poll(); if (rc__) return rc__; } while (0)

int64_t inrf_flat_param_count(int variant, int n_classes) {
  NetLayout L;
  int rc = make_layout(variant, n_classes, &L);
  return rc ? rc : L.flat_count;
}

int64_t inrf_packed_bytes(int variant, int n_classes) {
  NetLayout L;
  int rc = make_layout(variant, n_classes, &L);
  return rc ? rc : L.total_bytes;
}

int inrf_pack_weights(const float* flat_params, int variant, int n_classes, void* packed, int64_t packed_bytes, void* stream) {
  INRF_POLL();
  INRF_CHECK_ARG(flat_params && packed, "null pointer");
  return pack_weights(flat_params, variant, n_classes, packed, packed_bytes, (cudaStream_t)stream);
}

int inrf_embed(const float* x, int64_t M, int n_freqs, float scalar_factor, float* out, void* stream) {
  INRF_CHECK_ARG(M >= 0 && (M == 0 || (x && out)), "null pointer / negative size");
  INRF_CHECK_SUPPORTED(n_freqs >= 0 && n_freqs <= 16, "n_freqs outside [0,16]");
  INRF_CHECK_ARG(scalar_factor > 0.f, "scalar_factor must be positive");
  return launch_embed(x, M, n_freqs, scalar_factor, out, (cudaStream_t)stream);
}

int inrf_mlp_fwd(const void* packed, int variant, int n_classes, int endpoint_feat, float pe_scalar_factor,
                 const float* pts, const float* viewdirs, int64_t M, float* raw, int precision, void* stream) {
  INRF_POLL();
  INRF_CHECK_ARG(M >= 0 && packed && (M == 0 || (pts && viewdirs && raw)), "null pointer / negative size");
  INRF_CHECK_ARG(pe_scalar_factor > 0.f, "pe_scalar_factor must be positive");
  INRF_CHECK_ARG(!(variant == INRF_NET_OBJECT && endpoint_feat), "object network has no endpoint feature");
  MlpArgs a{};
  a.packed = packed; a.variant = variant; a.n_classes = n_classes; a.endpoint = endpoint_feat ? 1 : 0;
  a.pe_scale = pe_scalar_factor; a.pts = pts; a.viewdirs = viewdirs; a.M = M; a.raw = raw; a.S = 1;
  return run_mlp(a, precision, (cudaStream_t)stream);
}

int inrf_mlp_fwd_embedded(const void* packed, int variant, int n_classes, int endpoint_feat, const float* emb,
                          int64_t M, float* raw, int precision, void* stream) {
  INRF_POLL();
  INRF_CHECK_ARG(M >= 0 && packed && (M == 0 || (emb && raw)), "null pointer / negative size");
  INRF_CHECK_ARG(!(variant == INRF_NET_OBJECT && endpoint_feat), "object network has no endpoint feature");
  MlpArgs a{};
  a.packed = packed; a.variant = variant; a.n_classes = n_classes; a.endpoint = endpoint_feat ? 1 : 0;
  a.pe_scale = 1.f; a.emb = emb; a.M = M; a.raw = raw; a.S = 1;
  return run_mlp(a, precision, (cudaStream_t)stream);
}

int64_t inrf_stash_floats_per_row(void) { return STASH_LD; }

static int fill_addressing(MlpArgs& a, const float* pts, const float* viewdirs, const float* rays, const float* z, int S,
                           const float* emb) {
  const int modes = (pts != nullptr) + (rays != nullptr) + (emb != nullptr);
  if (modes != 1) { set_error("exactly one of (pts,viewdirs) / (rays,z) / emb must be given"); return INRF_EINVAL; }
  if (pts && !viewdirs) { set_error("viewdirs missing"); return INRF_EINVAL; }
  if (rays && (!z || S <= 0)) { set_error("z / S missing"); return INRF_EINVAL; }
  a.pts = pts; a.viewdirs = viewdirs; a.rays = rays; a.z = z; a.S = rays ? S : 1; a.emb = emb;
  return INRF_OK;
}

int inrf_mlp_fwd_train(const void* packed, int variant, int n_classes, int endpoint_feat, float pe_scalar_factor,
                       const float* pts, const float* viewdirs, const float* rays, const float* z, int S, const float* emb,
                       int64_t M, float* raw, float* stash, void* stream) {
  INRF_CHECK_ARG(M >= 0 && packed && (M == 0 || (raw && stash)), "null pointer / negative size");
  INRF_CHECK_ARG(pe_scalar_factor > 0.f, "pe_scalar_factor must be positive");
  INRF_CHECK_ARG(!(variant == INRF_NET_OBJECT && endpoint_feat), "object network has no endpoint feature");
  if (M == 0) return INRF_OK;
  MlpArgs a{};
  int rc = fill_addressing(a, pts, viewdirs, rays, z, S, emb);
  if (rc) return rc;
  a.packed = packed; a.variant = variant; a.n_classes = n_classes; a.endpoint = endpoint_feat ? 1 : 0;
  a.pe_scale = pe_scalar_factor; a.M = M; a.raw = raw; a.stash = stash;
  return launch_mlp_fp32(a, (cudaStream_t)stream);
}

int inrf_mlp_bwd(const float* flat_params, int variant, int n_classes, int endpoint_feat, float pe_scalar_factor,
                 const float* pts, const float* viewdirs, const float* rays, const float* z, int S, const float* emb,
                 int64_t M, const float* raw, const float* stash, const float* grad_raw, float* grad_flat, void* stream) {
  INRF_CHECK_ARG(M >= 0 && flat_params && grad_flat && (M == 0 || (raw && stash && grad_raw)), "null pointer / negative size");
  INRF_CHECK_ARG(pe_scalar_factor > 0.f, "pe_scalar_factor must be positive");
  INRF_CHECK_ARG(!(variant == INRF_NET_OBJECT && endpoint_feat), "object network has no endpoint feature");
  if (M == 0) return INRF_OK;
  MlpBwdArgs b{};
  int rc = fill_addressing(b.f, pts, viewdirs, rays, z, S, emb);
  if (rc) return rc;
  b.f.variant = variant; b.f.n_classes = n_classes; b.f.endpoint = endpoint_feat ? 1 : 0; b.f.pe_scale = pe_scalar_factor;
  b.f.M = M; b.f.raw = const_cast<float*>(raw);
  b.flat = flat_params; b.stash = stash; b.grad_raw = grad_raw; b.grad_flat = grad_flat;
  return launch_mlp_bwd_fp32(b, (cudaStream_t)stream);
}

int64_t inrf_mlp_stash_img_bytes(int64_t M) { return M <= 0 ? 0 : (M + 127) / 128 * IMG_STASH_SLOTS * (int64_t)IMG_BYTES; }
int64_t inrf_mlp_bwd_tc_workspace_bytes(int variant, int n_classes, int64_t M) { return tc_bwd_workspace_bytes(variant, n_classes, M < 0 ? 0 : M); }

int inrf_mlp_fwd_train_tc(const void* packed, int variant, int n_classes, int endpoint_feat, float pe_scalar_factor,
                          const float* pts, const float* viewdirs, const float* rays, const float* z, int S, const float* emb,
                          int64_t M, float* raw, void* stash_img, void* stream) {
  INRF_POLL();
  INRF_CHECK_ARG(M >= 0 && packed && (M == 0 || (raw && stash_img)), "null pointer / negative size");
  INRF_CHECK_ARG(pe_scalar_factor > 0.f, "pe_scalar_factor must be positive");
  INRF_CHECK_ARG(!(variant == INRF_NET_OBJECT && endpoint_feat), "object network has no endpoint feature");
  if (M == 0) return INRF_OK;
  MlpArgs a{};
  int rc = fill_addressing(a, pts, viewdirs, rays, z, S, emb);
  if (rc) return rc;
  a.packed = packed; a.variant = variant; a.n_classes = n_classes; a.endpoint = endpoint_feat ? 1 : 0;
  a.pe_scale = pe_scalar_factor; a.M = M; a.raw = raw; a.stash_img = static_cast<unsigned char*>(stash_img);
  return launch_mlp_tc(a, (cudaStream_t)stream);
}

int inrf_mlp_bwd_tc(const void* packed, const float* flat_params, int variant, int n_classes, int endpoint_feat, int64_t M,
                    const float* raw, const void* stash_img, const float* grad_raw, void* workspace, int64_t workspace_bytes,
                    float* grad_flat, void* stream) {
  INRF_POLL();
  INRF_CHECK_ARG(M >= 0 && packed && flat_params && grad_flat && (M == 0 || (raw && stash_img && grad_raw && workspace)),
                 "null pointer / negative size");
  INRF_CHECK_ARG(!(variant == INRF_NET_OBJECT && endpoint_feat), "object network has no endpoint feature");
  if (M == 0) return INRF_OK;
  if (workspace_bytes < tc_bwd_workspace_bytes(variant, n_classes, M)) { set_error("inrf_mlp_bwd_tc: workspace too small"); return INRF_EWORKSPACE; }
  INRF_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 255) == 0 && (reinterpret_cast<uintptr_t>(stash_img) & 255) == 0,
                 "stash / workspace must be 256-byte aligned");      // bulk copies need 16 B; the images carry their own swizzle
  TcBwdArgs b{};
  b.packed = packed; b.flat = flat_params; b.variant = variant; b.n_classes = n_classes; b.endpoint = endpoint_feat ? 1 : 0;
  b.M = M; b.raw = raw; b.grad_raw = grad_raw; b.stash_img = static_cast<const unsigned char*>(stash_img);
  b.work = static_cast<unsigned char*>(workspace); b.grad_flat = grad_flat;
  return launch_mlp_bwd_tc(b, (cudaStream_t)stream);
}

int inrf_mlp_fwd_rays(const void* packed, int variant, int n_classes, int endpoint_feat, float pe_scalar_factor,
                      const float* rays, const float* z, int64_t N, int S, float* raw, int precision, void* stream) {
  INRF_POLL();
  INRF_CHECK_ARG(N >= 0 && S > 0 && packed && (N == 0 || (rays && z && raw)), "null pointer / bad size");
  INRF_CHECK_ARG(pe_scalar_factor > 0.f, "pe_scalar_factor must be positive");
  INRF_CHECK_ARG(!(variant == INRF_NET_OBJECT && endpoint_feat), "object network has no endpoint feature");
  MlpArgs a{};
  a.packed = packed; a.variant = variant; a.n_classes = n_classes; a.endpoint = endpoint_feat ? 1 : 0;
  a.pe_scale = pe_scalar_factor; a.rays = rays; a.z = z; a.S = S; a.M = N * S; a.raw = raw;
  return run_mlp(a, precision, (cudaStream_t)stream);
}

int inrf_raw2outputs(const float* raw, const float* z, const float* rays_d, int ld_rays_d, const float* noise,
                     int64_t N, int S, int n_classes, int endpoint_feat, int white_bkgd, float* rec, float* weights,
                     void* stream) {
  INRF_CHECK_ARG(N >= 0 && S > 0 && (N == 0 || (raw && z && rays_d && rec)), "null pointer / bad size");
  INRF_CHECK_ARG(ld_rays_d >= 3, "ld_rays_d < 3");
  INRF_CHECK_SUPPORTED(n_classes >= 0 && n_classes <= MAX_CLASSES, "n_classes out of range");
  return launch_raw2outputs(raw, z, rays_d, ld_rays_d, noise, N, S, n_classes, endpoint_feat ? 1 : 0, white_bkgd,
                            rec, weights, (cudaStream_t)stream);
}

int inrf_raw2outputs_bwd(const float* raw, const float* z, const float* rays_d, int ld_rays_d, const float* noise,
                         int64_t N, int S, int n_classes, int endpoint_feat, int white_bkgd, const float* grad_rec,
                         const float* grad_weights, float* grad_raw, void* stream) {
  INRF_CHECK_ARG(N >= 0 && S > 0 && (N == 0 || (raw && z && rays_d && grad_rec && grad_raw)), "null pointer / bad size");
  INRF_CHECK_ARG(ld_rays_d >= 3, "ld_rays_d < 3");
  INRF_CHECK_SUPPORTED(n_classes >= 0 && n_classes <= MAX_CLASSES, "n_classes out of range");
  return launch_raw2outputs_bwd(raw, z, rays_d, ld_rays_d, noise, N, S, n_classes, endpoint_feat ? 1 : 0, white_bkgd,
                                grad_rec, grad_weights, grad_raw, (cudaStream_t)stream);
}

int inrf_sample_pdf(const float* bins, const float* weights, int ld_w, const float* u, const float* u_det, int64_t N,
                    int B, int n_samples, float* samples, int64_t* inds, float* cdf_out, void* stream) {
  INRF_CHECK_ARG(N >= 0 && n_samples > 0 && (N == 0 || (bins && weights && samples)), "null pointer / bad size");
  INRF_CHECK_ARG(u != nullptr || u_det != nullptr, "need u or u_det");
  INRF_CHECK_ARG(ld_w >= B - 1, "ld_w smaller than the number of weights per ray");
  return launch_sample_pdf(bins, weights, ld_w, nullptr, u, u_det, N, B, n_samples, samples, inds, cdf_out, (cudaStream_t)stream);
}

int inrf_invert_cdf(const float* bins, const float* cdf, const float* u, int64_t N, int B, int n_samples,
                    float* samples, int64_t* inds, void* stream) {
  INRF_CHECK_ARG(N >= 0 && n_samples > 0 && (N == 0 || (bins && cdf && u && samples)), "null pointer / bad size");
  return launch_sample_pdf(bins, nullptr, 0, cdf, u, nullptr, N, B, n_samples, samples, inds, nullptr, (cudaStream_t)stream);
}

int inrf_merge_sorted(const float* z_a, const float* z_b, int64_t N, int Sa, int Sb, float* z_out, float* z_std, void* stream) {
  INRF_CHECK_ARG(N >= 0 && Sa >= 0 && Sb >= 0 && (N == 0 || (z_out && (Sa == 0 || z_a) && (Sb == 0 || z_b))), "null pointer / bad size");
  INRF_CHECK_ARG(z_std == nullptr || Sb > 0, "z_std needs Sb > 0");
  return launch_merge_sorted(z_a, z_b, N, Sa, Sb, z_out, z_std, (cudaStream_t)stream);
}

int inrf_coarse_z(const float* rays, const float* t_vals, const float* t_rand, int64_t N, int S, int lindisp, float* z, void* stream) {
  INRF_CHECK_ARG(N >= 0 && S > 0 && (N == 0 || (rays && t_vals && z)), "null pointer / bad size");
  return launch_coarse_z(rays, t_vals, t_rand, N, S, lindisp, z, (cudaStream_t)stream);
}

int inrf_get_rays(int H, int W, float fx, float fy, float cx, float cy, const float* c2w_host, float near, float far,
                  float* rays, void* stream) {
  INRF_CHECK_ARG(H > 0 && W > 0 && c2w_host && rays, "null pointer / bad size");
  INRF_CHECK_ARG(fx != 0.f && fy != 0.f, "zero focal length");
  return launch_get_rays(H, W, fx, fy, cx, cy, c2w_host, 0, 0, near, far, nullptr, (int64_t)H * W, rays, (cudaStream_t)stream);
}

int inrf_rays_from_pixels(const int64_t* pix, int64_t N, int H, int W, float fx, float fy, float cx, float cy,
                          const float* c2w_host, int convention, int euclidean, float near, float far, float* rays,
                          void* stream) {
  INRF_CHECK_ARG(H > 0 && W > 0 && c2w_host && (rays || N == 0) && N >= 0, "null pointer / bad size");
  INRF_CHECK_ARG(fx != 0.f && fy != 0.f, "zero focal length");
  INRF_CHECK_ARG(convention == INRF_CAM_OPENGL || convention == INRF_CAM_OPENCV, "unknown camera convention");
  INRF_CHECK_ARG(pix != nullptr || N == (int64_t)H * W, "pix == NULL means the full image: N must be H*W");
  return launch_get_rays(H, W, fx, fy, cx, cy, c2w_host, convention == INRF_CAM_OPENCV, euclidean != 0, near, far, pix, N,
                         rays, (cudaStream_t)stream);
}

int64_t inrf_render_workspace_bytes(const InrfRenderCfg* cfg, int64_t N) {
  int rc = check_cfg(cfg);
  if (rc) return rc;
  if (N < 0) { set_error("negative ray count"); return INRF_EINVAL; }
  // worst case over the paths inrf_render_fwd may take (the stage path with its own raw tensors is the larger one)
  const int64_t a = plan_ws(*cfg, N, true, true, true).total, b = plan_ws_fused(*cfg, N, true).total;
  return (a > b ? a : b) + 256;
}

int inrf_render_fwd(const float* rays, int64_t N, const void* packed_coarse, const void* packed_fine,
                    const InrfRenderCfg* cfg, const float* t_vals, const float* u_det, const float* t_rand,
                    const float* u, const float* noise_coarse, const float* noise_fine, float* rec_coarse,
                    float* rec_fine, float* z_std, float* raw_coarse, float* raw_fine, float* z_fine,
                    float* weights_fine, void* workspace, int64_t workspace_bytes, void* stream) {
  INRF_POLL();
  int rc = check_cfg(cfg);
  if (rc) return rc;
  const InrfRenderCfg& c = *cfg;
  INRF_CHECK_ARG(N >= 0, "negative ray count");
  if (N == 0) return INRF_OK;
  INRF_CHECK_ARG(rays && packed_coarse && t_vals && rec_coarse && workspace, "null pointer");
  const int Sc = c.n_samples, Sf = c.n_importance, St = Sc + Sf;
  if (Sf > 0) INRF_CHECK_ARG(rec_fine && (u || u_det), "fine pass needs rec_fine and u or u_det");
  cudaStream_t st = (cudaStream_t)stream;
  const bool fused = can_fuse(c, raw_coarse != nullptr || raw_fine != nullptr || weights_fine != nullptr, u != nullptr);
  WsPlan p = fused ? plan_ws_fused(c, N, z_fine == nullptr) : plan_ws(c, N, raw_coarse == nullptr, raw_fine == nullptr, z_fine == nullptr);
  uintptr_t base = (reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255);
  int64_t usable = workspace_bytes - (int64_t)(base - reinterpret_cast<uintptr_t>(workspace));
  if (usable < p.total) { set_error("workspace too small: %lld < %lld", (long long)usable, (long long)p.total); return INRF_EWORKSPACE; }
  auto F = [&](int64_t off) { return reinterpret_cast<float*>(base + off); };
  if (fused) {
    MlpArgs a{};
    a.rays = rays;
    return render_fused(a, N, packed_coarse, packed_fine, c, t_vals, u_det, t_rand, noise_coarse, noise_fine, rec_coarse, rec_fine,
                        z_std, z_fine ? z_fine : F(p.z_f), F(p.ring), st);
  }
  float* z_c = F(p.z_c);
  float* w_c = F(p.w_c);
  float* zmid = F(p.zmid);
  float* zs = F(p.zs);
  float* z_f = z_fine ? z_fine : F(p.z_f);
  float* raw_c = raw_coarse ? raw_coarse : F(p.raw_c);
  float* raw_f = raw_fine ? raw_fine : F(p.raw_f);

  // coarse pass: depths -> field -> composite
  if ((rc = launch_coarse_z(rays, t_vals, t_rand, N, Sc, c.lindisp, z_c, st))) return rc;
  MlpArgs a{};
  a.packed = packed_coarse; a.variant = c.variant; a.n_classes = c.n_classes; a.endpoint = 0;
  a.pe_scale = c.pe_scalar_factor; a.rays = rays; a.z = z_c; a.S = Sc; a.M = N * Sc; a.raw = raw_c;
  if ((rc = run_mlp(a, c.precision, st))) return rc;
  if ((rc = launch_raw2outputs(raw_c, z_c, rays + 3, 11, noise_coarse, N, Sc, c.n_classes, 0, c.white_bkgd, rec_coarse, w_c, st))) return rc;
  if (Sf == 0) return INRF_OK;
  // hierarchical resampling on the interval mid-points with weights[1:-1] (run_nerf.py:499-503)
  if ((rc = launch_zmid(z_c, N, Sc, zmid, st))) return rc;
  if ((rc = launch_sample_pdf(zmid, w_c + 1, Sc, nullptr, u, u_det, N, Sc - 1, Sf, zs, nullptr, nullptr, st))) return rc;
  if ((rc = launch_merge_sorted(z_c, zs, N, Sc, Sf, z_f, z_std, st))) return rc;
  // fine pass
  a.packed = packed_fine ? packed_fine : packed_coarse;
  a.endpoint = c.endpoint_feat ? 1 : 0;
  a.z = z_f; a.S = St; a.M = N * St; a.raw = raw_f;
  if ((rc = run_mlp(a, c.precision, st))) return rc;
  return launch_raw2outputs(raw_f, z_f, rays + 3, 11, noise_fine, N, St, c.n_classes, a.endpoint, c.white_bkgd, rec_fine, weights_fine, st);
}

/* ---- training-mode draws generated inside the stage kernels (no generator launch, no [N,S] tensors) -------------------- */
int inrf_coarse_z_rng(const float* rays, const float* t_vals, uint64_t seed, int64_t N, int S, int lindisp, float* z, void* stream) {
  INRF_CHECK_ARG(N >= 0 && S > 0 && (N == 0 || (rays && t_vals && z)), "null pointer / bad size");
  return launch_coarse_z(rays, t_vals, nullptr, N, S, lindisp, z, (cudaStream_t)stream, Rng{seed, RNG_T_RAND, 1.f, 1, rng_epoch_dev()});
}

int inrf_sample_pdf_rng(const float* bins, const float* weights, int ld_w, uint64_t seed, int64_t N, int B, int n_samples,
                        float* samples, void* stream) {
  INRF_CHECK_ARG(N >= 0 && n_samples > 0 && (N == 0 || (bins && weights && samples)), "null pointer / bad size");
  INRF_CHECK_ARG(ld_w >= B - 1, "ld_w smaller than the number of weights per ray");
  return launch_sample_pdf(bins, weights, ld_w, nullptr, nullptr, nullptr, N, B, n_samples, samples, nullptr, nullptr,
                           (cudaStream_t)stream, Rng{seed, RNG_U, 1.f, 1, rng_epoch_dev()});
}

int inrf_raw2outputs_rng(const float* raw, const float* z, const float* rays_d, int ld_rays_d, float noise_std, uint64_t seed,
                         int fine_pass, int64_t N, int S, int n_classes, int endpoint_feat, int white_bkgd, float* rec,
                         float* weights, void* stream) {
  INRF_CHECK_ARG(N >= 0 && S > 0 && (N == 0 || (raw && z && rays_d && rec)), "null pointer / bad size");
  INRF_CHECK_ARG(ld_rays_d >= 3, "ld_rays_d < 3");
  INRF_CHECK_SUPPORTED(n_classes >= 0 && n_classes <= MAX_CLASSES, "n_classes out of range");
  return launch_raw2outputs(raw, z, rays_d, ld_rays_d, nullptr, N, S, n_classes, endpoint_feat ? 1 : 0, white_bkgd, rec, weights,
                            (cudaStream_t)stream, Rng{seed, (unsigned)(fine_pass ? RNG_NOISE_FINE : RNG_NOISE_COARSE), noise_std, noise_std > 0.f, rng_epoch_dev()});
}

int inrf_raw2outputs_bwd_rng(const float* raw, const float* z, const float* rays_d, int ld_rays_d, float noise_std, uint64_t seed,
                             int fine_pass, int64_t N, int S, int n_classes, int endpoint_feat, int white_bkgd,
                             const float* grad_rec, const float* grad_weights, float* grad_raw, void* stream) {
  INRF_CHECK_ARG(N >= 0 && S > 0 && (N == 0 || (raw && z && rays_d && grad_rec && grad_raw)), "null pointer / bad size");
  INRF_CHECK_ARG(ld_rays_d >= 3, "ld_rays_d < 3");
  INRF_CHECK_SUPPORTED(n_classes >= 0 && n_classes <= MAX_CLASSES, "n_classes out of range");
  return launch_raw2outputs_bwd(raw, z, rays_d, ld_rays_d, nullptr, N, S, n_classes, endpoint_feat ? 1 : 0, white_bkgd, grad_rec,
                                grad_weights, grad_raw, (cudaStream_t)stream,
                                Rng{seed, (unsigned)(fine_pass ? RNG_NOISE_FINE : RNG_NOISE_COARSE), noise_std, noise_std > 0.f, rng_epoch_dev()});
}

int inrf_render_fwd_camera(const InrfCamera* cam, int64_t pix0, int64_t N, const void* packed_coarse, const void* packed_fine,
                           const InrfRenderCfg* cfg, const float* t_vals, const float* u_det, float* rec_coarse, float* rec_fine,
                           float* z_std, float* z_fine, void* workspace, int64_t workspace_bytes, void* stream) {
  INRF_POLL();
  int rc = check_cfg(cfg);
  if (rc) return rc;
  const InrfRenderCfg& c = *cfg;
  INRF_CHECK_ARG(cam != nullptr && cam->H > 0 && cam->W > 0 && cam->fx != 0.f && cam->fy != 0.f, "bad camera");
  INRF_CHECK_ARG(cam->convention == INRF_CAM_OPENGL || cam->convention == INRF_CAM_OPENCV, "unknown camera convention");
  INRF_CHECK_ARG(N >= 0 && pix0 >= 0 && pix0 + N <= (int64_t)cam->H * cam->W, "pixel range outside the frame");
  if (N == 0) return INRF_OK;
  INRF_CHECK_ARG(packed_coarse && t_vals && rec_coarse && workspace, "null pointer");
  if (c.n_importance > 0) INRF_CHECK_ARG(rec_fine && u_det, "fine pass needs rec_fine and u_det");
  if (!can_fuse(c, false, false)) {
    set_error("inrf_render_fwd_camera: this configuration is not rendered by the fused kernel (needs INRF_PREC_TC, 32 | samples, "
              "64 + 128 when there is a fine pass, no endpoint feature); generate the rays (inrf_get_rays) and call inrf_render_fwd");
    return INRF_EUNSUPPORTED;
  }
  WsPlan p = plan_ws_fused(c, N, z_fine == nullptr);
  uintptr_t base = (reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255);
  int64_t usable = workspace_bytes - (int64_t)(base - reinterpret_cast<uintptr_t>(workspace));
  if (usable < p.total) { set_error("workspace too small: %lld < %lld", (long long)usable, (long long)p.total); return INRF_EWORKSPACE; }
  MlpArgs a{};
  a.cam_on = 1; a.cam_H = cam->H; a.cam_W = cam->W; a.cam_pix0 = pix0;
  a.cam.fx = cam->fx; a.cam.fy = cam->fy; a.cam.cx = cam->cx; a.cam.cy = cam->cy; a.cam.nearv = cam->near_; a.cam.farv = cam->far_;
  a.cam.opencv = cam->convention == INRF_CAM_OPENCV; a.cam.euclidean = cam->euclidean != 0;
  for (int i = 0; i < 12; ++i) a.cam.m[i] = cam->c2w[i];
  return render_fused(a, N, packed_coarse, packed_fine, c, t_vals, u_det, nullptr, nullptr, nullptr, rec_coarse, rec_fine, z_std,
                      z_fine ? z_fine : reinterpret_cast<float*>(base + p.z_f), reinterpret_cast<float*>(base + p.ring),
                      (cudaStream_t)stream);
}

}  // extern "C"
