// Fused intrinsic-decomposition losses of one training step and their gradient
// (object_level/run_nerf_helpers.py:11, 15-86 / SSR/training/training_utils.py:124-207; composed in
// run_nerf.py:975-1017, trainer.py:913-990).  The reference evaluates img2mse, compute_intrinsic_loss
// (chroma, residual, reflectance sparsity, shading smoothness, far reflectance, intensity) and the cluster
// term as ~40 elementwise / reduction launches on 1-4 k rays, plus as many again in autograd; here the
// forward is ONE launch producing the eight scalars and the backward ONE launch producing the gradient
// of any weighted sum of them with respect to the rendered maps.
//
// Ray pairing (compute_intrinsic_loss): split = N/2, pair i <-> N-split+i (a sampled pixel and its
// neighbour); "far" pairs inside the first half: split2 = split/2, i <-> split-split2+i.
// Pair weights depend on the ground-truth colours and masks/labels only (no gradient):
//   object fork: w = exp(-60 dc) m1 m2, w_inv = dc m1 m2;  SSR fork: w = exp(-60 dc) [l1 == l2], w_inv = dc
// with dc = (r1-r2)^2 + (g1-g2)^2 on chromaticities r = R/(R+G+B+1e-5), g = G/(R+G+B+1e-5).
// (compute_depth_weight is evaluated by the reference and then replaced by the constant 1.)
#include "common.cuh"

namespace inrf {

constexpr int LOSS_THREADS = 1024;
enum { L_IMG = 0, L_CHROMA, L_RESID, L_REFLECT, L_SHADE, L_FAR, L_INTENS, L_CLUSTER, L_NTERMS };

struct LossArgs {
  const float *rgb, *albedo, *shading, *residual;     // rendered maps (row strides below; rgb may be null)
  int ld_rgb, ld_alb, ld_sh, ld_res;
  const float* gt;            // [N,3]
  const float* label;         // [N] object mask (mode 0) or semantic label (mode 1)
  const float* target;        // [N,3] cluster target albedo or null
  int64_t N;
  int mode;
};

__device__ __forceinline__ void chroma(const float* c, float& r, float& g, float& s) {
  s = __fadd_rn(__fadd_rn(__fadd_rn(c[0], c[1]), c[2]), 1e-5f);
  r = __fdiv_rn(c[0], s);
  g = __fdiv_rn(c[1], s);
}
__device__ __forceinline__ void pair_weight(const LossArgs& a, int64_t i, int64_t j, float& w, float& w_inv) {
  float r1, g1, s1, r2, g2, s2;
  chroma(a.gt + 3 * i, r1, g1, s1);
  chroma(a.gt + 3 * j, r2, g2, s2);
  const float dr = r1 - r2, dg = g1 - g2;
  const float dc = dr * dr + dg * dg;
  const float l1 = a.label[i], l2 = a.label[j];
  if (a.mode == 0) { const float m = l1 * l2; w = expf(-60.f * dc) * m; w_inv = dc * m; }
  else { w = expf(-60.f * dc) * (l1 == l2 ? 1.f : 0.f); w_inv = dc; }
}

// block-wide sum of NV doubles per thread -> every thread gets the totals
template <int NV>
__device__ void block_sum(double* v, double* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    double x = v[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if (lane == 0) scratch[k * 32 + warp] = x;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    double x = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) x += scratch[k * 32 + w];   // fixed order: deterministic
    v[k] = x;
  }
  __syncthreads();
}

// sums needed by both passes: [0] sum(gt) [1] sum(albedo)
__device__ void mean_sums(const LossArgs& a, double* scratch, double& sum_gt, double& sum_alb) {
  double v[2] = {0.0, 0.0};
  for (int64_t i = threadIdx.x; i < a.N; i += blockDim.x) {
    const float* g = a.gt + 3 * i;
    const float* al = a.albedo + (int64_t)a.ld_alb * i;
    v[0] += (double)g[0] + (double)g[1] + (double)g[2];
    v[1] += (double)al[0] + (double)al[1] + (double)al[2];
  }
  block_sum<2>(v, scratch);
  sum_gt = v[0];
  sum_alb = v[1];
}

__global__ void __launch_bounds__(LOSS_THREADS) k_intrinsic_loss_fwd(LossArgs a, float* __restrict__ losses) {
  __shared__ double scratch[8 * 32];
  const int64_t N = a.N, split = N / 2, split2 = split / 2;
  double sum_gt, sum_alb;
  mean_sums(a, scratch, sum_gt, sum_alb);
  double v[7] = {0, 0, 0, 0, 0, 0, 0};     // img, chroma, resid, reflect, shade, far, cluster
  for (int64_t i = threadIdx.x; i < N; i += blockDim.x) {
    const float* al = a.albedo + (int64_t)a.ld_alb * i;
    const float* g = a.gt + 3 * i;
    if (a.rgb) {
      const float* c = a.rgb + (int64_t)a.ld_rgb * i;
#pragma unroll
      for (int k = 0; k < 3; ++k) { const float d = c[k] - g[k]; v[0] += (double)(d * d); }
    }
    float r1, g1, s1, r2, g2, s2;
    chroma(al, r1, g1, s1);
    chroma(g, r2, g2, s2);
    v[1] += (double)((r1 - r2) * (r1 - r2)) + (double)((g1 - g2) * (g1 - g2));
    const float* rs = a.residual + (int64_t)a.ld_res * i;
#pragma unroll
    for (int k = 0; k < 3; ++k) v[2] += (double)(rs[k] * rs[k]);
    if (a.target) {
#pragma unroll
      for (int k = 0; k < 3; ++k) { const float d = al[k] - a.target[3 * i + k]; v[6] += (double)(d * d); }
    }
    if (i < split) {
      const int64_t j = N - split + i;
      float w, w_inv;
      pair_weight(a, i, j, w, w_inv);
      const float* al2 = a.albedo + (int64_t)a.ld_alb * j;
      float n2 = 0.f;
#pragma unroll
      for (int k = 0; k < 3; ++k) { const float d = al[k] - al2[k]; n2 += d * d; }
      v[3] += (double)(w * n2);
      const float ds = a.shading[(int64_t)a.ld_sh * i] - a.shading[(int64_t)a.ld_sh * j];
      v[4] += (double)(w_inv * ds * ds);
    }
    if (i < split2) {
      const int64_t j = split - split2 + i;
      float w, w_inv;
      pair_weight(a, i, j, w, w_inv);
      const float* al2 = a.albedo + (int64_t)a.ld_alb * j;
      float n2 = 0.f;
#pragma unroll
      for (int k = 0; k < 3; ++k) { const float d = al[k] - al2[k]; n2 += d * d; }
      v[5] += (double)(w * n2);
    }
  }
  block_sum<7>(v, scratch);
  if (threadIdx.x == 0) {
    const double n3 = 3.0 * (double)N;
    losses[L_IMG] = a.rgb ? (float)(v[0] / n3) : 0.f;
    losses[L_CHROMA] = (float)(v[1] / (double)N);      // mean((r1-r2)^2) + mean((g1-g2)^2)
    losses[L_RESID] = (float)(v[2] / n3);
    losses[L_REFLECT] = split > 0 ? (float)(v[3] / (double)split) : __int_as_float(0x7fc00000);    // torch.mean of an empty tensor is NaN
    losses[L_SHADE] = split > 0 ? (float)(v[4] / (double)split) : __int_as_float(0x7fc00000);
    losses[L_FAR] = split2 > 0 ? (float)(v[5] / (double)split2) : __int_as_float(0x7fc00000);
    const double dm = sum_gt / n3 - sum_alb / n3;
    losses[L_INTENS] = (float)(dm * dm);
    losses[L_CLUSTER] = a.target ? (float)(v[6] / n3) : 0.f;
  }
}

struct LossGrads {
  float *g_rgb, *g_alb, *g_sh, *g_res;    // written (not accumulated); same strides as the inputs; g_rgb may be null
};

// gradient of sum_k w[k] * loss_k; one thread owns whole rows, pair terms are gathered from both partners
__global__ void __launch_bounds__(LOSS_THREADS) k_intrinsic_loss_bwd(LossArgs a, const float* __restrict__ w, LossGrads G) {
  __shared__ double scratch[2 * 32];
  const int64_t N = a.N, split = N / 2, split2 = split / 2;
  double sum_gt, sum_alb;
  mean_sums(a, scratch, sum_gt, sum_alb);
  const float n3 = 3.f * (float)N;
  const float k_int = -2.f * (float)(sum_gt / (3.0 * (double)N) - sum_alb / (3.0 * (double)N)) / n3 * w[L_INTENS];
  for (int64_t i = threadIdx.x; i < N; i += blockDim.x) {
    const float* al = a.albedo + (int64_t)a.ld_alb * i;
    const float* g = a.gt + 3 * i;
    if (G.g_rgb) {
      const float* c = a.rgb + (int64_t)a.ld_rgb * i;
#pragma unroll
      for (int k = 0; k < 3; ++k) G.g_rgb[(int64_t)a.ld_rgb * i + k] = w[L_IMG] * 2.f * (c[k] - g[k]) / n3;
    }
    const float* rs = a.residual + (int64_t)a.ld_res * i;
#pragma unroll
    for (int k = 0; k < 3; ++k) G.g_res[(int64_t)a.ld_res * i + k] = w[L_RESID] * 2.f * rs[k] / n3;
    // albedo: chroma + intensity + cluster
    float ga[3];
    {
      float r1, g1, s1, r2, g2, s2;
      chroma(al, r1, g1, s1);
      chroma(g, r2, g2, s2);
      const float er = 2.f * (r1 - r2) / (float)N, eg = 2.f * (g1 - g2) / (float)N;
      const float inv_s = 1.f / s1;
      // dr/da = (delta_0 - r)/s, dg/da = (delta_1 - g)/s
      ga[0] = w[L_CHROMA] * (er * (1.f - r1) - eg * g1) * inv_s;
      ga[1] = w[L_CHROMA] * (-er * r1 + eg * (1.f - g1)) * inv_s;
      ga[2] = w[L_CHROMA] * (-er * r1 - eg * g1) * inv_s;
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      ga[k] += k_int;
      if (a.target) ga[k] += w[L_CLUSTER] * 2.f * (al[k] - a.target[3 * i + k]) / n3;
    }
    float gs = 0.f;
    // near pair, this row as first partner (i < split) and/or as second partner (i >= N - split)
    if (split > 0) {
      if (i < split) {
        const int64_t j = N - split + i;
        float pw, pw_inv;
        pair_weight(a, i, j, pw, pw_inv);
        const float* al2 = a.albedo + (int64_t)a.ld_alb * j;
#pragma unroll
        for (int k = 0; k < 3; ++k) ga[k] += w[L_REFLECT] * 2.f * pw * (al[k] - al2[k]) / (float)split;
        gs += w[L_SHADE] * 2.f * pw_inv * (a.shading[(int64_t)a.ld_sh * i] - a.shading[(int64_t)a.ld_sh * j]) / (float)split;
      }
      if (i >= N - split) {
        const int64_t j = i - (N - split);
        float pw, pw_inv;
        pair_weight(a, j, i, pw, pw_inv);
        const float* al1 = a.albedo + (int64_t)a.ld_alb * j;
#pragma unroll
        for (int k = 0; k < 3; ++k) ga[k] -= w[L_REFLECT] * 2.f * pw * (al1[k] - al[k]) / (float)split;
        gs -= w[L_SHADE] * 2.f * pw_inv * (a.shading[(int64_t)a.ld_sh * j] - a.shading[(int64_t)a.ld_sh * i]) / (float)split;
      }
    }
    if (split2 > 0) {
      if (i < split2) {
        const int64_t j = split - split2 + i;
        float pw, pw_inv;
        pair_weight(a, i, j, pw, pw_inv);
        const float* al2 = a.albedo + (int64_t)a.ld_alb * j;
#pragma unroll
        for (int k = 0; k < 3; ++k) ga[k] += w[L_FAR] * 2.f * pw * (al[k] - al2[k]) / (float)split2;
      }
      if (i >= split - split2 && i < split) {
        const int64_t j = i - (split - split2);
        float pw, pw_inv;
        pair_weight(a, j, i, pw, pw_inv);
        const float* al1 = a.albedo + (int64_t)a.ld_alb * j;
#pragma unroll
        for (int k = 0; k < 3; ++k) ga[k] -= w[L_FAR] * 2.f * pw * (al1[k] - al[k]) / (float)split2;
      }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) G.g_alb[(int64_t)a.ld_alb * i + k] = ga[k];
    G.g_sh[(int64_t)a.ld_sh * i] = gs;
  }
}

static int check_loss_args(const LossArgs& a) {
  if (a.N < 0) { set_error("negative ray count"); return INRF_EINVAL; }
  if (a.N > 0 && (!a.albedo || !a.shading || !a.residual || !a.gt || !a.label)) { set_error("intrinsic loss: null pointer"); return INRF_EINVAL; }
  if (a.mode != 0 && a.mode != 1) { set_error("intrinsic loss: mode must be 0 (object masks) or 1 (semantic labels)"); return INRF_EINVAL; }
  if (a.ld_alb < 3 || a.ld_res < 3 || a.ld_sh < 1 || (a.rgb && a.ld_rgb < 3)) { set_error("intrinsic loss: row stride too small"); return INRF_EINVAL; }
  return INRF_OK;
}

int launch_intrinsic_loss_fwd(const LossArgs& a, float* losses, cudaStream_t st) {
  int rc = check_loss_args(a);
  if (rc) return rc;
  if (!losses) { set_error("intrinsic loss: null output"); return INRF_EINVAL; }
  k_intrinsic_loss_fwd<<<1, LOSS_THREADS, 0, st>>>(a, losses);
  INRF_LAUNCH_CHECK();
  return INRF_OK;
}

int launch_intrinsic_loss_bwd(const LossArgs& a, const float* w, const LossGrads& G, cudaStream_t st) {
  int rc = check_loss_args(a);
  if (rc) return rc;
  if (a.N == 0) return INRF_OK;
  if (!w || !G.g_alb || !G.g_sh || !G.g_res || (G.g_rgb && !a.rgb)) { set_error("intrinsic loss backward: null pointer"); return INRF_EINVAL; }
  k_intrinsic_loss_bwd<<<1, LOSS_THREADS, 0, st>>>(a, w, G);
  INRF_LAUNCH_CHECK();
  return INRF_OK;
}

}  // namespace inrf

extern "C" {

int inrf_intrinsic_loss_fwd(const float* rgb, int ld_rgb, const float* albedo, int ld_alb, const float* shading, int ld_sh,
                            const float* residual, int ld_res, const float* gt_rgb, const float* label,
                            const float* target_albedo, int64_t N, int mode, float* losses, void* stream) {
  inrf::LossArgs a{rgb, albedo, shading, residual, ld_rgb, ld_alb, ld_sh, ld_res, gt_rgb, label, target_albedo, N, mode};
  return inrf::launch_intrinsic_loss_fwd(a, losses, (cudaStream_t)stream);
}

int inrf_intrinsic_loss_bwd(const float* rgb, int ld_rgb, const float* albedo, int ld_alb, const float* shading, int ld_sh,
                            const float* residual, int ld_res, const float* gt_rgb, const float* label,
                            const float* target_albedo, int64_t N, int mode, const float* weights, float* g_rgb,
                            float* g_albedo, float* g_shading, float* g_residual, void* stream) {
  inrf::LossArgs a{rgb, albedo, shading, residual, ld_rgb, ld_alb, ld_sh, ld_res, gt_rgb, label, target_albedo, N, mode};
  inrf::LossGrads G{g_rgb, g_albedo, g_shading, g_residual};
  return inrf::launch_intrinsic_loss_bwd(a, weights, G, (cudaStream_t)stream);
}

}  // extern "C"
