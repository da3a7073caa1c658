// Per-ray stage kernels: positional encoding, coarse depths, alpha compositing,
// inverse-CDF resampling and merge-sort.  All of them are HBM/latency-bound scans with one
// warp per ray (warp-shuffle scans and reductions); none is GEMM shaped.
#include "common.cuh"
#include <math_constants.h>

namespace inrf {

constexpr int WARPS_PER_CTA = 8;
constexpr unsigned FULL = 0xffffffffu;
constexpr int MAX_EXTRA_PER_LANE = 8;   // (112 classes + 128 endpoint) / 32

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}

// ---------------------------------------------------------------------------------
// Embedder.embed  (run_nerf_helpers.py:195-225; semantic_nerf.py:14-65)
// ---------------------------------------------------------------------------------
__global__ void k_embed(const float* __restrict__ x, int64_t M, int L, float scale, float* __restrict__ out) {
  const int width = 3 + 6 * L;
  int64_t total = M * width;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t m = i / width;
    int c = (int)(i - m * width);
    float v;
    if (c < 3) {
      v = x[m * 3 + c];
      if (scale != 1.f) v = __fdiv_rn(v, scale);
    } else {
      int q = c - 3, k = q / 6, r = q % 6, ax = r % 3;
      float xv = x[m * 3 + ax];
      if (scale != 1.f) xv = __fdiv_rn(xv, scale);
      float arg = xv * (float)(1 << k);          // exact (power of two)
      v = (r < 3) ? sinf(arg) : cosf(arg);
    }
    out[i] = v;
  }
}

// ---------------------------------------------------------------------------------
// coarse sample depths (run_nerf.py:464-486; trainer.py:730-746)
// ---------------------------------------------------------------------------------
__global__ void k_coarse_z(const float* __restrict__ rays, const float* __restrict__ t_vals,
                           const float* __restrict__ t_rand, int64_t N, int S, int lindisp, float* __restrict__ z, Rng rng) {
  int64_t total = N * S;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t n = i / S;
    int s = (int)(i - n * S);
    float nearv = rays[n * 11 + 6], farv = rays[n * 11 + 7];
    const bool jit = t_rand != nullptr || rng.on;
    z[i] = coarse_z_sample(nearv, farv, t_vals, s, S, lindisp, jit, t_rand != nullptr ? t_rand[i] : (rng.on ? rng_uniform(rng, i) : 0.f));
  }
}

// ---------------------------------------------------------------------------------
// raw2outputs (run_nerf.py:359-412; model_utils.py:39-116): one warp per ray.
// Lane l owns samples l, l+32, ...; the exclusive transmittance product is a warp
// multiplicative scan per 32-sample group with a running carry.
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
k_raw2outputs(const float* __restrict__ raw, const float* __restrict__ z, const float* __restrict__ rays_d, int ld_d,
              const float* __restrict__ noise, int64_t N, int S, int ch, int n_extra, int white_bkgd,
              float* __restrict__ rec, int rec_ch, float* __restrict__ weights, Rng rng) {
  const int lane = threadIdx.x & 31;
  const int64_t ray = blockIdx.x * (int64_t)WARPS_PER_CTA + (threadIdx.x >> 5);
  if (ray >= N) return;
  const float* rw = raw + ray * (int64_t)S * ch;
  const float* zr = z + ray * (int64_t)S;
  const float dx = rays_d[ray * ld_d + 0], dy = rays_d[ray * ld_d + 1], dz = rays_d[ray * ld_d + 2];
  const float dnorm = sqrtf(dx * dx + dy * dy + dz * dz);

  float carry = 1.f;                 // prod_{j < group start} (1 - alpha_j + 1e-10)
  float acc[12];                     // rgb3 albedo3 shading residual3 depth acc
#pragma unroll
  for (int i = 0; i < 12; ++i) acc[i] = 0.f;
  float ex[MAX_EXTRA_PER_LANE];      // semantic logits / endpoint features, channel = lane + 32 i
#pragma unroll
  for (int i = 0; i < MAX_EXTRA_PER_LANE; ++i) ex[i] = 0.f;
  for (int g = 0; g < S; g += 32) {
    int s = g + lane;
    bool ok = s < S;
    float w = 0.f, one_minus = 1.f, zs = 0.f;
    float c[INRF_RAW_BASE];
    if (ok) {
      zs = zr[s];
      float dist = (s + 1 < S) ? __fsub_rn(zr[s + 1], zs) : 1e10f;
      dist = __fmul_rn(dist, dnorm);
#pragma unroll
      for (int i = 0; i < INRF_RAW_BASE; ++i) c[i] = rw[(int64_t)s * ch + i];
      float sig = c[3];
      if (noise != nullptr) sig = __fadd_rn(sig, noise[ray * S + s]);
      else if (rng.on) sig = __fadd_rn(sig, rng_normal(rng, ray * S + s));
      float alpha = __fsub_rn(1.f, expf(-__fmul_rn(fmaxf(sig, 0.f), dist)));
      one_minus = __fadd_rn(__fsub_rn(1.f, alpha), 1e-10f);
      w = alpha;
    }
    // inclusive product scan of one_minus over the warp
    float p = one_minus;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      float q = __shfl_up_sync(FULL, p, o);
      if (lane >= o) p *= q;
    }
    float excl = __shfl_up_sync(FULL, p, 1);
    if (lane == 0) excl = 1.f;
    float T = carry * excl;
    carry *= __shfl_sync(FULL, p, 31);
    w *= T;
    if (ok) {
      if (weights != nullptr) weights[ray * S + s] = w;
      acc[0] += w * c[0]; acc[1] += w * c[1]; acc[2] += w * c[2];
      acc[3] += w * c[4]; acc[4] += w * c[5]; acc[5] += w * c[6];
      acc[6] += w * c[7];
      acc[7] += w * c[8]; acc[8] += w * c[9]; acc[9] += w * c[10];
      acc[10] += w * zs;
      acc[11] += w;
    }
    // extra channels (semantic logits, endpoint features): lanes stride over channels and
    // every sample's weight is broadcast by shuffle (coalesced row reads)
    if (n_extra > 0) {
      for (int j = 0; j < 32 && g + j < S; ++j) {
        float wj = __shfl_sync(FULL, w, j);
        const float* row = rw + (int64_t)(g + j) * ch + INRF_RAW_BASE;
#pragma unroll
        for (int i = 0; i < MAX_EXTRA_PER_LANE; ++i) {
          int e = lane + 32 * i;
          if (e < n_extra) ex[i] += wj * row[e];
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 12; ++i) acc[i] = warp_sum(acc[i]);
  const float accw = acc[11];
  const float bg = white_bkgd ? (1.f - accw) : 0.f;
  float* r = rec + ray * rec_ch;
  if (lane == 0) {
    r[0] = acc[0] + bg; r[1] = acc[1] + bg; r[2] = acc[2] + bg;
    float depth = acc[10];
    float ratio = depth / accw;                         // 0/0 -> NaN, kept (appendix A8)
    // torch.max(1e-10, NaN) propagates NaN; fmaxf would drop it
    float m = (ratio != ratio) ? ratio : fmaxf(1e-10f, ratio);
    r[3] = 1.f / m;
    r[4] = accw;
    r[5] = acc[3] + bg; r[6] = acc[4] + bg; r[7] = acc[5] + bg;
    r[8] = acc[6] + bg;
    r[9] = acc[7]; r[10] = acc[8]; r[11] = acc[9];      // residual: no background (A9)
    r[12] = depth;
  }
#pragma unroll
  for (int i = 0; i < MAX_EXTRA_PER_LANE; ++i) {
    int e = lane + 32 * i;
    if (e < n_extra) r[INRF_REC_BASE + e] = ex[i];   // white bkgd for the semantic part: k_add_bg_sem
  }
}

// ---------------------------------------------------------------------------------
// backward of raw2outputs: one warp per ray, lane l owns samples l, l+32, ... (<= 8 groups).
//   w_i = a_i T_i,  T_i = prod_{j<i} om_j,  om_j = 1 - a_j + 1e-10,  a = 1 - exp(-relu(s+n) d)
//   dL/da_i = Gw_i T_i - (sum_{j>i} Gw_j w_j) / om_i,   da/ds = d (1-a) [s+n > 0]
// ---------------------------------------------------------------------------------
constexpr int BWD_MAXG = 8;

__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
k_raw2outputs_bwd(const float* __restrict__ raw, const float* __restrict__ z, const float* __restrict__ rays_d, int ld_d,
                  const float* __restrict__ noise, int64_t N, int S, int ch, int n_sem, int n_extra, int white_bkgd,
                  const float* __restrict__ grec, int rec_ch, const float* __restrict__ gweights, float* __restrict__ graw, Rng rng) {
  const int lane = threadIdx.x & 31;
  const int64_t ray = blockIdx.x * (int64_t)WARPS_PER_CTA + (threadIdx.x >> 5);
  if (ray >= N) return;
  const float* rw = raw + ray * (int64_t)S * ch;
  float* gr = graw + ray * (int64_t)S * ch;
  const float* zr = z + ray * (int64_t)S;
  const float* g = grec + ray * rec_ch;
  const float dx = rays_d[ray * ld_d + 0], dy = rays_d[ray * ld_d + 1], dz = rays_d[ray * ld_d + 2];
  const float dnorm = sqrtf(dx * dx + dy * dy + dz * dz);
  const int G = (S + 31) / 32;
  float al[BWD_MAXG], T[BWD_MAXG], dist[BWD_MAXG], pos[BWD_MAXG];
  float carry = 1.f, acc = 0.f, depth = 0.f;
#pragma unroll
  for (int gi = 0; gi < BWD_MAXG; ++gi) {
    al[gi] = 0.f; T[gi] = 0.f; dist[gi] = 0.f; pos[gi] = 0.f;
    if (gi < G) {
      const int s = gi * 32 + lane;
      float one_minus = 1.f, a = 0.f;
      if (s < S) {
        float d = (s + 1 < S) ? (zr[s + 1] - zr[s]) : 1e10f;
        d *= dnorm;
        float sig = rw[(int64_t)s * ch + 3] + (noise ? noise[ray * S + s] : (rng.on ? rng_normal(rng, ray * S + s) : 0.f));
        a = 1.f - expf(-fmaxf(sig, 0.f) * d);
        one_minus = 1.f - a + 1e-10f;
        dist[gi] = d;
        pos[gi] = sig > 0.f ? 1.f : 0.f;
      }
      float p = one_minus;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { float q = __shfl_up_sync(FULL, p, o); if (lane >= o) p *= q; }
      float excl = __shfl_up_sync(FULL, p, 1);
      if (lane == 0) excl = 1.f;
      al[gi] = a;
      T[gi] = carry * excl;
      carry *= __shfl_sync(FULL, p, 31);
      if (s < S) { float w = a * T[gi]; acc += w; depth += w * zr[s]; }
    }
  }
  acc = warp_sum(acc);
  depth = warp_sum(depth);
  // upstream gradients of the per-ray record
  const float g_disp = g[3], g_acc = g[4], g_depth = g[12];
  const float ratio = depth / acc;
  const float m = fmaxf(1e-10f, ratio);
  const float g_ratio = (ratio > 1e-10f) ? (-g_disp / (m * m)) : 0.f;   // NaN ratio (acc == 0): no gradient
  float bg = 0.f;
  if (white_bkgd) {
    bg = g[0] + g[1] + g[2] + g[5] + g[6] + g[7] + g[8];
    for (int e = 0; e < n_sem; ++e) bg += g[INRF_REC_BASE + e];
  }
  // reverse sweep: suffix sum of Gw_j w_j
  float suffix = 0.f;
#pragma unroll
  for (int gi = BWD_MAXG - 1; gi >= 0; --gi) {
    if (gi < G) {
      const int s = gi * 32 + lane;
      float gw = 0.f, w = 0.f;
      if (s < S) {
        const float* c = rw + (int64_t)s * ch;
        w = al[gi] * T[gi];
        gw = g[0] * c[0] + g[1] * c[1] + g[2] * c[2] + g[5] * c[4] + g[6] * c[5] + g[7] * c[6] + g[8] * c[7] +
             g[9] * c[8] + g[10] * c[9] + g[11] * c[10] + g_depth * zr[s] + g_acc - bg;
        for (int e = 0; e < n_extra; ++e) gw += g[INRF_REC_BASE + e] * c[INRF_RAW_BASE + e];
        if (acc != 0.f) gw += g_ratio * (zr[s] - ratio) / acc;
        if (gweights) gw += gweights[ray * S + s];
        // colour-like channels: dL/dc = g w
        float* o = gr + (int64_t)s * ch;
        o[0] = g[0] * w; o[1] = g[1] * w; o[2] = g[2] * w;
        o[4] = g[5] * w; o[5] = g[6] * w; o[6] = g[7] * w; o[7] = g[8] * w;
        o[8] = g[9] * w; o[9] = g[10] * w; o[10] = g[11] * w;
        for (int e = 0; e < n_extra; ++e) o[INRF_RAW_BASE + e] = g[INRF_REC_BASE + e] * w;
      }
      // inclusive suffix scan of gw*w inside the group (lanes above me), plus the groups after
      float v = gw * w;
      float incl = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { float q = __shfl_down_sync(FULL, incl, o); if (lane + o < 32) incl += q; }
      const float after = suffix + (incl - v);          // sum over samples strictly after s
      suffix += __shfl_sync(FULL, incl, 0);
      if (s < S) {
        const float om = 1.f - al[gi] + 1e-10f;
        const float dLda = gw * T[gi] - after / om;
        gr[(int64_t)s * ch + 3] = dLda * dist[gi] * (1.f - al[gi]) * pos[gi];
      }
    }
  }
}

// semantic white-background fix-up kept separate so the main kernel stays simple
__global__ void k_add_bg_sem(float* __restrict__ rec, int64_t N, int rec_ch, int n_sem) {
  int64_t total = N * n_sem;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t n = i / n_sem;
    int e = (int)(i - n * n_sem);
    float accw = rec[n * rec_ch + 4];
    rec[n * rec_ch + INRF_REC_BASE + e] += 1.f - accw;
  }
}

// ---------------------------------------------------------------------------------
// sample_pdf (run_nerf_helpers.py:402-445; rays.py:176-220): one warp per ray.
// The cdf is accumulated in fp64 and rounded per element, which is what ATen's CPU
// cumsum does for float input (acc_type<float,false> = double).
// ---------------------------------------------------------------------------------
constexpr int MAX_BINS = 256;

__device__ __forceinline__ void invert_one(const float* cdf_s, const float* bins, int B, float u, float* sample, int64_t* ind) {
  // searchsorted(cdf, u, right=True): number of entries <= u
  int lo = 0, hi = B;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (cdf_s[mid] <= u) lo = mid + 1; else hi = mid;
  }
  int below = max(lo - 1, 0), above = min(lo, B - 1);
  float c0 = cdf_s[below], c1 = cdf_s[above];
  float b0 = bins[below], b1 = bins[above];
  float denom = __fsub_rn(c1, c0);
  if (denom < 1e-5f) denom = 1.f;
  float t = __fdiv_rn(__fsub_rn(u, c0), denom);
  *sample = __fadd_rn(b0, __fmul_rn(t, __fsub_rn(b1, b0)));
  if (ind) *ind = lo;
}

__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
k_sample_pdf(const float* __restrict__ bins, const float* __restrict__ weights, int ld_w, const float* __restrict__ cdf_in,
             const float* __restrict__ u, const float* __restrict__ u_det, int64_t N, int B, int n_samples,
             float* __restrict__ samples, int64_t* __restrict__ inds, float* __restrict__ cdf_out, Rng rng) {
  __shared__ float s_cdf[WARPS_PER_CTA][MAX_BINS];
  __shared__ float s_bins[WARPS_PER_CTA][MAX_BINS];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t ray = blockIdx.x * (int64_t)WARPS_PER_CTA + wid;
  if (ray >= N) return;
  float* cdf_s = s_cdf[wid];
  float* bin_s = s_bins[wid];
  for (int i = lane; i < B; i += 32) bin_s[i] = bins[ray * B + i];
  if (cdf_in != nullptr) {
    for (int i = lane; i < B; i += 32) cdf_s[i] = cdf_in[ray * B + i];
  } else {
    const float* w = weights + ray * (int64_t)ld_w;
    float part = 0.f;
    for (int i = lane; i < B - 1; i += 32) part += __fadd_rn(w[i], 1e-5f);
    const float total = warp_sum(part);
    double carry = 0.0;
    if (lane == 0) cdf_s[0] = 0.f;
    for (int g = 0; g < B - 1; g += 32) {
      int i = g + lane;
      double p = (i < B - 1) ? (double)__fdiv_rn(__fadd_rn(w[i], 1e-5f), total) : 0.0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        double q = __shfl_up_sync(FULL, p, o);
        if (lane >= o) p += q;
      }
      if (i < B - 1) cdf_s[i + 1] = (float)(carry + p);
      carry += __shfl_sync(FULL, p, 31);
    }
  }
  __syncwarp();
  if (cdf_out != nullptr) for (int i = lane; i < B; i += 32) cdf_out[ray * B + i] = cdf_s[i];
  for (int j = lane; j < n_samples; j += 32) {
    float uj = (u != nullptr) ? u[ray * n_samples + j] : (rng.on ? rng_uniform(rng, ray * n_samples + j) : u_det[j]);
    float smp; int64_t ind;
    invert_one(cdf_s, bin_s, B, uj, &smp, &ind);
    samples[ray * n_samples + j] = smp;
    if (inds != nullptr) inds[ray * n_samples + j] = ind;
  }
}

// ---------------------------------------------------------------------------------
// sort(cat(z_a, z_b)) + std(z_b, unbiased=False)  (run_nerf.py:503,519): one warp per
// ray, bitonic network in shared memory on the next power of two (padding = +inf).
// ---------------------------------------------------------------------------------
constexpr int MERGE_WARPS = 4;
constexpr int MERGE_MAX = 1024;

__global__ void __launch_bounds__(MERGE_WARPS * 32)
k_merge_sorted(const float* __restrict__ za, const float* __restrict__ zb, int64_t N, int Sa, int Sb, int P2,
               float* __restrict__ zout, float* __restrict__ zstd) {
  extern __shared__ float s_buf[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t ray = blockIdx.x * (int64_t)MERGE_WARPS + wid;
  if (ray >= N) return;
  float* v = s_buf + wid * P2;
  const int S = Sa + Sb;
  for (int i = lane; i < P2; i += 32) {
    float x = CUDART_INF_F;
    if (i < Sa) x = za[ray * Sa + i];
    else if (i < S) x = zb[ray * Sb + (i - Sa)];
    v[i] = x;
  }
  if (zstd != nullptr) {
    float s = 0.f;
    for (int i = lane; i < Sb; i += 32) s += zb[ray * Sb + i];
    float mean = warp_sum(s) / (float)Sb;
    float q = 0.f;
    for (int i = lane; i < Sb; i += 32) { float d = zb[ray * Sb + i] - mean; q += d * d; }
    q = warp_sum(q);
    if (lane == 0) zstd[ray] = sqrtf(q / (float)Sb);
  }
  __syncwarp();
  // Fast path (every inference call, and training too: the coarse depths are ascending by construction and
  // sample_pdf's output is ascending whenever its u is): both lists already sorted -> each element's final position
  // is its own index plus its rank in the other list (binary search in shared memory); ties put `za` first, which
  // yields the same VALUES as any other tie rule.  NaNs fail the sortedness test and take the general path.
  {
    bool ok = true;
    for (int i = lane; i < S - 1; i += 32)
      if (i != Sa - 1) ok = ok && (v[i] <= v[i + 1]);
    if (__all_sync(0xffffffffu, ok)) {
      const float* a = v;
      const float* b = v + Sa;
      for (int i = lane; i < S; i += 32) {
        const float x = v[i];
        int lo = 0, hi, pos;
        if (i < Sa) {                                   // rank of a_i in b: #(b < x)
          hi = Sb;
          while (lo < hi) { const int mid = (lo + hi) >> 1; if (b[mid] < x) lo = mid + 1; else hi = mid; }
          pos = i + lo;
        } else {                                        // rank of b_j in a: #(a <= x)
          hi = Sa;
          while (lo < hi) { const int mid = (lo + hi) >> 1; if (a[mid] <= x) lo = mid + 1; else hi = mid; }
          pos = (i - Sa) + lo;
        }
        zout[ray * S + pos] = x;
      }
      return;
    }
  }
  for (int k = 2; k <= P2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = lane; t < P2 / 2; t += 32) {
        int i = 2 * t - (t & (j - 1));      // index with bit j cleared
        int p = i + j;
        bool up = (i & k) == 0;
        float a = v[i], b = v[p];
        if ((a > b) == up) { v[i] = b; v[p] = a; }
      }
      __syncwarp();
    }
  }
  for (int i = lane; i < S; i += 32) zout[ray * S + i] = v[i];
}

// z_mid = .5 * (z[1:] + z[:-1])  (run_nerf.py:499)
__global__ void k_zmid(const float* __restrict__ z, int64_t N, int S, float* __restrict__ zmid) {
  int64_t total = N * (S - 1);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t n = i / (S - 1);
    int s = (int)(i - n * (S - 1));
    zmid[i] = __fmul_rn(0.5f, __fadd_rn(z[n * S + s + 1], z[n * S + s]));
  }
}

// get_rays + render()'s ray packing (run_nerf_helpers.py:359-368, run_nerf.py:100-128) and its SSR twin
// create_rays (SSR/models/rays.py:48-76 get_rays_camera, :79-84 get_rays_world, :223-256), for a full image
// (pix == nullptr) or for selected pixels pix[n] = row * W + column (training batches: sampling_index,
// rays.py:153-172; run_nerf.py:913-932 - the random draws stay with the caller).
__global__ void k_get_rays(int H, int W, Cam c, const int64_t* __restrict__ pix, int64_t total, float* __restrict__ rays) {
  for (int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; n < total; n += (int64_t)gridDim.x * blockDim.x) {
    float o[11];
    cam_ray(c, H, W, pix ? pix[n] : n, o);
    float* dst = rays + n * 11;
#pragma unroll
    for (int i = 0; i < 11; ++i) dst[i] = o[i];
  }
}

// ---------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------
static inline int grid_for(int64_t total, int block, int cap = 148 * 16) {
  int64_t g = (total + block - 1) / block;
  if (g < 1) g = 1;
  if (g > cap) g = cap;
  return (int)g;
}

int launch_get_rays(int H, int W, float fx, float fy, float cx, float cy, const float* c2w, int opencv, int euclidean,
                    float nearv, float farv, const int64_t* pix, int64_t n, float* rays, cudaStream_t st) {
  if (n == 0) return INRF_OK;
  Cam c;
  c.fx = fx; c.fy = fy; c.cx = cx; c.cy = cy; c.nearv = nearv; c.farv = farv; c.opencv = opencv; c.euclidean = euclidean;
  for (int i = 0; i < 12; ++i) c.m[i] = c2w[i];
  k_get_rays<<<grid_for(n, 256), 256, 0, st>>>(H, W, c, pix, n, rays);
  INRF_LAUNCH_CHECK();
  return INRF_OK;
}

int launch_zmid(const float* z, int64_t N, int S, float* zmid, cudaStream_t st) {
  if (N == 0) return INRF_OK;
  k_zmid<<<grid_for(N * (S - 1), 256), 256, 0, st>>>(z, N, S, zmid);
  INRF_LAUNCH_CHECK();
  return INRF_OK;
}

int launch_embed(const float* x, int64_t M, int L, float scale, float* out, cudaStream_t st) {
  if (M == 0) return INRF_OK;
  k_embed<<<grid_for(M * (3 + 6 * L), 256), 256, 0, st>>>(x, M, L, scale, out);
  INRF_LAUNCH_CHECK();
  return INRF_OK;
}

int launch_coarse_z(const float* rays, const float* t_vals, const float* t_rand, int64_t N, int S, int lindisp,
                    float* z, cudaStream_t st, Rng rng) {
  if (N == 0) return INRF_OK;
  k_coarse_z<<<grid_for(N * S, 256), 256, 0, st>>>(rays, t_vals, t_rand, N, S, lindisp, z, rng);
  INRF_LAUNCH_CHECK();
  return INRF_OK;
}

int launch_raw2outputs(const float* raw, const float* z, const float* rays_d, int ld_d, const float* noise, int64_t N,
                       int S, int n_classes, int endpoint, int white_bkgd, float* rec, float* weights, cudaStream_t st, Rng rng) {
  if (N == 0) return INRF_OK;
  int ch = raw_channels(n_classes, endpoint), rc = rec_channels(n_classes, endpoint);
  int64_t blocks = (N + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
  k_raw2outputs<<<(unsigned)blocks, WARPS_PER_CTA * 32, 0, st>>>(raw, z, rays_d, ld_d, noise, N, S, ch, ch - INRF_RAW_BASE,
                                                               white_bkgd, rec, rc, weights, rng);
  INRF_LAUNCH_CHECK();
  if (white_bkgd && n_classes > 0) {
    k_add_bg_sem<<<grid_for(N * n_classes, 256), 256, 0, st>>>(rec, N, rc, n_classes);
    INRF_LAUNCH_CHECK();
  }
  return INRF_OK;
}

int launch_raw2outputs_bwd(const float* raw, const float* z, const float* rays_d, int ld_d, const float* noise, int64_t N,
                           int S, int n_classes, int endpoint, int white_bkgd, const float* grec, const float* gweights,
                           float* graw, cudaStream_t st, Rng rng) {
  if (N == 0) return INRF_OK;
  if (S > BWD_MAXG * 32) { set_error("raw2outputs_bwd: %d samples per ray > %d", S, BWD_MAXG * 32); return INRF_EUNSUPPORTED; }
  int ch = raw_channels(n_classes, endpoint), rc = rec_channels(n_classes, endpoint);
  int64_t blocks = (N + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
  k_raw2outputs_bwd<<<(unsigned)blocks, WARPS_PER_CTA * 32, 0, st>>>(raw, z, rays_d, ld_d, noise, N, S, ch, n_classes,
                                                                   ch - INRF_RAW_BASE, white_bkgd, grec, rc, gweights, graw, rng);
  INRF_LAUNCH_CHECK();
  return INRF_OK;
}

int launch_sample_pdf(const float* bins, const float* weights, int ld_w, const float* cdf_in, const float* u,
                      const float* u_det, int64_t N, int B, int n_samples, float* samples, int64_t* inds,
                      float* cdf_out, cudaStream_t st, Rng rng) {
  if (N == 0) return INRF_OK;
  if (B > MAX_BINS || B < 2) { set_error("sample_pdf: bins per ray %d outside [2,%d]", B, MAX_BINS); return INRF_EUNSUPPORTED; }
  int64_t blocks = (N + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
  k_sample_pdf<<<(unsigned)blocks, WARPS_PER_CTA * 32, 0, st>>>(bins, weights, ld_w, cdf_in, u, u_det, N, B, n_samples,
                                                              samples, inds, cdf_out, rng);
  INRF_LAUNCH_CHECK();
  return INRF_OK;
}

int launch_merge_sorted(const float* za, const float* zb, int64_t N, int Sa, int Sb, float* zout, float* zstd,
                        cudaStream_t st) {
  if (N == 0) return INRF_OK;
  int S = Sa + Sb;
  if (S > MERGE_MAX || S < 1) { set_error("merge_sorted: %d samples per ray outside [1,%d]", S, MERGE_MAX); return INRF_EUNSUPPORTED; }
  int P2 = 2;
  while (P2 < S) P2 <<= 1;
  int64_t blocks = (N + MERGE_WARPS - 1) / MERGE_WARPS;
  k_merge_sorted<<<(unsigned)blocks, MERGE_WARPS * 32, MERGE_WARPS * P2 * sizeof(float), st>>>(za, zb, N, Sa, Sb, P2, zout, zstd);
  INRF_LAUNCH_CHECK();
  return INRF_OK;
}

}  // namespace inrf
