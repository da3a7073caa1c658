// Reflectance clustering kernels (object_level/cluster.py, SSR/training/cluster.py).
// All of them are scans over pixels x {anchors, seeds, neighbours}: FP32-ALU / L2 bound, no
// tensor cores.  Arg-min style reductions use 64-bit keys (orderable distance << 32 | index)
// with atomicMin so that results are deterministic and ties resolve to the lowest index,
// which is torch.argmin's rule and the deterministic statement of the reference's
// sort-then-scatter in choose_anchors.
#include "common.cuh"
#include <math_constants.h>

namespace inrf {

typedef unsigned long long u64;

__device__ __forceinline__ unsigned orderable(float f) {   // monotone float -> uint (NaN sorts last)
  unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__device__ __forceinline__ void map_color(float r, float g, float b, float f, float* o) {
  // Cluster.mapping_color (cluster.py:266-275): I = r+g+b; (I/3*f, g/I, b/I); no guard for I=0
  float I = __fadd_rn(__fadd_rn(r, g), b);
  o[0] = __fmul_rn(__fdiv_rn(I, 3.0f), f);
  o[1] = __fdiv_rn(g, I);
  o[2] = __fdiv_rn(b, I);
}

__global__ void k_mapping_color(const float* __restrict__ rgb, int64_t P, float f, float* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < P; i += (int64_t)gridDim.x * blockDim.x) {
    float o[3];
    map_color(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2], f, o);
    out[3 * i] = o[0]; out[3 * i + 1] = o[1]; out[3 * i + 2] = o[2];
  }
}

// ---------------------------------------------------------------------------------
// nearest anchor: grid = (pixel blocks, anchor splits); anchors staged through smem as
// (x,y,z,|a|^2); every thread owns PIX pixels and scans the tile (broadcast smem reads).
// ---------------------------------------------------------------------------------
constexpr int NA_THREADS = 256;
constexpr int NA_PIX = 2;
constexpr int NA_TILE = 1024;

__global__ void k_fill_u64(u64* p, int64_t n, u64 v) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}

__global__ void __launch_bounds__(NA_THREADS)
k_nearest_anchor(const float* __restrict__ rgb, int64_t P, const float* __restrict__ anchors, int64_t A,
                 int map, float f, int64_t anchors_per_split, u64* __restrict__ keys) {
  __shared__ float4 s_a[NA_TILE];
  float px[NA_PIX][3], pp[NA_PIX], best[NA_PIX];
  unsigned bidx[NA_PIX];
  int64_t pid[NA_PIX];
#pragma unroll
  for (int q = 0; q < NA_PIX; ++q) {
    pid[q] = (blockIdx.x * (int64_t)NA_PIX + q) * NA_THREADS + threadIdx.x;
    int64_t i = pid[q] < P ? pid[q] : P - 1;
    float r = rgb[3 * i], g = rgb[3 * i + 1], b = rgb[3 * i + 2];
    if (map) map_color(r, g, b, f, px[q]); else { px[q][0] = r; px[q][1] = g; px[q][2] = b; }
    pp[q] = __fadd_rn(__fadd_rn(__fmul_rn(px[q][0], px[q][0]), __fmul_rn(px[q][1], px[q][1])), __fmul_rn(px[q][2], px[q][2]));
    best[q] = CUDART_INF_F;
    bidx[q] = 0xffffffffu;
  }
  const int64_t a_begin = blockIdx.y * anchors_per_split;
  const int64_t a_end = min(A, a_begin + anchors_per_split);
  for (int64_t t0 = a_begin; t0 < a_end; t0 += NA_TILE) {
    int n = (int)min((int64_t)NA_TILE, a_end - t0);
    __syncthreads();
    for (int j = threadIdx.x; j < n; j += NA_THREADS) {
      float x = anchors[3 * (t0 + j)], y = anchors[3 * (t0 + j) + 1], z = anchors[3 * (t0 + j) + 2];
      s_a[j] = make_float4(x, y, z, __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
    }
    __syncthreads();
    for (int j = 0; j < n; ++j) {
      const float4 a = s_a[j];
#pragma unroll
      for (int q = 0; q < NA_PIX; ++q) {
        // compute_dist (cluster.py:241-247): |a|^2 + |p|^2 - 2 a.p
        float dot = fmaf(a.z, px[q][2], fmaf(a.y, px[q][1], __fmul_rn(a.x, px[q][0])));
        float d = __fsub_rn(__fadd_rn(a.w, pp[q]), __fmul_rn(2.f, dot));
        if (d < best[q]) { best[q] = d; bidx[q] = (unsigned)(t0 + j); }
      }
    }
  }
#pragma unroll
  for (int q = 0; q < NA_PIX; ++q) {
    if (pid[q] < P && bidx[q] != 0xffffffffu) {
      u64 key = ((u64)orderable(best[q]) << 32) | bidx[q];
      atomicMin(&keys[pid[q]], key);
    }
  }
}

// key -> index (all-NaN rows never update the key: torch.argmin returns the first NaN = 0)
__global__ void k_finish_nearest(u64* __restrict__ keys, int64_t P, const int64_t* __restrict__ links,
                                 const float* __restrict__ centers, float* __restrict__ out_rgb,
                                 int64_t* __restrict__ out_class, int64_t* __restrict__ out_idx) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < P; i += (int64_t)gridDim.x * blockDim.x) {
    u64 k = keys[i];
    int64_t idx = (k == ~0ull) ? 0 : (int64_t)(k & 0xffffffffu);
    if (out_idx) out_idx[i] = idx;
    if (links) {
      int64_t c = links[idx];
      if (out_class) out_class[i] = c;
      if (out_rgb) { out_rgb[3 * i] = centers[3 * c]; out_rgb[3 * i + 1] = centers[3 * c + 1]; out_rgb[3 * i + 2] = centers[3 * c + 2]; }
    }
  }
}

// ---------------------------------------------------------------------------------
// choose_anchors (cluster.py:150-176)
// ---------------------------------------------------------------------------------
constexpr int VOX = 100;
constexpr int VOX_TOTAL = VOX * VOX * VOX;

__device__ __forceinline__ int voxel_coord(float p) {
  float q = __fdiv_rn(p, 0.01f);
  if (!(q == q)) return 0;                 // NaN -> INT64_MIN under .long() -> clamped to 0
  long long v = (long long)q;              // truncation toward zero like Tensor.long()
  if (q >= 9.2e18f) v = 0x7fffffffffffffffLL;
  if (q <= -9.2e18f) v = 0;
  return (int)max(0LL, min((long long)(VOX - 1), v));
}

__global__ void k_voxel_min(const float* __restrict__ pix, int64_t P, u64* __restrict__ vox) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < P; i += (int64_t)gridDim.x * blockDim.x) {
    float d = 0.f;
    int id[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float p = pix[3 * i + c];
      id[c] = voxel_coord(p);
      float center = __fadd_rn(__fmul_rn((float)id[c], 0.01f), 0.005f);
      float e = __fsub_rn(center, p);
      d = (c == 0) ? __fmul_rn(e, e) : __fadd_rn(d, __fmul_rn(e, e));
    }
    u64 key = ((u64)orderable(d) << 32) | (unsigned)i;
    atomicMin(&vox[(id[0] * VOX + id[1]) * VOX + id[2]], key);
  }
}

// ordered compaction of the occupied voxels by one CTA (1e6 entries, ~1k block scans)
__global__ void __launch_bounds__(1024)
k_compact_anchors(const u64* __restrict__ vox, const float* __restrict__ pix, const int64_t* __restrict__ labels,
                  float* __restrict__ anchors, int64_t* __restrict__ links, int32_t* __restrict__ n_out) {
  __shared__ int s_warp[32];
  __shared__ int s_base, s_total;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_base = 0;
  __syncthreads();
  for (int v0 = 0; v0 < VOX_TOTAL; v0 += 1024) {
    int v = v0 + threadIdx.x;
    u64 k = (v < VOX_TOTAL) ? vox[v] : ~0ull;
    int has = (k != ~0ull);
    unsigned bal = __ballot_sync(0xffffffffu, has);
    int pre = __popc(bal & ((1u << lane) - 1));
    if (lane == 0) s_warp[wid] = __popc(bal);
    __syncthreads();
    if (wid == 0) {
      int c = s_warp[lane];
      int x = c;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
      s_warp[lane] = x - c;              // exclusive offset of each warp
      if (lane == 31) s_total = x;
    }
    __syncthreads();
    if (has) {
      int off = s_base + s_warp[wid] + pre;
      int64_t i = (int64_t)(k & 0xffffffffu);
      anchors[3 * (int64_t)off] = pix[3 * i];
      anchors[3 * (int64_t)off + 1] = pix[3 * i + 1];
      anchors[3 * (int64_t)off + 2] = pix[3 * i + 2];
      links[off] = labels[i];
    }
    __syncthreads();
    if (threadIdx.x == 0) s_base += s_total;
    __syncthreads();
  }
  if (threadIdx.x == 0) *n_out = s_base;
}

// ---------------------------------------------------------------------------------
// flat-kernel mean shift for a group of seeds per CTA (sklearn _mean_shift_single_seed)
// ---------------------------------------------------------------------------------
constexpr int MS_SEEDS = 8;
constexpr int MS_THREADS = 256;

__global__ void __launch_bounds__(MS_THREADS)
k_meanshift(const float* __restrict__ pts, int64_t P, const float* __restrict__ seeds, int64_t Q, float bw,
            int max_iter, float* __restrict__ centers, int32_t* __restrict__ n_within, int32_t* __restrict__ n_iter) {
  __shared__ double s_sum[MS_THREADS / 32][MS_SEEDS][3];
  __shared__ int s_cnt[MS_THREADS / 32][MS_SEEDS];
  __shared__ float s_mean[MS_SEEDS][3];
  __shared__ int s_active[MS_SEEDS], s_iters[MS_SEEDS], s_count[MS_SEEDS];
  __shared__ int s_any;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t q0 = blockIdx.x * (int64_t)MS_SEEDS;
  if (threadIdx.x < MS_SEEDS) {
    int64_t q = q0 + threadIdx.x;
    bool ok = q < Q;
    for (int c = 0; c < 3; ++c) s_mean[threadIdx.x][c] = ok ? seeds[3 * q + c] : 0.f;
    s_active[threadIdx.x] = ok; s_iters[threadIdx.x] = 0; s_count[threadIdx.x] = 0;
  }
  __syncthreads();
  const float bw2 = bw * bw;
  const float stop = 1e-3f * bw;
  for (int it = 0; it < max_iter + 1; ++it) {
    float m[MS_SEEDS][3];
    int act[MS_SEEDS];
#pragma unroll
    for (int s = 0; s < MS_SEEDS; ++s) { act[s] = s_active[s]; m[s][0] = s_mean[s][0]; m[s][1] = s_mean[s][1]; m[s][2] = s_mean[s][2]; }
    double sum[MS_SEEDS][3];
    int cnt[MS_SEEDS];
#pragma unroll
    for (int s = 0; s < MS_SEEDS; ++s) { sum[s][0] = sum[s][1] = sum[s][2] = 0.0; cnt[s] = 0; }
    for (int64_t i = threadIdx.x; i < P; i += MS_THREADS) {
      float x = pts[3 * i], y = pts[3 * i + 1], z = pts[3 * i + 2];
#pragma unroll
      for (int s = 0; s < MS_SEEDS; ++s) {
        float dx = x - m[s][0], dy = y - m[s][1], dz = z - m[s][2];
        float d2 = dx * dx + dy * dy + dz * dz;
        if (act[s] && d2 <= bw2) { sum[s][0] += x; sum[s][1] += y; sum[s][2] += z; cnt[s]++; }
      }
    }
#pragma unroll
    for (int s = 0; s < MS_SEEDS; ++s) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        double v = sum[s][c];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) s_sum[wid][s][c] = v;
      }
      int n = cnt[s];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
      if (lane == 0) s_cnt[wid][s] = n;
    }
    __syncthreads();
    if (threadIdx.x < MS_SEEDS && s_active[threadIdx.x]) {
      const int s = threadIdx.x;
      double t[3] = {0, 0, 0};
      int n = 0;
      for (int w = 0; w < MS_THREADS / 32; ++w) { t[0] += s_sum[w][s][0]; t[1] += s_sum[w][s][1]; t[2] += s_sum[w][s][2]; n += s_cnt[w][s]; }
      s_count[s] = n;
      if (n == 0) {                       // sklearn: "if len(points_within) == 0: break"
        s_active[s] = 0;
      } else {
        float nm[3] = {(float)(t[0] / n), (float)(t[1] / n), (float)(t[2] / n)};
        float dx = nm[0] - s_mean[s][0], dy = nm[1] - s_mean[s][1], dz = nm[2] - s_mean[s][2];
        float shift = sqrtf(dx * dx + dy * dy + dz * dz);
        s_mean[s][0] = nm[0]; s_mean[s][1] = nm[1]; s_mean[s][2] = nm[2];
        // converged or out of iterations: sklearn returns the NEW mean, len(points_within)
        // of this sweep and completed_iterations (incremented only when it continues)
        if (shift <= stop || s_iters[s] == max_iter) s_active[s] = 0;
        else s_iters[s]++;
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) { int a = 0; for (int s = 0; s < MS_SEEDS; ++s) a |= s_active[s]; s_any = a; }
    __syncthreads();
    if (!s_any) break;
  }
  if (threadIdx.x < MS_SEEDS && q0 + threadIdx.x < Q) {
    int64_t q = q0 + threadIdx.x;
    for (int c = 0; c < 3; ++c) centers[3 * q + c] = s_mean[threadIdx.x][c];
    n_within[q] = s_count[threadIdx.x];
    n_iter[q] = s_iters[threadIdx.x];
  }
}

// ---------------------------------------------------------------------------------
// k-th nearest neighbour distance (estimate_bandwidth): one CTA per query, distances in smem,
// bitonic sort, pick element k-1
// ---------------------------------------------------------------------------------
constexpr int KN_THREADS = 256;

__global__ void __launch_bounds__(KN_THREADS)
k_kth_dist(const float* __restrict__ pts, int P, int P2, const float* __restrict__ qs, int k, float* __restrict__ out) {
  extern __shared__ float s_d[];
  const int q = blockIdx.x;
  const float x = qs[3 * q], y = qs[3 * q + 1], z = qs[3 * q + 2];
  for (int i = threadIdx.x; i < P2; i += KN_THREADS) {
    float d = CUDART_INF_F;
    if (i < P) { float dx = pts[3 * i] - x, dy = pts[3 * i + 1] - y, dz = pts[3 * i + 2] - z; d = dx * dx + dy * dy + dz * dz; }
    s_d[i] = d;
  }
  __syncthreads();
  for (int kk = 2; kk <= P2; kk <<= 1) {
    for (int j = kk >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < P2 / 2; t += KN_THREADS) {
        int i = 2 * t - (t & (j - 1));
        int p = i + j;
        bool up = (i & kk) == 0;
        float a = s_d[i], b = s_d[p];
        if ((a > b) == up) { s_d[i] = b; s_d[p] = a; }
      }
      __syncthreads();
    }
  }
  if (threadIdx.x == 0) out[q] = sqrtf(s_d[k - 1]);
}

static inline int grid1d(int64_t n, int block, int cap = 148 * 8) {
  int64_t g = (n + block - 1) / block;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

int nearest_common(const float* rgb, int64_t P, const float* anchors, int64_t A, int map, float f, u64* keys,
                   cudaStream_t st) {
  k_fill_u64<<<grid1d(P, 256), 256, 0, st>>>(keys, P, ~0ull);
  INRF_LAUNCH_CHECK();
  int64_t pix_blocks = (P + NA_THREADS * NA_PIX - 1) / (NA_THREADS * NA_PIX);
  // few pixels and many anchors (a training step): split the anchor range across CTAs
  int splits = 1;
  while (pix_blocks * splits < 148 * 2 && (A + splits - 1) / splits > NA_TILE) splits *= 2;
  int64_t per = ((A + splits - 1) / splits + NA_TILE - 1) / NA_TILE * NA_TILE;
  splits = (int)((A + per - 1) / per);
  dim3 grid((unsigned)pix_blocks, (unsigned)splits);
  k_nearest_anchor<<<grid, NA_THREADS, 0, st>>>(rgb, P, anchors, A, map, f, per, keys);
  INRF_LAUNCH_CHECK();
  return INRF_OK;
}

}  // namespace inrf

using namespace inrf;

extern "C" {

int inrf_mapping_color(const float* rgb, int64_t P, float intensity_factor, float* out, void* stream) {
  INRF_CHECK_ARG(P >= 0 && (P == 0 || (rgb && out)), "null pointer / negative size");
  if (P == 0) return INRF_OK;
  k_mapping_color<<<grid1d(P, 256), 256, 0, (cudaStream_t)stream>>>(rgb, P, intensity_factor, out);
  INRF_LAUNCH_CHECK();
  return INRF_OK;
}

int inrf_nearest_anchor(const float* rgb, int64_t P, const float* anchors, int64_t A, int map_color_, float intensity_factor,
                        int64_t* idx, void* stream) {
  INRF_CHECK_ARG(P >= 0 && A > 0 && A < 0xffffffffLL && (P == 0 || (rgb && anchors && idx)), "null pointer / bad size");
  if (P == 0) return INRF_OK;
  cudaStream_t st = (cudaStream_t)stream;
  u64* keys = reinterpret_cast<u64*>(idx);            // the output buffer doubles as the key array
  int rc = nearest_common(rgb, P, anchors, A, map_color_, intensity_factor, keys, st);
  if (rc) return rc;
  k_finish_nearest<<<grid1d(P, 256), 256, 0, st>>>(keys, P, nullptr, nullptr, nullptr, nullptr, idx);
  INRF_LAUNCH_CHECK();
  return INRF_OK;
}

int inrf_dest_color(const float* rgb, int64_t P, const float* anchors, const int64_t* links, int64_t A,
                    const float* rgb_centers, int64_t K, float intensity_factor, float* out_rgb, int64_t* out_class,
                    void* stream) {
  INRF_CHECK_ARG(P >= 0 && A > 0 && A < 0xffffffffLL && K > 0, "bad size");
  INRF_CHECK_ARG(P == 0 || (rgb && anchors && links && (out_rgb || out_class)), "null pointer");
  INRF_CHECK_ARG(out_rgb == nullptr || rgb_centers != nullptr, "rgb_centers missing");
  INRF_CHECK_ARG(out_class != nullptr || out_rgb != nullptr, "no output requested");
  if (P == 0) return INRF_OK;
  cudaStream_t st = (cudaStream_t)stream;
  // key scratch: out_class when present, otherwise the first 8 bytes per pixel of out_rgb (12 B/pixel)
  // would alias rows - so a class buffer is required as scratch for the colour-only call.
  INRF_CHECK_ARG(out_class != nullptr, "out_class doubles as the arg-min scratch and must be provided");
  u64* keys = reinterpret_cast<u64*>(out_class);
  int rc = nearest_common(rgb, P, anchors, A, 1, intensity_factor, keys, st);
  if (rc) return rc;
  k_finish_nearest<<<grid1d(P, 256), 256, 0, st>>>(keys, P, links, rgb_centers, out_rgb, out_class, nullptr);
  INRF_LAUNCH_CHECK();
  return INRF_OK;
}

int inrf_choose_anchors(const float* pixels, const int64_t* labels, int64_t P, unsigned long long* voxel_key,
                        float* anchors, int64_t* links, int32_t* n_anchors, void* stream) {
  INRF_CHECK_ARG(P > 0 && P < 0xffffffffLL && pixels && labels && voxel_key && anchors && links && n_anchors, "null pointer / bad size");
  cudaStream_t st = (cudaStream_t)stream;
  k_fill_u64<<<grid1d(VOX_TOTAL, 256), 256, 0, st>>>(voxel_key, VOX_TOTAL, ~0ull);
  INRF_LAUNCH_CHECK();
  k_voxel_min<<<grid1d(P, 256), 256, 0, st>>>(pixels, P, voxel_key);
  INRF_LAUNCH_CHECK();
  k_compact_anchors<<<1, 1024, 0, st>>>(voxel_key, pixels, labels, anchors, links, n_anchors);
  INRF_LAUNCH_CHECK();
  return INRF_OK;
}

int inrf_meanshift_seeds(const float* points, int64_t P, const float* seeds, int64_t Q, float bandwidth, int max_iter,
                         float* centers, int32_t* n_within, int32_t* n_iter, void* stream) {
  INRF_CHECK_ARG(P > 0 && Q >= 0 && points && (Q == 0 || (seeds && centers && n_within && n_iter)), "null pointer / bad size");
  INRF_CHECK_ARG(bandwidth > 0.f && max_iter > 0, "bandwidth / max_iter must be positive");
  if (Q == 0) return INRF_OK;
  int64_t blocks = (Q + MS_SEEDS - 1) / MS_SEEDS;
  k_meanshift<<<(unsigned)blocks, MS_THREADS, 0, (cudaStream_t)stream>>>(points, P, seeds, Q, bandwidth, max_iter, centers, n_within, n_iter);
  INRF_LAUNCH_CHECK();
  return INRF_OK;
}

int inrf_kth_neighbor_dist(const float* points, int64_t P, const float* queries, int64_t Q, int k, float* kth_dist, void* stream) {
  INRF_CHECK_ARG(P > 0 && Q >= 0 && points && (Q == 0 || (queries && kth_dist)), "null pointer / bad size");
  INRF_CHECK_ARG(k >= 1 && k <= P, "k outside [1,P]");
  INRF_CHECK_SUPPORTED(P <= 32768, "more than 32768 reference points (estimate_bandwidth subsamples to n_samples)");
  if (Q == 0) return INRF_OK;
  int P2 = 2;
  while (P2 < P) P2 <<= 1;
  size_t smem = (size_t)P2 * sizeof(float);
  INRF_CUDA(cudaFuncSetAttribute(k_kth_dist, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_kth_dist<<<(unsigned)Q, KN_THREADS, smem, (cudaStream_t)stream>>>(points, (int)P, P2, queries, k, kth_dist);
  INRF_LAUNCH_CHECK();
  return INRF_OK;
}

}  // extern "C"
