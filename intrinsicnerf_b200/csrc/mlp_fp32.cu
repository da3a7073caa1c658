// CUDA-core fp32 evaluation of the intrinsic field network (INRF_PREC_FP32).
//
// Literal layer-by-layer statement of NeRF.forward (object_level/run_nerf_helpers.py:284-325)
// and Semantic_NeRF.forward (SSR/models/semantic_nerf.py:123-181) fused with the positional
// encoding (Embedder, run_nerf_helpers.py:195-225) and run_network's per-sample view
// direction expansion (run_nerf.py:42-56).  FFMA only - this is the strict-fp32 mode and
// the yardstick the tensor-core kernel (mlp_tc.cu) is validated against on the GPU.
//
// One CTA = 64 sample rows, 256 threads.  Activations ping-pong between two 64x256 fp32
// shared-memory tiles; weights are read transposed ([K][N], coalesced over N) straight from
// L2 (2.6 MB per network, resident).  Warp w owns rows 8w..8w+7, lane l owns output columns
// l, l+32, ...; the activation reads are warp-wide broadcasts.
#include "common.cuh"

namespace inrf {

constexpr int FT_ROWS = 64;
constexpr int FT_THREADS = 256;
constexpr int LD_PE = 64;
constexpr int LD_DIR = 28;
constexpr int LD_H = 256;

struct Fp32Params {
  NetLayout L;
  MlpArgs a;
  int out_ch;
};

struct Piece { const float* in; int ld; int K; };

// out[r][n] = act(bias[n] + sum over pieces sum_k in[r][k] * Wt[koff+k][n])
template <int TN, bool RELU>
__device__ __forceinline__ void dense(const Piece* pieces, int n_pieces, const float* __restrict__ Wt,
                                      const float* __restrict__ bias, float* out, int ldo, int col0) {
  constexpr int N = TN * 32;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float acc[8][TN];
#pragma unroll
  for (int j = 0; j < TN; ++j) {
    float b = __ldg(bias + lane + 32 * j);
#pragma unroll
    for (int r = 0; r < 8; ++r) acc[r][j] = b;
  }
  int koff = 0;
  for (int p = 0; p < n_pieces; ++p) {
    const float* in = pieces[p].in + (8 * w) * pieces[p].ld;
    const int ld = pieces[p].ld, K = pieces[p].K;
    int k = 0;
    for (; k + 4 <= K; k += 4) {
      float wv[4][TN];
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
#pragma unroll
        for (int j = 0; j < TN; ++j) wv[kk][j] = __ldg(Wt + (int64_t)(koff + k + kk) * N + lane + 32 * j);
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const float4 a4 = *reinterpret_cast<const float4*>(in + r * ld + k);
#pragma unroll
        for (int j = 0; j < TN; ++j) {
          acc[r][j] = fmaf(a4.x, wv[0][j], acc[r][j]);
          acc[r][j] = fmaf(a4.y, wv[1][j], acc[r][j]);
          acc[r][j] = fmaf(a4.z, wv[2][j], acc[r][j]);
          acc[r][j] = fmaf(a4.w, wv[3][j], acc[r][j]);
        }
      }
    }
    for (; k < K; ++k) {
      float wv[TN];
#pragma unroll
      for (int j = 0; j < TN; ++j) wv[j] = __ldg(Wt + (int64_t)(koff + k) * N + lane + 32 * j);
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        float a = in[r * ld + k];
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[r][j] = fmaf(a, wv[j], acc[r][j]);
      }
    }
    koff += K;
  }
#pragma unroll
  for (int r = 0; r < 8; ++r)
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      float v = acc[r][j];
      if (RELU) v = fmaxf(v, 0.f);
      out[(8 * w + r) * ldo + col0 + lane + 32 * j] = v;
    }
}

// narrow head: out[r][n] = bias[n] + sum_k in[r][k] * W[n][k]  (W row-major [N][K]); lanes split K
__device__ __forceinline__ float head_dot(const float* in_row, const float* __restrict__ Wn, int K) {
  const int lane = threadIdx.x & 31;
  float s = 0.f;
  for (int k = lane; k < K; k += 32) s = fmaf(in_row[k], __ldg(Wn + k), s);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  return s;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// copy a [64 x width] activation tile to the training stash (rows beyond M are skipped)
__device__ __forceinline__ void stash_tile(float* stash, int64_t row0, int64_t M, const float* tile, int ld, int col0,
                                           int width, int stash_col) {
  if (stash == nullptr) return;
  for (int i = threadIdx.x; i < FT_ROWS * width; i += FT_THREADS) {
    const int r = i / width, c = i - r * width;
    if (row0 + r < M) stash[(row0 + r) * STASH_LD + stash_col + c] = tile[r * ld + col0 + c];
  }
}

__global__ void __launch_bounds__(FT_THREADS, 1) k_mlp_fp32(const __grid_constant__ Fp32Params P) {
  extern __shared__ __align__(16) float smem[];
  float* s_pe = smem;                          // [64][64]  (col 63 = 0)
  float* s_dir = s_pe + FT_ROWS * LD_PE;       // [64][28]  (col 27 = 0)
  float* s_a = s_dir + FT_ROWS * LD_DIR;       // [64][256]
  float* s_b = s_a + FT_ROWS * LD_H;           // [64][256]
  float* s_out = s_b + FT_ROWS * LD_H;         // [64][12]: sigma, albedo3, shading, residual3
  const unsigned char* blob = static_cast<const unsigned char*>(P.a.packed);
  auto WT = [&](int l) { return reinterpret_cast<const float*>(blob + P.L.f32_wt[l]); };
  auto BI = [&](int l) { return reinterpret_cast<const float*>(blob + P.L.f32_b[l]); };
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int C = P.L.n_classes;
  const int64_t n_tiles = (P.a.M + FT_ROWS - 1) / FT_ROWS;

  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t row0 = tile * FT_ROWS;
    __syncthreads();
    // ---- positional encoding of 64 sample points and their view directions --------------
    for (int r = threadIdx.x; r < FT_ROWS; r += FT_THREADS) {
      int64_t m = min(row0 + r, P.a.M - 1);
      float x[3], d[3];
      if (P.a.emb != nullptr) {            // pre-embedded rows: copy, no sin/cos
        const float* e = P.a.emb + m * (PE_PTS + PE_DIR);
        for (int i = 0; i < PE_PTS; ++i) s_pe[r * LD_PE + i] = e[i];
        s_pe[r * LD_PE + 63] = 0.f;
        for (int i = 0; i < PE_DIR; ++i) s_dir[r * LD_DIR + i] = e[PE_PTS + i];
        s_dir[r * LD_DIR + 27] = 0.f;
        continue;
      }
      if (P.a.rays != nullptr) {
        int64_t n = m / P.a.S;
        const float* ray = P.a.rays + n * 11;
        float zv = P.a.z[m];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          x[i] = __fadd_rn(ray[i], __fmul_rn(ray[3 + i], zv));   // o + d*z, two roundings like ATen
          d[i] = ray[8 + i];
        }
      } else {
#pragma unroll
        for (int i = 0; i < 3; ++i) { x[i] = P.a.pts[m * 3 + i]; d[i] = P.a.viewdirs[m * 3 + i]; }
      }
      if (P.a.pe_scale != 1.f) {
#pragma unroll
        for (int i = 0; i < 3; ++i) x[i] = __fdiv_rn(x[i], P.a.pe_scale);
      }
      float* pe = s_pe + r * LD_PE;
      float* de = s_dir + r * LD_DIR;
#pragma unroll
      for (int i = 0; i < 3; ++i) { pe[i] = x[i]; de[i] = d[i]; }
      for (int k = 0; k < 10; ++k) {
        float f = (float)(1 << k);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          float sv, cv;
          sincosf(x[i] * f, &sv, &cv);
          pe[3 + 6 * k + i] = sv;
          pe[3 + 6 * k + 3 + i] = cv;
        }
      }
      pe[63] = 0.f;
      for (int k = 0; k < 4; ++k) {
        float f = (float)(1 << k);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          float sv, cv;
          sincosf(d[i] * f, &sv, &cv);
          de[3 + 6 * k + i] = sv;
          de[3 + 6 * k + 3 + i] = cv;
        }
      }
      de[27] = 0.f;
    }
    __syncthreads();
    // ---- trunk ---------------------------------------------------------------------------
    Piece pc[2];
    pc[0] = {s_pe, LD_PE, PE_PTS};
    dense<8, true>(pc, 1, WT(L_T0), BI(L_T0), s_a, LD_H, 0);
    __syncthreads();
    stash_tile(P.a.stash, row0, P.a.M, s_a, LD_H, 0, W_HID, ST_H);
    float* cur = s_a;
    float* nxt = s_b;
    for (int l = 1; l < 8; ++l) {
      if (l == 5) {
        pc[0] = {s_pe, LD_PE, PE_PTS};
        pc[1] = {cur, LD_H, W_HID};
        dense<8, true>(pc, 2, WT(L_T0 + l), BI(L_T0 + l), nxt, LD_H, 0);
      } else {
        pc[0] = {cur, LD_H, W_HID};
        dense<8, true>(pc, 1, WT(L_T0 + l), BI(L_T0 + l), nxt, LD_H, 0);
      }
      __syncthreads();
      stash_tile(P.a.stash, row0, P.a.M, nxt, LD_H, 0, W_HID, ST_H + l * W_HID);
      float* t = cur; cur = nxt; nxt = t;
    }
    // cur = h (trunk output), nxt = scratch
    // ---- sigma, albedo, shading heads ------------------------------------------------------
    pc[0] = {cur, LD_H, W_HID};
    dense<4, true>(pc, 1, WT(L_ALB1), BI(L_ALB1), nxt, LD_H, 0);
    dense<4, true>(pc, 1, WT(L_SH1), BI(L_SH1), nxt, LD_H, 128);
    for (int r = 0; r < 8; ++r) {
      float sg = head_dot(cur + (8 * w + r) * LD_H, WT(L_ALPHA), W_HID);
      if (lane == 0) s_out[(8 * w + r) * 12 + 0] = sg + __ldg(BI(L_ALPHA));
    }
    __syncthreads();
    stash_tile(P.a.stash, row0, P.a.M, nxt, LD_H, 0, W_HID, ST_AS);
    for (int r = 0; r < 8; ++r) {
      const float* row = nxt + (8 * w + r) * LD_H;
      for (int n = 0; n < 3; ++n) {
        float v = head_dot(row, WT(L_ALB2) + n * 128, 128);
        if (lane == 0) s_out[(8 * w + r) * 12 + 1 + n] = sigmoidf_(v + __ldg(BI(L_ALB2) + n));
      }
      float v = head_dot(row + 128, WT(L_SH2), 128);
      if (lane == 0) s_out[(8 * w + r) * 12 + 4] = sigmoidf_(v + __ldg(BI(L_SH2)));
    }
    __syncthreads();
    // ---- semantic head (SSR): Linear+ReLU(128) -> Linear(C) on the trunk output ------------
    if (C > 0) {
      dense<4, true>(pc, 1, WT(L_SEM1), BI(L_SEM1), nxt, LD_H, 0);
      __syncthreads();
      stash_tile(P.a.stash, row0, P.a.M, nxt, LD_H, 0, 128, ST_SEM1);
      for (int r = 0; r < 8; ++r) {
        int64_t m = row0 + 8 * w + r;
        const float* row = nxt + (8 * w + r) * LD_H;
        for (int n = 0; n < C; ++n) {
          float v = head_dot(row, WT(L_SEM2) + n * 128, 128);
          if (lane == 0 && m < P.a.M) P.a.raw[m * P.out_ch + INRF_RAW_BASE + n] = v + __ldg(BI(L_SEM2) + n);
        }
      }
      __syncthreads();
    }
    // ---- feature -> views -> residual ----------------------------------------------------------
    dense<8, false>(pc, 1, WT(L_FEAT), BI(L_FEAT), nxt, LD_H, 0);
    __syncthreads();
    stash_tile(P.a.stash, row0, P.a.M, nxt, LD_H, 0, W_HID, ST_FEAT);
    pc[0] = {nxt, LD_H, W_HID};
    pc[1] = {s_dir, LD_DIR, PE_DIR};
    dense<4, true>(pc, 2, WT(L_VIEWS), BI(L_VIEWS), cur, LD_H, 0);   // h is dead now
    __syncthreads();
    stash_tile(P.a.stash, row0, P.a.M, cur, LD_H, 0, 128, ST_V);
    for (int r = 0; r < 8; ++r) {
      const float* row = cur + (8 * w + r) * LD_H;
      for (int n = 0; n < 3; ++n) {
        float v = head_dot(row, WT(L_RES) + n * 128, 128);
        if (lane == 0) s_out[(8 * w + r) * 12 + 5 + n] = sigmoidf_(v + __ldg(BI(L_RES) + n));
      }
    }
    __syncthreads();
    // ---- assemble raw rows: rgb3 sigma albedo3 shading residual3 [sem] [endpoint] ----------------
    for (int i = threadIdx.x; i < FT_ROWS * INRF_RAW_BASE; i += FT_THREADS) {
      int r = i / INRF_RAW_BASE, c = i % INRF_RAW_BASE;
      int64_t m = row0 + r;
      if (m >= P.a.M) continue;
      const float* o = s_out + r * 12;
      float v;
      if (c < 3) v = __fadd_rn(__fmul_rn(o[1 + c], o[4]), o[5 + c]);   // albedo*shading + residual
      else if (c == 3) v = o[0];
      else if (c < 7) v = o[1 + (c - 4)];
      else if (c == 7) v = o[4];
      else v = o[5 + (c - 8)];
      P.a.raw[m * P.out_ch + c] = v;
    }
    if (P.a.endpoint) {
      for (int i = threadIdx.x; i < FT_ROWS * 128; i += FT_THREADS) {
        int r = i >> 7, c = i & 127;
        int64_t m = row0 + r;
        if (m < P.a.M) P.a.raw[m * P.out_ch + INRF_RAW_BASE + C + c] = cur[r * LD_H + c];
      }
    }
  }
}

int launch_mlp_fp32(const MlpArgs& a, cudaStream_t st) {
  if (a.M == 0) return INRF_OK;
  Fp32Params P;
  int rc = make_layout(a.variant, a.n_classes, &P.L);
  if (rc) return rc;
  P.a = a;
  P.out_ch = raw_channels(a.n_classes, a.endpoint);
  size_t smem = (size_t)(FT_ROWS * LD_PE + FT_ROWS * LD_DIR + 2 * FT_ROWS * LD_H + FT_ROWS * 12) * sizeof(float);
  INRF_CUDA(cudaFuncSetAttribute(k_mlp_fp32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int dev = 0, sms = 148;
  INRF_CUDA(cudaGetDevice(&dev));
  INRF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  int64_t tiles = (a.M + FT_ROWS - 1) / FT_ROWS;
  int grid = (int)(tiles < sms ? tiles : sms);
  k_mlp_fp32<<<grid, FT_THREADS, smem, st>>>(P);
  INRF_LAUNCH_CHECK();
  return INRF_OK;
}

}  // namespace inrf
