// Backward of the intrinsic field network (fp32 CUDA cores): parameter gradients of NeRF.forward
// (object_level/run_nerf_helpers.py:284-325) / Semantic_NeRF.forward (SSR/models/semantic_nerf.py:123-181)
// given dL/d(raw).  The sample positions carry no gradient (the reference never differentiates rays or
// depths: z_samples is detached, run_nerf.py:501), so only dL/d(parameters) is produced.
//
// One CTA = 64 sample rows.  The training forward (mlp_fp32.cu with a stash pointer) saved every
// post-activation tile in HBM; here each layer is walked in reverse with three 64x256 fp32 tiles in
// shared memory:  dZ = dY (.) [Y > 0],  dW += dZ^T X (64-accumulator register tiles, flushed with
// fp32 atomics into the flat gradient vector),  db += colsum(dZ),  dX = dZ W.
#include "common.cuh"

namespace inrf {

constexpr int BT_ROWS = 64;
constexpr int BT_THREADS = 256;
constexpr int LDT = 256;         // tile leading dimension
constexpr int LDPE = 64;
constexpr int LDDIR = 32;        // 27 used, padded to 32 so 16-wide k blocks stay in bounds

struct BwdParams {
  NetLayout L;
  MlpBwdArgs a;
  int out_ch;
};

// tile[r][c] = src[(row0+r)*src_ld + col0 + c] (zeros past M)
__device__ __forceinline__ void load_tile(float* tile, int ld, const float* __restrict__ src, int64_t row0, int64_t M,
                                          int src_ld, int col0, int width) {
  for (int i = threadIdx.x; i < BT_ROWS * width; i += BT_THREADS) {
    const int r = i / width, c = i - r * width;
    tile[r * ld + c] = (row0 + r < M) ? __ldg(src + (row0 + r) * (int64_t)src_ld + col0 + c) : 0.f;
  }
}

// tile[r][c] = 0 where y[(row0+r)][col0+c] <= 0 (ReLU backward against the stashed activation)
__device__ __forceinline__ void relu_mask(float* tile, int ld, const float* __restrict__ y, int64_t row0, int64_t M, int y_ld,
                                          int col0, int width) {
  for (int i = threadIdx.x; i < BT_ROWS * width; i += BT_THREADS) {
    const int r = i / width, c = i - r * width;
    const float yv = (row0 + r < M) ? __ldg(y + (row0 + r) * (int64_t)y_ld + col0 + c) : 0.f;
    if (!(yv > 0.f)) tile[r * ld + c] = 0.f;
  }
}

// gb[c] += sum_r dz[r][c]
__device__ __forceinline__ void accum_bias(const float* dz, int ld, int N, float* __restrict__ gb) {
  for (int c = threadIdx.x; c < N; c += BT_THREADS) {
    float s = 0.f;
#pragma unroll 8
    for (int r = 0; r < BT_ROWS; ++r) s += dz[r * ld + c];
    atomicAdd(gb + c, s);
  }
}

// gW[n][koff + k] += sum_r dz[r][dz_col0 + n] * x[r][k]   for n < N (multiple of 64), k < K (<= 256)
// thread (ty, tx): 4 n x 16 k register tile per 64-wide n block
__device__ __forceinline__ void accum_dw(const float* dz, int ldz, int dz_col0, int N, const float* x, int ldx, int K,
                                         float* __restrict__ gW, int ldw, int koff) {
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  const int k0 = tx * 16;
  if (k0 >= K) return;
  for (int nb = 0; nb < N; nb += 64) {
    const int n0 = nb + ty * 4;
    float acc[4][16];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[i][j] = 0.f;
#pragma unroll 2
    for (int r = 0; r < BT_ROWS; ++r) {
      const float4 d4 = *reinterpret_cast<const float4*>(dz + r * ldz + dz_col0 + n0);
      float xv[16];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 t = *reinterpret_cast<const float4*>(x + r * ldx + k0 + 4 * q);
        xv[4 * q] = t.x; xv[4 * q + 1] = t.y; xv[4 * q + 2] = t.z; xv[4 * q + 3] = t.w;
      }
      const float dv[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[i][j] = fmaf(dv[i], xv[j], acc[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (k0 + j < K) atomicAdd(gW + (int64_t)(n0 + i) * ldw + koff + k0 + j, acc[i][j]);
  }
}

// out[r][c] (+)= sum_{j<J} in[r][in_col0 + j] * W[j*ldw + c]   for c < 32*TN;  warp w owns rows 8w..8w+7
template <int TN, bool ACCUM>
__device__ __forceinline__ void dense_jk(const float* in, int ldi, int in_col0, int J, const float* __restrict__ W, int ldw,
                                         float* out, int ldo) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float acc[8][TN];
#pragma unroll
  for (int r = 0; r < 8; ++r)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[r][j] = ACCUM ? out[(8 * w + r) * ldo + lane + 32 * j] : 0.f;
  const float* ip = in + (8 * w) * ldi + in_col0;
  for (int k = 0; k < J; k += 4) {
    float wv[4][TN];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk)
#pragma unroll
      for (int j = 0; j < TN; ++j) wv[kk][j] = __ldg(W + (int64_t)(k + kk) * ldw + lane + 32 * j);
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const float4 a4 = *reinterpret_cast<const float4*>(ip + r * ldi + k);
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        acc[r][j] = fmaf(a4.x, wv[0][j], acc[r][j]);
        acc[r][j] = fmaf(a4.y, wv[1][j], acc[r][j]);
        acc[r][j] = fmaf(a4.z, wv[2][j], acc[r][j]);
        acc[r][j] = fmaf(a4.w, wv[3][j], acc[r][j]);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < 8; ++r)
#pragma unroll
    for (int j = 0; j < TN; ++j) out[(8 * w + r) * ldo + lane + 32 * j] = acc[r][j];
}

// narrow head (n_out <= 4 rows of W[n][K], K = 128 or 256): g[r][n] are the pre-activation grads
//   gW[n][k] += sum_r g[r][g_col0+n] x[r][x_col0+k];  gb[n] += sum_r g;  dx[r][dx_col0+k] (+)= sum_n g[r][n] W[n][k]
template <bool ACCUM>
__device__ __forceinline__ void narrow_head(const float* g, int ldg, int g_col0, int n_out, const float* x, int ldx, int x_col0,
                                            int K, const float* __restrict__ W, float* __restrict__ gW, float* __restrict__ gb,
                                            float* dx, int lddx, int dx_col0) {
  for (int i = threadIdx.x; i < n_out * K; i += BT_THREADS) {
    const int n = i / K, k = i - n * K;
    float s = 0.f;
#pragma unroll 8
    for (int r = 0; r < BT_ROWS; ++r) s = fmaf(g[r * ldg + g_col0 + n], x[r * ldx + x_col0 + k], s);
    atomicAdd(gW + n * K + k, s);
  }
  if (threadIdx.x < n_out) {
    float s = 0.f;
    for (int r = 0; r < BT_ROWS; ++r) s += g[r * ldg + g_col0 + threadIdx.x];
    atomicAdd(gb + threadIdx.x, s);
  }
  if (dx != nullptr) {
    for (int i = threadIdx.x; i < BT_ROWS * K; i += BT_THREADS) {
      const int r = i / K, k = i - r * K;
      float s = ACCUM ? dx[r * lddx + dx_col0 + k] : 0.f;
      for (int n = 0; n < n_out; ++n) s = fmaf(g[r * ldg + g_col0 + n], __ldg(W + n * K + k), s);
      dx[r * lddx + dx_col0 + k] = s;
    }
  }
}

__global__ void __launch_bounds__(BT_THREADS, 1) k_mlp_bwd_fp32(const __grid_constant__ BwdParams P) {
  extern __shared__ __align__(16) float smem[];
  float* bufA = smem;                        // [64][256]
  float* bufB = bufA + BT_ROWS * LDT;        // [64][256]
  float* bufX = bufB + BT_ROWS * LDT;        // [64][256]
  float* s_pe = bufX + BT_ROWS * LDT;        // [64][64]
  float* s_dir = s_pe + BT_ROWS * LDPE;      // [64][32] (+ pad)
  float* s_g = s_dir + BT_ROWS * LDDIR + 32; // [64][8]: g_sigma, g_albedo_pre[3], g_shading_pre, g_residual_pre[3]
  const NetLayout& L = P.L;
  const float* flat = P.a.flat;
  float* gflat = P.a.grad_flat;
  auto Wp = [&](int l) { return flat + L.flat_w[l]; };
  auto GW = [&](int l) { return gflat + L.flat_w[l]; };
  auto GB = [&](int l) { return gflat + L.flat_b[l]; };
  const MlpArgs& F = P.a.f;
  const int C = L.n_classes;
  const int64_t M = F.M;
  const int64_t n_tiles = (M + BT_ROWS - 1) / BT_ROWS;
  const float* stash = P.a.stash;

  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t row0 = tile * BT_ROWS;
    __syncthreads();
    // ---- positional encodings of the tile (same arithmetic as the forward kernel) and head gradients ----
    for (int r = threadIdx.x; r < BT_ROWS; r += BT_THREADS) {
      const int64_t m = min(row0 + r, M - 1);
      const bool valid = row0 + r < M;
      float* pe = s_pe + r * LDPE;
      float* de = s_dir + r * LDDIR;
      if (F.emb != nullptr) {
        const float* e = F.emb + m * (PE_PTS + PE_DIR);
        for (int i = 0; i < PE_PTS; ++i) pe[i] = e[i];
        for (int i = 0; i < PE_DIR; ++i) de[i] = e[PE_PTS + i];
      } else {
        float x[3], d[3];
        if (F.rays != nullptr) {
          const int64_t n = m / F.S;
          const float* ray = F.rays + n * 11;
          const float zv = F.z[m];
          for (int i = 0; i < 3; ++i) { x[i] = __fadd_rn(ray[i], __fmul_rn(ray[3 + i], zv)); d[i] = ray[8 + i]; }
        } else {
          for (int i = 0; i < 3; ++i) { x[i] = F.pts[m * 3 + i]; d[i] = F.viewdirs[m * 3 + i]; }
        }
        if (F.pe_scale != 1.f) for (int i = 0; i < 3; ++i) x[i] = __fdiv_rn(x[i], F.pe_scale);
        for (int i = 0; i < 3; ++i) { pe[i] = x[i]; de[i] = d[i]; }
        for (int k = 0; k < 10; ++k)
          for (int i = 0; i < 3; ++i) { float sv, cv; sincosf(x[i] * (float)(1 << k), &sv, &cv); pe[3 + 6 * k + i] = sv; pe[3 + 6 * k + 3 + i] = cv; }
        for (int k = 0; k < 4; ++k)
          for (int i = 0; i < 3; ++i) { float sv, cv; sincosf(d[i] * (float)(1 << k), &sv, &cv); de[3 + 6 * k + i] = sv; de[3 + 6 * k + 3 + i] = cv; }
      }
      pe[63] = 0.f;
      for (int i = PE_DIR; i < LDDIR; ++i) de[i] = 0.f;
      // raw = [rgb3, sigma, albedo3, shading, residual3, ...]; rgb = albedo*shading + residual
      float* g = s_g + r * 8;
      if (valid) {
        const float* go = P.a.grad_raw + m * P.out_ch;
        const float* ro = F.raw + m * P.out_ch;
        const float sh = ro[7];
        float g_sh = go[7];
        g[0] = go[3];
        for (int i = 0; i < 3; ++i) {
          const float alb = ro[4 + i], res = ro[8 + i];
          const float g_alb = go[4 + i] + go[i] * sh;
          const float g_res = go[8 + i] + go[i];
          g_sh += go[i] * alb;
          g[1 + i] = g_alb * alb * (1.f - alb);
          g[5 + i] = g_res * res * (1.f - res);
        }
        g[4] = g_sh * sh * (1.f - sh);
      } else {
        for (int i = 0; i < 8; ++i) g[i] = 0.f;
      }
    }
    __syncthreads();
    // ---- residual head on V = relu(views): dV -> bufA[:, :128] ---------------------------------------------
    load_tile(bufX, LDT, stash, row0, M, STASH_LD, ST_V, 128);
    __syncthreads();
    narrow_head<false>(s_g, 8, 5, 3, bufX, LDT, 0, 128, Wp(L_RES), GW(L_RES), GB(L_RES), bufA, LDT, 0);
    __syncthreads();
    if (F.endpoint) {                        // the endpoint feature rows are relu(views) themselves
      for (int i = threadIdx.x; i < BT_ROWS * 128; i += BT_THREADS) {
        const int r = i >> 7, c = i & 127;
        if (row0 + r < M) bufA[r * LDT + c] += P.a.grad_raw[(row0 + r) * P.out_ch + INRF_RAW_BASE + C + c];
      }
      __syncthreads();
    }
    relu_mask(bufA, LDT, stash, row0, M, STASH_LD, ST_V, 128);          // dZ_views
    // ---- views layer: X = [feature (256) | gamma(d) (27)] -------------------------------------------------------
    load_tile(bufX, LDT, stash, row0, M, STASH_LD, ST_FEAT, 256);
    __syncthreads();
    accum_dw(bufA, LDT, 0, 128, bufX, LDT, 256, GW(L_VIEWS), W_HID + PE_DIR, 0);
    accum_dw(bufA, LDT, 0, 128, s_dir, LDDIR, PE_DIR, GW(L_VIEWS), W_HID + PE_DIR, W_HID);
    accum_bias(bufA, LDT, 128, GB(L_VIEWS));
    dense_jk<8, false>(bufA, LDT, 0, 128, Wp(L_VIEWS), W_HID + PE_DIR, bufB, LDT);   // dFeature
    __syncthreads();
    // ---- feature layer (no activation): X = h8 -> dh in bufA --------------------------------------------------------
    load_tile(bufX, LDT, stash, row0, M, STASH_LD, ST_H + 7 * W_HID, 256);
    __syncthreads();
    accum_dw(bufB, LDT, 0, 256, bufX, LDT, 256, GW(L_FEAT), W_HID, 0);
    accum_bias(bufB, LDT, 256, GB(L_FEAT));
    dense_jk<8, false>(bufB, LDT, 0, 256, Wp(L_FEAT), W_HID, bufA, LDT);            // dh = dFeature W_f
    __syncthreads();
    // ---- sigma head: dh += g_sigma w_alpha -----------------------------------------------------------------------------
    narrow_head<true>(s_g, 8, 0, 1, bufX, LDT, 0, 256, Wp(L_ALPHA), GW(L_ALPHA), GB(L_ALPHA), bufA, LDT, 0);
    __syncthreads();
    // ---- albedo / shading heads on AS = relu(albedo1 | shading1) (tile in bufB) -------------------------------------------
    load_tile(bufB, LDT, stash, row0, M, STASH_LD, ST_AS, 256);
    __syncthreads();
    // second layers: gradients of W2 / b2 from AS, then dAS in place (the ReLU mask is AS > 0 itself)
    for (int i = threadIdx.x; i < 4 * 128; i += BT_THREADS) {
      const int n = i >> 7, k = i & 127;      // n < 3: albedo2 row n over AS[:, k]; n == 3: shading2 over AS[:, 128+k]
      float s = 0.f;
      for (int r = 0; r < BT_ROWS; ++r) s = fmaf(s_g[r * 8 + 1 + n], bufB[r * LDT + (n == 3 ? 128 : 0) + k], s);
      atomicAdd((n == 3 ? GW(L_SH2) : GW(L_ALB2) + n * 128) + k, s);
    }
    if (threadIdx.x < 4) {
      float s = 0.f;
      for (int r = 0; r < BT_ROWS; ++r) s += s_g[r * 8 + 1 + threadIdx.x];
      atomicAdd(threadIdx.x == 3 ? GB(L_SH2) : GB(L_ALB2) + threadIdx.x, s);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < BT_ROWS * 256; i += BT_THREADS) {
      const int r = i >> 8, c = i & 255;
      float v = 0.f;
      if (bufB[r * LDT + c] > 0.f) {
        if (c < 128) v = s_g[r * 8 + 1] * __ldg(Wp(L_ALB2) + c) + s_g[r * 8 + 2] * __ldg(Wp(L_ALB2) + 128 + c) + s_g[r * 8 + 3] * __ldg(Wp(L_ALB2) + 256 + c);
        else v = s_g[r * 8 + 4] * __ldg(Wp(L_SH2) + c - 128);
      }
      bufB[r * LDT + c] = v;                 // dZ of albedo1 | shading1
    }
    __syncthreads();
    accum_dw(bufB, LDT, 0, 128, bufX, LDT, 256, GW(L_ALB1), W_HID, 0);
    accum_dw(bufB, LDT, 128, 128, bufX, LDT, 256, GW(L_SH1), W_HID, 0);
    accum_bias(bufB, LDT, 128, GB(L_ALB1));
    accum_bias(bufB + 128, LDT, 128, GB(L_SH1));
    dense_jk<8, true>(bufB, LDT, 0, 128, Wp(L_ALB1), W_HID, bufA, LDT);
    __syncthreads();
    dense_jk<8, true>(bufB, LDT, 128, 128, Wp(L_SH1), W_HID, bufA, LDT);
    __syncthreads();
    // ---- semantic head (SSR): logits = W2 relu(W1 h8 + b1) + b2 ---------------------------------------------------------------
    if (C > 0) {
      load_tile(bufB, LDT, stash, row0, M, STASH_LD, ST_SEM1, 128);       // relu(sem1)
      __syncthreads();
      for (int i = threadIdx.x; i < C * 128; i += BT_THREADS) {
        const int n = i >> 7, k = i & 127;
        float s = 0.f;
        for (int r = 0; r < BT_ROWS; ++r)
          if (row0 + r < M) s = fmaf(P.a.grad_raw[(row0 + r) * P.out_ch + INRF_RAW_BASE + n], bufB[r * LDT + k], s);
        atomicAdd(GW(L_SEM2) + n * 128 + k, s);
      }
      for (int n = threadIdx.x; n < C; n += BT_THREADS) {
        float s = 0.f;
        for (int r = 0; r < BT_ROWS; ++r)
          if (row0 + r < M) s += P.a.grad_raw[(row0 + r) * P.out_ch + INRF_RAW_BASE + n];
        atomicAdd(GB(L_SEM2) + n, s);
      }
      __syncthreads();
      for (int i = threadIdx.x; i < BT_ROWS * 128; i += BT_THREADS) {
        const int r = i >> 7, k = i & 127;
        float v = 0.f;
        if (bufB[r * LDT + k] > 0.f && row0 + r < M) {
          const float* go = P.a.grad_raw + (row0 + r) * P.out_ch + INRF_RAW_BASE;
          for (int n = 0; n < C; ++n) v = fmaf(go[n], __ldg(Wp(L_SEM2) + n * 128 + k), v);
        }
        bufB[r * LDT + 128 + k] = v;        // dZ of sem1, kept beside the activation tile
      }
      __syncthreads();
      accum_dw(bufB, LDT, 128, 128, bufX, LDT, 256, GW(L_SEM1), W_HID, 0);
      accum_bias(bufB + 128, LDT, 128, GB(L_SEM1));
      dense_jk<8, true>(bufB, LDT, 128, 128, Wp(L_SEM1), W_HID, bufA, LDT);
      __syncthreads();
    }
    // ---- trunk, layers 7..0: bufA = dh_l ---------------------------------------------------------------------------------------
    float* dh = bufA;
    float* other = bufB;
    for (int l = 7; l >= 0; --l) {
      relu_mask(dh, LDT, stash, row0, M, STASH_LD, ST_H + l * W_HID, 256);       // dZ_l
      if (l >= 1) load_tile(bufX, LDT, stash, row0, M, STASH_LD, ST_H + (l - 1) * W_HID, 256);   // X_l = h_{l-1}
      __syncthreads();
      const int Kl = (l == 0) ? PE_PTS : (l == 5 ? PE_PTS + W_HID : W_HID);
      if (l == 0 || l == 5) accum_dw(dh, LDT, 0, 256, s_pe, LDPE, PE_PTS, GW(L_T0 + l), Kl, 0);
      if (l >= 1) accum_dw(dh, LDT, 0, 256, bufX, LDT, 256, GW(L_T0 + l), Kl, l == 5 ? PE_PTS : 0);
      accum_bias(dh, LDT, 256, GB(L_T0 + l));
      if (l >= 1) {
        dense_jk<8, false>(dh, LDT, 0, 256, Wp(L_T0 + l) + (l == 5 ? PE_PTS : 0), Kl, other, LDT);   // dh_{l-1}
        __syncthreads();
        float* t = dh; dh = other; other = t;
      }
    }
  }
}

int launch_mlp_bwd_fp32(const MlpBwdArgs& a, cudaStream_t st) {
  if (a.f.M == 0) return INRF_OK;
  BwdParams P;
  int rc = make_layout(a.f.variant, a.f.n_classes, &P.L);
  if (rc) return rc;
  P.a = a;
  P.out_ch = raw_channels(a.f.n_classes, a.f.endpoint);
  const size_t smem = (size_t)(3 * BT_ROWS * LDT + BT_ROWS * LDPE + BT_ROWS * LDDIR + 32 + BT_ROWS * 8) * sizeof(float);
  INRF_CUDA(cudaFuncSetAttribute(k_mlp_bwd_fp32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int dev = 0, sms = 148;
  INRF_CUDA(cudaGetDevice(&dev));
  INRF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int64_t tiles = (a.f.M + BT_ROWS - 1) / BT_ROWS;
  const int grid = (int)(tiles < sms ? tiles : sms);
  k_mlp_bwd_fp32<<<grid, BT_THREADS, smem, st>>>(P);
  INRF_LAUNCH_CHECK();
  return INRF_OK;
}

}  // namespace inrf
