// Weight packing: canonical flat fp32 parameters -> packed blob consumed by the kernels.
//   section F32 : per layer fp32 weights (transposed [K][N] when N>=128, original [N][K]
//                 for the narrow heads) + bias              -> CUDA-core path (mlp_fp32.cu)
//   section COMP: views' = views_linears.0[:, :256] @ feature_linear (feature_linear has no
//                 activation, run_nerf_helpers.py:307-309 / semantic_nerf.py:155-161, so the
//                 two linear maps compose exactly; evaluated once per pack in fp64)
//   section TCB : fp32 bias table in tensor-core epilogue order
//   section TCW : fp16 (RN) operand blocks, UMMA K-major SWIZZLE_128B, in MMA issue order
#include "common.cuh"
#include <stdarg.h>

namespace inrf {

static thread_local char g_err[512] = "no error";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return INRF_ECUDA;
}
const char* last_error() { return g_err; }

static inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

int make_layout(int variant, int n_classes, NetLayout* L) {
  if (variant != INRF_NET_OBJECT && variant != INRF_NET_SSR) { set_error("unknown network variant %d", variant); return INRF_EINVAL; }
  if (n_classes < 0 || n_classes > MAX_CLASSES) { set_error("n_classes %d outside [0,%d]", n_classes, MAX_CLASSES); return INRF_EUNSUPPORTED; }
  if (variant == INRF_NET_OBJECT && n_classes != 0) { set_error("object network has no semantic head"); return INRF_EINVAL; }
  memset(L, 0, sizeof(*L));
  L->variant = variant;
  L->n_classes = n_classes;
  L->n_layers = n_classes > 0 ? L_COUNT : L_SEM1;
  int64_t off = 0;
  for (int l = 0; l < L->n_layers; ++l) {
    LayerDims d = layer_dims(l, n_classes);
    L->flat_w[l] = off; off += (int64_t)d.K * d.N;
    L->flat_b[l] = off; off += d.N;
  }
  L->flat_count = off;
  int64_t b = 0;
  for (int l = 0; l < L->n_layers; ++l) {
    LayerDims d = layer_dims(l, n_classes);
    L->f32_wt[l] = b; b = align_up(b + (int64_t)d.K * d.N * 4, 256);
    L->f32_b[l] = b;  b = align_up(b + (int64_t)d.N * 4, 256);
  }
  // composed views' weight [128][256] + bias [128] (fp32)
  L->comp = b;
  b = align_up(b + (128 * 256 + 128) * 4, 256);
  L->tc_bias = b;
  b = align_up(b + TC_BIAS_FLOATS * 4, 1024);
  TcProgram prog;
  int rc = make_tc_program(variant, n_classes, &prog);
  if (rc) return rc;
  L->tc_blocks = b;
  L->tc_blocks_bytes = prog.bytes;
  b += prog.bytes;
  L->total_bytes = align_up(b, 1024);
  return INRF_OK;
}


// ---------------------------------------------------------------------------------
// tensor-core block program (must match the issue order in mlp_tc.cu)
// ---------------------------------------------------------------------------------
static void push(TcProgram* p, int layer, int rows, int n0, int k0, int kcols, int kind) {
  TcBlock& b = p->blk[p->n_blocks++];
  b.layer = (int16_t)layer; b.rows = (int16_t)rows; b.n0 = (int16_t)n0; b.k0 = (int16_t)k0;
  b.kcols = (int16_t)kcols; b.kind = (int16_t)kind; b.byte_off = p->bytes;
  b.bytes = (kind == 3) ? rows * 32 : rows * 128;
  p->bytes += b.bytes;
}
// bias of `layer` rows [n0, n0+128) as a K=16 operand (accumulator initialisation by MMA)
static void push_bias(TcProgram* p, int layer, int n0) { push(p, layer, 128, n0, 0, 3, 3); }
// everything pushed since `start_bytes` becomes one ring fill
static void close_fill(TcProgram* p, int start_bytes, int step, int block0) {
  p->fill_step[p->n_fills] = (int16_t)step;
  p->fill_block0[p->n_fills] = (int16_t)block0;
  p->fill_off[p->n_fills] = start_bytes;
  p->fill_bytes[p->n_fills] = p->bytes - start_bytes;
  p->n_fills++;
  p->bytes = (p->bytes + 1023) / 1024 * 1024;      // next fill starts 1024-byte aligned
}

int make_tc_program(int variant, int n_classes, TcProgram* p) {
  p->n_blocks = 0; p->bytes = 0; p->n_fills = 0;
  const bool sem = n_classes > 0;
  int f, fb, step;
#define FILL(stmts) do { f = p->bytes; fb = p->n_blocks; stmts; close_fill(p, f, step, fb); } while (0)
  // trunk layers: N = 256 operand tiles (rows 0..127 | 128..255), one fill per K chunk
  for (int l = 0; l < 8; ++l) {
    step = l;
    FILL(push_bias(p, L_T0 + l, 0); push_bias(p, L_T0 + l, 128));
    if (l == 0 || l == 5) FILL(push(p, L_T0 + l, 128, 0, 0, PE_PTS, 0); push(p, L_T0 + l, 128, 128, 0, PE_PTS, 0));
    if (l == 0) continue;
    const int base = (l == 5) ? PE_PTS : 0;
    for (int c = 0; c < 4; ++c)
      FILL(push(p, L_T0 + l, 128, 0, base + 64 * c, 64, 0); push(p, L_T0 + l, 128, 128, base + 64 * c, 64, 0));
  }
  // views' (composed with feature_linear) [| semantic hidden layer]: N = 128 [256], K = 256, then
  // the 27 direction-encoding columns (K = 32) for the views' rows only
  step = TS_VIEWS;
  FILL(push_bias(p, -1, 0); if (sem) push_bias(p, L_SEM1, 0));
  for (int c = 0; c < 4; ++c) FILL(push(p, -1, 128, 0, 64 * c, 64, 1); if (sem) push(p, L_SEM1, 128, 0, 64 * c, 64, 0));
  FILL(push(p, L_VIEWS, 128, 0, W_HID, PE_DIR, 0));
  // albedo1 | shading1: one 256-wide GEMM on the trunk output
  step = TS_ALBSH;
  FILL(push_bias(p, L_ALB1, 0); push_bias(p, L_SH1, 0));
  for (int c = 0; c < 4; ++c) FILL(push(p, L_ALB1, 128, 0, 64 * c, 64, 0); push(p, L_SH1, 128, 0, 64 * c, 64, 0));
  // residual head on relu(views'):  16 x 128 (two K chunks in one fill)
  step = TS_RES;
  FILL(for (int c = 0; c < 2; ++c) push(p, L_RES, 16, 0, 64 * c, 64, 0));
  // albedo2 (rows 0..2, K 0..127) + shading2 (row 3, K 128..255): block-diagonal 16 x 256
  step = TS_ALB2SH2;
  FILL(for (int c = 0; c < 4; ++c) push(p, L_ALB2, 16, 0, 64 * c, 64, 2));
  // semantic logits on relu(sem1): ceil16(C) x 128
  if (sem) {
    const int rows = (n_classes + 15) / 16 * 16;
    step = TS_SEM2;
    FILL(for (int c = 0; c < 2; ++c) push(p, L_SEM2, rows, 0, 64 * c, 64, 0));
  }
#undef FILL
  if (p->n_blocks > TC_MAX_BLOCKS || p->n_fills > TC_MAX_FILLS) { set_error("tc program too long"); return INRF_EUNSUPPORTED; }
  return INRF_OK;
}

// ---------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------
struct PackParams {
  NetLayout L;
  int* status;     // deferred status record (common.cuh)
};

// fp32 -> fp16 (RN) for a tensor-core operand; a finite value beyond +-65504 saturates and is reported.
// `tiny` counts non-zero values below 2^-17: there fp16 is subnormal with fewer than 8 significant bits (or zero).
__device__ __forceinline__ __half to_operand_half(float v, bool& overflow, int* tiny = nullptr) {
  __half h = __float2half_rn(v);
  if (__hisinf(h) && isfinite(v)) {
    overflow = true;
    h = __float2half_rn(v > 0.f ? 65504.f : -65504.f);
  }
  if (tiny != nullptr && v != 0.f && fabsf(v) < 7.62939453125e-06f) ++*tiny;
  return h;
}

__global__ void k_pack_f32(const float* __restrict__ flat, unsigned char* __restrict__ packed, PackParams P) {
  int l = blockIdx.y;
  if (l >= P.L.n_layers) return;
  LayerDims d = layer_dims(l, P.L.n_classes);
  const float* w = flat + P.L.flat_w[l];
  float* wt = reinterpret_cast<float*>(packed + P.L.f32_wt[l]);
  int64_t total = (int64_t)d.K * d.N;
  bool transpose = d.N >= 128;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    if (transpose) {           // out[k][n] = w[n][k]; i enumerates the output
      int k = (int)(i / d.N), n = (int)(i % d.N);
      wt[i] = w[(int64_t)n * d.K + k];
    } else {
      wt[i] = w[i];
    }
  }
  float* b = reinterpret_cast<float*>(packed + P.L.f32_b[l]);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < d.N; i += gridDim.x * blockDim.x) b[i] = flat[P.L.flat_b[l] + i];
}

// views'[n][k] = sum_j views_w[n][j] * feat_w[j][k]   (n<128, k<256, j<256), fp64 accumulate
// views'_b[n]  = sum_j views_w[n][j] * feat_b[j] + views_b[n]
__global__ void k_compose_views(const float* __restrict__ flat, unsigned char* __restrict__ packed, PackParams P) {
  const float* vw = flat + P.L.flat_w[L_VIEWS];   // [128][283]
  const float* fw = flat + P.L.flat_w[L_FEAT];    // [256][256]
  const float* fb = flat + P.L.flat_b[L_FEAT];
  const float* vb = flat + P.L.flat_b[L_VIEWS];
  float* cw = reinterpret_cast<float*>(packed + P.L.comp);
  float* cb = cw + 128 * 256;
  int n = blockIdx.x;       // 128 blocks
  int k = threadIdx.x;      // 256 threads
  // four interleaved partial sums: the dependent fp64 FMA chain, not the loads, paced the single-accumulator loop
  double a4[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll 2
  for (int j = 0; j < W_HID; j += 4) {
#pragma unroll
    for (int i = 0; i < 4; ++i) a4[i] += (double)vw[n * (W_HID + PE_DIR) + j + i] * (double)fw[(j + i) * W_HID + k];
  }
  cw[n * W_HID + k] = (float)((a4[0] + a4[1]) + (a4[2] + a4[3]));
  if (k < 32) {                        // bias: one warp, lane-strided partial sums + butterfly
    double b = 0.0;
    for (int j = k; j < W_HID; j += 32) b += (double)vw[n * (W_HID + PE_DIR) + j] * (double)fb[j];
    for (int o = 16; o > 0; o >>= 1) b += __shfl_xor_sync(0xffffffffu, b, o);
    if (k == 0) cb[n] = (float)(b + (double)vb[n]);
  }
}

// bias table layout (floats): see mlp_tc.cu TCB_* constants
__global__ void k_pack_tc_bias(const float* __restrict__ flat, unsigned char* __restrict__ packed, PackParams P) {
  float* t = reinterpret_cast<float*>(packed + P.L.tc_bias);
  const float* cb = reinterpret_cast<const float*>(packed + P.L.comp) + 128 * 256;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= TC_BIAS_FLOATS) return;
  float v = 0.f;
  if (i < 2048) v = flat[P.L.flat_b[i / 256] + (i % 256)];
  else if (i < 2176) v = cb[i - 2048];
  else if (i < 2304) v = (P.L.n_classes > 0) ? flat[P.L.flat_b[L_SEM1] + (i - 2176)] : 0.f;
  else if (i < 2432) v = flat[P.L.flat_b[L_ALB1] + (i - 2304)];
  else if (i < 2560) v = flat[P.L.flat_b[L_SH1] + (i - 2432)];
  else if (i < 2816) v = flat[P.L.flat_w[L_ALPHA] + (i - 2560)];
  else if (i == 2816) v = flat[P.L.flat_b[L_ALPHA]];
  else if (i < 2820) v = flat[P.L.flat_b[L_ALB2] + (i - 2817)];
  else if (i == 2820) v = flat[P.L.flat_b[L_SH2]];
  else if (i < 2824) v = flat[P.L.flat_b[L_RES] + (i - 2821)];
  else if (i < 2824 + P.L.n_classes) v = flat[P.L.flat_b[L_SEM2] + (i - 2824)];
  else if (i >= 4096 && i < 4096 + 512) {          // residual head, float4 per hidden unit k: (w0k, w1k, w2k, 0)
    const int k = (i - 4096) >> 2, c = (i - 4096) & 3;
    v = c < 3 ? flat[P.L.flat_w[L_RES] + c * 128 + k] : 0.f;
  } else if (i >= 4608 && i < 4608 + 1024) {       // albedo2 (k < 128: w0k, w1k, w2k, 0) | shading2 (k >= 128: wk, 0, 0, 0)
    const int k = (i - 4608) >> 2, c = (i - 4608) & 3;
    if (k < 128) v = c < 3 ? flat[P.L.flat_w[L_ALB2] + c * 128 + k] : 0.f;
    else v = c == 0 ? flat[P.L.flat_w[L_SH2] + (k - 128)] : 0.f;
  }
  t[i] = v;
}

__global__ void k_pack_tc_blocks(const float* __restrict__ flat, unsigned char* __restrict__ packed, PackParams P,
                                 const __grid_constant__ TcProgram prog) {
  const TcBlock b = prog.blk[blockIdx.x];
  __half* dst = reinterpret_cast<__half*>(packed + P.L.tc_blocks + b.byte_off);
  const float* comp = reinterpret_cast<const float*>(packed + P.L.comp);
  bool overflow = false;
  if (b.kind == 3) {
    // no-swizzle K-major core matrices: 8 rows x 8 halves = 128 B contiguous; K halves 128 B apart
    // (LBO), 8-row groups 256 B apart (SBO).  bias = hi + lo + lo2 in three fp16 columns.
    for (int r = threadIdx.x; r < 128; r += blockDim.x) {
      float bv = (b.layer < 0) ? comp[128 * W_HID + b.n0 + r] : flat[P.L.flat_b[b.layer] + b.n0 + r];
      __half hi = to_operand_half(bv, overflow);
      float r1 = bv - __half2float(hi);
      __half lo = __float2half_rn(r1);
      __half lo2 = __float2half_rn(r1 - __half2float(lo));
      for (int k = 0; k < 16; ++k) {
        __half v = k == 0 ? hi : (k == 1 ? lo : (k == 2 ? lo2 : __float2half_rn(0.f)));
        dst[((r >> 3) * 256 + (k >> 3) * 128 + (r & 7) * 16 + (k & 7) * 2) >> 1] = v;
      }
    }
    if (__syncthreads_or(overflow) && threadIdx.x == 0) status_raise(P.status, DST_F16_WEIGHT, blockIdx.x);
    return;
  }
  int tiny = 0, nonzero = 0;
  for (int e = threadIdx.x; e < b.rows * 64; e += blockDim.x) {
    int r = e >> 6, kk = e & 63;
    float v = 0.f;
    if (b.kind == 0) {
      LayerDims d = layer_dims(b.layer, P.L.n_classes);
      int n = b.n0 + r;
      if (n < d.N && kk < b.kcols) v = flat[P.L.flat_w[b.layer] + (int64_t)n * d.K + b.k0 + kk];
    } else if (b.kind == 1) {
      v = comp[r * W_HID + b.k0 + kk];
    } else {   // block-diagonal albedo2 / shading2 over K = [relu(albedo1) | relu(shading1)]
      int k = b.k0 + kk;
      if (r < 3 && k < 128) v = flat[P.L.flat_w[L_ALB2] + r * 128 + k];
      else if (r == 3 && k >= 128) v = flat[P.L.flat_w[L_SH2] + (k - 128)];
    }
    // UMMA K-major SWIZZLE_128B: 8-row atom = 1024 B, row = 128 B, 16 B unit index ^ (row & 7)
    int unit = kk >> 3, within = kk & 7;
    int off_bytes = (r >> 3) * 1024 + (r & 7) * 128 + ((unit ^ (r & 7)) << 4) + within * 2;
    dst[off_bytes >> 1] = to_operand_half(v, overflow, &tiny);
    nonzero += (v != 0.f);
  }
  if (__syncthreads_or(overflow) && threadIdx.x == 0) status_raise(P.status, DST_F16_WEIGHT, blockIdx.x, 0);
  // a block whose weights mostly sit below fp16's normal range would silently lose the layer: report it as well
  __shared__ int s_cnt[2];
  if (threadIdx.x == 0) s_cnt[0] = s_cnt[1] = 0;
  __syncthreads();
  for (int o = 16; o > 0; o >>= 1) { tiny += __shfl_xor_sync(0xffffffffu, tiny, o); nonzero += __shfl_xor_sync(0xffffffffu, nonzero, o); }
  if ((threadIdx.x & 31) == 0) { atomicAdd(&s_cnt[0], tiny); atomicAdd(&s_cnt[1], nonzero); }
  __syncthreads();
  if (threadIdx.x == 0 && s_cnt[0] * 4 > s_cnt[1]) status_raise(P.status, DST_F16_WEIGHT, blockIdx.x, 1);
}

int pack_weights(const float* flat, int variant, int n_classes, void* packed, int64_t packed_bytes, cudaStream_t st) {
  PackParams P;
  int rc = make_layout(variant, n_classes, &P.L);
  if (rc) return rc;
  if (packed_bytes < P.L.total_bytes) { set_error("packed buffer too small: %lld < %lld", (long long)packed_bytes, (long long)P.L.total_bytes); return INRF_EINVAL; }
  TcProgram prog;
  rc = make_tc_program(variant, n_classes, &prog);
  if (rc) return rc;
  P.status = status_flag_dev();
  if (P.status == nullptr) return INRF_ECUDA;
  unsigned char* out = static_cast<unsigned char*>(packed);
  k_pack_f32<<<dim3(32, P.L.n_layers), 256, 0, st>>>(flat, out, P);
  INRF_LAUNCH_CHECK();
  k_compose_views<<<128, 256, 0, st>>>(flat, out, P);
  INRF_LAUNCH_CHECK();
  k_pack_tc_bias<<<TC_BIAS_FLOATS / 256, 256, 0, st>>>(flat, out, P);
  INRF_LAUNCH_CHECK();
  k_pack_tc_blocks<<<prog.n_blocks, 256, 0, st>>>(flat, out, P, prog);   // 2 KB table by value
  INRF_LAUNCH_CHECK();
  return INRF_OK;
}

}  // namespace inrf
