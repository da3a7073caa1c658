// Tensor-core backward pass of the intrinsic field network (training path, INRF_PREC_TC).
//
// The forward (mlp_tc.cu, STASH instantiation) leaves every post-activation tile of a 128-sample tile in HBM as
// a 16 KB image of the shared-memory UMMA operand chunk (common.cuh: IMG_*).  The backward is a short sequence
// of launches over those images - nothing is recomputed, and no operand is ever re-laid-out:
//
//   k_bwd_heads   per sample: d(loss)/d(pre-sigmoid heads, sigma, logits) from grad_raw and raw -> image G
//   k_gemm_dx     dZ_out = relu'(H_out) * (sum_k dZ_in[k] W_k)   tcgen05 GEMM, A = gradient images (K-major),
//                 B = transposed weight tiles (packed per call), epilogue masks with the stashed activation
//                 image and writes the next gradient image; 10 launches walk heads -> trunk layer 7 -> ... -> 0
//   k_gemm_dw     dW = dZ^T X for every layer in ONE launch: both operands are read MN-major from the same
//                 images (rows of the tile are the K dimension), split-K over tiles into partial tiles that
//                 k_dw_reduce adds in a fixed order (bit-reproducible gradients); an all-ones operand tile adds the
//                 bias gradient as 16 extra accumulator columns
//   k_unfold_comp views' = views_linears.0[:, :256] o feature_linear was composed at pack time (pack.cu); its
//                 gradient is unfolded onto the two original matrices.
//
// Arithmetic: fp16 operands (RN) x fp32 accumulation, like the forward.  Gradients are scaled by a power of two
// chosen from max|grad_raw| (device-side, no host sync) so that fp16 holds them without underflow; the scale is
// removed in fp32 when dW is written.  Reference semantics: autograd through NeRF.forward
// (object_level/run_nerf_helpers.py:284-325) / Semantic_NeRF.forward (SSR/models/semantic_nerf.py:123-181).
#include "common.cuh"
#include "tc_ptx.cuh"
#include <stdlib.h>

namespace inrf {
namespace ttc {
using namespace ptx;

constexpr int TILE = 128;

struct ImgRef {                 // image of tile t: base + (t * tslots + slot) * IMG_BYTES
  const unsigned char* base;
  int tslots, slot;
  __host__ __device__ const unsigned char* at(int64_t t, int c = 0) const { return base + (t * tslots + slot + c) * (int64_t)IMG_BYTES; }
};

__device__ int g_dbg[8];

// ------------------------------------------------------------------------------------------
// gradient scale: S = 2^(7 - e) with max|grad_raw| in [2^(e-1), 2^e)  ->  scaled maximum in [64, 128)
// ------------------------------------------------------------------------------------------
__global__ void k_amax(const float* __restrict__ g, int64_t n, unsigned int* amax_bits) {
  float m = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = fabsf(g[i]);
    if (v < 3.0e38f) m = fmaxf(m, v);             // inf / NaN gradients do not pick the scale
  }
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(amax_bits, __float_as_uint(m));
}
__device__ __forceinline__ float grad_scale(const unsigned int* amax_bits) {
  const float a = __uint_as_float(*amax_bits);
  if (!(a > 0.f)) return 1.f;
  int e;
  frexpf(a, &e);
  return ldexpf(1.f, 7 - e);
}

__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
  __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void st_global_v4(unsigned char* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.global.v4.b32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 ld_global_nc_v4(const unsigned char* p) {
  uint4 r;
  asm volatile("ld.global.nc.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
// byte offset of 16-byte unit `u` of row `row` inside a chunk image
__device__ __forceinline__ uint32_t img_unit(int row, int u) {
  return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((u ^ (row & 7)) << 4));
}

// ------------------------------------------------------------------------------------------
// head gradients -> image G (2 chunks): cols 0..2 albedo, 3 shading, 4..6 residual (pre-sigmoid), 7 sigma,
// 8..8+C semantic logits.  rgb = albedo*shading + residual (run_nerf_helpers.py:320), so
//   d albedo = g_rgb*shading + g_albedo, d shading = sum g_rgb*albedo + g_shading, d residual = g_rgb + g_residual.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_bwd_heads(const float* __restrict__ grad_raw, const float* __restrict__ raw, int out_ch,
                                                   int C, int64_t M, unsigned char* work, const unsigned int* amax_bits) {
  const int row = threadIdx.x;
  const int64_t tile = blockIdx.x, m = tile * TILE + row;
  const float S = grad_scale(amax_bits);
  float g[128];
#pragma unroll
  for (int i = 0; i < 128; ++i) g[i] = 0.f;
  if (m < M) {
    const float* gr = grad_raw + m * out_ch;
    const float* rw = raw + m * out_ch;
    const float sh = rw[7];
    float dsh = gr[7];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const float alb = rw[4 + i], res = rw[8 + i], grgb = gr[i];
      g[i] = (grgb * sh + gr[4 + i]) * alb * (1.f - alb) * S;
      g[4 + i] = (grgb + gr[8 + i]) * res * (1.f - res) * S;
      dsh = fmaf(grgb, alb, dsh);
    }
    g[3] = dsh * sh * (1.f - sh) * S;
    g[7] = gr[3] * S;
#pragma unroll
    for (int c = 0; c < MAX_CLASSES; ++c)
      if (c < C) g[8 + c] = gr[INRF_RAW_BASE + c] * S;
  }
  unsigned char* img = work + (tile * IMG_BWD_SLOTS + IB_G) * (int64_t)IMG_BYTES;
#pragma unroll
  for (int ch = 0; ch < 2; ++ch)
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const float* s = g + ch * 64 + u * 8;
      st_global_v4(img + ch * IMG_BYTES + img_unit(row, u), pack_h2(s[0], s[1]), pack_h2(s[2], s[3]), pack_h2(s[4], s[5]), pack_h2(s[6], s[7]));
    }
}

// ------------------------------------------------------------------------------------------
// transposed weight tiles for the dX GEMMs: tile = [N rows (input unit n)] x [64 K columns (output unit k)],
// K-major SWIZZLE_128B, value = d(pre-activation k of the source)/d(input n)
// ------------------------------------------------------------------------------------------
enum { BW_PLAIN = 0, BW_COMP, BW_HEAD_AS, BW_HEAD_V, BW_HEAD_SEM, BW_ALPHA };
struct BwTile { int16_t kind, layer, out0, in0, N, pad; int32_t off; };
constexpr int BW_MAX_TILES = 48;
struct BwProgram { int n; int bytes; BwTile t[BW_MAX_TILES]; };

__global__ void k_pack_bwd(const float* __restrict__ flat, const unsigned char* __restrict__ packed, NetLayout L,
                           const __grid_constant__ BwProgram prog, unsigned char* blob) {
  const BwTile t = prog.t[blockIdx.x];
  __half* dst = reinterpret_cast<__half*>(blob + t.off);
  const float* comp = reinterpret_cast<const float*>(packed + L.comp);
  for (int e = blockIdx.y * blockDim.x + threadIdx.x; e < t.N * 64; e += blockDim.x * gridDim.y) {   // grid.y slices a tile
    const int n = e >> 6, k = e & 63;
    float v = 0.f;
    switch (t.kind) {
      case BW_PLAIN: {
        const LayerDims d = layer_dims(t.layer, L.n_classes);
        const int o = t.out0 + k, i = t.in0 + n;
        if (o < d.N && i < d.K) v = flat[L.flat_w[t.layer] + (int64_t)o * d.K + i];
      } break;
      case BW_COMP: v = comp[(t.out0 + k) * W_HID + n]; break;
      case BW_HEAD_AS:          // G cols 0..2 -> albedo hidden units (n < 128), col 3 -> shading hidden units
        if (n < 128 && k < 3) v = flat[L.flat_w[L_ALB2] + k * 128 + n];
        else if (n >= 128 && k == 3) v = flat[L.flat_w[L_SH2] + (n - 128)];
        break;
      case BW_HEAD_V: if (k >= 4 && k < 7) v = flat[L.flat_w[L_RES] + (k - 4) * 128 + n]; break;
      case BW_HEAD_SEM: {
        const int c = t.out0 + k - 8;
        if (c >= 0 && c < L.n_classes) v = flat[L.flat_w[L_SEM2] + c * 128 + n];
      } break;
      case BW_ALPHA: if (k == 7) v = flat[L.flat_w[L_ALPHA] + n]; break;
    }
    dst[(img_unit(n, k >> 3) + (k & 7) * 2) >> 1] = __float2half_rn(v);
  }
}

// ------------------------------------------------------------------------------------------
// dX GEMM: out[128 x N] = mask * (sum_kc A_kc[128 x 64] . B_kc[N x 64]^T (+ addend)), persistent over tiles.
// warps 0-7 epilogue (TMEM lanes 32*(w&3).., 32-column half (w>>2) of every 64-column chunk), warp 8 producer,
// warp 9 MMA issuer; 3-stage ring of (A chunk 16 KB + B tile <= 32 KB); two 256-column accumulators so that the
// epilogue of tile t overlaps the MMAs of tile t+1.  The epilogue touches global memory with ONE coalesced word per
// thread and chunk (the ReLU mask bits the forward stored) and builds the output images in shared memory, from where a
// single bulk store writes them: the earlier per-thread 16-byte image loads / stores cost 2 x 4096 LSU wavefronts per
// tile - 4.4 of the 5.9 us a tile took, against 1.3 us of MMAs.
// ------------------------------------------------------------------------------------------
constexpr int DX_MAX_K = 10, DX_NS = 3, DX_STAGE = IMG_BYTES + 32768, DX_THREADS = 320;
constexpr int DX_OUT = DX_NS * DX_STAGE;                 // staging of the tile's output images (4 x 16 KB) for the bulk store
constexpr int DX_BAR = DX_OUT + 4 * IMG_BYTES, DX_SMEM = DX_BAR + 256;
struct DxParams {
  int n_k, N, n_tiles, n_iter;
  ImgRef a[DX_MAX_K];
  const unsigned char* b;
  ImgRef mask, out;              // mask: the forward stash slot whose ReLU decisions gate this GEMM's output (bit words, IS_MASK)
  const float* addend;           // optional fp32 rows (first column already applied), row stride addend_ld
  int addend_ld;
  int64_t M;
  const unsigned int* amax_bits;
  int* dbg;
  int* status;                   // deferred status record (common.cuh)
};
enum { DXB_FULL = 0, DXB_EMPTY = DX_NS, DXB_ACC_FULL = 2 * DX_NS, DXB_ACC_EMPTY = 2 * DX_NS + 2, DXB_COUNT = 2 * DX_NS + 4 };

#define TTC_WAIT(addr, par, code)                                          \
  do {                                                                     \
    if (!dead && !mbar_wait((addr), (par))) { dead = true; if (atomicCAS(P.dbg, 0, (code)) == 0) { P.dbg[1] = blockIdx.x; P.dbg[2] = threadIdx.x; \
      status_raise(P.status, DST_WATCHDOG, (code), threadIdx.x >> 5, -1, blockIdx.x, 2); } } \
  } while (0)

__global__ void __launch_bounds__(DX_THREADS, 1) k_gemm_dx(const __grid_constant__ DxParams P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sb = smem_u32(smem);
  const uint32_t bar0 = sb + DX_BAR;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + DX_BAR + 8 * DXB_COUNT);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  bool dead = false;
  if (threadIdx.x == 0) {
    for (int s = 0; s < DX_NS; ++s) { mbar_init(bar0 + 8 * (DXB_FULL + s), 1); mbar_init(bar0 + 8 * (DXB_EMPTY + s), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(bar0 + 8 * (DXB_ACC_FULL + i), 1); mbar_init(bar0 + 8 * (DXB_ACC_EMPTY + i), 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 9) tmem_alloc512(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t b_bytes = (uint32_t)P.N * 128u;

  if (warp == 8) {                                     // ---- producer
    const bool leader = elect_one();
    int slot = 0;
    uint32_t par = (1u << DX_NS) - 1;                   // "empty" barriers start released
    for (int it = 0; it < P.n_iter; ++it) {
      const int64_t tile = (int64_t)it * gridDim.x + blockIdx.x;
      if (tile >= P.n_tiles) break;
      for (int kc = 0; kc < P.n_k; ++kc) {
        TTC_WAIT(bar0 + 8 * (DXB_EMPTY + slot), (par >> slot) & 1u, 1);
        par ^= 1u << slot;
        if (leader && !dead) {
          const uint32_t full = bar0 + 8 * (DXB_FULL + slot), dst = sb + slot * DX_STAGE;
          mbar_expect_tx(full, IMG_BYTES + b_bytes);
          bulk_g2s(dst, P.a[kc].at(tile), IMG_BYTES, full);
          bulk_g2s(dst + IMG_BYTES, P.b + (int64_t)kc * b_bytes, b_bytes, full);
        }
        __syncwarp();
        slot = (slot + 1 == DX_NS) ? 0 : slot + 1;
      }
    }
  } else if (warp == 9) {                              // ---- MMA issuer
    const bool leader = elect_one();
    int slot = 0, buf = 0;
    uint32_t par_full = 0, par_acc = 3u;               // accumulators start drained
    const uint32_t idesc = idesc_f16(128, P.N, 0, 0);
    for (int it = 0; it < P.n_iter; ++it) {
      const int64_t tile = (int64_t)it * gridDim.x + blockIdx.x;
      if (tile >= P.n_tiles) break;
      TTC_WAIT(bar0 + 8 * (DXB_ACC_EMPTY + buf), (par_acc >> buf) & 1u, 2);
      par_acc ^= 1u << buf;
      tc_fence_after();
      for (int kc = 0; kc < P.n_k; ++kc) {
        TTC_WAIT(bar0 + 8 * (DXB_FULL + slot), (par_full >> slot) & 1u, 3);
        par_full ^= 1u << slot;
        tc_fence_after();
        if (leader && !dead) {
          const uint64_t ad = desc_k_sw128(sb + slot * DX_STAGE), bd = desc_k_sw128(sb + slot * DX_STAGE + IMG_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k) tc_mma(tmem + buf * 256, ad + 2 * k, bd + 2 * k, idesc, (kc | k) ? 1u : 0u);
          tc_commit(bar0 + 8 * (DXB_EMPTY + slot));
        }
        __syncwarp();
        slot = (slot + 1 == DX_NS) ? 0 : slot + 1;
      }
      if (leader && !dead) tc_commit(bar0 + 8 * (DXB_ACC_FULL + buf));
      __syncwarp();
      buf ^= 1;
    }
  } else {                                             // ---- epilogue
    const int q = warp & 3, jj = warp >> 2, row = q * 32 + lane;
    const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
    const float S = P.addend ? grad_scale(P.amax_bits) : 1.f;
    uint32_t uoff[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) uoff[u] = img_unit(row, jj * 4 + u);
    int buf = 0;
    uint32_t par_acc = 0;
    const int n_chunks = P.N >> 6;
    const uint32_t stage = sb + DX_OUT;
    for (int it = 0; it < P.n_iter; ++it) {
      const int64_t tile = (int64_t)it * gridDim.x + blockIdx.x;
      if (tile >= P.n_tiles) break;
      const int64_t m = tile * TILE + row;
      // ReLU masks of the tile: one word per chunk, requested before the accumulator is waited for
      const uint32_t* mw = reinterpret_cast<const uint32_t*>(P.mask.base + (tile * P.mask.tslots + IS_MASK) * (int64_t)IMG_BYTES) +
                           (P.mask.slot - IS_H) * 256 + jj * 128 + row;
      uint32_t mbits[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) mbits[c] = (c < n_chunks) ? __ldg(mw + c * 256) : 0u;
      TTC_WAIT(bar0 + 8 * (DXB_ACC_FULL + buf), (par_acc >> buf) & 1u, 4);
      par_acc ^= 1u << buf;
      tc_fence_after();
      // the previous tile's bulk store has finished reading the staging buffer
      if (threadIdx.x == 0) bulk_wait_read0();
      asm volatile("bar.sync 1, 256;" ::: "memory");
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (c >= n_chunks) break;
        uint32_t v[32];
        tmem_ld32(lane_addr + buf * 256 + c * 64 + jj * 32, v);
        tmem_ld_wait();
        float f[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
        if (P.addend != nullptr && m < P.M) {
          const float* ad = P.addend + m * P.addend_ld + c * 64 + jj * 32;
#pragma unroll
          for (int i = 0; i < 32; ++i) f[i] = fmaf(ad[i], S, f[i]);
        }
        const uint32_t mb = mbits[c];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          uint32_t pk[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float lo = ((mb >> (4 * u + i)) & 1u) ? f[8 * u + 2 * i] : 0.f;            // bit j: column 2j
            const float hi = ((mb >> (16 + 4 * u + i)) & 1u) ? f[8 * u + 2 * i + 1] : 0.f;   // bit 16 + j: column 2j + 1
            pk[i] = pack_h2(lo, hi);
          }
          st_shared_v4(stage + c * IMG_BYTES + uoff[u], pk[0], pk[1], pk[2], pk[3]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar0 + 8 * (DXB_ACC_EMPTY + buf));       // the accumulator is drained: next tile's MMAs may start
      fence_async_smem();
      asm volatile("bar.sync 1, 256;" ::: "memory");                      // all 8 warps have written (and fenced) their part
      if (threadIdx.x == 0 && !dead)
        bulk_s2g(const_cast<unsigned char*>(P.out.at(tile, 0)), stage, (uint32_t)n_chunks * IMG_BYTES);
      buf ^= 1;
    }
    if (threadIdx.x == 0) bulk_wait_all0();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc512(tmem);
}

// ------------------------------------------------------------------------------------------
// The seven trunk dX GEMMs (layers 7..1: dZ_{l-1} = relu'(H_{l-1}) * (dZ_l W_l)) as ONE persistent launch.  A tile's dZ
// never returns from HBM between layers: the epilogue builds dZ_{l-1} in shared memory as the operand image (in place
// over dZ_l, whose MMAs are complete), the next layer's MMAs read it from there, and a bulk store sends the same bytes to
// the gradient-image slot the dW GEMM reads later.  Two tiles (X, Y) are in flight per CTA - one 64 KB image buffer and
// one 256-column accumulator each - so that the epilogue of one overlaps the MMAs of the other; the weight tiles stream
// through a 3-stage ring of 32 KB (from L2: 896 KB per network).  warps 0-7 epilogue, 8 producer, 9 issuer.
// Measured (B200, 1536 tiles): 231 us against 7 x 42 us for the per-layer launches = 2.75 us per tile-layer.  The bound
// is SHARED-MEMORY bandwidth, as for the SS forward: per tile-layer the MMAs read A 64 KB + B 128 KB, the ring takes
// 128 KB of weight writes, the epilogue stores 64 KB and the bulk store reads them again - 448 KB = 3500 cycles at
// 128 B/clk against 2176 cycles of MMAs.  It is not HBM (0.70 GB of gradient images leave at 3.0 TB/s; a fill kernel
// writes 7.4 TB/s on this board) and not L2: CL = 2 (INRF_DX_CHAIN=2, A/B only), where the two CTAs of a cluster run
// in lock-step and share every weight fetch by multicast, halves the 6 TB/s of L2 -> SM weight traffic and changes
// nothing (240 us) - the multicast still writes every byte into both rings.  Lock-step needs the same trip counts, so
// a CTA without a (second) tile computes on a clamped one and drops the result.
// ------------------------------------------------------------------------------------------
constexpr int CH_LAYERS = 7, CH_NSB = 3, CH_BUF = 4 * IMG_BYTES;
constexpr int CH_RING = 2 * CH_BUF, CH_BAR = CH_RING + CH_NSB * 32768, CH_SMEM = CH_BAR + 256;
struct ChainParams {
  int n_tiles, n_pairs, n_iter;
  ImgRef in;                         // dZ_7: 4 chunk images per tile
  const unsigned char* b[CH_LAYERS]; // per layer: 4 K-major weight tiles of 32 KB (k_pack_bwd)
  ImgRef mask[CH_LAYERS];            // forward stash slot whose ReLU bits gate the layer's output
  ImgRef out[CH_LAYERS];             // gradient-image slot of the layer's output (4 chunks)
  int* dbg;
  int* status;
};
enum { CHB_BFULL = 0, CHB_BEMPTY = CH_NSB, CHB_IN_FULL = 2 * CH_NSB, CHB_A_READY = 2 * CH_NSB + 2, CHB_ACC_FULL = 2 * CH_NSB + 4,
       CHB_BUF_FREE = 2 * CH_NSB + 6, CHB_COUNT = 2 * CH_NSB + 8 };

template <int CL>
__global__ void __launch_bounds__(DX_THREADS, 1) k_dx_chain(const __grid_constant__ ChainParams P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sb = smem_u32(smem);
  const uint32_t bar0 = sb + CH_BAR;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + CH_BAR + 8 * CHB_COUNT);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (CL > 1) ? (int)cluster_ctarank() : 0;
  bool dead = false;
  if (threadIdx.x == 0) {
    for (int s = 0; s < CH_NSB; ++s) { mbar_init(bar0 + 8 * (CHB_BFULL + s), 1); mbar_init(bar0 + 8 * (CHB_BEMPTY + s), CL); }
    for (int t = 0; t < 2; ++t) {
      mbar_init(bar0 + 8 * (CHB_IN_FULL + t), 1); mbar_init(bar0 + 8 * (CHB_A_READY + t), 8);
      mbar_init(bar0 + 8 * (CHB_ACC_FULL + t), 1); mbar_init(bar0 + 8 * (CHB_BUF_FREE + t), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 9) tmem_alloc512(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();                      // the peer's barriers exist before any multicast / remote arrive
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  // pair `it` of this CTA: tiles 2p and 2p + 1 with p = it * gridDim.x + blockIdx.x.  Every CTA runs n_iter pairs of two
  // tiles (cluster lock-step); tiles past the end are computed on the last tile's data and not stored.
  auto pair_of = [&](int it) { return (int64_t)it * gridDim.x + blockIdx.x; };
  auto clamp_tile = [&](int64_t t) { return t < P.n_tiles ? t : (int64_t)P.n_tiles - 1; };
  constexpr int nt = 2;

  if (warp == 8) {                                     // ---- producer: tile inputs, then the weight tiles layer by layer
    const bool leader = elect_one();
    int slot = 0;
    uint32_t par_empty = (1u << CH_NSB) - 1, par_free = 3u;        // "empty" / "free" barriers start released
    for (int it = 0; it < P.n_iter; ++it) {
      const int64_t p = pair_of(it);
      if (CL == 1 && p >= P.n_pairs) break;
      for (int t = 0; t < nt; ++t) {
        TTC_WAIT(bar0 + 8 * (CHB_BUF_FREE + t), (par_free >> t) & 1u, 11);
        par_free ^= 1u << t;
        if (leader && !dead) {
          const uint32_t full = bar0 + 8 * (CHB_IN_FULL + t);
          mbar_expect_tx(full, CH_BUF);
          bulk_g2s(sb + t * CH_BUF, P.in.at(clamp_tile(2 * p + t)), CH_BUF, full);
        }
        __syncwarp();
      }
      for (int li = 0; li < CH_LAYERS; ++li)
        for (int t = 0; t < nt; ++t)
          for (int kc = 0; kc < 4; ++kc) {
            TTC_WAIT(bar0 + 8 * (CHB_BEMPTY + slot), (par_empty >> slot) & 1u, 12);
            par_empty ^= 1u << slot;
            if (leader && !dead) {
              const uint32_t full = bar0 + 8 * (CHB_BFULL + slot), dst = sb + CH_RING + slot * 32768;
              mbar_expect_tx(full, 32768);
              if (CL == 1) bulk_g2s(dst, P.b[li] + kc * 32768, 32768, full);
              else bulk_g2s_mc(dst + rank * 16384, P.b[li] + kc * 32768 + rank * 16384, 16384, full, (uint16_t)3);
            }
            __syncwarp();
            slot = (slot + 1 == CH_NSB) ? 0 : slot + 1;
          }
    }
  } else if (warp == 9) {                              // ---- MMA issuer
    const bool leader = elect_one();
    int slot = 0;
    uint32_t par_full = 0, par_in = 0, par_ready = 0;
    const uint32_t idesc = idesc_f16(128, 256, 0, 0);
    for (int it = 0; it < P.n_iter; ++it) {
      if (CL == 1 && pair_of(it) >= P.n_pairs) break;
      for (int li = 0; li < CH_LAYERS; ++li)
        for (int t = 0; t < nt; ++t) {
          if (li == 0) { TTC_WAIT(bar0 + 8 * (CHB_IN_FULL + t), (par_in >> t) & 1u, 13); par_in ^= 1u << t; }
          else { TTC_WAIT(bar0 + 8 * (CHB_A_READY + t), (par_ready >> t) & 1u, 14); par_ready ^= 1u << t; }
          tc_fence_after();
          for (int kc = 0; kc < 4; ++kc) {
            TTC_WAIT(bar0 + 8 * (CHB_BFULL + slot), (par_full >> slot) & 1u, 15);
            par_full ^= 1u << slot;
            tc_fence_after();
            if (leader && !dead) {
              const uint64_t ad = desc_k_sw128(sb + t * CH_BUF + kc * IMG_BYTES), bd = desc_k_sw128(sb + CH_RING + slot * 32768);
#pragma unroll
              for (int k = 0; k < 4; ++k) tc_mma(tmem + t * 256, ad + 2 * k, bd + 2 * k, idesc, (kc | k) ? 1u : 0u);
              if (CL == 1) tc_commit(bar0 + 8 * (CHB_BEMPTY + slot));
              else tc_commit_mc(bar0 + 8 * (CHB_BEMPTY + slot), (uint16_t)3);      // both CTAs' producers learn that this CTA is done with the slot
            }
            __syncwarp();
            slot = (slot + 1 == CH_NSB) ? 0 : slot + 1;
          }
          if (leader && !dead) tc_commit(bar0 + 8 * (CHB_ACC_FULL + t));
          __syncwarp();
        }
      // the last layer's output of both tiles still has to be consumed by the epilogue: its A_READY arrivals of that
      // layer are waited for here so that the parity bookkeeping stays aligned with the next pair
      for (int t = 0; t < nt; ++t) { TTC_WAIT(bar0 + 8 * (CHB_A_READY + t), (par_ready >> t) & 1u, 16); par_ready ^= 1u << t; }
    }
  } else {                                             // ---- epilogue (both tiles, alternately)
    const int q = warp & 3, jj = warp >> 2, row = q * 32 + lane;
    const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
    uint32_t uoff[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) uoff[u] = img_unit(row, jj * 4 + u);
    uint32_t par_acc = 0;
    for (int it = 0; it < P.n_iter; ++it) {
      const int64_t p = pair_of(it);
      if (CL == 1 && p >= P.n_pairs) break;
      for (int li = 0; li < CH_LAYERS; ++li)
        for (int t = 0; t < nt; ++t) {
          const bool live = 2 * p + t < P.n_tiles;
          const int64_t tile = clamp_tile(2 * p + t);
          const uint32_t stage = sb + t * CH_BUF;
          const uint32_t* mw = reinterpret_cast<const uint32_t*>(P.mask[li].base + (tile * P.mask[li].tslots + IS_MASK) * (int64_t)IMG_BYTES) +
                               (P.mask[li].slot - IS_H) * 256 + jj * 128 + row;
          uint32_t mbits[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) mbits[c] = __ldg(mw + c * 256);
          TTC_WAIT(bar0 + 8 * (CHB_ACC_FULL + t), (par_acc >> t) & 1u, 17);
          par_acc ^= 1u << t;
          tc_fence_after();
          // this buffer's previous image has left shared memory: the stores alternate X, Y, X, ... so only the newest
          // one (the other tile's) may still be reading
          if (threadIdx.x == 0) bulk_wait_read1();
          asm volatile("bar.sync 1, 256;" ::: "memory");
#pragma unroll
          for (int cp = 0; cp < 4; cp += 2) {                            // two chunks per TMEM load batch
            uint32_t v0[32], v1[32];
            tmem_ld32(lane_addr + t * 256 + cp * 64 + jj * 32, v0);
            tmem_ld32(lane_addr + t * 256 + (cp + 1) * 64 + jj * 32, v1);
            tmem_ld_wait();
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const uint32_t* v = h ? v1 : v0;
              const uint32_t mb = mbits[cp + h];
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                uint32_t pk[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const float lo = ((mb >> (4 * u + i)) & 1u) ? __uint_as_float(v[8 * u + 2 * i]) : 0.f;
                  const float hi = ((mb >> (16 + 4 * u + i)) & 1u) ? __uint_as_float(v[8 * u + 2 * i + 1]) : 0.f;
                  pk[i] = pack_h2(lo, hi);
                }
                st_shared_v4(stage + (cp + h) * IMG_BYTES + uoff[u], pk[0], pk[1], pk[2], pk[3]);
              }
            }
          }
          tc_fence_before();
          fence_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar0 + 8 * (CHB_A_READY + t));      // operand of the next layer (and: accumulator drained)
          asm volatile("bar.sync 1, 256;" ::: "memory");
          if (threadIdx.x == 0 && !dead) {
            if (live) bulk_s2g(const_cast<unsigned char*>(P.out[li].at(tile, 0)), stage, CH_BUF);
            else asm volatile("cp.async.bulk.commit_group;" ::: "memory");           // keeps the X, Y, X, ... group order
            if (li == CH_LAYERS - 1) {                                  // the buffer goes back to the producer for the next pair
              bulk_wait_read0();
              mbar_arrive(bar0 + 8 * (CHB_BUF_FREE + t));
            }
          }
        }
    }
    if (threadIdx.x == 0) bulk_wait_all0();
  }
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();                      // nobody leaves while the peer may still multicast into this CTA
  if (warp == 9) tmem_dealloc512(tmem);
}

// ------------------------------------------------------------------------------------------
// dW GEMM: D[128 out x N in] += sum over the tiles of a split of dZ^T X, both operands MN-major from the images.
// grid = (items, splits); warps 0-3 epilogue, warp 4 producer, warp 5 issuer; 2-stage ring of (2 + n_x) chunks.
// Accumulator columns [256, 272) hold dZ^T 1 = the bias gradient.
// ------------------------------------------------------------------------------------------
struct DwRect { int row0, nrows, col0, ncols, ld, pad; float* dst; float* bias; };
struct DwItem { ImgRef a, x; int x_bytes, N, n_rect, pad; DwRect r[4]; };
constexpr int DW_MAX_ITEMS = 40, DW_STAGE = 6 * IMG_BYTES, DW_THREADS = 192;
constexpr int DW_MAX_SPLITS = 16, DW_PART_LD = 288;          // partial tile of one (item, split): 128 rows x (256 + 32) fp32
constexpr int DW_ONES = 2 * DW_STAGE, DW_BAR = DW_ONES + 1024, DW_SMEM = DW_BAR + 128;
struct DwParams {
  int n_items, n_tiles, n_splits, pad;
  const unsigned int* amax_bits;
  int* dbg;
  int* status;                   // deferred status record (common.cuh)
  float* part;                   // [n_items][n_splits][128][DW_PART_LD] split-K partial accumulators
  DwItem item[DW_MAX_ITEMS];
};
enum { DWB_FULL = 0, DWB_EMPTY = 2, DWB_ACC = 4, DWB_COUNT = 5 };

__global__ void __launch_bounds__(DW_THREADS, 1) k_gemm_dw(const __grid_constant__ DwParams P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const DwItem& I = P.item[blockIdx.x];
  const int t0 = (int)((int64_t)blockIdx.y * P.n_tiles / P.n_splits), t1 = (int)((int64_t)(blockIdx.y + 1) * P.n_tiles / P.n_splits);
  if (t0 >= t1) return;
  const uint32_t sb = smem_u32(smem), bar0 = sb + DW_BAR;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + DW_BAR + 8 * DWB_COUNT);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  bool dead = false;
  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) { mbar_init(bar0 + 8 * (DWB_FULL + s), 1); mbar_init(bar0 + 8 * (DWB_EMPTY + s), 1); }
    mbar_init(bar0 + 8 * DWB_ACC, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {                                     // all-ones operand tile (layout-agnostic)
    uint32_t* ones = reinterpret_cast<uint32_t*>(smem + DW_ONES);
    for (int i = lane; i < 256; i += 32) ones[i] = 0x3c003c00u;
    fence_async_smem();
  }
  if (warp == 5) tmem_alloc512(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 4) {                                     // ---- producer
    const bool leader = elect_one();
    uint32_t par = 3u;
    for (int t = t0; t < t1; ++t) {
      const int s = (t - t0) & 1;
      TTC_WAIT(bar0 + 8 * (DWB_EMPTY + s), (par >> s) & 1u, 11);
      par ^= 1u << s;
      if (leader && !dead) {
        const uint32_t full = bar0 + 8 * (DWB_FULL + s), dst = sb + s * DW_STAGE;
        mbar_expect_tx(full, 2 * IMG_BYTES + (uint32_t)I.x_bytes);
        bulk_g2s(dst, I.a.at(t), 2 * IMG_BYTES, full);
        bulk_g2s(dst + 2 * IMG_BYTES, I.x.at(t), (uint32_t)I.x_bytes, full);
      }
      __syncwarp();
    }
  } else if (warp == 5) {                              // ---- MMA issuer
    const bool leader = elect_one();
    uint32_t par = 0;
    const uint32_t idesc = idesc_f16(128, I.N, 1, 1), idesc_b = idesc_f16(128, 16, 1, 1);
    const uint64_t ones = desc_flat(sb + DW_ONES, 256, 128);
    for (int t = t0; t < t1; ++t) {
      const int s = (t - t0) & 1;
      TTC_WAIT(bar0 + 8 * (DWB_FULL + s), (par >> s) & 1u, 12);
      par ^= 1u << s;
      tc_fence_after();
      if (leader && !dead) {
        const uint32_t st = sb + s * DW_STAGE;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint64_t ad = desc_mn_sw128(st + k * 2048, IMG_BYTES), xd = desc_mn_sw128(st + 2 * IMG_BYTES + k * 2048, IMG_BYTES);
          const uint32_t acc = (t > t0 || k > 0) ? 1u : 0u;
          tc_mma(tmem, ad, xd, idesc, acc);
          tc_mma(tmem + 256, ad, ones, idesc_b, acc);
        }
        tc_commit(bar0 + 8 * (DWB_EMPTY + s));
      }
      __syncwarp();
    }
    if (leader && !dead) tc_commit(bar0 + 8 * DWB_ACC);
    __syncwarp();
  } else {                                             // ---- epilogue: accumulator -> this split's partial tile
    // (no atomics: k_dw_reduce adds the splits in a fixed order, so the weight gradients are bit-reproducible)
    TTC_WAIT(bar0 + 8 * DWB_ACC, 0u, 13);
    tc_fence_after();
    const int row = warp * 32 + lane;
    const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
    float* part = P.part + (((int64_t)blockIdx.x * P.n_splits + blockIdx.y) * 128 + row) * DW_PART_LD;
    uint32_t v[32];
    const int n_groups = (I.N + 31) >> 5;
    for (int g = 0; g <= n_groups; ++g) {                // the last round reads the bias columns [256, 272)
      const int col = g < n_groups ? g * 32 : 256;
      tmem_ld32(lane_addr + col, v);
      tmem_ld_wait();
      float4* d = reinterpret_cast<float4*>(part + col);
#pragma unroll
      for (int i = 0; i < 8; ++i)
        d[i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc512(tmem);
}

// split-K reduction of the dW partial tiles: every gradient element has exactly one (item, rectangle) that owns it, its
// splits are added in index order, the gradient scale is removed in fp32 and the result accumulated into the caller's
// buffer - deterministic, unlike fp32 atomics
__global__ void __launch_bounds__(256) k_dw_reduce(const __grid_constant__ DwParams P) {
  const DwItem& I = P.item[blockIdx.x];
  const float inv = 1.f / grad_scale(P.amax_bits);
  bool bad = false;               // a non-finite partial sum: some fp16 gradient image overflowed upstream
  const float* part = P.part + (int64_t)blockIdx.x * P.n_splits * 128 * DW_PART_LD;
  const int64_t sstride = (int64_t)128 * DW_PART_LD;
  for (int ri = 0; ri < I.n_rect; ++ri) {
    const DwRect R = I.r[ri];
    const int total = R.nrows * R.ncols;
    for (int e = blockIdx.y * blockDim.x + threadIdx.x; e < total; e += blockDim.x * gridDim.y) {   // grid.y slices every rectangle
      const int r = e / R.ncols, c = e - r * R.ncols;
      const float* p = part + (int64_t)(R.row0 + r) * DW_PART_LD + R.col0 + c;
      float acc = 0.f;
      for (int sp = 0; sp < P.n_splits; ++sp) acc += p[sp * sstride];
      if (!isfinite(acc)) bad = true;
      R.dst[(int64_t)r * R.ld + c] += acc * inv;
    }
    if (R.bias != nullptr) {
      for (int r = blockIdx.y * blockDim.x + threadIdx.x; r < R.nrows; r += blockDim.x * gridDim.y) {
        const float* p = part + (int64_t)(R.row0 + r) * DW_PART_LD + 256;
        float acc = 0.f;
        for (int sp = 0; sp < P.n_splits; ++sp) acc += p[sp * sstride];
        if (!isfinite(acc)) bad = true;
        R.bias[r] += acc * inv;
      }
    }
  }
  if (__syncthreads_or(bad) && threadIdx.x == 0 && atomicCAS(P.dbg + 4, 0, 1) == 0) status_raise(P.status, DST_F16_GRAD, blockIdx.x);
}

// ------------------------------------------------------------------------------------------
// unfold the gradient of the composed views' matrix: Wc = Wv1 Wf, bc = Wv1 bf + bv  (Wv1 = views weight[:, :256])
//   dWv1 += dWc Wf^T + dbc (x) bf;  dWf += Wv1^T dWc;  dbf += Wv1^T dbc;  dbv += dbc
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_unfold_comp(const float* __restrict__ flat, NetLayout L, const float* __restrict__ dcomp, float* grad) {
  const float* vw = flat + L.flat_w[L_VIEWS];
  const float* fw = flat + L.flat_w[L_FEAT];
  const float* fb = flat + L.flat_b[L_FEAT];
  const float* dwc = dcomp;
  const float* dbc = dcomp + 128 * W_HID;
  if (blockIdx.x < 128) {
    // dWv1[n][j], block = 8 rows n x 32 columns j; thread = (column j, eighth kq of the k range): Wf is read along k
    // (32 consecutive floats per thread), the eight partial sums of a column meet in a 3-step butterfly
    __shared__ float s_dw[8][8][33];       // [row][k eighth][k within]: the 8 threads of a column hit 8 different banks
    const int n0 = (blockIdx.x >> 3) * 8, j0 = (blockIdx.x & 7) * 32;
    const int jl = threadIdx.x >> 3, kq = threadIdx.x & 7, j = j0 + jl;
    for (int e = threadIdx.x; e < 8 * W_HID; e += 256) s_dw[e >> 8][(e >> 5) & 7][e & 31] = dwc[(n0 + (e >> 8)) * W_HID + (e & 255)];
    __syncthreads();
    float acc[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) acc[r] = 0.f;
    const float* wrow = fw + j * W_HID + kq * 32;      // (the flat parameter vector is only 4-byte aligned per tensor)
#pragma unroll
    for (int i4 = 0; i4 < 8; ++i4) {
      const float4 w = make_float4(__ldg(wrow + 4 * i4), __ldg(wrow + 4 * i4 + 1), __ldg(wrow + 4 * i4 + 2), __ldg(wrow + 4 * i4 + 3));
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const float* d = &s_dw[r][kq][4 * i4];
        acc[r] = fmaf(d[0], w.x, acc[r]);
        acc[r] = fmaf(d[1], w.y, acc[r]);
        acc[r] = fmaf(d[2], w.z, acc[r]);
        acc[r] = fmaf(d[3], w.w, acc[r]);
      }
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int o = 1; o < 8; o <<= 1) acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], o);
    }
    float mine = acc[0];                   // lane kq of a column's group writes row n0 + kq
#pragma unroll
    for (int r = 1; r < 8; ++r) mine = (kq == r) ? acc[r] : mine;
    grad[L.flat_w[L_VIEWS] + (n0 + kq) * (W_HID + PE_DIR) + j] += fmaf(dbc[n0 + kq], fb[j], mine);
    if ((blockIdx.x & 7) == 0 && threadIdx.x < 8) grad[L.flat_b[L_VIEWS] + n0 + threadIdx.x] += dbc[n0 + threadIdx.x];
  } else {                                 // dWf[jf][k = thread], dbf[jf]
    const int jf = blockIdx.x - 128, k = threadIdx.x;
    float acc[4] = {0.f, 0.f, 0.f, 0.f}, accb = 0.f;
#pragma unroll 4
    for (int n = 0; n < 128; n += 4) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float w = vw[(n + i) * (W_HID + PE_DIR) + jf];
        acc[i] = fmaf(w, dwc[(n + i) * W_HID + k], acc[i]);
        if (k == 0) accb = fmaf(w, dbc[n + i], accb);
      }
    }
    grad[L.flat_w[L_FEAT] + jf * W_HID + k] += (acc[0] + acc[1]) + (acc[2] + acc[3]);
    if (k == 0) grad[L.flat_b[L_FEAT] + jf] += accb;
  }
}

}  // namespace ttc

// ==========================================================================================
// host side
// ==========================================================================================
static inline int64_t n_tiles_of(int64_t M) { return (M + ttc::TILE - 1) / ttc::TILE; }
static inline int64_t up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }
constexpr int64_t WS_HEAD = 4096;                                     // amax word + padding
constexpr int64_t WS_COMP = (128 * W_HID + 128) * 4;                 // gradient of the composed views' matrix + bias
constexpr int64_t WS_BLOB = ttc::BW_MAX_TILES * 32768;               // transposed weight tiles
constexpr int64_t WS_PART = (int64_t)ttc::DW_MAX_ITEMS * ttc::DW_MAX_SPLITS * 128 * ttc::DW_PART_LD * 4;   // dW split-K partials

int64_t tc_bwd_workspace_bytes(int variant, int n_classes, int64_t M) {
  (void)variant; (void)n_classes;
  return up(WS_HEAD + WS_COMP, 1024) + WS_BLOB + WS_PART + n_tiles_of(M) * IMG_BWD_SLOTS * (int64_t)IMG_BYTES;
}

int launch_mlp_bwd_tc(const TcBwdArgs& a, cudaStream_t st) {
  using namespace ttc;
  if (a.M == 0) return INRF_OK;
  NetLayout L;
  int rc = make_layout(a.variant, a.n_classes, &L);
  if (rc) return rc;
  const bool sem = a.n_classes > 0;
  const int C = a.n_classes, out_ch = raw_channels(C, a.endpoint);
  const int64_t T = n_tiles_of(a.M);
  if (T > 0x7fffffff / IMG_BWD_SLOTS) { set_error("batch too large for one backward call"); return INRF_EUNSUPPORTED; }
  unsigned char* ws = a.work;
  unsigned int* amax = reinterpret_cast<unsigned int*>(ws);
  float* dcomp = reinterpret_cast<float*>(ws + WS_HEAD);
  unsigned char* blob = ws + up(WS_HEAD + WS_COMP, 1024);
  float* part = reinterpret_cast<float*>(blob + WS_BLOB);
  unsigned char* img = blob + WS_BLOB + WS_PART;
  int* dbg = nullptr;
  INRF_CUDA(cudaGetSymbolAddress((void**)&dbg, ttc::g_dbg));
  INRF_CUDA(cudaMemsetAsync(ws, 0, WS_HEAD + WS_COMP, st));
  static const bool checked = getenv("INRF_TC_CHECK") != nullptr && getenv("INRF_TC_CHECK")[0] == '1';
  INRF_CUDA(cudaMemsetAsync(dbg, 0, 8 * sizeof(int), st));      // per-launch claim words
  int* status = status_flag_dev();
  if (status == nullptr) return INRF_ECUDA;
  int dev = 0, sms = 148;
  INRF_CUDA(cudaGetDevice(&dev));
  INRF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));

  k_amax<<<sms * 4, 256, 0, st>>>(a.grad_raw, a.M * out_ch, amax);
  INRF_LAUNCH_CHECK();
  k_bwd_heads<<<(unsigned)T, 128, 0, st>>>(a.grad_raw, a.raw, out_ch, C, a.M, img, amax);
  INRF_LAUNCH_CHECK();

  // ---- dX GEMM sequence and its weight tiles ---------------------------------------------------------
  auto W = [&](int slot) { return ImgRef{img, IMG_BWD_SLOTS, slot}; };
  auto F = [&](int slot) { return ImgRef{a.stash_img, IMG_STASH_SLOTS, slot}; };
  BwProgram prog;
  prog.n = 0; prog.bytes = 0;
  DxParams G[12];
  int ng = 0;
  auto begin = [&](int N, ImgRef mask, ImgRef out) {
    DxParams& g = G[ng];
    memset(&g, 0, sizeof(g));
    g.N = N; g.mask = mask; g.out = out; g.b = blob + prog.bytes; g.M = a.M; g.amax_bits = amax; g.dbg = dbg; g.status = status; g.n_tiles = (int)T;
  };
  auto add = [&](ImgRef src, int kind, int layer, int out0, int in0) {
    DxParams& g = G[ng];
    g.a[g.n_k++] = src;
    BwTile& t = prog.t[prog.n++];
    t.kind = (int16_t)kind; t.layer = (int16_t)layer; t.out0 = (int16_t)out0; t.in0 = (int16_t)in0; t.N = (int16_t)g.N; t.pad = 0;
    t.off = prog.bytes;
    prog.bytes += g.N * 128;
  };
  // relu'(albedo1|shading1) * (G . [albedo2; shading2])
  begin(256, F(IS_AS), W(IB_DAS)); add(W(IB_G), BW_HEAD_AS, 0, 0, 0); ++ng;
  // relu'(views') * (G . residual head (+ endpoint-feature gradient))
  begin(128, F(IS_V), W(IB_DV)); add(W(IB_G), BW_HEAD_V, 0, 0, 0);
  if (a.endpoint) { G[ng].addend = a.grad_raw + INRF_RAW_BASE + C; G[ng].addend_ld = out_ch; }
  ++ng;
  if (sem) {
    begin(128, F(IS_S1), W(IB_DS1));
    add(W(IB_G), BW_HEAD_SEM, 0, 0, 0); add(W(IB_G + 1), BW_HEAD_SEM, 0, 64, 0);
    ++ng;
  }
  // trunk output: albedo1|shading1, views' (composed), sem1 and sigma branches meet
  begin(256, F(IS_H + 28), W(IB_DZ + 28));
  for (int c = 0; c < 4; ++c) add(W(IB_DAS + c), BW_PLAIN, c < 2 ? L_ALB1 : L_SH1, (c & 1) * 64, 0);
  for (int c = 0; c < 2; ++c) add(W(IB_DV + c), BW_COMP, -1, c * 64, 0);
  if (sem) for (int c = 0; c < 2; ++c) add(W(IB_DS1 + c), BW_PLAIN, L_SEM1, c * 64, 0);
  add(W(IB_G), BW_ALPHA, 0, 0, 0);
  ++ng;
  for (int l = 7; l >= 1; --l) {          // dZ_{l-1} = relu'(H_{l-1}) * (dZ_l W_l[:, hidden part])
    begin(256, F(IS_H + 4 * (l - 1)), W(IB_DZ + 4 * (l - 1)));
    for (int c = 0; c < 4; ++c) add(W(IB_DZ + 4 * l + c), BW_PLAIN, L_T0 + l, c * 64, l == 5 ? PE_PTS : 0);
    ++ng;
  }
  if (prog.n > BW_MAX_TILES || prog.bytes > WS_BLOB) { set_error("internal: backward weight program too long"); return INRF_EINVAL; }
  k_pack_bwd<<<dim3(prog.n, 8), 256, 0, st>>>(a.flat, static_cast<const unsigned char*>(a.packed), L, prog, blob);
  INRF_LAUNCH_CHECK();
  INRF_CUDA(cudaFuncSetAttribute(k_gemm_dx, cudaFuncAttributeMaxDynamicSharedMemorySize, DX_SMEM));
  const int grid = (int)(T < sms ? T : sms);
  // the head GEMMs and the trunk-output GEMM one launch each, the seven trunk layers as one chained launch
  // (INRF_DX_CHAIN=0: one launch per layer, for A/B)
  static const int chain_mode = getenv("INRF_DX_CHAIN") != nullptr ? atoi(getenv("INRF_DX_CHAIN")) : 1;   // 0 per-layer launches, 2 clusters
  const bool chain_env = chain_mode != 0;
  const int n_single = chain_env ? ng - CH_LAYERS : ng;
  for (int i = 0; i < n_single; ++i) {
    G[i].n_iter = (int)((T + grid - 1) / grid);
    k_gemm_dx<<<grid, DX_THREADS, DX_SMEM, st>>>(G[i]);
    INRF_LAUNCH_CHECK();
  }
  if (chain_env) {
    ChainParams Cp;
    memset(&Cp, 0, sizeof(Cp));
    Cp.n_tiles = (int)T; Cp.n_pairs = (int)((T + 1) / 2); Cp.dbg = dbg; Cp.status = status;
    Cp.in = G[n_single].a[0];                       // dZ_7: the trunk-output GEMM's four output images
    for (int li = 0; li < CH_LAYERS; ++li) {
      const DxParams& g = G[n_single + li];
      Cp.b[li] = g.b; Cp.mask[li] = g.mask; Cp.out[li] = g.out;
    }
    int cl = (chain_mode == 2 && Cp.n_pairs >= 2) ? 2 : 1;
    int cgrid = Cp.n_pairs < sms ? Cp.n_pairs : sms;
    cgrid = cgrid / cl * cl;                         // whole clusters only
    Cp.n_iter = (Cp.n_pairs + cgrid - 1) / cgrid;
    void (*kern)(ChainParams) = cl == 2 ? k_dx_chain<2> : k_dx_chain<1>;
    INRF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, CH_SMEM));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)cgrid);
    cfg.blockDim = dim3(DX_THREADS);
    cfg.dynamicSmemBytes = CH_SMEM;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cl;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    INRF_CUDA(cudaLaunchKernelEx(&cfg, kern, Cp));
    note_launch();
  }

  // ---- dW: one launch, one item per (GEMM, 128 output rows) --------------------------------------------
  DwParams D;                              // 9 KB launch table, rebuilt per call (pointers depend on the caller's buffers)
  memset(&D, 0, sizeof(D));
  D.n_tiles = (int)T; D.amax_bits = amax; D.dbg = dbg; D.status = status; D.part = part;
  float* gf = a.grad_flat;
  auto item = [&](ImgRef az, ImgRef x, int n_x, int N) -> DwItem& {
    DwItem& it = D.item[D.n_items++];
    it.a = az; it.x = x; it.x_bytes = n_x * IMG_BYTES; it.N = N; it.n_rect = 0;
    return it;
  };
  auto rect = [&](DwItem& it, int row0, int nrows, int col0, int ncols, float* dst, int ld, float* bias) {
    DwRect& r = it.r[it.n_rect++];
    r.row0 = row0; r.nrows = nrows; r.col0 = col0; r.ncols = ncols; r.ld = ld; r.dst = dst; r.bias = bias;
  };
  for (int l = 0; l < 8; ++l) {
    const LayerDims d = layer_dims(L_T0 + l, C);
    for (int h = 0; h < 2; ++h) {
      float* w = gf + L.flat_w[L_T0 + l] + (int64_t)128 * h * d.K;
      float* b = gf + L.flat_b[L_T0 + l] + 128 * h;
      if (l == 0 || l == 5) rect(item(W(IB_DZ + 4 * l + 2 * h), F(IS_PE), 1, 64), 0, 128, 0, PE_PTS, w, d.K, l == 0 ? b : nullptr);
      if (l > 0) rect(item(W(IB_DZ + 4 * l + 2 * h), F(IS_H + 4 * (l - 1)), 4, 256), 0, 128, 0, 256, w + (l == 5 ? PE_PTS : 0), d.K, b);
    }
  }
  rect(item(W(IB_DAS), F(IS_H + 28), 4, 256), 0, 128, 0, 256, gf + L.flat_w[L_ALB1], 256, gf + L.flat_b[L_ALB1]);
  rect(item(W(IB_DAS + 2), F(IS_H + 28), 4, 256), 0, 128, 0, 256, gf + L.flat_w[L_SH1], 256, gf + L.flat_b[L_SH1]);
  rect(item(W(IB_DV), F(IS_H + 28), 4, 256), 0, 128, 0, 256, dcomp, 256, dcomp + 128 * W_HID);
  rect(item(W(IB_DV), F(IS_DIR), 1, 32), 0, 128, 0, PE_DIR, gf + L.flat_w[L_VIEWS] + W_HID, W_HID + PE_DIR, nullptr);
  if (sem) rect(item(W(IB_DS1), F(IS_H + 28), 4, 256), 0, 128, 0, 256, gf + L.flat_w[L_SEM1], 256, gf + L.flat_b[L_SEM1]);
  {
    DwItem& it = item(W(IB_G), F(IS_AS), 4, 256);            // G^T relu(albedo1|shading1)
    rect(it, 0, 3, 0, 128, gf + L.flat_w[L_ALB2], 128, gf + L.flat_b[L_ALB2]);
    rect(it, 3, 1, 128, 128, gf + L.flat_w[L_SH2], 128, gf + L.flat_b[L_SH2]);
    rect(it, 4, 3, 0, 0, nullptr, 0, gf + L.flat_b[L_RES]);
    rect(it, 7, 1, 0, 0, nullptr, 0, gf + L.flat_b[L_ALPHA]);
  }
  rect(item(W(IB_G), F(IS_V), 2, 128), 4, 3, 0, 128, gf + L.flat_w[L_RES], 128, nullptr);
  rect(item(W(IB_G), F(IS_H + 28), 4, 256), 7, 1, 0, 256, gf + L.flat_w[L_ALPHA], 256, nullptr);
  if (sem) rect(item(W(IB_G), F(IS_S1), 2, 128), 8, C, 0, 128, gf + L.flat_w[L_SEM2], 128, gf + L.flat_b[L_SEM2]);
  if (D.n_items > DW_MAX_ITEMS) { set_error("internal: too many dW items"); return INRF_EINVAL; }
  int splits = (2 * sms) / D.n_items;           // at most two full waves of CTAs (one per SM): a third, nearly empty wave costs a whole CTA time
  if (splits > DW_MAX_SPLITS) splits = DW_MAX_SPLITS;
  if (splits > T) splits = (int)T;              // every split owns at least one tile, so every partial tile is written
  if (splits < 1) splits = 1;
  D.n_splits = splits;
  INRF_CUDA(cudaFuncSetAttribute(k_gemm_dw, cudaFuncAttributeMaxDynamicSharedMemorySize, DW_SMEM));
  k_gemm_dw<<<dim3(D.n_items, splits), DW_THREADS, DW_SMEM, st>>>(D);
  INRF_LAUNCH_CHECK();
  k_dw_reduce<<<dim3(D.n_items, 16), 256, 0, st>>>(D);      // (item, slice): ~450 CTAs instead of ~28
  INRF_LAUNCH_CHECK();
  k_unfold_comp<<<128 + 256, 256, 0, st>>>(a.flat, L, dcomp, gf);
  INRF_LAUNCH_CHECK();
  if (checked) {      // debug mode (INRF_TC_CHECK=1): synchronise and report this call's status record right away
    INRF_CUDA(cudaStreamSynchronize(st));
    return status_poll();
  }
  return INRF_OK;
}

}  // namespace inrf
