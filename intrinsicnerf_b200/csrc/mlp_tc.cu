// Tensor-core evaluation of the intrinsic field network on sm_100a (INRF_PREC_TC).
//
// One persistent CTA per SM walks 128-sample tiles.  Per tile the whole network
// (NeRF.forward, object_level/run_nerf_helpers.py:284-325 / Semantic_NeRF.forward,
// SSR/models/semantic_nerf.py:123-181, fused with the Embedder and run_network's
// per-sample view-direction expansion) runs without touching HBM in between:
//
//   warp 14     weight producer : streams the pre-swizzled fp16 operand tiles of the packed blob
//                                 (pack.cu, MMA issue order) into a 4 x 32 KB shared-memory ring with
//                                 cp.async.bulk + mbarrier complete_tx (TMA bulk copies; optionally
//                                 multicast across a 2-CTA cluster)
//   warp 15     MMA issuer      : converged warp, one elected lane issues tcgen05.mma (kind::f16,
//                                 M=128, N=256 for the 256-wide layers, K=16) with fp32 accumulators
//                                 in TMEM; the bias enters as one extra K=16 MMA of a constant "ones"
//                                 tile with hi/lo/lo2 bias columns (accumulator initialisation);
//                                 tcgen05.commit signals "accumulator ready", "ring slot free",
//                                 "A chunk free"
//   warps 8-11  front end       : sample position (o + d z), range-reduced sin/cos positional
//                                 encoding of the NEXT tile, written as fp16 UMMA operand tiles
//   warps 0-7   epilogue        : tcgen05.ld accumulator -> ReLU -> fp16 -> the next layer's A operand
//                                 (SWIZZLE_128B K-major), in place, K chunk by K chunk so that the
//                                 next layer's MMAs start as soon as chunk 0 exists; sigma head as an
//                                 fp32 dot product on the un-rounded trunk output; sigmoid heads;
//                                 packed raw rows to HBM (TS variant: the head values wait in registers
//                                 and the rows are finished under the NEXT tile's layer 1, PendingHeads)
//   warps 12-13 ray back end    : fused renderer only - compositing and hierarchical resampling of the
//                                 previous tile's rows (ray_backend); warp 13 is the stash warp in the
//                                 training instantiation, warp 12 also writes the constant operands
//
// The issuer warp paces the kernel (profiles/r02_issuer_timeline.md): ~590 cycles of dependent bookkeeping
// per ring fill against 512 cycles of MMAs.  Do not add work to Issuer::acquire / finish / probe without measuring -
// two run-time flags there once cost 7 % of the launch.  The sy.stamp() calls compile to nothing unless
// -DINRF_TC_TIMELINE (python -m intrinsicnerf_b200.build --timeline).
//
// Why N=256 instructions: consecutive MMAs into the same accumulator are dependent; a 64-cycle
// N=128 instruction cannot hide the accumulate latency (measured: 1.3x slower than N=256).
//
// Arithmetic: operands rounded to fp16 (RN, 11-bit significand), products and sums in fp32.
// feature_linear has no activation, so views_linears.0 o feature_linear is composed into one
// 128x256 matrix at pack time (pack.cu) - one 256x256 GEMM per sample less than the literal graph.
#include "common.cuh"
#include <stdlib.h>

namespace inrf {
namespace tc {

constexpr int TILE_M = 128;
constexpr int NUM_THREADS = 512;
constexpr int NS = 4;                       // weight ring stages of TC_SLOT_BYTES (32 KB), shared-memory-activation kernel
constexpr int NS_MAX = 6;                   // TS variant: no H buffer, so the ring gets its 64 KB (runtime Params.ns)
constexpr int CHUNK = 16384;                // 128 rows x 64 fp16, SWIZZLE_128B
// shared memory map (bytes)
constexpr int SM_H = 0;                     // 4 chunks: hidden activations (A operand), in place
constexpr int SM_PE = SM_H + 4 * CHUNK;     // gamma(x): 63 cols + zero
constexpr int SM_DIR = SM_PE + CHUNK;       // gamma(d): 27 cols + zeros (only K=32 is multiplied)
constexpr int SM_V = SM_PE;                 // relu(views') (2 chunks) reuses the PE|DIR region at the tile tail
constexpr int SM_RING = SM_DIR + CHUNK;
constexpr int SM_SIG = SM_RING + NS * TC_SLOT_BYTES;   // [128][2] fp32 sigma partials
constexpr int SM_ALPHA = SM_SIG + 1024;                // alpha_linear weight row, fp32 [256]
constexpr int SM_ONES = SM_ALPHA + 1024;               // 8 x 16 fp16 "ones" A operand for the bias MMAs
constexpr int SM_BAR = SM_ONES + 256;
constexpr int SM_TOTAL = SM_BAR + 512;
static_assert(SM_TOTAL <= 232448, "shared memory budget");

// bias table offsets (floats), written by pack.cu:k_pack_tc_bias
constexpr int TCB_VIEWS = 2048, TCB_SEM1 = 2176, TCB_ALBSH = 2304, TCB_ALPHA_W = 2560, TCB_ALPHA_B = 2816,
              TCB_ALB2 = 2817, TCB_SH2 = 2820, TCB_RES = 2821, TCB_SEM2 = 2824;

// barrier ids
enum {
  B_WFULL = 0,                 // [ns] ring slot filled (tx bytes)
  B_WEMPTY = B_WFULL + NS_MAX, // [ns] ring slot consumed (tcgen05.commit)
  B_F_READY = B_WEMPTY + NS_MAX,   // PE|DIR tiles of the next tile written (128 front-end threads)
  B_F_FREE,                    // PE|DIR|V region no longer read by the tensor core
  B_A_READY,                   // [4] H chunk c written by all 8 epilogue warps (256 threads)
  B_H_FREE = B_A_READY + 4,    // tail only: the tensor core finished reading H (before it is overwritten
                               // by relu(albedo1|shading1) / relu(sem1)); in the trunk "accumulator
                               // complete" already implies that every read of H by that layer is done
  B_ACC_FULL,                  // [2] accumulator (TMEM columns parity*256 ..) complete
  B_V_READY = B_ACC_FULL + 2,  // relu(views') written (256 threads)
  B_SMALL_FULL, B_SEM2_FULL, B_TAIL_DONE,
  B_RAW_READY,                 // [2] fused mode: the raw rows of a tile are in ring slot (it & 1) (8 epilogue warps)
  B_RAW_FREE = B_RAW_READY + 2,  // [2] both back-end warps have consumed the slot
  B_ST_DONE = B_RAW_FREE + 2,    // [4] training: the stash copy of H chunk c has left shared memory (the chunk may be rewritten)
  B_STV_DONE = B_ST_DONE + 4,    // training: the same for the two relu(views') chunks in the PE|DIR region
  B_ACC_FIRST,                 // [2] split hand-off (Params.split_nf): accumulator columns [0, split_nf) of a trunk layer complete
  B_COUNT = B_ACC_FIRST + 2
};
static_assert(B_COUNT <= 48, "barrier area");

struct Params {
  MlpArgs a;
  const unsigned char* blocks;    // fp16 operand blob
  const float* bias;              // fp32 bias table
  int n_fills;
  int fill_off[TC_MAX_FILLS];     // byte offsets / sizes of the ring fills of one tile
  int fill_bytes[TC_MAX_FILLS];
  int out_ch, C, sem_rows;
  int bias_mma;                   // 1: accumulators are initialised with the bias by an MMA (default)
  int n_iter;                     // tile iterations per CTA (identical for every CTA: cluster lock-step)
  int ts_fine;                    // TS variant: 1 = one K chunk per TMEM load batch (INRF_TC_TS=2)
  int ns;                         // weight ring stages in use (4, or 6 in the TS variant)
  int split_nf;                   // TS / HY trunk layers 1..7: the MMAs of the LAST K chunk are issued as N = split_nf (output
                                  // columns [0, split_nf), committed to B_ACC_FIRST) + N = 256 - split_nf, so the drain of the
                                  // first output chunk starts split_nf/256 of a K chunk after the last input chunk arrived
                                  // instead of a whole K chunk (512 cycles) after it.  0 = off, 64 or 128
  int fuse;                       // 1: composite (and resample) in-kernel, CTAs walk CONTIGUOUS tiles (whole rays per CTA)
  FuseArgs f;
  int* dbg;                       // [16] per-launch abort / claim words (device, cleared before every launch)
  int* status;                    // pinned host record the next API call reads (common.cuh: status_raise)
  long long watchdog;             // cycles a blocking barrier wait may take before the launch is abandoned
  int fault;                      // test hook (INRF_TC_FAULT=n, first n launches): the weight producer stops after three fills
  long long* prof;                // optional wait-cycle counters of CTA 0 (INRF_TC_PROF=1)
  int no_weights;                 // timing experiment: do not wait for / stream weights (results are garbage)
  int exp_flags;                  // A/B switch (INRF_TC_EXP): 8 = ray back end on front-end warps 0 and 2 (its first placement)
  int stash_abl;                  // timing experiment (INRF_TC_STASH_ABL): 1 no mask words, 2 no bulk copies, 4 no "copy has read" waits
};

__device__ int g_dbg[16];
__device__ long long g_prof[5 * 128];
#ifdef INRF_TC_TIMELINE
// development build (tools/gpu_timeline.sh): clock64 stamps of the issuer and of two epilogue warps of CTA 0 for two
// steady-state tiles; printed by the launcher.  (stamp << 4) | kind
constexpr int TL_CAP = 1024, TL_IT0 = 10, TL_IT1 = 11;
__device__ long long g_tl[4 * TL_CAP];
__device__ int g_tl_n[4];
#endif

// ------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t addr, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(addr), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t addr) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t addr, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(addr), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try(uint32_t addr, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
  return ok;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}
__device__ __forceinline__ void bulk_g2s_mc(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar, uint16_t mask) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar), "h"(mask) : "memory");
}
__device__ __forceinline__ void tc_commit_mc(uint32_t mbar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(mbar), "h"(mask) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t mbar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory");
}
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
               "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
               "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                 "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                 "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                 "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
               : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// A operand from tensor memory ("TS" form): lanes = the 128 rows, 32-bit column j of the operand holds K elements 2j (low
// half) and 2j+1 of the row, so one K=16 step reads 8 columns (cute::SM100_MMA_F16BF16_TS / tmem_frg_1sm<half, half>)
__device__ __forceinline__ void tc_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
               "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
               ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
                 "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor, K-major, SWIZZLE_128B (cute/arch/mma_sm100_desc.hpp):
//  [0,14) start>>4 | [16,30) LBO>>4 (unused for swizzled K-major, 1) | [32,46) SBO>>4 = 1024>>4 |
//  [46,48) version = 1 | [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// no-swizzle K-major core-matrix layout (layout type 0): 8 rows x 16 B contiguous, the two K halves
// LBO bytes apart, 8-row groups SBO bytes apart.  SBO = 0 replays one 8-row group for all 128 rows.
__device__ __forceinline__ uint64_t make_desc_flat(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
// instruction descriptor: D=f32 (bit 4), A=B=f16 (0), K-major both, N>>3 at [17,23), M>>4 at [24,29)
__device__ __forceinline__ constexpr uint32_t make_idesc(int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
}

// ------------------------------------------------------------------------------------------
// barrier bookkeeping with a watchdog: a stuck wait records who/where, raises a global abort
// flag and lets every role run to the end (garbage out, but no hung GPU and no lost context)
// ------------------------------------------------------------------------------------------
// out-of-line spin with watchdog; returns true when the wait was abandoned (dead)
__device__ __noinline__ bool slow_wait_impl(uint32_t bar_addr, uint32_t parity, int id, int tile, int* dbg, long long* prof,
                                            int* status, long long watchdog) {
  const long long t0 = clock64();
  uint32_t spins = 0;
  bool dead = false;
  while (!mbar_try(bar_addr, parity)) {
    if ((++spins & 0x3ff) == 0) {
      if (*(volatile int*)dbg != 0) { dead = true; break; }
      if (clock64() - t0 > watchdog) {
        if (atomicCAS(dbg, 0, 1) == 0) {        // first to give up: tell the host (next API call) and every other role
          dbg[1] = id; dbg[2] = threadIdx.x >> 5; dbg[3] = tile; dbg[4] = blockIdx.x; dbg[5] = (int)parity;
          __threadfence();
          status_raise(status, DST_WATCHDOG, id, threadIdx.x >> 5, tile, blockIdx.x, 1);
        }
        dead = true;
        break;
      }
    }
  }
  if (prof) { prof[id] += clock64() - t0; prof[id + 64] += 1; }
  return dead;
}

__device__ __forceinline__ uint32_t mbar_test(uint32_t addr, uint32_t parity) {   // non-blocking probe
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
  return ok;
}

struct Sync {             // lives in registers (never escapes by address)
  uint32_t bar0;          // smem address of barrier 0
  uint64_t phase;         // one parity bit per barrier id
  int* dbg;
  int* status;
  long long watchdog;
  long long* prof;        // wait-cycle counters of this role (CTA 0, one thread per role) or nullptr
  bool dead;
  int tile;
#ifdef INRF_TC_TIMELINE
  long long* tl; int tl_n; int it;
  __device__ __forceinline__ void stamp(int kind) {
    if (tl != nullptr && (it == TL_IT0 || it == TL_IT1) && tl_n < TL_CAP) tl[tl_n++] = (clock64() << 4) | kind;
  }
  __device__ __forceinline__ void set_it(int i) { it = i; }
#else
  __device__ __forceinline__ void stamp(int) {}
  __device__ __forceinline__ void set_it(int) {}
#endif
  __device__ __forceinline__ uint32_t addr(int id) const { return bar0 + 8u * id; }
  __device__ __forceinline__ uint32_t take_parity(int id) {       // consume the next phase of barrier id
    const uint32_t parity = (uint32_t)((phase >> id) & 1ull);
    phase ^= (1ull << id);
    return parity;
  }
  __device__ __forceinline__ void slow(int id, uint32_t parity) {
    if (!dead) dead = slow_wait_impl(addr(id), parity, id, tile, dbg, prof, status, watchdog);
  }
  __device__ __forceinline__ void wait(int id) {
    const uint32_t parity = take_parity(id);
    if (dead) return;
    if (mbar_try(addr(id), parity)) return;
    slow(id, parity);
  }
};

// One arrival per warp: every lane has fenced its own shared-memory writes, the warp converges and
// lane 0 arrives (a 32-lane arrive on one mbarrier word serialises like a same-address atomic).
__device__ __forceinline__ void warp_arrive(uint32_t bar, int lane) {
  __syncwarp();
  if (lane == 0) mbar_arrive(bar);
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}

// one K-major SWIZZLE_128B row address: 8-row atoms of 1024 B, 16-byte unit XOR (row & 7)
__device__ __forceinline__ uint32_t swz(uint32_t chunk_base, int row, int unit) {
  return chunk_base + (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((unit ^ (row & 7)) << 4));
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void st_global_v4(unsigned char* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.global.v4.b32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
  __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

// ------------------------------------------------------------------------------------------
// front end: positional encoding with one shared range reduction per coordinate.
// sin(2^k x) = sin(2 pi frac(2^k x/(2 pi))): x/(2 pi) is formed as a two-float product, scaling by
// 2^k and taking the fractional part are exact, so the argument error does not grow with k.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void pe_coord(float x, int n_freqs, float* s, float* c) {
  const float HI = 0.15915494f, LO = 6.4206383e-09f, TWO_PI = 6.2831855f;
  float p = x * HI;
  float e = fmaf(x, HI, -p);
  float lo = fmaf(x, LO, e);
  float sc = 1.f;
#pragma unroll
  for (int k = 0; k < 10; ++k) {
    if (k < n_freqs) {
      float ph = p * sc;
      float r = (ph - rintf(ph)) + lo * sc;
      float ang = r * TWO_PI;
      s[k] = __sinf(ang);
      c[k] = __cosf(ang);
    }
    sc *= 2.f;
  }
}

// relu + round-to-nearest fp16 pair, saturating at +-65504 instead of producing inf (one F2FP instruction)
__device__ __forceinline__ uint32_t pack_relu_sat_h2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint32_t hmax2_u32(uint32_t a, uint32_t b) {
  __half2 x = *reinterpret_cast<__half2*>(&a), y = *reinterpret_cast<__half2*>(&b);
  x = __hmax2(x, y);
  return *reinterpret_cast<uint32_t*>(&x);
}

// tile walked by this CTA in iteration `it`.  Fused mode: contiguous runs (n_iter is a multiple of the tiles per ray
// group, so every ray is completed by the CTA that started it); otherwise interleaved.
__device__ __forceinline__ int64_t tile_of(const Params& P, int it) {
  return P.fuse ? (int64_t)blockIdx.x * P.n_iter + it : (int64_t)it * gridDim.x + blockIdx.x;
}
// ray record n (o3 d3 near far viewdir3): read from the caller's table or generated from the frame's camera
__device__ __forceinline__ void get_ray(const Params& P, int64_t n, float* r) {
  if (P.a.cam_on) { cam_ray(P.a.cam, P.a.cam_H, P.a.cam_W, P.a.cam_pix0 + n, r); return; }
  const float* ray = P.a.rays + n * 11;
#pragma unroll
  for (int i = 0; i < 11; ++i) r[i] = __ldg(ray + i);
}
// depth of sample s of ray n: the caller's array, or (coarse pass of the fused renderer) generated here
__device__ __forceinline__ float sample_depth(const Params& P, int64_t n, int s) {
  if (P.a.z != nullptr) return __ldg(P.a.z + n * P.a.S + s);
  float nearv, farv;
  if (P.a.cam_on) { nearv = P.a.cam.nearv; farv = P.a.cam.farv; }
  else { nearv = __ldg(P.a.rays + n * 11 + 6); farv = __ldg(P.a.rays + n * 11 + 7); }
  const bool jit = P.f.t_rand != nullptr;
  return coarse_z_sample(nearv, farv, P.f.t_vals, s, P.a.S, P.f.lindisp, jit, jit ? __ldg(P.f.t_rand + n * P.a.S + s) : 0.f);
}

// ------------------------------------------------------------------------------------------
// ray back end (fused renderer): one warp turns the raw rows of the previous tile - handed over through a small
// global ring that stays in L2 - into per-ray records while the tensor pipe works on the next tile.
//   compositing  raw2outputs, run_nerf.py:359-412 / model_utils.py:39-116 - same arithmetic, in the same order, as
//                k_raw2outputs (stages.cu): lane l owns samples l, l+32, ...; 32-sample segments with a running carry
//   resampling   sample_pdf + sort(cat(z, z_samples)) + std, run_nerf.py:499-503,519 - the arithmetic of k_zmid,
//                k_sample_pdf and k_merge_sorted with the 63-entry cdf, the bins and both sorted lists held in
//                registers and searched with warp shuffles (no shared memory: the CTA has none left)
// ------------------------------------------------------------------------------------------
constexpr unsigned FULLMASK = 0xffffffffu;
__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULLMASK, v, o);
  return v;
}
struct RayState {             // per-lane state of the ray the back-end warp is compositing
  float carry;                // prod (1 - alpha + 1e-10) over the segments already seen
  float acc[12];              // rgb3 albedo3 shading residual3 depth acc, this lane's samples
  float ex[4];                // semantic logits, channel lane + 32 i
  float zc[2], wc[2];         // coarse pass with resampling (S == 64): depth / weight of samples lane, lane + 32
};

// searchsorted-style fetches from tables spread over the warp (entry k lives in lane k & 31 of register k >> 5)
__device__ __forceinline__ float fetch2(float r0, float r1, int k) {
  const float a = __shfl_sync(FULLMASK, r0, k & 31), b = __shfl_sync(FULLMASK, r1, k & 31);
  return (k < 32) ? a : b;
}
__device__ __forceinline__ float fetch4(const float* r, int k) {
  const float a = __shfl_sync(FULLMASK, r[0], k & 31), b = __shfl_sync(FULLMASK, r[1], k & 31);
  const float c = __shfl_sync(FULLMASK, r[2], k & 31), d = __shfl_sync(FULLMASK, r[3], k & 31);
  return (k < 64) ? ((k < 32) ? a : b) : ((k < 96) ? c : d);
}

// coarse ray finished: z_samples = sample_pdf(z_mid, weights[1:-1], 128, det) ; z_out = sort(cat(z, z_samples)) ; z_std
__device__ __noinline__ void resample_ray(const Params& P, int64_t ray, const RayState& rs, int lane) {
  // pdf weights i = 0..61 <-> coarse weights i+1 (k_sample_pdf with weights + 1, B = 63)
  const float w_lo = __shfl_down_sync(FULLMASK, rs.wc[0], 1), w_32 = __shfl_sync(FULLMASK, rs.wc[1], 0);
  const float pw0 = (lane == 31) ? w_32 : w_lo;                        // i = lane
  const float pw1 = __shfl_down_sync(FULLMASK, rs.wc[1], 1);           // i = lane + 32 (valid for lane < 30)
  float part = __fadd_rn(pw0, 1e-5f);
  if (lane + 32 < 62) part += __fadd_rn(pw1, 1e-5f);
  const float total = wsum(part);
  // cdf: fp64 inclusive scan rounded per element (ATen's CPU cumsum), cdf[0] = 0
  double p0 = (double)__fdiv_rn(__fadd_rn(pw0, 1e-5f), total);
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const double q = __shfl_up_sync(FULLMASK, p0, o); if (lane >= o) p0 += q; }
  const float cv0 = (float)(0.0 + p0);                                  // cdf[lane + 1]
  const double carry = __shfl_sync(FULLMASK, p0, 31);
  double p1 = (lane + 32 < 62) ? (double)__fdiv_rn(__fadd_rn(pw1, 1e-5f), total) : 0.0;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const double q = __shfl_up_sync(FULLMASK, p1, o); if (lane >= o) p1 += q; }
  const float cv1 = (float)(carry + p1);                                // cdf[lane + 33] (lane < 30)
  // bins = z_mid[k] = .5 (z[k+1] + z[k]), k = 0..62
  const float z_lo = __shfl_down_sync(FULLMASK, rs.zc[0], 1), z_32 = __shfl_sync(FULLMASK, rs.zc[1], 0);
  const float zm0 = __fmul_rn(0.5f, __fadd_rn((lane == 31) ? z_32 : z_lo, rs.zc[0]));             // k = lane
  const float zm1 = __fmul_rn(0.5f, __fadd_rn(__shfl_down_sync(FULLMASK, rs.zc[1], 1), rs.zc[1])); // k = lane + 32 (lane < 31)
  auto cdf_at = [&](int k) { const float v = fetch2(cv0, cv1, (k - 1) & 63); return k == 0 ? 0.f : v; };
  const int B = 63;
  float smp[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const float u = __ldg(P.f.u_det + lane + 32 * r);
    int lo = 0, hi = B;
#pragma unroll
    for (int it = 0; it < 6; ++it) {                                    // searchsorted(cdf, u, right=True)
      const bool act = lo < hi;
      const int mid = act ? (lo + hi) >> 1 : 0;
      const float v = cdf_at(mid);
      if (act) { if (v <= u) lo = mid + 1; else hi = mid; }
    }
    const int below = max(lo - 1, 0), above = min(lo, B - 1);
    const float c0 = cdf_at(below), c1 = cdf_at(above);
    const float b0 = fetch2(zm0, zm1, below), b1 = fetch2(zm0, zm1, above);
    float denom = __fsub_rn(c1, c0);
    if (denom < 1e-5f) denom = 1.f;
    const float t = __fdiv_rn(__fsub_rn(u, c0), denom);
    smp[r] = __fadd_rn(b0, __fmul_rn(t, __fsub_rn(b1, b0)));
  }
  // z_std = std(z_samples, unbiased=False)
  if (P.f.z_std != nullptr) {
    float sm = 0.f;
#pragma unroll
    for (int r = 0; r < 4; ++r) sm += smp[r];
    const float mean = wsum(sm) / 128.f;
    float q = 0.f;
#pragma unroll
    for (int r = 0; r < 4; ++r) { const float d = smp[r] - mean; q += d * d; }
    q = wsum(q);
    if (lane == 0) P.f.z_std[ray] = sqrtf(q / 128.f);
  }
  // merge: both lists ascending (depths by construction, samples because u and the cdf are) -> final position =
  // own index + rank in the other list (ties: depths first), exactly k_merge_sorted's fast path
  float* zo = P.f.z_out + ray * 192;
  bool ok = true;
  {
    const float a_next0 = (lane == 31) ? z_32 : z_lo;
    ok = ok && (rs.zc[0] <= a_next0);
    if (lane < 31) ok = ok && (rs.zc[1] <= __shfl_down_sync(FULLMASK, rs.zc[1], 1)); else (void)__shfl_down_sync(FULLMASK, rs.zc[1], 1);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const float nxt = __shfl_down_sync(FULLMASK, smp[r], 1);
      const float wrap = (r < 3) ? __shfl_sync(FULLMASK, smp[r < 3 ? r + 1 : 3], 0) : 0.f;
      if (lane < 31) ok = ok && (smp[r] <= nxt);
      else if (r < 3) ok = ok && (smp[r] <= wrap);
    }
  }
  if (__all_sync(FULLMASK, ok)) {
#pragma unroll
    for (int g = 0; g < 2; ++g) {                                       // depths: index i, rank = #(samples < x)
      const float x = rs.zc[g];
      int lo = 0, hi = 128;
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const bool act = lo < hi;
        const int mid = act ? (lo + hi) >> 1 : 0;
        const float v = fetch4(smp, mid);
        if (act) { if (v < x) lo = mid + 1; else hi = mid; }
      }
      zo[lane + 32 * g + lo] = x;
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {                                       // samples: index j, rank = #(depths <= x)
      const float x = smp[r];
      int lo = 0, hi = 64;
#pragma unroll
      for (int it = 0; it < 7; ++it) {
        const bool act = lo < hi;
        const int mid = act ? (lo + hi) >> 1 : 0;
        const float v = fetch2(rs.zc[0], rs.zc[1], mid);
        if (act) { if (v <= x) lo = mid + 1; else hi = mid; }
      }
      zo[lane + 32 * r + lo] = x;
    }
    return;
  }
  // general case (NaN weights, non-monotone input): rank sort over all 192 values, NaNs last, ties by source index
  float v[6] = {rs.zc[0], rs.zc[1], smp[0], smp[1], smp[2], smp[3]};
  int rank[6] = {0, 0, 0, 0, 0, 0};
  for (int sr = 0; sr < 6; ++sr) {
    for (int sl = 0; sl < 32; ++sl) {
      const float y = __shfl_sync(FULLMASK, v[sr], sl);
      const int yi = sr * 32 + sl;
#pragma unroll
      for (int e = 0; e < 6; ++e) {
        const float x = v[e];
        const int xi = e * 32 + lane;
        const bool xn = x != x, yn = y != y;
        const bool less = yn ? (xn && yi < xi) : (xn || y < x || (y == x && yi < xi));
        rank[e] += less ? 1 : 0;
      }
    }
  }
#pragma unroll
  for (int e = 0; e < 6; ++e) zo[rank[e]] = v[e];
}

// `which` = 0 / 1: the two back-end warps share the work by ray parity (a ray is always handled by one warp, so its
// running transmittance stays in that warp's registers)
template <bool SAMPLER>
__device__ __forceinline__ void ray_backend(const Params& P, Sync& sy, int it, int lane, RayState& rs, int which) {
  const int slot = it & 1;
  const int64_t tile = tile_of(P, it);
  sy.tile = (int)tile;
  sy.set_it(it);
  sy.stamp(3);
  sy.wait(B_RAW_READY + slot);
  sy.stamp(1);
  const float* ring = P.f.ring + ((size_t)blockIdx.x * 2 + slot) * TILE_M * P.out_ch;   // plain (coherent) loads: written by this CTA
  const int S = P.a.S;
  for (int q = 0; q < 4; ++q) {
    const int64_t m0 = tile * TILE_M + q * 32;
    if (m0 >= P.a.M) break;                             // 32 | S | M: a segment is entirely inside or outside the batch
    const int64_t ray = m0 / S;
    if ((int)(ray & 1) != which) continue;
    const int s0 = (int)(m0 - ray * S), s = s0 + lane;
    if (s0 == 0) {
      rs.carry = 1.f;
#pragma unroll
      for (int i = 0; i < 12; ++i) rs.acc[i] = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) rs.ex[i] = 0.f;
    }
    float rr[11];
    get_ray(P, ray, rr);
    const float dx = rr[3], dy = rr[4], dz = rr[5];
    const float dnorm = sqrtf(dx * dx + dy * dy + dz * dz);
    const float zs = sample_depth(P, ray, s);
    float znext = __shfl_down_sync(FULLMASK, zs, 1);
    if (lane == 31 && s + 1 < S) znext = sample_depth(P, ray, s + 1);
    float dist = (s + 1 < S) ? __fsub_rn(znext, zs) : 1e10f;
    dist = __fmul_rn(dist, dnorm);
    const float* row = ring + (size_t)(q * 32 + lane) * P.out_ch;
    float c[INRF_RAW_BASE];
#pragma unroll
    for (int i = 0; i < INRF_RAW_BASE; ++i) c[i] = row[i];
    float sig = c[3];
    if (P.f.noise != nullptr) sig = __fadd_rn(sig, __ldg(P.f.noise + ray * S + s));
    const float alpha = __fsub_rn(1.f, expf(-__fmul_rn(fmaxf(sig, 0.f), dist)));
    float p = __fadd_rn(__fsub_rn(1.f, alpha), 1e-10f);
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const float t = __shfl_up_sync(FULLMASK, p, o); if (lane >= o) p *= t; }
    float excl = __shfl_up_sync(FULLMASK, p, 1);
    if (lane == 0) excl = 1.f;
    const float T = rs.carry * excl;
    rs.carry *= __shfl_sync(FULLMASK, p, 31);
    const float w = alpha * T;
    rs.acc[0] += w * c[0]; rs.acc[1] += w * c[1]; rs.acc[2] += w * c[2];
    rs.acc[3] += w * c[4]; rs.acc[4] += w * c[5]; rs.acc[5] += w * c[6];
    rs.acc[6] += w * c[7];
    rs.acc[7] += w * c[8]; rs.acc[8] += w * c[9]; rs.acc[9] += w * c[10];
    rs.acc[10] += w * zs;
    rs.acc[11] += w;
    if (SAMPLER) {
      if (s0 == 0) { rs.zc[0] = zs; rs.wc[0] = w; } else { rs.zc[1] = zs; rs.wc[1] = w; }
    }
    if (P.C > 0) {                                       // semantic logits: lanes stride over channels, weights by shuffle
      const float* r0 = ring + (size_t)(q * 32) * P.out_ch + INRF_RAW_BASE;
      const int nslab = (P.C + 31) >> 5;                 // channel slabs of 32 (C <= 112: at most 4)
      for (int j0 = 0; j0 < 32; j0 += 8) {               // 8 rows per batch: their loads are independent and issue back to back
        float v[8][4];
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          const float* rj = r0 + (size_t)(j0 + jj) * P.out_ch;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int e = lane + 32 * i;
            v[jj][i] = (i < nslab && e < P.C) ? rj[e] : 0.f;
          }
        }
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {                 // same accumulation order as k_raw2outputs: samples ascending
          const float wj = __shfl_sync(FULLMASK, w, j0 + jj);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int e = lane + 32 * i;
            if (i < nslab && e < P.C) rs.ex[i] += wj * v[jj][i];
          }
        }
      }
    }
    if (s0 + 32 == S) {                                  // ray complete: record (k_raw2outputs' epilogue)
      float a[12];
#pragma unroll
      for (int i = 0; i < 12; ++i) a[i] = wsum(rs.acc[i]);
      const float accw = a[11];
      const float bg = P.f.white_bkgd ? (1.f - accw) : 0.f;
      float* r = P.f.rec + ray * (INRF_REC_BASE + P.C);
      if (lane == 0) {
        r[0] = a[0] + bg; r[1] = a[1] + bg; r[2] = a[2] + bg;
        const float depth = a[10];
        const float ratio = depth / accw;                // 0/0 -> NaN, kept (appendix A8)
        const float mx = (ratio != ratio) ? ratio : fmaxf(1e-10f, ratio);
        r[3] = 1.f / mx;
        r[4] = accw;
        r[5] = a[3] + bg; r[6] = a[4] + bg; r[7] = a[5] + bg;
        r[8] = a[6] + bg;
        r[9] = a[7]; r[10] = a[8]; r[11] = a[9];          // residual: no background (A9)
        r[12] = depth;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int e = lane + 32 * i;
        if (e < P.C) r[INRF_REC_BASE + e] = P.f.white_bkgd ? rs.ex[i] + (1.f - accw) : rs.ex[i];   // model_utils.py:113-114
      }
      sy.stamp(5);
      if (SAMPLER) resample_ray(P, ray, rs, lane);
      sy.stamp(6);
    }
    sy.stamp(2);
  }
  __syncwarp();
  if (lane == 0) mbar_arrive(sy.addr(B_RAW_FREE + slot));
  sy.stamp(4);
}

// swizzled byte offsets of the 8 16-byte units of this thread's row inside a chunk
struct RowAddr {
  uint32_t unit[8];
};

template <bool STASH>
__device__ __forceinline__ void front_end(const Params& P, Sync& sy, uint32_t smem_base, int row) {
  // INRF_TC_EXP & 8 (A/B): the ray back end on front-end warps 0 and 2, after each tile's encoding (its first placement)
  RayState rs;
  rs.carry = 1.f;
  const bool backend = !STASH && P.fuse && (P.exp_flags & 8) && (row < 32 || (row >= 64 && row < 96));
  const int which = row >= 64 ? 1 : 0;
  const int blane = row & 31;
  const bool sampler = P.f.n_importance > 0;
  RowAddr ra;
#pragma unroll
  for (int u = 0; u < 8; ++u) ra.unit[u] = (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((u ^ (row & 7)) << 4));
  for (int it = 0; it < P.n_iter; ++it) {
    const int64_t tile = tile_of(P, it);                        // may run past the end: rows clamp, stores are masked
    sy.tile = (int)tile;
    int64_t m = tile * TILE_M + row;
    if (m >= P.a.M) m = P.a.M - 1;
    float x[3], d[3];
    const bool pre_embedded = P.a.emb != nullptr;
    const float* e = pre_embedded ? P.a.emb + m * (PE_PTS + PE_DIR) : nullptr;
    if (!pre_embedded) {
      if (P.a.rays != nullptr || P.a.cam_on) {
        const int64_t n = m / P.a.S;
        float ray[11];
        get_ray(P, n, ray);
        const float zv = sample_depth(P, n, (int)(m - n * P.a.S));
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          x[i] = __fadd_rn(ray[i], __fmul_rn(ray[3 + i], zv));   // o + d z (run_nerf.py:488)
          d[i] = ray[8 + i];
        }
      } else {
#pragma unroll
        for (int i = 0; i < 3; ++i) { x[i] = __ldg(P.a.pts + m * 3 + i); d[i] = __ldg(P.a.viewdirs + m * 3 + i); }
      }
      if (P.a.pe_scale != 1.f) {
#pragma unroll
        for (int i = 0; i < 3; ++i) x[i] = __fdiv_rn(x[i], P.a.pe_scale);
      }
    }
    // encode into packed halves (registers) while the previous tile still owns the PE|DIR region
    uint32_t pw[32], dw[16];
    {
      float pe[64];
      if (pre_embedded) {
#pragma unroll
        for (int i = 0; i < 63; ++i) pe[i] = __ldg(e + i);
      } else {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          float s[10], c[10];
          pe_coord(x[i], 10, s, c);
          pe[i] = x[i];
#pragma unroll
          for (int k = 0; k < 10; ++k) { pe[3 + 6 * k + i] = s[k]; pe[3 + 6 * k + 3 + i] = c[k]; }
        }
      }
      pe[63] = 0.f;
#pragma unroll
      for (int i = 0; i < 32; ++i) pw[i] = pack_h2(pe[2 * i], pe[2 * i + 1]);
    }
    {
      float de[32];
      if (pre_embedded) {
#pragma unroll
        for (int i = 0; i < 27; ++i) de[i] = __ldg(e + 63 + i);
      } else {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          float s[10], c[10];
          pe_coord(d[i], 4, s, c);
          de[i] = d[i];
#pragma unroll
          for (int k = 0; k < 4; ++k) { de[3 + 6 * k + i] = s[k]; de[3 + 6 * k + 3 + i] = c[k]; }
        }
      }
#pragma unroll
      for (int i = 27; i < 32; ++i) de[i] = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) dw[i] = pack_h2(de[2 * i], de[2 * i + 1]);
    }
    sy.wait(B_F_FREE);
    if (STASH) sy.wait(B_STV_DONE);            // the stash copy of the previous tile's relu(views') has been read out of this region
#pragma unroll
    for (int u = 0; u < 8; ++u) st_shared_v4(smem_base + SM_PE + ra.unit[u], pw[4 * u], pw[4 * u + 1], pw[4 * u + 2], pw[4 * u + 3]);
#pragma unroll
    for (int u = 0; u < 4; ++u) st_shared_v4(smem_base + SM_DIR + ra.unit[u], dw[4 * u], dw[4 * u + 1], dw[4 * u + 2], dw[4 * u + 3]);
    fence_async_smem();
    warp_arrive(sy.addr(B_F_READY), row & 31);
    if (STASH && tile * TILE_M < P.a.M) {        // training: the same operand images go to the stash
      unsigned char* g = P.a.stash_img + (tile * IMG_STASH_SLOTS + IS_PE) * (int64_t)IMG_BYTES;
#pragma unroll
      for (int u = 0; u < 8; ++u) st_global_v4(g + ra.unit[u], pw[4 * u], pw[4 * u + 1], pw[4 * u + 2], pw[4 * u + 3]);
      g += (IS_DIR - IS_PE) * IMG_BYTES;
#pragma unroll
      for (int u = 0; u < 4; ++u) st_global_v4(g + ra.unit[u], dw[4 * u], dw[4 * u + 1], dw[4 * u + 2], dw[4 * u + 3]);
    }
    if (!STASH && backend && it > 0) {
      if (sampler) ray_backend<true>(P, sy, it - 1, blane, rs, which); else ray_backend<false>(P, sy, it - 1, blane, rs, which);
    }
  }
  if (!STASH && backend && P.n_iter > 0) {
    if (sampler) ray_backend<true>(P, sy, P.n_iter - 1, blane, rs, which); else ray_backend<false>(P, sy, P.n_iter - 1, blane, rs, which);
  }
}

// ray back end of the fused renderer: warps 12 and 13 (idle otherwise in the inference kernels) share the rays by parity.
// It used to run on front-end warps 0 and 2 after each tile's encoding; the encoding alone keeps a front-end warp busy
// for ~11 600 of a tile's ~30 000 cycles, and once the lean issuer had shortened the tile the sum no longer fitted:
// the fused launches lost 3 us per tile waiting for gamma(x) of the next tile.
__device__ __forceinline__ void backend_role(const Params& P, Sync& sy, int which, int lane) {
  RayState rs;
  rs.carry = 1.f;
  const bool sampler = P.f.n_importance > 0;
  for (int it = 0; it < P.n_iter; ++it) {
    if (sampler) ray_backend<true>(P, sy, it, lane, rs, which); else ray_backend<false>(P, sy, it, lane, rs, which);
  }
}

// ------------------------------------------------------------------------------------------
// weight producer (converged warp, elected lane issues the bulk copies)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void producer(const Params& P, Sync& sy, uint32_t smem_base, int cl, int rank) {
  int slot = 0;
  const uint16_t mask = (uint16_t)((1u << cl) - 1u);
  const bool leader = elect_one();
  for (int it = 0; it < P.n_iter; ++it) {
    sy.tile = it;
    for (int b = 0; b < P.n_fills; ++b) {
      if (P.fault && it == 0 && b == 3) return;      // test hook: starve the ring -> the issuer's wait must trip the watchdog
      sy.wait(B_WEMPTY + slot);           // every CTA of the cluster has consumed this slot
      if (leader && !sy.dead) {
        const uint32_t bytes = (uint32_t)P.fill_bytes[b];
        const uint32_t dst = smem_base + SM_RING + slot * TC_SLOT_BYTES;
        mbar_expect_tx(sy.addr(B_WFULL + slot), bytes);
        if (cl == 1) {
          bulk_g2s(dst, P.blocks + P.fill_off[b], bytes, sy.addr(B_WFULL + slot));
        } else {
          // each CTA fetches 1/cl of the tile from L2 and multicasts it into every CTA's slot
          const uint32_t part = bytes / (uint32_t)cl;
          bulk_g2s_mc(dst + rank * part, P.blocks + P.fill_off[b] + rank * part, part, sy.addr(B_WFULL + slot), mask);
        }
      }
      __syncwarp();
      slot = (slot + 1 == P.ns) ? 0 : slot + 1;
    }
  }
}

// ------------------------------------------------------------------------------------------
// stash warp (training forward): every fp16 operand chunk the epilogue builds in shared memory is, as it stands, the
// 16 KB image the backward reads - one elected lane copies it to HBM with a bulk store as soon as the chunk's
// "written" barrier completes (the same barrier that releases it to the MMA issuer).  The epilogue warps used to
// write these images themselves: 4 x 16 B per thread and chunk into 32 different lines per instruction, more LSU
// wavefronts per tile than the tensor pipe has cycles (the training forward ran at 47 us per tile, 2.3x the inference
// kernel).  B_ST_DONE[c] / B_STV_DONE tell the epilogue (and the front end, for the PE|DIR|V region) that the copy has
// finished READING the chunk, so it may be rewritten in place.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ void stasher(const Params& P, Sync& sy, uint32_t smem_base) {
  const bool leader = elect_one();
  const bool sem = P.C > 0;
  const uint32_t H = smem_base + SM_H, V = smem_base + SM_V;
  int pending = -1;                          // "read done" barrier of the copy in flight
  for (int it = 0; it < P.n_iter; ++it) {
    const int64_t tile = tile_of(P, it);
    sy.tile = (int)tile;
    const bool live = tile * TILE_M < P.a.M && !(P.stash_abl & 2);   // tiles past the end of the batch are not stored
    unsigned char* simg = P.a.stash_img + (live ? tile : 0) * IMG_STASH_SLOTS * (int64_t)IMG_BYTES;
    // One copy stays in flight: the "has been read" signal of a chunk is given when the NEXT copy has been issued (the
    // wait for a copy's own read would put ~0.5 us per chunk, 40 chunks per tile, on the epilogue's critical path -
    // measured: 41 us per tile).  The next copy never depends on that signal, so the chain cannot deadlock.
    auto chunk = [&](int ready_bar, int done_bar, uint32_t src, int slot, uint32_t bytes) {
      sy.wait(ready_bar);                    // the writers fenced (fence.proxy.async) before arriving
      if (leader && !sy.dead) {
        if (live) {
          bulk_s2g(simg + (int64_t)slot * IMG_BYTES, src, bytes);
          bulk_wait_read1();                 // every copy but the one just issued has read its source
        } else {
          bulk_wait_read0();
        }
        if (pending >= 0) mbar_arrive(sy.addr(pending));
      }
      __syncwarp();
      pending = done_bar;
    };
    for (int l = 0; l < 8; ++l)
      for (int c = 0; c < 4; ++c) chunk(B_A_READY + c, B_ST_DONE + c, H + c * CHUNK, IS_H + 4 * l + c, CHUNK);
    chunk(B_V_READY, B_STV_DONE, V, IS_V, 2 * CHUNK);
    for (int c = 0; c < 4; ++c) chunk(B_A_READY + c, B_ST_DONE + c, H + c * CHUNK, IS_AS + c, CHUNK);
    if (sem)
      for (int c = 0; c < 2; ++c) chunk(B_A_READY + c, B_ST_DONE + c, H + c * CHUNK, IS_S1 + c, CHUNK);
  }
  if (leader) {
    bulk_wait_all0();                        // the images are complete in global memory before the CTA retires
    if (pending >= 0 && !sy.dead) mbar_arrive(sy.addr(pending));
  }
  __syncwarp();
}

// ------------------------------------------------------------------------------------------
// MMA issuer: converged warp (descriptor arithmetic stays on the uniform datapath), one elected
// lane issues tcgen05.mma / tcgen05.commit
// ------------------------------------------------------------------------------------------
struct Issuer {
  Sync& sy;
  uint32_t smem_base, tmem;
  int slot;
  int ns;
  uint32_t ones;          // shared-memory address of the constant "ones" A tile (absolute: smem_base may be a shifted role base)
  int cl;
  int bias_mma;
  int no_weights;
  bool leader;
  // probe-ahead state: the barriers of the NEXT fill are tested (non-blocking) before the MMAs of
  // the current fill are issued, so the mbarrier round trip overlaps the issue of tcgen05.mma
  uint32_t pw_ok, pa_ok, pw_par, pa_par;
  int pa_bar;             // activation barrier of the probed fill or -1
  __device__ __forceinline__ uint32_t slot_addr() const { return smem_base + SM_RING + slot * TC_SLOT_BYTES; }
  __device__ __forceinline__ void commit(int bar) {
    if (leader) tc_commit(sy.addr(bar));
    __syncwarp();
  }
  // start probing the fill that will be consumed next (ring slot `slot`), optionally with an activation barrier
  __device__ __forceinline__ void probe(int act_bar) {
    pa_bar = act_bar;
    sy.stamp(14);
    pw_par = sy.take_parity(B_WFULL + slot);
    pw_ok = no_weights ? 1u : mbar_test(sy.addr(B_WFULL + slot), pw_par);
    pa_ok = 1u;
    if (act_bar >= 0) {
      pa_par = sy.take_parity(act_bar);
      pa_ok = mbar_test(sy.addr(act_bar), pa_par);
    }
    sy.stamp(15);
  }
  // the probed fill is usable (falls back to blocking waits when the probe said "not yet")
  __device__ __forceinline__ void acquire() {
    if (!pw_ok) sy.slow(B_WFULL + slot, pw_par);
    if (!pa_ok) sy.slow(pa_bar, pa_par);
    sy.stamp(13);                                     // barriers passed
    tc_fence_after();
    sy.stamp(1);                                      // fill acquired (weights + activation chunk)
  }
  __device__ __forceinline__ void release_slot(int s) {   // MMAs reading slot s are done -> refill
    if (!no_weights && leader) {
      if (cl == 1) tc_commit(sy.addr(B_WEMPTY + s));
      else tc_commit_mc(sy.addr(B_WEMPTY + s), (uint16_t)((1u << cl) - 1u));
    }
    __syncwarp();
  }
  __device__ __forceinline__ void release() { release_slot(slot); }
  // end of a fill: release the slot, move on, probe the next fill's barriers (their round trip overlaps MMA execution).
  // The issuer warp paces the kernel - ~90 dependent instructions per fill against 512 cycles of MMAs (profiles/
  // r02_issuer_timeline.md) - so nothing is added here lightly: two run-time experiment flags in this path cost 7 % of the tile
  __device__ __forceinline__ void finish(int next_act) {
    release();
    advance();
    if (next_act != -2) probe(next_act);
  }
  __device__ __forceinline__ void advance() { slot = (slot + 1 == ns) ? 0 : slot + 1; sy.stamp(2); }   // fill issued
  // K = 16*KSTEPS of A chunk `a_chunk` times the operand tile at byte offset `b_off` of slot `sl`
  template <int KSTEPS>
  __device__ __forceinline__ void mma(uint32_t a_chunk, uint32_t b_addr, int n, uint32_t col, bool first) {
    const uint64_t ad = make_desc(a_chunk);
    const uint64_t bd = make_desc(b_addr);
    const uint32_t id = make_idesc(n);
    if (leader) {
#pragma unroll
      for (int k = 0; k < KSTEPS; ++k)
        tc_mma(tmem + col, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), id, (first && k == 0) ? 0u : 1u);
    }
    __syncwarp();
  }
  // the same with the A operand in tensor memory: 16*KSTEPS K elements starting at TMEM column a_col (8 columns per step)
  template <int KSTEPS>
  __device__ __forceinline__ void mma_ts(uint32_t a_col, uint32_t b_addr, int n, uint32_t col, bool first) {
    const uint64_t bd = make_desc(b_addr);
    const uint32_t id = make_idesc(n);
    if (leader) {
#pragma unroll
      for (int k = 0; k < KSTEPS; ++k)
        tc_mma_ts(tmem + col, tmem + a_col + 8u * k, bd + (uint64_t)(2 * k), id, (first && k == 0) ? 0u : 1u);
    }
    __syncwarp();
  }
  template <int KSTEPS>
  __device__ __forceinline__ void fill_mma_ts(uint32_t a_col, int n, uint32_t col, bool first, int next_act) {
    acquire();
    mma_ts<KSTEPS>(a_col, slot_addr(), n, col, first);
    finish(next_act);
  }
  // last K chunk of a trunk layer, split hand-off: output columns [0, nf) first (N = nf, weight-tile rows [0, nf)), commit
  // `first_bar`, then the other 256 - nf columns.  Same products, same accumulation order per column as one N = 256 MMA.
  template <int KSTEPS>
  __device__ __forceinline__ void fill_mma_ts_split(uint32_t a_col, int nf, uint32_t col, bool first, int first_bar, int next_act) {
    acquire();
    mma_ts<KSTEPS>(a_col, slot_addr(), nf, col, first);
    commit(first_bar);
    mma_ts<KSTEPS>(a_col, slot_addr() + (uint32_t)nf * 128u, 256 - nf, col + (uint32_t)nf, first);
    finish(next_act);
  }
  // whole fill = one operand tile.  `next_act`: activation barrier of the FOLLOWING fill (-1 none,
  // -2: do not probe ahead - the caller probes after doing something else)
  template <int KSTEPS>
  __device__ __forceinline__ void fill_mma(uint32_t a_chunk, int n, uint32_t col, bool first, int next_act) {
    acquire();
    mma<KSTEPS>(a_chunk, slot_addr(), n, col, first);    // MMAs first: their issue is on the critical path
    finish(next_act);                                     // the probe's round trip overlaps MMA execution
  }
  // two N=128 accumulators initialised from the two 128-row bias sub-blocks of ONE fill (albedo1 | shading1 in TS mode)
  __device__ __forceinline__ bool bias2(uint32_t col_a, uint32_t col_b, int next_act) {
    acquire();
    if (bias_mma && leader) {
      tc_mma(tmem + col_a, make_desc_flat(ones, 128, 0), make_desc_flat(slot_addr(), 128, 256), make_idesc(128), 0u);
      tc_mma(tmem + col_b, make_desc_flat(ones, 128, 0), make_desc_flat(slot_addr() + 4096, 128, 256), make_idesc(128), 0u);
    }
    __syncwarp();
    release();
    advance();
    if (next_act != -2) probe(next_act);
    return bias_mma != 0;
  }
  // accumulator columns [col, col+n) := bias (one K=16 MMA of the constant "ones" tile)
  __device__ __forceinline__ bool bias(int n, uint32_t col, int next_act) {
    acquire();
    if (bias_mma && leader)
      tc_mma(tmem + col, make_desc_flat(ones, 128, 0), make_desc_flat(slot_addr(), 128, 256), make_idesc(n), 0u);
    __syncwarp();
    finish(next_act);
    return bias_mma != 0;
  }
};

// HY (hybrid, round 2): trunk layers 1..7 take their A operand from tensor memory like the TS variant (the epilogue packs
// layers 0..6 into the drained accumulator), layer 7's output goes to the shared-memory H buffer and the tail below runs
// unchanged - the networks whose tail does not fit the TS layout (Semantic_NeRF: views' | sem1 | albedo1 | shading1 need
// 512 accumulator columns next to the packed trunk output; endpoint feature) still drop seven layers of activation
// traffic from shared memory.
template <bool HY>
__device__ __forceinline__ void issuer(const Params& P, Sync& sy, uint32_t smem_base, uint32_t tmem, int cl, uint32_t ones) {
  Issuer I{sy, smem_base, tmem, 0, P.ns, ones, cl, P.bias_mma, P.no_weights, elect_one(), 1u, 1u, 0u, 0u, -1};
  const uint32_t H = smem_base + SM_H, PE = smem_base + SM_PE, DIR = smem_base + SM_DIR, V = smem_base + SM_V;
  const bool sem = P.C > 0;
  const int nv = sem ? 256 : 128;       // views' [| sem1] width
  I.probe(-1);                          // first fill of the first tile (layer-0 bias)
  for (int it = 0; it < P.n_iter; ++it) {
    sy.tile = it;
    // The two 256-column accumulators swap roles every tile: A0 (even layers, views', the narrow heads) is the
    // accumulator the PREVIOUS tile used for its odd layers and albedo1|shading1, which is fully drained as soon as
    // that tile's albedo2/shading2 MMAs have been issued - so layer 0 of this tile runs while the previous tile's
    // epilogue warps are still reading the narrow heads out of its own A0 and writing the rows (the tile tail).
    const uint32_t A0 = (it & 1) ? 256u : 0u, A1 = 256u - A0;
    sy.wait(B_F_READY);                 // gamma(x), gamma(d) of this tile
    // ---- trunk layer 0: K = 64 (gamma(x)) -> accumulator A0 --------------------------------------
    {
      const bool init = I.bias(256, A0, -1);
      I.fill_mma<4>(PE, 256, A0, !init, -1);
      I.commit(B_ACC_FULL + 0);
    }
    sy.wait(B_TAIL_DONE);               // previous tile's narrow-head / logit columns (in this tile's A1) have been read
    // ---- trunk layers 1..7 ------------------------------------------------------------------------
    for (int l = 1; l < 8; ++l) {
      const uint32_t acc = (l & 1) ? A1 : A0;
      bool first = !I.bias(256, acc, l == 5 ? -1 : B_A_READY + 0);
      if (l == 5) {                      // skip connection: [gamma(x), h] -> K = 64 + 256
        I.fill_mma<4>(PE, 256, acc, first, B_A_READY + 0);
        first = false;
      }
      for (int c = 0; c < 4; ++c) {
        if (HY && c == 3 && P.split_nf) I.fill_mma_ts_split<4>(((l & 1) ? A0 : A1) + 96u, P.split_nf, acc, first, B_ACC_FIRST + (l & 1), -1);
        else if (HY) I.fill_mma_ts<4>(((l & 1) ? A0 : A1) + 32u * c, 256, acc, first, c < 3 ? B_A_READY + c + 1 : -1);   // packed output of layer l-1
        else I.fill_mma<4>(H + c * CHUNK, 256, acc, first, c < 3 ? B_A_READY + c + 1 : -1);
        first = false;
      }
      I.commit(B_ACC_FULL + (l & 1));
    }
    // ---- views' [| sem1] on the trunk output -> accumulator A0 (drained since layer 6) --------------
    {
      bool first = !I.bias(nv, A0, B_A_READY + 0);
      for (int c = 0; c < 4; ++c) {       // all four waits also prove layer 7's accumulator (A1) is drained
        I.fill_mma<4>(H + c * CHUNK, nv, A0, first, c < 3 ? B_A_READY + c + 1 : -1);
        first = false;
      }
      I.fill_mma<2>(DIR, 128, A0, false, -1);
      I.commit(B_ACC_FULL + 0);
    }
    // ---- albedo1 | shading1 on the trunk output -> accumulator A1 -----------------------------------
    {
      bool first = !I.bias(256, A1, -1);
      for (int c = 0; c < 4; ++c) {
        I.fill_mma<4>(H + c * CHUNK, 256, A1, first, -1);
        first = false;
      }
      I.commit(B_H_FREE);                // last reader of the trunk output: relu(albedo1|shading1) may land in H
      I.commit(B_ACC_FULL + 1);
    }
    // ---- residual head on relu(views'): 16 x 128 -> accumulator A0 cols [0,16) -----------------------
    sy.wait(B_V_READY);
    {
      I.acquire();
      const uint32_t b_addr = I.slot_addr();
      I.mma<4>(V, b_addr, 16, A0, true);
      I.mma<4>(V + CHUNK, b_addr + 2048, 16, A0, false);
      I.release();
      I.advance();
      I.probe(-1);
    }
    I.commit(B_F_FREE);                  // PE | DIR | V region may be rewritten by the front end
    // ---- albedo2 / shading2 on relu(albedo1 | shading1): 16 x 256 -> accumulator A0 cols [16,32) -----
    //      (its four A_READY waits also prove that A1 is drained: the next tile's layer 0 may overwrite it)
    {
      I.acquire();
      const uint32_t b_addr = I.slot_addr();
      for (int c = 0; c < 4; ++c) {
        sy.wait(B_A_READY + c);
        tc_fence_after();
        I.mma<4>(H + c * CHUNK, b_addr + 2048 * c, 16, A0 + 16, c == 0);
      }
      I.release();
      I.advance();
      I.probe(-1);
    }
    if (sem) I.commit(B_H_FREE);         // relu(sem1) may overwrite H chunks 0,1
    I.commit(B_SMALL_FULL);
    // ---- semantic logits on relu(sem1): C x 128 -> accumulator A0 cols [32, 32 + sem_rows) (the views' | sem1
    //      columns there were drained by the V / sem1 epilogues) ---------------------------------------------
    if (sem) {
      I.acquire();
      const uint32_t b_addr = I.slot_addr();
      for (int c = 0; c < 2; ++c) {
        sy.wait(B_A_READY + c);
        tc_fence_after();
        I.mma<4>(H + c * CHUNK, b_addr + (uint32_t)(P.sem_rows * 128 * c), P.sem_rows, A0 + 32, c == 0);
      }
      I.release();
      I.advance();
      I.probe(-1);
      I.commit(B_SEM2_FULL);
    }
  }
}

// ------------------------------------------------------------------------------------------
// epilogue warps.  Warp (q, jj): TMEM lanes 32q..32q+31 (one row per thread); of every 64-column
// K chunk it owns columns jj*32..jj*32+31, so all 8 warps finish chunk 0 first, then 1, 2, 3.
// ------------------------------------------------------------------------------------------
// 32 accumulator columns -> (+bias) -> ReLU -> fp16 -> 4 swizzled 16-byte stores
template <bool ADD_BIAS, bool SIGMA, bool GOUT>
__device__ __forceinline__ void epi_store32(const uint32_t* v, const float* __restrict__ bias, uint32_t dst_chunk,
                                            const RowAddr& ra, int unit0, const float* alpha_w_smem, float* sigma_acc,
                                            float* gout, uint32_t* gmask, uint32_t& amax) {
  float f[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
  if (ADD_BIAS) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(bias) + i);
      f[4 * i] += t.x; f[4 * i + 1] += t.y; f[4 * i + 2] += t.z; f[4 * i + 3] += t.w;
    }
  }
  if (SIGMA) {                                // sigma head: fp32 dot on the un-rounded ReLU output
    float s = *sigma_acc;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 t = *reinterpret_cast<const float4*>(alpha_w_smem + 4 * i);
      s = fmaf(fmaxf(f[4 * i], 0.f), t.x, s);
      s = fmaf(fmaxf(f[4 * i + 1], 0.f), t.y, s);
      s = fmaf(fmaxf(f[4 * i + 2], 0.f), t.z, s);
      s = fmaf(fmaxf(f[4 * i + 3], 0.f), t.w, s);
    }
    *sigma_acc = s;
  }
  if (GOUT && gout != nullptr) {              // endpoint feature rows (fp32, post-ReLU)
#pragma unroll
    for (int i = 0; i < 32; ++i) gout[i] = fmaxf(f[i], 0.f);
  }
  // training: ReLU decisions of these 32 columns as one word - bit j = column 2j, bit 16 + j = column 2j + 1 (the order
  // three instructions per packed pair produce: a stored half h >= 0 is non-zero iff bit 15 of h + 0x7fff is set)
  uint32_t bits = 0u;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    uint32_t pk[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      pk[i] = pack_relu_sat_h2(f[8 * u + 2 * i], f[8 * u + 2 * i + 1]);
      amax = hmax2_u32(amax, pk[i]);          // running maximum of the stored halves: fp16-range check at the tile end
      if (gmask != nullptr) bits = (bits >> 1) | ((pk[i] + 0x7fff7fffu) & 0x80008000u);
    }
    st_shared_v4(dst_chunk + ra.unit[unit0 + u], pk[0], pk[1], pk[2], pk[3]);
  }
  if (gmask != nullptr) *gmask = bits;         // lanes = consecutive rows: one coalesced 128-byte store per warp
}

// A 256-column accumulator -> 4 H chunks, two chunks per TMEM load batch.
// taddr: this thread's lane + accumulator base column; bias: 256 floats (fallback path only).
// MODE 0: plain layer, 1: also accumulate the sigma head (trunk layer 7), 2: also emit fp32 rows (endpoint)
template <int MODE>
__device__ __forceinline__ void epi_layer(bool add_bias, uint32_t taddr, const float* bias, uint32_t dst0, int n_chunks,
                                          const RowAddr& ra, int jj, int lane, Sync& sy, int free_bar0, int ready_bar0,
                                          const float* alpha_smem, float* sigma_acc, float* gout0, uint32_t* gmask0,
                                          uint32_t& amax, int st_bar0 = -1, bool st_per_chunk = true) {
#pragma unroll 1
  for (int cp = 0; cp < n_chunks; cp += 2) {
    uint32_t v0[32], v1[32];
    tmem_ld32(taddr + cp * 64 + jj * 32, v0);
    tmem_ld32(taddr + (cp + 1) * 64 + jj * 32, v1);
    tmem_ld_wait();
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int c = cp + half;
      const int col = c * 64 + jj * 32;
      const uint32_t* v = half == 0 ? v0 : v1;
      if (free_bar0 >= 0 && c == 0) sy.wait(free_bar0);     // one "H free" barrier for the whole layer
      // training: the stash warp's bulk copy of the chunk's previous contents has finished reading shared memory
      if (st_bar0 >= 0 && (st_per_chunk || c == 0)) sy.wait(st_bar0 + (st_per_chunk ? c : 0));
      float* g = (MODE == 2 && gout0) ? gout0 + col : nullptr;
      const float* aw = (MODE == 1) ? alpha_smem + col : nullptr;
      uint32_t* gi = gmask0 ? gmask0 + c * 256 : nullptr;            // mask words of image slot +c
      if (add_bias) epi_store32<true, MODE == 1, MODE == 2>(v, bias + col, dst0 + c * CHUNK, ra, jj * 4, aw, sigma_acc, g, gi, amax);
      else epi_store32<false, MODE == 1, MODE == 2>(v, bias + col, dst0 + c * CHUNK, ra, jj * 4, aw, sigma_acc, g, gi, amax);
      fence_async_smem();
      tc_fence_before();
      if (ready_bar0 >= 0) warp_arrive(sy.addr(ready_bar0 + c), lane);
    }
  }
}

__device__ __forceinline__ float sigmoid_(float x) { return 1.f / (1.f + expf(-x)); }

template <int MODE>
__device__ __forceinline__ void epi_acc_ts(int fine, uint32_t src, uint32_t dst, int n_chunks, int jj, int lane, Sync& sy, int ready_bar0,
                                           const float* alpha_smem, float* sigma_acc, uint32_t& amax, int pair_bar, int c_begin = 0);
template <int MODE>
__device__ __forceinline__ void epi_trunk_ts(const Params& P, int l, uint32_t R, int jj, int lane, Sync& sy,
                                             const float* alpha_smem, float* sigma_acc, uint32_t& amax, int pair_bar);

template <bool STASH, bool HY>
__device__ __forceinline__ void epilogue(const Params& P, Sync& sy, uint8_t* smem, uint32_t smem_base, uint32_t tmem,
                                         int q, int jj, int lane) {
  const int row = q * 32 + lane;
  const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
  float* s_sig = reinterpret_cast<float*>(smem + SM_SIG);
  const float* s_alpha = reinterpret_cast<const float*>(smem + SM_ALPHA);
  const uint32_t H = smem_base + SM_H, V = smem_base + SM_V;
  const bool add_bias = P.bias_mma == 0;
  const bool sem = P.C > 0;
  RowAddr ra;
#pragma unroll
  for (int u = 0; u < 8; ++u) ra.unit[u] = (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((u ^ (row & 7)) << 4));
  for (int it = 0; it < P.n_iter; ++it) {
    const int64_t tile = tile_of(P, it);                        // may run past the end: stores are masked
    sy.tile = (int)tile;
    const int64_t m = tile * TILE_M + row;
    const bool valid = P.fuse ? true : (m < P.a.M);             // fused: every row of the ring slot is written
    // raw row destination: the caller's [M, out_ch] tensor, or (fused) this tile's slot of the L2-resident ring
    float* grow = P.fuse ? P.f.ring + (((size_t)blockIdx.x * 2 + (it & 1)) * TILE_M + row) * P.out_ch
                         : P.a.raw + (m < P.a.M ? m : 0) * P.out_ch;
    // training: every operand chunk built here also goes to the stash - copied out of shared memory by the stash warp
    // (stasher), which needs the chunk intact until its bulk copy has read it: B_ST_DONE / B_STV_DONE
    const int ST = (STASH && !(P.stash_abl & 4)) ? B_ST_DONE : -1, STV = (STASH && !(P.stash_abl & 4)) ? B_STV_DONE : -1;
    // ... and its ReLU mask goes out as one bit word per thread and chunk (slots IS_MASK.., common.cuh)
    uint32_t* smask = (STASH && tile * TILE_M < P.a.M && !(P.stash_abl & 1))
        ? reinterpret_cast<uint32_t*>(P.a.stash_img + (tile * IMG_STASH_SLOTS + IS_MASK) * (int64_t)IMG_BYTES) + jj * 128 + row : nullptr;
#define MSLOT(s) (smask ? smask + ((s) - IS_H) * 256 : nullptr)
    // ---- trunk: accumulator of layer l -> A operand of layer l+1 (in place in H) ------------------
    float sig = 0.f;
    uint32_t amax = 0u;                                  // running max of every fp16 activation this thread stored
    const uint32_t A0 = (it & 1) ? 256u : 0u, A1 = 256u - A0;   // accumulator roles swap every tile (see issuer)
    for (int l = 0; l < 8; ++l) {
      if (HY && l < 7) {                               // hybrid: packed halves back into the drained accumulator (see issuer)
        epi_trunk_ts<0>(P, l, lane_addr + ((l & 1) ? A1 : A0), jj, lane, sy, nullptr, nullptr, amax, 2 + q);
        continue;
      }
      if (HY && P.split_nf) sy.wait(B_ACC_FIRST + 1);  // layer 7 of the hybrid goes to shared memory two chunks per batch
      sy.wait(B_ACC_FULL + (l & 1));
      tc_fence_after();
      if (l == 7) epi_layer<1>(add_bias, lane_addr + A1, P.bias + l * 256, H, 4, ra, jj, lane, sy, -1, B_A_READY, s_alpha, &sig, nullptr, MSLOT(IS_H + 28), amax, ST);
      else epi_layer<0>(add_bias, lane_addr + ((l & 1) ? A1 : A0), P.bias + l * 256, H, 4, ra, jj, lane, sy, -1, B_A_READY, nullptr, nullptr, nullptr, MSLOT(IS_H + 4 * l), amax, ST);
    }
    s_sig[row * 2 + jj] = sig;                         // fixed-order sum later: deterministic sigma
    // ---- relu(views') -> V (PE|DIR region; its last readers finished with accumulator 0); runs while
    //      the tensor core works on albedo1|shading1 ---------------------------------------------------
    sy.wait(B_ACC_FULL + 0);
      tc_fence_after();
    if (P.a.endpoint)
      epi_layer<2>(add_bias, lane_addr + A0, P.bias + TCB_VIEWS, V, 2, ra, jj, lane, sy, -1, -1, nullptr, nullptr,
                   valid ? grow + INRF_RAW_BASE + P.C : nullptr, MSLOT(IS_V), amax, STV, false);
    else
      epi_layer<0>(add_bias, lane_addr + A0, P.bias + TCB_VIEWS, V, 2, ra, jj, lane, sy, -1, -1, nullptr, nullptr, nullptr, MSLOT(IS_V), amax, STV, false);
    warp_arrive(sy.addr(B_V_READY), lane);
    // ---- relu(albedo1 | shading1) -> H (in place over the trunk output) ---------------------------------
    sy.wait(B_ACC_FULL + 1);
      tc_fence_after();
    epi_layer<0>(add_bias, lane_addr + A1, P.bias + TCB_ALBSH, H, 4, ra, jj, lane, sy, B_H_FREE, B_A_READY, nullptr, nullptr, nullptr, MSLOT(IS_AS), amax, ST);
    // ---- relu(sem1) -> H chunks 0,1 (after the albedo2/shading2 MMAs released them) -------------------
    if (sem)
      epi_layer<0>(add_bias, lane_addr + A0 + 128, P.bias + TCB_SEM1, H, 2, ra, jj, lane, sy, B_H_FREE, B_A_READY, nullptr, nullptr, nullptr, MSLOT(IS_S1), amax, ST);
#undef MSLOT
    // ---- heads -> raw row ------------------------------------------------------------------------------
    sy.wait(B_SMALL_FULL);
    if (sem) sy.wait(B_SEM2_FULL);
    if (P.fuse) sy.wait(B_RAW_FREE + (it & 1));           // the back-end warp is done with this slot (two tiles ago)
    tc_fence_after();
    __syncwarp();
    asm volatile("bar.sync 1, 256;" ::: "memory");     // both sigma partials of every row are in smem
    if (jj == 0) {
      uint32_t v[32];
      tmem_ld32(lane_addr + A0, v);
      tmem_ld_wait();
      if (valid) {
        float res[3], alb[3], sh;
#pragma unroll
        for (int i = 0; i < 3; ++i) res[i] = sigmoid_(__uint_as_float(v[i]) + __ldg(P.bias + TCB_RES + i));
#pragma unroll
        for (int i = 0; i < 3; ++i) alb[i] = sigmoid_(__uint_as_float(v[16 + i]) + __ldg(P.bias + TCB_ALB2 + i));
        sh = sigmoid_(__uint_as_float(v[19]) + __ldg(P.bias + TCB_SH2));
        const float sigma = (s_sig[row * 2] + s_sig[row * 2 + 1]) + __ldg(P.bias + TCB_ALPHA_B);
#pragma unroll
        for (int i = 0; i < 3; ++i) grow[i] = __fadd_rn(__fmul_rn(alb[i], sh), res[i]);
        grow[3] = sigma;
#pragma unroll
        for (int i = 0; i < 3; ++i) grow[4 + i] = alb[i];
        grow[7] = sh;
#pragma unroll
        for (int i = 0; i < 3; ++i) grow[8 + i] = res[i];
      }
    } else if (sem) {
      for (int c0 = 0; c0 < P.C; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(lane_addr + A0 + 32 + c0, v);
        tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c0 + i < P.C) grow[INRF_RAW_BASE + c0 + i] = __uint_as_float(v[i]) + __ldg(P.bias + TCB_SEM2 + c0 + i);
        }
      }
    }
    // fp16 range: a stored activation at the saturation value 65504 (0x7bff) means the tile left the range the
    // tensor-core path can represent; report it (next API call returns INRF_ERANGE) instead of passing inf on
    if (m < P.a.M && !sy.dead && ((amax & 0xffffu) >= 0x7bffu || (amax >> 16) >= 0x7bffu)) {   // (an abandoned launch computes garbage)
      if (atomicCAS(P.dbg + 8, 0, 1) == 0) status_raise(P.status, DST_F16_ACT, 0, (int)tile, blockIdx.x);
    }
    tc_fence_before();
    asm volatile("bar.sync 1, 256;" ::: "memory");     // s_sig may be rewritten by the next tile
    warp_arrive(sy.addr(B_TAIL_DONE), lane);
    if (P.fuse) {
      // hand the rows to the back-end warps AFTER releasing the tensor pipe: the fence waits for the ring stores to
      // be visible, which would otherwise sit on the tile-to-tile critical path
      __threadfence_block();
      warp_arrive(sy.addr(B_RAW_READY + (it & 1)), lane);
    }
  }
}

// ------------------------------------------------------------------------------------------
// "TS" variant (object network, inference): the hidden activations never touch shared memory.  The epilogue writes
// relu(accumulator) as packed fp16 pairs back INTO the tensor-memory columns it has just drained (tcgen05.st), and the next
// layer's MMAs take their A operand from there (tcgen05.mma with a TMEM A operand).  Per layer-tile this removes the 68 KB
// of A-operand reads and the 64 KB of activation stores from the 404 KB of shared-memory traffic that bound the SS kernel
// (DESIGN.md section 4b), and fence.proxy.async from the layer-to-layer chain.
//   layer l accumulates in R(l) = A0 (even l) / A1 (odd l); its packed output lives in R(l)[0,128) and is dead once layer
//   l+1 has been issued, i.e. before layer l+2 overwrites R(l).  Tail: h7 sits in A1[0,128); views' -> A0[0,128),
//   albedo1 -> A1[128,256), shading1 -> A0[128,256) (two N=128 accumulators fed from the same weight fill); their packed
//   ReLUs go to A0[0,64), A1[128,192), A0[128,192); residual -> A0[64,80), albedo2|shading2 -> A0[80,96).
// Two warps share each 32-lane quadrant, and a warp's packed store may land on columns the OTHER warp of the quadrant has
// only just loaded, so the pair meets at a 64-thread named barrier between "all loads of the batch are in registers" and
// the stores.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void issuer_ts(const Params& P, Sync& sy, uint32_t smem_base, uint32_t tmem, int cl, uint32_t ones) {
  Issuer I{sy, smem_base, tmem, 0, P.ns, ones, cl, P.bias_mma, P.no_weights, elect_one(), 1u, 1u, 0u, 0u, -1};
  const uint32_t PE = smem_base + SM_PE, DIR = smem_base + SM_DIR;
  I.probe(-1);
  for (int it = 0; it < P.n_iter; ++it) {
    sy.tile = it;
    sy.set_it(it);
    sy.stamp(3);                                      // tile begins
    const uint32_t A0 = (it & 1) ? 256u : 0u, A1 = 256u - A0;
    sy.wait(B_F_READY);
    sy.stamp(4);
    {
      const bool init = I.bias(256, A0, -1);
      I.fill_mma<4>(PE, 256, A0, !init, -1);
      I.commit(B_ACC_FULL + 0);
    }
    sy.wait(B_TAIL_DONE);
    sy.stamp(5);
    for (int l = 1; l < 8; ++l) {
      const uint32_t acc = (l & 1) ? A1 : A0, src = (l & 1) ? A0 : A1;      // input = packed output of layer l-1
      bool first = !I.bias(256, acc, l == 5 ? -1 : B_A_READY + 0);
      if (l == 5) {
        I.fill_mma<4>(PE, 256, acc, first, B_A_READY + 0);
        first = false;
      }
      for (int c = 0; c < 4; ++c) {
        if (c == 3 && P.split_nf) I.fill_mma_ts_split<4>(src + 96u, P.split_nf, acc, first, B_ACC_FIRST + (l & 1), -1);
        else I.fill_mma_ts<4>(src + 32u * c, 256, acc, first, c < 3 ? B_A_READY + c + 1 : -1);
        first = false;
      }
      I.commit(B_ACC_FULL + (l & 1));
    }
    // ---- views' on h7 (A1[0,128)) + gamma(d) -> A0[0,128) -------------------------------------------------
    {
      bool first = !I.bias(128, A0, B_A_READY + 0);
      for (int c = 0; c < 4; ++c) {
        I.fill_mma_ts<4>(A1 + 32u * c, 128, A0, first, c < 3 ? B_A_READY + c + 1 : -1);
        first = false;
      }
      I.fill_mma<2>(DIR, 128, A0, false, -1);
      I.commit(B_ACC_FULL + 0);
    }
    // ---- albedo1 -> A1[128,256), shading1 -> A0[128,256): both halves of every albedo1|shading1 weight fill -----
    {
      const bool init = I.bias2(A1 + 128, A0 + 128, -1);
      for (int c = 0; c < 4; ++c) {
        I.acquire();
        I.mma_ts<4>(A1 + 32u * c, I.slot_addr(), 128, A1 + 128, !init && c == 0);
        I.mma_ts<4>(A1 + 32u * c, I.slot_addr() + 16384, 128, A0 + 128, !init && c == 0);
        I.release();
        I.advance();
        I.probe(-1);
      }
      I.commit(B_ACC_FULL + 1);
    }
    // ---- residual head on relu(views') (A0[0,64)) -> A0[64,80) ---------------------------------------------------
    sy.wait(B_V_READY);
    sy.stamp(6);
    {
      I.acquire();
      const uint32_t b_addr = I.slot_addr();
      I.mma_ts<4>(A0 + 0, b_addr, 16, A0 + 64, true);
      I.mma_ts<4>(A0 + 32, b_addr + 2048, 16, A0 + 64, false);
      I.release();
      I.advance();
      I.probe(-1);
    }
    I.commit(B_F_FREE);
    // ---- albedo2 / shading2 on [relu(albedo1) (A1[128,192)) | relu(shading1) (A0[128,192))] -> A0[80,96) -----------
    {
      I.acquire();
      const uint32_t b_addr = I.slot_addr();
      for (int c = 0; c < 4; ++c) {
        sy.wait(B_A_READY + c);
        sy.stamp(7);
        tc_fence_after();
        I.mma_ts<4>((c < 2 ? A1 + 128 + 32u * c : A0 + 128 + 32u * (c - 2)), b_addr + 2048 * c, 16, A0 + 80, c == 0);
      }
      I.release();
      I.advance();
      I.probe(-1);
    }
    I.commit(B_SMALL_FULL);
  }
}

// 2 chunks (64 fp32 columns each) of the accumulator at `src` -> relu -> packed halves -> TMEM columns `dst` + 32 c + 16 jj.
// MODE 1 also accumulates the sigma head.  `pair_bar`: the quadrant's named barrier (two warps).
template <int MODE>
__device__ __forceinline__ void epi_batch_ts(uint32_t src, uint32_t dst, int col0, int jj, int lane, Sync& sy, int ready_bar0,
                                             const float* alpha_smem, float* sigma_acc, uint32_t& amax, int pair_bar) {
  uint32_t v0[32], v1[32];
  tmem_ld32(src + jj * 32, v0);
  tmem_ld32(src + 64 + jj * 32, v1);
  tmem_ld_wait();
  // both warps of the quadrant hold their columns in registers (immediate barrier ids keep ptxas from reserving all 16)
  switch (pair_bar) {
    case 2: asm volatile("bar.sync 2, 64;" ::: "memory"); break;
    case 3: asm volatile("bar.sync 3, 64;" ::: "memory"); break;
    case 4: asm volatile("bar.sync 4, 64;" ::: "memory"); break;
    default: asm volatile("bar.sync 5, 64;" ::: "memory"); break;
  }
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const uint32_t* v = half == 0 ? v0 : v1;
    float f[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
    if (MODE == 1) {
      float s = *sigma_acc;
      const float* aw = alpha_smem + col0 + half * 64 + jj * 32;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 t = *reinterpret_cast<const float4*>(aw + 4 * i);
        s = fmaf(fmaxf(f[4 * i], 0.f), t.x, s);
        s = fmaf(fmaxf(f[4 * i + 1], 0.f), t.y, s);
        s = fmaf(fmaxf(f[4 * i + 2], 0.f), t.z, s);
        s = fmaf(fmaxf(f[4 * i + 3], 0.f), t.w, s);
      }
      *sigma_acc = s;
    }
    uint32_t pk[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      pk[i] = pack_relu_sat_h2(f[2 * i], f[2 * i + 1]);
      amax = hmax2_u32(amax, pk[i]);
    }
    tmem_st16(dst + half * 32 + jj * 16, pk);
    tmem_st_wait();
    tc_fence_before();
    if (ready_bar0 >= 0) warp_arrive(sy.addr(ready_bar0 + half), lane);
  }
}

// one chunk per batch: a thread cannot convert while its own load is in flight (wait::ld waits for all of them), but the
// eight warps reach this point staggered by the read port (4 KB each, 64 B/clk), so with ONE x32 load per wait the port
// stays busy with other warps' loads while a warp converts, and K chunk c of the next layer is released after (c+1)/4 of
// the drain instead of after 1/2 and 1/1 of it
template <int MODE>
__device__ __forceinline__ void epi_chunk_ts(uint32_t src, uint32_t dst, int col, int jj, int lane, Sync& sy, int ready_bar,
                                             const float* alpha_smem, float* sigma_acc, uint32_t& amax, int pair_bar) {
  uint32_t v[32];
  tmem_ld32(src + jj * 32, v);
  tmem_ld_wait();
  sy.stamp(8);                                        // chunk in registers
  switch (pair_bar) {
    case 2: asm volatile("bar.sync 2, 64;" ::: "memory"); break;
    case 3: asm volatile("bar.sync 3, 64;" ::: "memory"); break;
    case 4: asm volatile("bar.sync 4, 64;" ::: "memory"); break;
    default: asm volatile("bar.sync 5, 64;" ::: "memory"); break;
  }
  float f[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
  if (MODE == 1) {
    float s = *sigma_acc;
    const float* aw = alpha_smem + col + jj * 32;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 t = *reinterpret_cast<const float4*>(aw + 4 * i);
      s = fmaf(fmaxf(f[4 * i], 0.f), t.x, s);
      s = fmaf(fmaxf(f[4 * i + 1], 0.f), t.y, s);
      s = fmaf(fmaxf(f[4 * i + 2], 0.f), t.z, s);
      s = fmaf(fmaxf(f[4 * i + 3], 0.f), t.w, s);
    }
    *sigma_acc = s;
  }
  uint32_t pk[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    pk[i] = pack_relu_sat_h2(f[2 * i], f[2 * i + 1]);
    amax = hmax2_u32(amax, pk[i]);
  }
  sy.stamp(9);                                        // converted
  tmem_st16(dst + jj * 16, pk);
  tmem_st_wait();
  tc_fence_before();
  if (ready_bar >= 0) warp_arrive(sy.addr(ready_bar), lane);
  sy.stamp(10);                                       // stored + arrived
}

// chunks [c_begin, n_chunks) (64 columns each) of the accumulator at `src` -> packed into `dst`, either two chunks per load
// batch or one
template <int MODE>
__device__ __forceinline__ void epi_acc_ts(int fine, uint32_t src, uint32_t dst, int n_chunks, int jj, int lane, Sync& sy, int ready_bar0,
                                           const float* alpha_smem, float* sigma_acc, uint32_t& amax, int pair_bar, int c_begin) {
  if (fine) {
#pragma unroll 1
    for (int c = c_begin; c < n_chunks; ++c)
      epi_chunk_ts<MODE>(src + 64 * c, dst + 32 * c, 64 * c, jj, lane, sy, ready_bar0 >= 0 ? ready_bar0 + c : -1, alpha_smem, sigma_acc, amax, pair_bar);
  } else {
#pragma unroll 1
    for (int c = c_begin; c < n_chunks; c += 2)
      epi_batch_ts<MODE>(src + 64 * c, dst + 32 * c, 64 * c, jj, lane, sy, ready_bar0 >= 0 ? ready_bar0 + c : -1, alpha_smem, sigma_acc, amax, pair_bar);
  }
}

// accumulator R of trunk layer l -> packed halves in place (the A operand of layer l+1).  With the split hand-off
// (Params.split_nf, layers 1..7) columns [0, split_nf) are complete - and drained - before the rest of the accumulator.
template <int MODE>
__device__ __forceinline__ void epi_trunk_ts(const Params& P, int l, uint32_t R, int jj, int lane, Sync& sy,
                                             const float* alpha_smem, float* sigma_acc, uint32_t& amax, int pair_bar) {
  int c0 = 0;
  if (l > 0 && P.split_nf) {
    sy.wait(B_ACC_FIRST + (l & 1));
    sy.stamp(11);                                     // first columns complete
    tc_fence_after();
    c0 = P.split_nf >> 6;
    epi_acc_ts<MODE>(P.ts_fine, R, R, c0, jj, lane, sy, B_A_READY, alpha_smem, sigma_acc, amax, pair_bar, 0);
  }
  sy.wait(B_ACC_FULL + (l & 1));
  sy.stamp(12);                                       // accumulator complete
  tc_fence_after();
  epi_acc_ts<MODE>(P.ts_fine, R, R, 4, jj, lane, sy, B_A_READY, alpha_smem, sigma_acc, amax, pair_bar, c0);
}

// Head values of a finished tile, held in registers until the tensor pipe is busy with the NEXT tile's layer 1: the sigmoids
// and the 11 strided row stores (~2 500 cycles) used to sit between "narrow heads complete" and the drain of the next
// tile's layer 0, i.e. on the tile-to-tile critical path with the tensor pipe idle (clock64 timeline).
struct PendingHeads {
  float h[7];          // pre-sigmoid residual3, albedo3, shading
  float sigma;         // sigma head incl. bias
  float* grow;         // destination row (caller's raw tensor or the CTA's ring slot)
  int slot;            // fused: ring slot (it & 1)
  bool valid, on;
};

__device__ __forceinline__ void finish_heads(const Params& P, Sync& sy, PendingHeads& pd, int jj, int lane) {
  if (!pd.on) return;
  pd.on = false;
  if (jj == 0) {
    if (P.fuse) sy.wait(B_RAW_FREE + pd.slot);        // the back end is done with this slot (two tiles ago)
    sy.stamp(7);
    if (pd.valid) {
      float res[3], alb[3], sh;
#pragma unroll
      for (int i = 0; i < 3; ++i) res[i] = sigmoid_(pd.h[i] + __ldg(P.bias + TCB_RES + i));
#pragma unroll
      for (int i = 0; i < 3; ++i) alb[i] = sigmoid_(pd.h[3 + i] + __ldg(P.bias + TCB_ALB2 + i));
      sh = sigmoid_(pd.h[6] + __ldg(P.bias + TCB_SH2));
      float* grow = pd.grow;
#pragma unroll
      for (int i = 0; i < 3; ++i) grow[i] = __fadd_rn(__fmul_rn(alb[i], sh), res[i]);
      grow[3] = pd.sigma;
#pragma unroll
      for (int i = 0; i < 3; ++i) grow[4 + i] = alb[i];
      grow[7] = sh;
#pragma unroll
      for (int i = 0; i < 3; ++i) grow[8 + i] = res[i];
    }
  }
  if (P.fuse) {
    __threadfence_block();
    warp_arrive(sy.addr(B_RAW_READY + pd.slot), lane);
  }
  sy.stamp(6);
}

__device__ __forceinline__ void epilogue_ts(const Params& P, Sync& sy, uint8_t* smem, uint32_t tmem, int q, int jj, int lane) {
  const int row = q * 32 + lane;
  const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
  float* s_sig = reinterpret_cast<float*>(smem + SM_SIG);
  const float* s_alpha = reinterpret_cast<const float*>(smem + SM_ALPHA);
  const int pair_bar = 2 + q;
  const int fine = P.ts_fine;
  PendingHeads pd;
  pd.on = false;
  for (int it = 0; it < P.n_iter; ++it) {
    const int64_t tile = tile_of(P, it);
    sy.tile = (int)tile;
    sy.set_it(it);
    sy.stamp(3);
    const int64_t m = tile * TILE_M + row;
    float sig = 0.f;
    uint32_t amax = 0u;
    const uint32_t A0 = (it & 1) ? 256u : 0u, A1 = 256u - A0;
    for (int l = 0; l < 8; ++l) {
      const uint32_t R = lane_addr + ((l & 1) ? A1 : A0);
      if (l == 7) epi_trunk_ts<1>(P, l, R, jj, lane, sy, s_alpha, &sig, amax, pair_bar);
      else epi_trunk_ts<0>(P, l, R, jj, lane, sy, nullptr, nullptr, amax, pair_bar);
      // layer 1 now keeps the tensor pipe busy for ~2 200 cycles: the previous tile's rows are finished here
      if (l == 0) finish_heads(P, sy, pd, jj, lane);
    }
    s_sig[row * 2 + jj] = sig;
    // relu(views') : A0[0,128) -> A0[0,64)
    sy.wait(B_ACC_FULL + 0);
    sy.stamp(13);                                     // views' complete
    tc_fence_after();
    epi_acc_ts<0>(fine, lane_addr + A0, lane_addr + A0, 2, jj, lane, sy, -1, nullptr, nullptr, amax, pair_bar);
    warp_arrive(sy.addr(B_V_READY), lane);
    // relu(albedo1) : A1[128,256) -> A1[128,192) ; relu(shading1) : A0[128,256) -> A0[128,192)
    sy.wait(B_ACC_FULL + 1);
    sy.stamp(14);                                     // albedo1 | shading1 complete
    tc_fence_after();
    epi_acc_ts<0>(fine, lane_addr + A1 + 128, lane_addr + A1 + 128, 2, jj, lane, sy, B_A_READY, nullptr, nullptr, amax, pair_bar);
    epi_acc_ts<0>(fine, lane_addr + A0 + 128, lane_addr + A0 + 128, 2, jj, lane, sy, B_A_READY + 2, nullptr, nullptr, amax, pair_bar);
    // heads: out of tensor memory into registers, then release the accumulator (and s_sig) for the next tile at once
    sy.wait(B_SMALL_FULL);
    sy.stamp(15);                                     // narrow heads complete
    tc_fence_after();
    __syncwarp();
    asm volatile("bar.sync 1, 256;" ::: "memory");    // both sigma partials of every row are in shared memory
    if (jj == 0) {
      uint32_t v[32];
      tmem_ld32(lane_addr + A0 + 64, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 3; ++i) { pd.h[i] = __uint_as_float(v[i]); pd.h[3 + i] = __uint_as_float(v[16 + i]); }
      pd.h[6] = __uint_as_float(v[19]);
      pd.sigma = (s_sig[row * 2] + s_sig[row * 2 + 1]) + __ldg(P.bias + TCB_ALPHA_B);
    }
    pd.valid = P.fuse ? true : (m < P.a.M);
    pd.grow = P.fuse ? P.f.ring + (((size_t)blockIdx.x * 2 + (it & 1)) * TILE_M + row) * P.out_ch
                     : P.a.raw + (m < P.a.M ? m : 0) * P.out_ch;
    pd.slot = it & 1;
    pd.on = true;
    if (m < P.a.M && !sy.dead && ((amax & 0xffffu) >= 0x7bffu || (amax >> 16) >= 0x7bffu)) {
      if (atomicCAS(P.dbg + 8, 0, 1) == 0) status_raise(P.status, DST_F16_ACT, 0, (int)tile, blockIdx.x);
    }
    // s_sig is next written after the NEXT tile's layer 7, which the issuer cannot reach before all eight arrivals below
    tc_fence_before();
    warp_arrive(sy.addr(B_TAIL_DONE), lane);
    sy.stamp(0);                                      // tile tail done
  }
  finish_heads(P, sy, pd, jj, lane);                  // the CTA's last tile
}

// ------------------------------------------------------------------------------------------
// kernel
// ------------------------------------------------------------------------------------------
template <int CL, bool STASH, bool TS = false, bool HY = false>
__global__ void __launch_bounds__(NUM_THREADS, 1) k_mlp_tc(const __grid_constant__ Params P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t smem_base = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SM_BAR + 8 * B_COUNT);
  const int rank = (CL > 1) ? (int)cluster_ctarank() : 0;

  Sync sy;
  sy.bar0 = smem_base + SM_BAR;
  sy.dbg = P.dbg;
  sy.status = P.status;
  sy.watchdog = P.watchdog;
  sy.prof = nullptr;
  sy.dead = false;
  sy.tile = -1;
  sy.phase = 0;
  if ((smem_base & 1023u) != 0) {                    // SWIZZLE_128B atoms need 1024 B alignment
    if (threadIdx.x == 0 && atomicCAS(P.dbg, 0, 2) == 0) status_raise(P.status, DST_SMEM_ALIGN, (int)smem_base);
    return;                                            // same for every CTA of the launch
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < NS_MAX; ++s) { mbar_init(sy.addr(B_WFULL + s), 1); mbar_init(sy.addr(B_WEMPTY + s), CL); }
    mbar_init(sy.addr(B_F_READY), 4); mbar_init(sy.addr(B_F_FREE), 1);
    for (int c = 0; c < 4; ++c) mbar_init(sy.addr(B_A_READY + c), 8);
    mbar_init(sy.addr(B_H_FREE), 1);
    mbar_init(sy.addr(B_ACC_FULL + 0), 1); mbar_init(sy.addr(B_ACC_FULL + 1), 1);
    mbar_init(sy.addr(B_ACC_FIRST + 0), 1); mbar_init(sy.addr(B_ACC_FIRST + 1), 1);
    mbar_init(sy.addr(B_V_READY), 8);
    mbar_init(sy.addr(B_SMALL_FULL), 1); mbar_init(sy.addr(B_SEM2_FULL), 1);
    mbar_init(sy.addr(B_TAIL_DONE), 8);
    for (int k = 0; k < 2; ++k) { mbar_init(sy.addr(B_RAW_READY + k), 8); mbar_init(sy.addr(B_RAW_FREE + k), 2); }
    for (int c = 0; c < 4; ++c) mbar_init(sy.addr(B_ST_DONE + c), 1);
    mbar_init(sy.addr(B_STV_DONE), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 12) {
    // constant operands: alpha_linear row (fp32) and the "ones" tile of the bias MMAs:
    // 2 core matrices of 8 rows x 8 halves; K columns 0..2 are 1.0, the rest 0
    float* s_alpha = reinterpret_cast<float*>(smem + SM_ALPHA);
    for (int i = lane; i < 256; i += 32) s_alpha[i] = __ldg(P.bias + TCB_ALPHA_W + i);
    __half* ones = reinterpret_cast<__half*>(smem + SM_ONES);
    for (int i = lane; i < 128; i += 32) {
      const int k = (i >> 6) * 8 + (i & 7);          // element i = core*64 + row*8 + kk
      ones[i] = __float2half_rn(k < 3 ? 1.f : 0.f);
    }
    fence_async_smem();
  }
  if (warp == 15) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();      // peers' barriers exist before any multicast / remote arrive
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  // "free"-type barriers start released: the first wait must pass on a fresh barrier
  const uint64_t released = (((1ull << NS_MAX) - 1) << B_WEMPTY) | (1ull << B_F_FREE) | (1ull << B_TAIL_DONE) | (3ull << B_RAW_FREE) |
                            (15ull << B_ST_DONE) | (1ull << B_STV_DONE);
  sy.phase = released;

  if (P.prof != nullptr && blockIdx.x == 0) {
    const int role = (threadIdx.x == 448) ? 0 : (threadIdx.x == 480 ? 1 : (threadIdx.x == 256 ? 2 : (threadIdx.x == 0 ? 3 : (threadIdx.x == 416 ? 4 : -1))));
    if (role >= 0) sy.prof = P.prof + role * 128;
  }
#ifdef INRF_TC_TIMELINE
  sy.tl = nullptr; sy.tl_n = 0; sy.it = -1;
  if (blockIdx.x == 0) {
    const int tr = (threadIdx.x == 480) ? 0 : (threadIdx.x == 0 ? 1 : (threadIdx.x == 128 ? 2 : (threadIdx.x == 384 ? 3 : -1)));
    if (tr >= 0 && P.n_iter > TL_IT1 + 2) sy.tl = g_tl + tr * TL_CAP;
  }
#endif
  const long long t_start = clock64();
  // warp roles.  The scheduler favours the highest warp id of each sub-partition, so the MMA issuer
  // (15) and the weight producer (14) sit on top of their sub-partitions; both run converged on all
  // 32 lanes (addresses and descriptors stay on the uniform datapath) and elect one lane to issue.
  // TS variant: there is no H buffer; PE | DIR | ring slide down to offset 0 (the roles address them relative to a base
  // shifted by -SM_PE), which gives the ring six 32 KB stages below the same SM_SIG.. tail of the map
  const uint32_t role_base = TS ? smem_base - (uint32_t)SM_PE : smem_base;
  if (warp == 14) {
    if (!P.no_weights) producer(P, sy, role_base, CL, rank);
  } else if (STASH && warp == 13) {
    stasher(P, sy, smem_base);
  } else if (!STASH && P.fuse && !(P.exp_flags & 8) && (warp == 12 || warp == 13)) {
    backend_role(P, sy, warp - 12, lane);
  } else if (warp == 15) {
    if (TS) issuer_ts(P, sy, role_base, tmem, CL, smem_base + SM_ONES); else issuer<HY>(P, sy, smem_base, tmem, CL, smem_base + SM_ONES);
  } else if (warp >= 8 && warp < 12) {
    front_end<STASH>(P, sy, role_base, (warp - 8) * 32 + lane);
  } else if (warp < 8) {
    if (TS) epilogue_ts(P, sy, smem, tmem, warp & 3, warp >> 2, lane);
    else epilogue<STASH, HY>(P, sy, smem, smem_base, tmem, warp & 3, warp >> 2, lane);
  }
  if (sy.prof != nullptr) sy.prof[63] = clock64() - t_start;
#ifdef INRF_TC_TIMELINE
  if (sy.tl != nullptr) g_tl_n[(int)((sy.tl - g_tl) / TL_CAP)] = sy.tl_n;
#endif
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();      // nobody leaves while a peer may still multicast into this CTA
  if (warp == 15) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

}  // namespace tc

int64_t mlp_tc_ring_bytes(int n_classes) {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return (int64_t)sms * 2 * tc::TILE_M * raw_channels(n_classes, 0) * (int64_t)sizeof(float);
}

int launch_mlp_tc(const MlpArgs& a, cudaStream_t st, const FuseArgs* fuse) {
  if (a.M == 0) return INRF_OK;
  if (fuse != nullptr) {
    if ((a.rays == nullptr && !a.cam_on) || a.S <= 0 || (a.S & 31) || a.endpoint || a.stash_img || fuse->rec == nullptr || fuse->ring == nullptr ||
        (a.z == nullptr && fuse->t_vals == nullptr)) {
      set_error("internal: fused tensor-core launch needs ray addressing, S %% 32 == 0, no endpoint feature, rec and ring");
      return INRF_EINVAL;
    }
    if (fuse->n_importance > 0 && (a.S != 64 || fuse->n_importance != 128 || fuse->u_det == nullptr || fuse->z_out == nullptr)) {
      set_error("internal: the in-kernel resampler implements 64 + 128 samples with deterministic u");
      return INRF_EINVAL;
    }
  }
  NetLayout L;
  int rc = make_layout(a.variant, a.n_classes, &L);
  if (rc) return rc;
  TcProgram prog;
  rc = make_tc_program(a.variant, a.n_classes, &prog);
  if (rc) return rc;
  tc::Params P;
  P.a = a;
  const unsigned char* blob = static_cast<const unsigned char*>(a.packed);
  P.blocks = blob + L.tc_blocks;
  P.bias = reinterpret_cast<const float*>(blob + L.tc_bias);
  P.n_fills = prog.n_fills;
  for (int i = 0; i < prog.n_fills; ++i) {
    P.fill_off[i] = prog.fill_off[i];
    P.fill_bytes[i] = prog.fill_bytes[i];
    if (prog.fill_bytes[i] > TC_SLOT_BYTES || (prog.fill_bytes[i] & 31)) { set_error("internal: ring fill %d has %d bytes", i, prog.fill_bytes[i]); return INRF_EINVAL; }
  }
  P.out_ch = raw_channels(a.n_classes, a.endpoint);
  P.C = a.n_classes;
  P.sem_rows = (a.n_classes + 15) / 16 * 16;
  static const int bias_mma_env = getenv("INRF_TC_BIASMMA") ? atoi(getenv("INRF_TC_BIASMMA")) : 1;
  P.bias_mma = bias_mma_env ? 1 : 0;
  static const bool prof_env = getenv("INRF_TC_PROF") != nullptr && getenv("INRF_TC_PROF")[0] == '1';
  static const bool now_env = getenv("INRF_TC_NOWEIGHTS") != nullptr && getenv("INRF_TC_NOWEIGHTS")[0] == '1';
  P.no_weights = now_env ? 1 : 0;
  static const int abl_env = getenv("INRF_TC_STASH_ABL") ? atoi(getenv("INRF_TC_STASH_ABL")) : 0;
  P.stash_abl = abl_env;
  static const int exp_env = getenv("INRF_TC_EXP") ? atoi(getenv("INRF_TC_EXP")) : 0;
  P.exp_flags = exp_env;
  P.prof = nullptr;
  if (prof_env) {
    long long* pp = nullptr;
    INRF_CUDA(cudaGetSymbolAddress((void**)&pp, tc::g_prof));
    INRF_CUDA(cudaMemsetAsync(pp, 0, 5 * 128 * sizeof(long long), st));
    P.prof = pp;
  }
  int* dbg = nullptr;
  INRF_CUDA(cudaGetSymbolAddress((void**)&dbg, tc::g_dbg));
  P.dbg = dbg;
  P.status = status_flag_dev();
  if (P.status == nullptr) return INRF_ECUDA;
  static const long long wd_env = getenv("INRF_TC_WATCHDOG_CYCLES") ? atoll(getenv("INRF_TC_WATCHDOG_CYCLES")) : 3000000000LL;
  static int faults_left = getenv("INRF_TC_FAULT") ? atoi(getenv("INRF_TC_FAULT")) : 0;   // test hook: fault the first n launches
  P.watchdog = wd_env > 0 ? wd_env : 3000000000LL;
  P.fault = faults_left > 0 ? 1 : 0;
  if (faults_left > 0) --faults_left;
  static const bool checked = getenv("INRF_TC_CHECK") != nullptr && getenv("INRF_TC_CHECK")[0] == '1';
  // the abort / claim words are per launch: a tripped watchdog must not poison later launches
  INRF_CUDA(cudaMemsetAsync(dbg, 0, 16 * sizeof(int), st));
  int dev = 0, sms = 148;
  INRF_CUDA(cudaGetDevice(&dev));
  INRF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  static const int cl_env = getenv("INRF_TC_CLUSTER") ? atoi(getenv("INRF_TC_CLUSTER")) : 2;
  const int64_t tiles = (a.M + tc::TILE_M - 1) / tc::TILE_M;
  int cl = (cl_env == 1) ? 1 : 2;     // default: 2-CTA clusters share every weight fetch (multicast)
  P.fuse = fuse != nullptr ? 1 : 0;
  if (fuse != nullptr) P.f = *fuse; else memset(&P.f, 0, sizeof(P.f));
  int grid;
  if (fuse == nullptr) {
    if (tiles < cl) cl = 1;
    grid = (int)(tiles < sms ? tiles : sms);
    grid = grid / cl * cl;                       // whole clusters only
    P.n_iter = (int)((tiles + grid - 1) / grid);
  } else {
    // contiguous runs of whole ray groups per CTA: lcm(S, 128) rows = unit_tiles tiles hold a whole number of rays, so
    // the back-end warp's running transmittance never crosses a CTA boundary
    int g = a.S, h = tc::TILE_M;
    while (h) { const int t = g % h; g = h; h = t; }
    const int unit_tiles = a.S / g;                                  // lcm(S,128)/128
    const int64_t units = (tiles + unit_tiles - 1) / unit_tiles;
    if (units < cl) cl = 1;
    grid = (int)(units < sms ? units : sms);
    grid = grid / cl * cl;
    P.n_iter = (int)((units + grid - 1) / grid) * unit_tiles;
  }
  // Activations in tensor memory (issuer_ts / epilogue_ts) wherever that variant is implemented: object network,
  // inference, bias by MMA.  INRF_TC_TS=0 selects the shared-memory-activation kernel everywhere (A/B, debugging),
  // =1 the two-chunks-per-load-batch drain.
  static const int ts_mode = getenv("INRF_TC_TS") ? atoi(getenv("INRF_TC_TS")) : 2;
  static const bool ts_env = ts_mode == 1 || ts_mode == 2;
  P.ts_fine = ts_mode == 2 ? 1 : 0;
  static const int ns_env = getenv("INRF_TC_NS") ? atoi(getenv("INRF_TC_NS")) : 0;
  const bool ts = ts_env && !a.stash_img && a.n_classes == 0 && !a.endpoint && a.variant == INRF_NET_OBJECT && P.bias_mma;
  // every other inference launch: TS trunk + shared-memory tail (INRF_TC_HY=0 keeps those on the SS kernel)
  static const bool hy_env = !(getenv("INRF_TC_HY") != nullptr && getenv("INRF_TC_HY")[0] == '0');
  const bool hy = ts_env && hy_env && !ts && !a.stash_img && P.bias_mma;
  P.ns = ts ? ((ns_env >= 2 && ns_env <= tc::NS_MAX) ? ns_env : tc::NS_MAX) : tc::NS;
  // split hand-off of the trunk layers (TS and HY kernels): INRF_TC_SPLIT = 0 (off) / 64 (default) / 128.  Only with the
  // chunk-per-batch drain (the default, INRF_TC_TS=2): the two-chunks-per-batch drain (INRF_TC_TS=1) is an A/B relic
  static const int split_env = getenv("INRF_TC_SPLIT") ? atoi(getenv("INRF_TC_SPLIT")) : 64;
  P.split_nf = ((ts || hy) && P.ts_fine && (split_env == 64 || split_env == 128)) ? split_env : 0;
  void (*kern)(tc::Params) = a.stash_img ? (cl == 2 ? tc::k_mlp_tc<2, true> : tc::k_mlp_tc<1, true>)
                             : ts ? (cl == 2 ? tc::k_mlp_tc<2, false, true> : tc::k_mlp_tc<1, false, true>)
                             : hy ? (cl == 2 ? tc::k_mlp_tc<2, false, false, true> : tc::k_mlp_tc<1, false, false, true>)
                                  : (cl == 2 ? tc::k_mlp_tc<2, false> : tc::k_mlp_tc<1, false>);
  INRF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SM_TOTAL));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(tc::NUM_THREADS);
  cfg.dynamicSmemBytes = tc::SM_TOTAL;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cl;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  INRF_CUDA(cudaLaunchKernelEx(&cfg, kern, P));
  note_launch();
  if (prof_env) {
    static const char* bar_names[tc::B_COUNT] = {"WFULL0","WFULL1","WFULL2","WFULL3","WFULL4","WFULL5","WEMPTY0","WEMPTY1","WEMPTY2","WEMPTY3",
      "WEMPTY4","WEMPTY5","F_READY","F_FREE","A_READY0","A_READY1","A_READY2","A_READY3","H_FREE","ACC_FULL0","ACC_FULL1","V_READY",
      "SMALL_FULL","SEM2_FULL","TAIL_DONE","RAW_READY0","RAW_READY1","RAW_FREE0","RAW_FREE1","ST_DONE0","ST_DONE1","ST_DONE2","ST_DONE3","STV_DONE","ACC_FIRST0","ACC_FIRST1"};
    static const char* roles[] = {"producer", "issuer", "frontend", "epilogue", "stasher"};
    long long h[5 * 128];
    INRF_CUDA(cudaStreamSynchronize(st));
    INRF_CUDA(cudaMemcpyFromSymbol(h, tc::g_prof, sizeof(h)));
    for (int r = 0; r < 5; ++r) {
      fprintf(stderr, "TCPROF role=%s total_cycles=%lld n_iter=%d\n", roles[r], h[r * 128 + 63], P.n_iter);
      for (int b = 0; b < tc::B_COUNT; ++b) {
        if (h[r * 128 + 64 + b] == 0) continue;
        fprintf(stderr, "TCPROF   %-10s waited %12lld cycles over %8lld slow waits\n", bar_names[b], h[r * 128 + b], h[r * 128 + 64 + b]);
      }
    }
  }
#ifdef INRF_TC_TIMELINE
  if (P.n_iter > tc::TL_IT1 + 2) {
    static long long h[4 * tc::TL_CAP];
    int hn[4];
    INRF_CUDA(cudaStreamSynchronize(st));
    INRF_CUDA(cudaMemcpyFromSymbol(h, tc::g_tl, sizeof(h)));
    INRF_CUDA(cudaMemcpyFromSymbol(hn, tc::g_tl_n, sizeof(hn)));
    long long t0 = -1;
    for (int r = 0; r < 4; ++r) for (int i = 0; i < hn[r]; ++i) { const long long t = h[r * tc::TL_CAP + i] >> 4; if (t0 < 0 || t < t0) t0 = t; }
    for (int r = 0; r < 4; ++r) {
      fprintf(stderr, "TCTL fuse=%d S=%d role=%d n=%d:", P.fuse, P.a.S, r, hn[r]);
      for (int i = 0; i < hn[r]; ++i) fprintf(stderr, " %lld:%d", (h[r * tc::TL_CAP + i] >> 4) - t0, (int)(h[r * tc::TL_CAP + i] & 15));
      fprintf(stderr, "\n");
    }
    const int zero[4] = {0, 0, 0, 0};
    INRF_CUDA(cudaMemcpyToSymbol(tc::g_tl_n, zero, sizeof(zero)));
  }
#endif
  if (checked) {      // debug mode (INRF_TC_CHECK=1): synchronise and report this launch's status record right away
    INRF_CUDA(cudaStreamSynchronize(st));
    return status_poll();
  }
  return INRF_OK;
}

}  // namespace inrf
