// Tensor-core evaluation of the intrinsic field network on sm_100a (INRF_PREC_TC), CTA-pair version.
//
// Two CTAs of one TPC form a pair (2-CTA cluster, tcgen05 cta_group::2): every MMA is M=256 (128 sample
// rows per CTA) x N<=256 and each CTA keeps only HALF of every weight tile in shared memory (N/2 rows),
// which leaves room for TWO 128-row tiles per CTA ("streams" 0 and 1).  The streams are half a tile
// apart: while the epilogue warps turn the accumulator of stream s into the next layer's A operand, the
// tensor pipe runs a layer of stream 1-s, so the accumulator drain / epilogue / hand-off latency that
// idled the pipe with one tile in flight is covered by useful MMAs.
//
// The whole network (NeRF.forward, object_level/run_nerf_helpers.py:284-325 / Semantic_NeRF.forward,
// SSR/models/semantic_nerf.py:123-181, fused with the Embedder and run_network's per-sample
// view-direction expansion) runs per tile without touching HBM in between:
//
//   warp 20     weight producer : streams this CTA's half of the pre-swizzled fp16 operand tiles (pack.cu)
//                                 into a 4 x 16 KB ring with cp.async.bulk + mbarrier complete_tx
//   warp 21     leader CTA: MMA issuer (converged warp, one elected lane issues tcgen05.mma
//                                 cta_group::2 kind::f16, M=256, K=16, fp32 accumulators in TMEM, 256
//                                 columns per stream; bias = one extra K=16 MMA of a constant "ones"
//                                 tile against (hi, lo, lo2) fp16 bias columns)
//               peer CTA:   relay (forwards "my half of ring slot i has landed" to the leader's barrier)
//   warps 16-19 front end       : o + d z, range-reduced sin/cos encodings as fp16 UMMA operand tiles;
//                                 gamma(x) and gamma(d) share one 16 KB tile per stream (gamma(d) is
//                                 written once layer 5 has consumed gamma(x))
//   warps 0-15  epilogue        : all 16 on whichever stream completes; tcgen05.ld -> ReLU -> fp16 -> next layer's A operand in
//                                 place (SWIZZLE_128B K-major); the narrow heads (sigma, albedo2,
//                                 shading2, residual) are fp32 dot products on the un-rounded
//                                 accumulators, so no hidden activation of the tail ever goes back to
//                                 shared memory; sigmoids; packed raw rows to HBM
//
// Arithmetic: operands rounded to fp16 (RN, 11-bit significand), products and sums in fp32.
// feature_linear has no activation, so views_linears.0 o feature_linear is composed into one
// 128x256 matrix at pack time (pack.cu) - one 256x256 GEMM per sample less than the literal graph.
#include "common.cuh"
#include <stdlib.h>

namespace inrf {
namespace tc2 {

constexpr int TILE_M = 128;
constexpr int N_EPI_WARPS = 16;             // 8 per stream
constexpr int W_FE0 = 16, N_FE_WARPS = 4, W_PROD = 20, W_MMA = 21;
constexpr int NUM_THREADS = 22 * 32;
constexpr int NS = 4;                       // weight ring stages
constexpr int SLOT = TC_SLOT_BYTES / 2;     // this CTA's half of a 256-row x 64-K operand tile
constexpr int CHUNK = 16384;                // 128 rows x 64 fp16, SWIZZLE_128B
constexpr int LAG = 5;                      // stream 1 runs LAG steps behind stream 0
// shared memory map (bytes)
constexpr int SM_STREAM = 5 * CHUNK;        // per stream: H (4 chunks, in place) + X (gamma(x) / gamma(d))
constexpr int SM_X = 4 * CHUNK;             // offset of X inside a stream
constexpr int SM_SCRATCH = 3 * CHUNK;       // [3][128][8] fp32 head partials, H chunk 3 at the tile tail
constexpr int SM_RING = 2 * SM_STREAM;
constexpr int SM_ONES = SM_RING + NS * SLOT;           // 8 x 16 fp16 "ones" A operand for the bias MMAs
constexpr int SM_BAR = SM_ONES + 256;
constexpr int SM_TOTAL = SM_BAR + 256;
static_assert(SM_TOTAL <= 232448, "shared memory budget");

// bias table offsets (floats), written by pack.cu:k_pack_tc_bias
constexpr int TCB_VIEWS = 2048, TCB_SEM1 = 2176, TCB_ALBSH = 2304, TCB_ALPHA_W = 2560, TCB_ALPHA_B = 2816,
              TCB_ALB2 = 2817, TCB_SH2 = 2820, TCB_RES = 2821, TCB_SEM2 = 2824, TCB_HW_RES = 4096, TCB_HW_AS = 4608;

// barrier ids
enum {
  B_WFULL = 0,                 // [NS] ring slot filled: local tx bytes (+ the peer's relay arrival in the leader)
  B_WEMPTY = B_WFULL + NS,     // [NS] ring slot consumed (tcgen05.commit multicast to both CTAs)
  B_STREAM = B_WEMPTY + NS,    // per stream: + 4 * s
  B_ACC_FULL = 0,              //   accumulator of the stream complete (commit multicast)
  B_EPI_DONE = 1,              //   leader only: both CTAs' epilogue warps are done with the accumulator (32 arrivals)
  B_XREADY = 2,                //   leader only: gamma(x) / gamma(d) written in both CTAs (8 arrivals)
  B_XFREE = 3,                 //   X tile no longer read by the tensor core (commit multicast)
  B_COUNT = B_STREAM + 8
};

// steps of one tile in issue order
enum { K_ALBSH = 8, K_VIEWS = 9, K_SEM2 = 10 };
constexpr int MAX_FILLS = 64;

struct Params {
  MlpArgs a;
  const unsigned char* blocks;    // fp16 operand blob
  const float* bias;              // fp32 table
  int n_steps;                    // 10, or 11 with the semantic head
  int step_fill0[12];             // first ring fill of each step / number of fills
  int step_nfill[12];
  int fill_src[2][MAX_FILLS];     // per pair rank: blob offset of this CTA's half
  int fill_len[MAX_FILLS];
  int fill_src2[2][MAX_FILLS];    // optional second piece (two-block fills), length 0 = none
  int fill_len2[MAX_FILLS];
  int out_ch, C, sem_rows;
  int n_iter;                     // tiles per stream per CTA (identical everywhere: the pair runs in lock-step)
  int* dbg;                       // [16] watchdog record (device)
  int no_weights;                 // timing experiment (INRF_TC_NOWEIGHTS=1): no weight streaming, results are garbage
  int exp;                        // timing experiments (INRF_TC_EXP): 1 = epilogue does not touch TMEM/smem, 2 = no MMA issue
};

__device__ int g_dbg[16];
#ifdef TC2_PROF
// profiling build (-DTC2_PROF): per role (one lane of a few warps of the first pair) cycles spent in
// blocking barrier waits [role*64 + id], number of such waits [role*64 + 32 + id], busy cycles [role*64 + 30],
// total [role*64 + 31]
__device__ long long g_prof[8 * 64];
__device__ __forceinline__ int prof_role() {
  if (blockIdx.x > 1 || (threadIdx.x & 31) != 0) return -1;
  const int w = threadIdx.x >> 5;
  if (blockIdx.x == 0) return w == 21 ? 0 : (w == 20 ? 1 : (w == 16 ? 2 : (w == 0 ? 3 : (w == 8 ? 4 : -1))));
  return w == 21 ? 5 : (w == 20 ? 6 : (w == 0 ? 7 : -1));
}
#endif

// ------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t addr, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(addr), "r"(count));
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
__device__ __forceinline__ uint32_t mbar_test(uint32_t addr, uint32_t parity) {   // non-blocking probe
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
  return ok;
}
// address of the same shared-memory location in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
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
// completion of every MMA issued so far -> the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void tc_commit_pair(uint32_t mbar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(mbar), "h"((uint16_t)3) : "memory");
}
// ... -> the leader's barrier only
__device__ __forceinline__ void tc_commit_leader(uint32_t mbar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(mbar), "h"((uint16_t)1) : "memory");
}
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
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
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
               "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                 "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
               : "r"(taddr) : "memory");
}
// wait for the outstanding TMEM loads; the registers are in/out operands so that no use of them is
// scheduled above the wait
__device__ __forceinline__ void tmem_wait16(uint32_t* v) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                 "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15])
               :: "memory");
}
// Walk NBLK 16-column blocks of this warp's accumulator slice (block b = columns (b>>1)*64 + (b&1)*16 ..)
// with two loads in flight: the TMEM read port (64 B/clk per SM - the epilogue's real bound) keeps
// streaming while the previous block is converted.
template <int NBLK, class F>
__device__ __forceinline__ void tmem_stream16(uint32_t taddr, F&& f) {
  static_assert(NBLK % 2 == 0, "pairs of blocks");
  uint32_t va[16], vb[16];
  tmem_ld16(taddr, va);
#pragma unroll
  for (int b = 0; b < NBLK; b += 2) {
    tmem_wait16(va);
    tmem_ld16(taddr + (uint32_t)(((b + 1) >> 1) * 64 + ((b + 1) & 1) * 16), vb);
    f(b, va);
    tmem_wait16(vb);
    if (b + 2 < NBLK) tmem_ld16(taddr + (uint32_t)(((b + 2) >> 1) * 64 + ((b + 2) & 1) * 16), va);
    f(b + 1, vb);
  }
}

// UMMA shared-memory descriptor, K-major, SWIZZLE_128B (cute/arch/mma_sm100_desc.hpp):
//  [0,14) start>>4 | [16,30) LBO>>4 (unused for swizzled K-major, 1) | [32,46) SBO>>4 = 1024>>4 |
//  [46,48) version = 1 | [61,64) layout = 2 (SWIZZLE_128B).  With cta_group::2 the same descriptor
//  addresses each CTA's own shared memory: its 128 rows of A, its N/2 rows of B.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// no-swizzle K-major core-matrix layout (layout type 0): 8 rows x 16 B contiguous, the two K halves
// LBO bytes apart, 8-row groups SBO bytes apart.  SBO = 0 replays one 8-row group for all rows.
__device__ __forceinline__ uint64_t make_desc_flat(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
// instruction descriptor: D=f32 (bit 4), A=B=f16 (0), K-major both, N>>3 at [17,23), M>>4 at [24,29); M = 256 (pair)
__device__ __forceinline__ constexpr uint32_t make_idesc(int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)((2 * TILE_M) >> 4) << 24);
}

// ------------------------------------------------------------------------------------------
// barrier bookkeeping with a watchdog: a stuck wait records who/where, raises a global abort
// flag and lets every role run to the end (garbage out, but no hung GPU and no lost context)
// ------------------------------------------------------------------------------------------
__device__ __noinline__ bool slow_wait_impl(uint32_t bar_addr, uint32_t parity, int id, int tile, int* dbg) {
  const long long t0 = clock64();
  uint32_t spins = 0;
  bool dead = false;
  while (!mbar_try(bar_addr, parity)) {
    if ((++spins & 0x3ff) == 0) {
      if (*(volatile int*)dbg != 0) { dead = true; break; }
      if (clock64() - t0 > 3000000000LL) {
        if (atomicCAS(dbg, 0, 1) == 0) {
          dbg[1] = id; dbg[2] = threadIdx.x >> 5; dbg[3] = tile; dbg[4] = blockIdx.x; dbg[5] = (int)parity;
          __threadfence();
        }
        dead = true;
        break;
      }
    }
  }
  return dead;
}

struct Sync {             // lives in registers (never escapes by address)
  uint32_t bar0;          // smem address of barrier 0
  uint32_t phase;         // one parity bit per barrier id
  int* dbg;
  bool dead;
  int tile;
  __device__ __forceinline__ uint32_t addr(int id) const { return bar0 + 8u * id; }
  __device__ __forceinline__ uint32_t take_parity(int id) {       // consume the next phase of barrier id
    const uint32_t parity = (phase >> id) & 1u;
    phase ^= (1u << id);
    return parity;
  }
#ifdef TC2_PROF
  __device__ __forceinline__ void record(int id, long long t0) {
    const int role = prof_role();
    if (role >= 0) { g_prof[role * 64 + id] += clock64() - t0; g_prof[role * 64 + 32 + id] += 1; }
  }
#endif
  __device__ __forceinline__ void slow(int id, uint32_t parity) {
#ifdef TC2_PROF
    const long long t0 = clock64();
#endif
    if (!dead) dead = slow_wait_impl(addr(id), parity, id, tile, dbg);
#ifdef TC2_PROF
    record(id, t0);
#endif
  }
  __device__ __forceinline__ void wait(int id) {
    const uint32_t parity = take_parity(id);
    if (dead) return;
#ifdef TC2_PROF
    const long long t0 = clock64();
    if (!mbar_try(addr(id), parity)) dead = slow_wait_impl(addr(id), parity, id, tile, dbg);
    record(id, t0);
#else
    if (mbar_try(addr(id), parity)) return;
    slow(id, parity);
#endif
  }
};
static_assert(B_COUNT <= 32, "Sync::phase is 32 bits");

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
// One arrival per warp on a barrier of the LEADER CTA: every lane has fenced its own shared-memory
// writes towards the async proxy (they are consumed by this CTA's own tensor core), the warp converges
// and lane 0 arrives remotely (same default semantics as cutlass::arch::ClusterBarrier::arrive(cta_id)).
__device__ __forceinline__ void warp_arrive_leader(uint32_t leader_bar, int lane) {
  __syncwarp();
  if (lane == 0) mbar_arrive_cluster(leader_bar);
}

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
  __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t relu_h2(uint32_t h2) {
  __half2 v = *reinterpret_cast<__half2*>(&h2);
  v = __hmax2(v, __float2half2_rn(0.f));
  return *reinterpret_cast<uint32_t*>(&v);
}

// swizzled byte offsets of the 8 16-byte units of this thread's row inside a chunk
// (K-major SWIZZLE_128B: 8-row atoms of 1024 B, 16-byte unit index XOR (row & 7))
struct RowAddr {
  uint32_t base, x;       // row offset inside the chunk, (row & 7) << 4
  __device__ __forceinline__ void init(int row) {
    base = (uint32_t)((row >> 3) * 1024 + (row & 7) * 128);
    x = (uint32_t)((row & 7) << 4);
  }
  __device__ __forceinline__ uint32_t unit(int u) const { return base + (((uint32_t)u << 4) ^ x); }
};

// the pair's static schedule: stream 0 runs step j % n_steps of its tile j / n_steps in slot j,
// stream 1 the same LAG slots later.  Producer, relay and issuer walk it identically.
template <class F>
__device__ __forceinline__ void for_each_step(int n_iter, int n_steps, F f) {
  const int total = n_iter * n_steps;
  int k0 = 0, k1 = 0;
  for (int j = 0; j < total + LAG; ++j) {
    if (j < total) { f(0, k0); k0 = (k0 + 1 == n_steps) ? 0 : k0 + 1; }
    if (j >= LAG) { f(1, k1); k1 = (k1 + 1 == n_steps) ? 0 : k1 + 1; }
  }
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

struct FrontEnd {
  const Params& P;
  Sync& sy;
  uint32_t smem_base;
  int row, lane;
  RowAddr ra;
  uint32_t xready[2];     // leader-CTA addresses of the two streams' XREADY barriers

  __device__ __forceinline__ int64_t sample_row(int s, int it) const {
    const int64_t tile = ((int64_t)it * 2 + s) * gridDim.x + blockIdx.x;   // may run past the end: rows clamp
    int64_t m = tile * TILE_M + row;
    return m < P.a.M ? m : P.a.M - 1;
  }
  // gamma(x) of one sample row -> X tile of stream s (64 halves: 63 + zero)
  __device__ __forceinline__ void points(int s, int it) {
    const int64_t m = sample_row(s, it);
    uint32_t pw[32];
    {
      float pe[64];
      if (P.a.emb != nullptr) {
        const float* e = P.a.emb + m * (PE_PTS + PE_DIR);
#pragma unroll
        for (int i = 0; i < 63; ++i) pe[i] = __ldg(e + i);
      } else {
        float x[3];
        if (P.a.rays != nullptr) {
          const float* ray = P.a.rays + (m / P.a.S) * 11;
          const float zv = __ldg(P.a.z + m);
#pragma unroll
          for (int i = 0; i < 3; ++i) x[i] = __fadd_rn(__ldg(ray + i), __fmul_rn(__ldg(ray + 3 + i), zv));   // o + d z (run_nerf.py:488)
        } else {
#pragma unroll
          for (int i = 0; i < 3; ++i) x[i] = __ldg(P.a.pts + m * 3 + i);
        }
        if (P.a.pe_scale != 1.f) {
#pragma unroll
          for (int i = 0; i < 3; ++i) x[i] = __fdiv_rn(x[i], P.a.pe_scale);
        }
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          float sn[10], cs[10];
          pe_coord(x[i], 10, sn, cs);
          pe[i] = x[i];
#pragma unroll
          for (int k = 0; k < 10; ++k) { pe[3 + 6 * k + i] = sn[k]; pe[3 + 6 * k + 3 + i] = cs[k]; }
        }
      }
      pe[63] = 0.f;
#pragma unroll
      for (int i = 0; i < 32; ++i) pw[i] = pack_h2(pe[2 * i], pe[2 * i + 1]);
    }
    sy.tile = it;
    sy.wait(B_STREAM + 4 * s + B_XFREE);
    const uint32_t X = smem_base + s * SM_STREAM + SM_X;
#pragma unroll
    for (int u = 0; u < 8; ++u) st_shared_v4(X + ra.unit(u), pw[4 * u], pw[4 * u + 1], pw[4 * u + 2], pw[4 * u + 3]);
    fence_async_smem();
    warp_arrive_leader(xready[s], lane);
  }
  // gamma(d) -> the first 32 K columns of the same tile (layer 5 has consumed gamma(x))
  __device__ __forceinline__ void dirs(int s, int it) {
    const int64_t m = sample_row(s, it);
    uint32_t dw[16];
    {
      float de[32];
      if (P.a.emb != nullptr) {
        const float* e = P.a.emb + m * (PE_PTS + PE_DIR) + PE_PTS;
#pragma unroll
        for (int i = 0; i < 27; ++i) de[i] = __ldg(e + i);
      } else {
        float d[3];
        if (P.a.rays != nullptr) {
          const float* ray = P.a.rays + (m / P.a.S) * 11;
#pragma unroll
          for (int i = 0; i < 3; ++i) d[i] = __ldg(ray + 8 + i);
        } else {
#pragma unroll
          for (int i = 0; i < 3; ++i) d[i] = __ldg(P.a.viewdirs + m * 3 + i);
        }
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          float sn[10], cs[10];
          pe_coord(d[i], 4, sn, cs);
          de[i] = d[i];
#pragma unroll
          for (int k = 0; k < 4; ++k) { de[3 + 6 * k + i] = sn[k]; de[3 + 6 * k + 3 + i] = cs[k]; }
        }
      }
#pragma unroll
      for (int i = 27; i < 32; ++i) de[i] = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) dw[i] = pack_h2(de[2 * i], de[2 * i + 1]);
    }
    sy.tile = it;
    sy.wait(B_STREAM + 4 * s + B_XFREE);
    const uint32_t X = smem_base + s * SM_STREAM + SM_X;
#pragma unroll
    for (int u = 0; u < 4; ++u) st_shared_v4(X + ra.unit(u), dw[4 * u], dw[4 * u + 1], dw[4 * u + 2], dw[4 * u + 3]);
    fence_async_smem();
    warp_arrive_leader(xready[s], lane);
  }
  // Event order = the order in which the issuer's static schedule frees the X tiles (LAG = 5, 10 or 11
  // steps): gamma(x) of stream 0 tile k, gamma(d) of stream 1 tile k-1, gamma(x) of stream 1 tile k,
  // gamma(d) of stream 0 tile k.
  __device__ __forceinline__ void run() {
    for (int k = 0; k <= P.n_iter; ++k) {
      if (k < P.n_iter) points(0, k);
      if (k >= 1) dirs(1, k - 1);
      if (k < P.n_iter) { points(1, k); dirs(0, k); }
    }
  }
};

// ------------------------------------------------------------------------------------------
// weight producer (converged warp, elected lane issues the bulk copies of this CTA's halves)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void producer(const Params& P, Sync& sy, uint32_t smem_base, int rank) {
  int slot = 0;
  const bool leader_lane = elect_one();
  for_each_step(P.n_iter, P.n_steps, [&](int, int k) {
    const int f0 = P.step_fill0[k], nf = P.step_nfill[k];
    for (int f = f0; f < f0 + nf; ++f) {
      sy.wait(B_WEMPTY + slot);
      if (leader_lane && !sy.dead) {
        const uint32_t len = (uint32_t)P.fill_len[f], len2 = (uint32_t)P.fill_len2[f];
        const uint32_t dst = smem_base + SM_RING + slot * SLOT;
        const uint32_t bar = sy.addr(B_WFULL + slot);
        mbar_expect_tx(bar, len + len2);
        bulk_g2s(dst, P.blocks + P.fill_src[rank][f], len, bar);
        if (len2) bulk_g2s(dst + len, P.blocks + P.fill_src2[rank][f], len2, bar);
      }
      __syncwarp();
      slot = (slot + 1 == NS) ? 0 : slot + 1;
    }
  });
}

// peer CTA: tell the leader that this CTA's half of each ring slot has landed
__device__ __forceinline__ void relay(const Params& P, Sync& sy) {
  int slot = 0;
  const bool leader_lane = elect_one();
  const uint32_t remote0 = mapa(sy.addr(B_WFULL), 0);      // barriers are 8 bytes apart in every CTA's window
  for_each_step(P.n_iter, P.n_steps, [&](int, int k) {
    const int nf = P.step_nfill[k];
    for (int f = 0; f < nf; ++f) {
      sy.wait(B_WFULL + slot);
      if (leader_lane) mbar_arrive_cluster(remote0 + 8u * slot);
      __syncwarp();
      slot = (slot + 1 == NS) ? 0 : slot + 1;
    }
  });
}

// ------------------------------------------------------------------------------------------
// MMA issuer (leader CTA): converged warp (descriptor arithmetic stays on the uniform datapath), one
// elected lane issues tcgen05.mma / tcgen05.commit
// ------------------------------------------------------------------------------------------
#ifdef TC2_PROF
#define ISSUE_T0() const long long ti0 = clock64()
#define ISSUE_T1() t_issue += clock64() - ti0
#else
#define ISSUE_T0()
#define ISSUE_T1()
#endif
struct Issuer {
  Sync& sy;
  uint32_t smem_base, tmem;
  int slot;
  bool leader;   // elected lane (cleared by the "no MMA" timing experiment)
  bool no_weights, no_mma;
  // probe-ahead: the "slot filled" barrier of the NEXT fill is tested (non-blocking) right after the MMAs
  // of the current fill were issued, so the mbarrier round trip overlaps tensor execution
  uint32_t pw_ok, pw_par;
#ifdef TC2_PROF
  long long t_issue;
#endif
  __device__ __forceinline__ uint32_t slot_addr() const { return smem_base + SM_RING + slot * SLOT; }
  __device__ __forceinline__ void commit_pair(int bar) {
    ISSUE_T0();
    if (leader) tc_commit_pair(sy.addr(bar));
    __syncwarp();
    ISSUE_T1();
  }
  __device__ __forceinline__ void probe() {
    pw_par = sy.take_parity(B_WFULL + slot);
    pw_ok = no_weights ? 1u : mbar_test(sy.addr(B_WFULL + slot), pw_par);
  }
  __device__ __forceinline__ void acquire() {
    if (!pw_ok) sy.slow(B_WFULL + slot, pw_par);
    tc_fence_after();
  }
  __device__ __forceinline__ void release() {        // MMAs reading the slot are done -> both producers refill
    ISSUE_T0();
    if (leader && !no_weights) tc_commit_pair(sy.addr(B_WEMPTY + slot));
    __syncwarp();
    ISSUE_T1();
    slot = (slot + 1 == NS) ? 0 : slot + 1;
    probe();
  }
  // K = 16*KSTEPS of A chunk `a_chunk` times the operand tile at `b_addr`
  template <int KSTEPS>
  __device__ __forceinline__ void mma(uint32_t a_chunk, uint32_t b_addr, int n, uint32_t col, bool first) {
    const uint64_t ad = make_desc(a_chunk);
    const uint64_t bd = make_desc(b_addr);
    const uint32_t id = make_idesc(n);
    ISSUE_T0();
    if (leader && !no_mma) {
#pragma unroll
      for (int k = 0; k < KSTEPS; ++k)
        tc_mma(tmem + col, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), id, (first && k == 0) ? 0u : 1u);
    }
    __syncwarp();
    ISSUE_T1();
  }
  template <int KSTEPS>
  __device__ __forceinline__ void fill_mma(uint32_t a_chunk, int n, uint32_t col) {
    acquire();
    mma<KSTEPS>(a_chunk, slot_addr(), n, col, false);
    release();
  }
  // accumulator columns [col, col+n) := bias (one K=16 MMA of the constant "ones" tile)
  __device__ __forceinline__ void bias(int n, uint32_t col) {
    acquire();
    ISSUE_T0();
    if (leader && !no_mma)
      tc_mma(tmem + col, make_desc_flat(smem_base + SM_ONES, 128, 0), make_desc_flat(slot_addr(), 128, 256), make_idesc(n), 0u);
    __syncwarp();
    ISSUE_T1();
    release();
  }
};

__device__ __forceinline__ void issuer(const Params& P, Sync& sy, uint32_t smem_base, uint32_t tmem, uint8_t* smem_u8) {
#ifdef TC2_PROF
  long long lat_seen = 0;
  if (threadIdx.x % 32 == 0) for (int i = 0; i < 4; ++i) reinterpret_cast<volatile long long*>(smem_u8 + SM_BAR + 192)[i] = clock64();
#endif
#ifdef TC2_PROF
  Issuer I{sy, smem_base, tmem, 0, elect_one(), P.no_weights != 0, (P.exp & 2) != 0, 1u, 0u, 0ll};
#else
  Issuer I{sy, smem_base, tmem, 0, elect_one(), P.no_weights != 0, (P.exp & 2) != 0, 1u, 0u};
#endif
  const bool sem = P.C > 0;
  const int nv = sem ? 256 : 128;       // views' [| sem1] width
  I.probe();
  for_each_step(P.n_iter, P.n_steps, [&](int s, int k) {
    const uint32_t H = smem_base + s * SM_STREAM, X = H + SM_X;
    const uint32_t acc = (uint32_t)s * 256;
    const int sb = B_STREAM + 4 * s;
    sy.wait(sb + B_EPI_DONE);           // accumulator drained and the A operand (if any) written, both CTAs
#ifdef TC2_PROF
    { volatile long long* ts = reinterpret_cast<volatile long long*>(smem_u8 + SM_BAR + 192) + 2 * s; lat_seen += clock64() - ts[1]; }
#endif
    if (k == 0) {                       // trunk layer 0: K = 64 (gamma(x))
      sy.wait(sb + B_XREADY);
      I.bias(256, acc);
      I.fill_mma<4>(X, 256, acc);
    } else if (k < 8) {                 // trunk layers 1..7
      I.bias(256, acc);
      if (k == 5) {                     // skip connection: [gamma(x), h] -> K = 64 + 256
        I.fill_mma<4>(X, 256, acc);
        I.commit_pair(sb + B_XFREE);    // gamma(d) may replace gamma(x)
      }
      for (int c = 0; c < 4; ++c) I.fill_mma<4>(H + c * CHUNK, 256, acc);
    } else if (k == K_ALBSH) {          // albedo1 | shading1 on the trunk output
      I.bias(256, acc);
      for (int c = 0; c < 4; ++c) I.fill_mma<4>(H + c * CHUNK, 256, acc);
    } else if (k == K_VIEWS) {          // views' [| sem1] on the trunk output + gamma(d) for the views' rows
      sy.wait(sb + B_XREADY);
      I.bias(nv, acc);
      for (int c = 0; c < 4; ++c) I.fill_mma<4>(H + c * CHUNK, nv, acc);
      I.fill_mma<2>(X, 128, acc);
      I.commit_pair(sb + B_XFREE);      // next tile's gamma(x) may land
    } else {                            // semantic logits on relu(sem1) (H chunks 0,1): ceil16(C) x 128
      I.acquire();
      const uint32_t b_addr = I.slot_addr();
      I.mma<4>(H, b_addr, P.sem_rows, acc, true);
      I.mma<4>(H + CHUNK, b_addr + (uint32_t)(P.sem_rows * 64), P.sem_rows, acc, false);
      I.release();
    }
    I.commit_pair(sb + B_ACC_FULL);
#ifdef TC2_PROF
    { volatile long long* ts = reinterpret_cast<volatile long long*>(smem_u8 + SM_BAR + 192) + 2 * s; ts[0] = clock64(); }
#endif
  });
#ifdef TC2_PROF
  { const int role = prof_role(); if (role >= 0) { g_prof[role * 64 + 30] = I.t_issue; g_prof[role * 64 + 29] = lat_seen; } }
#endif
}

// ------------------------------------------------------------------------------------------
// epilogue warps.  All 16 warps serve whichever stream's accumulator completes next (the streams
// alternate, so the drain of one accumulator gets twice the warps and finishes in half the time).
// Warp (q, cq): TMEM lanes 32q..32q+31 (one row per thread); of every 64-column chunk it owns columns
// cq*16..cq*16+15.
// ------------------------------------------------------------------------------------------
// 16 accumulator columns -> ReLU -> fp16 -> 2 swizzled 16-byte stores
__device__ __forceinline__ void store16(const uint32_t* v, uint32_t dst_chunk, const RowAddr& ra, int unit0) {
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    uint32_t pk[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
      pk[i] = relu_h2(pack_h2(__uint_as_float(v[8 * u + 2 * i]), __uint_as_float(v[8 * u + 2 * i + 1])));
    st_shared_v4(dst_chunk + ra.unit(unit0 + u), pk[0], pk[1], pk[2], pk[3]);
  }
}

__device__ __forceinline__ float sigmoid_(float x) { return 1.f / (1.f + expf(-x)); }

// NBLK 16-column blocks (block b = this warp's columns of chunk b) with two TMEM loads in flight
template <int NBLK, class F>
__device__ __forceinline__ void tmem_chunks16(uint32_t taddr, F&& f) {
  static_assert(NBLK % 2 == 0, "pairs of blocks");
  uint32_t va[16], vb[16];
  tmem_ld16(taddr, va);
#pragma unroll
  for (int b = 0; b < NBLK; b += 2) {
    tmem_wait16(va);
    tmem_ld16(taddr + (uint32_t)((b + 1) * 64), vb);
    f(b, va);
    tmem_wait16(vb);
    if (b + 2 < NBLK) tmem_ld16(taddr + (uint32_t)((b + 2) * 64), va);
    f(b + 1, vb);
  }
}

struct HeadAcc {          // per-stream partial dot products of the narrow heads (this thread's columns)
  float sig, alb[3], sh, res[3];
};

__device__ __forceinline__ void epilogue(const Params& P, Sync& sy, uint8_t* smem, uint32_t smem_base, uint32_t tmem,
                                         int q, int cq, int lane) {
  const int row = q * 32 + lane;
  const uint32_t lane_col = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)cq * 16;
  const bool sem = P.C > 0;
  const float4* hw_res = reinterpret_cast<const float4*>(P.bias + TCB_HW_RES);
  const float4* hw_as = reinterpret_cast<const float4*>(P.bias + TCB_HW_AS);
  const uint32_t done0 = mapa(sy.addr(B_STREAM + B_EPI_DONE), 0);     // stream 1: + 32 bytes
  RowAddr ra;
  ra.init(row);
  HeadAcc hs[2];
  int it_of[2] = {0, 0};
#ifdef TC2_PROF
  long long busy = 0, tb = 0, lat_wake = 0;
  const bool prof_me = prof_role() >= 0 && blockIdx.x == 0;
#define PROF_BEGIN() do { tb = clock64(); if (prof_me) lat_wake += tb - (reinterpret_cast<volatile long long*>(smem + SM_BAR + 192) + 2 * s)[0]; } while (0)
#define PROF_END() do { const long long te = clock64(); busy += te - tb; if (prof_me) (reinterpret_cast<volatile long long*>(smem + SM_BAR + 192) + 2 * s)[1] = te; } while (0)
#else
#define PROF_BEGIN()
#define PROF_END()
#endif
  for_each_step(P.n_iter, P.n_steps, [&](int s, int k) {
    const uint32_t acc = lane_col + (uint32_t)s * 256;
    const uint32_t H = smem_base + s * SM_STREAM;
    const int sb = B_STREAM + 4 * s;
    const uint32_t done = done0 + 32u * s;
    HeadAcc& h = hs[s];
    sy.tile = it_of[s];
    sy.wait(sb + B_ACC_FULL);
    PROF_BEGIN();
    tc_fence_after();
    if (k < 8) {
      // ---- trunk: accumulator of layer k -> A operand of layer k+1 (in place in H) --------------------
      if (P.exp & 1) {
      } else if (k < 7) {
        tmem_chunks16<4>(acc, [&](int b, const uint32_t* v) { store16(v, H + b * CHUNK, ra, cq * 2); });
      } else {                          // + sigma head: fp32 dot on the un-rounded ReLU output (alpha_linear)
        const float4* aw = reinterpret_cast<const float4*>(P.bias + TCB_ALPHA_W + cq * 16);
        float sg = 0.f;
        tmem_chunks16<4>(acc, [&](int b, const uint32_t* v) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 t = __ldg(aw + b * 16 + i);
            sg = fmaf(fmaxf(__uint_as_float(v[4 * i]), 0.f), t.x, sg);
            sg = fmaf(fmaxf(__uint_as_float(v[4 * i + 1]), 0.f), t.y, sg);
            sg = fmaf(fmaxf(__uint_as_float(v[4 * i + 2]), 0.f), t.z, sg);
            sg = fmaf(fmaxf(__uint_as_float(v[4 * i + 3]), 0.f), t.w, sg);
          }
          store16(v, H + b * CHUNK, ra, cq * 2);
        });
        h.sig = sg;
      }
      fence_async_smem();
    } else if (k == K_ALBSH) {
      // ---- relu(albedo1 | shading1) -> albedo2 / shading2 partial dot products (fp32) -------------------
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, sh = 0.f;
      if (!(P.exp & 1))
      tmem_chunks16<4>(acc, [&](int b, const uint32_t* v) {     // chunks 0,1: albedo_linear1 hidden units, 2,3: shading
        const float4* w = hw_as + b * 64 + cq * 16;
        if (b < 2) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float4 t = __ldg(w + i);
            const float a = fmaxf(__uint_as_float(v[i]), 0.f);
            a0 = fmaf(a, t.x, a0); a1 = fmaf(a, t.y, a1); a2 = fmaf(a, t.z, a2);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i)
            sh = fmaf(fmaxf(__uint_as_float(v[i]), 0.f), __ldg(reinterpret_cast<const float*>(w + i)), sh);
        }
      });
      h.alb[0] = a0; h.alb[1] = a1; h.alb[2] = a2; h.sh = sh;
    } else if (k == K_VIEWS) {
      // ---- relu(views') -> residual partial dot products [endpoint features]; relu(sem1) -> H chunks 0,1;
      //      then combine the four column quarters of every row, sigmoids, raw row --------------------------
      const int64_t tile = ((int64_t)it_of[s] * 2 + s) * gridDim.x + blockIdx.x;   // may run past the end: stores are masked
      const int64_t m = tile * TILE_M + row;
      const bool valid = m < P.a.M;
      float* grow = P.a.raw + (valid ? m : 0) * P.out_ch;
      float r0 = 0.f, r1 = 0.f, r2 = 0.f;
      {
        float* g = (P.a.endpoint && valid) ? grow + INRF_RAW_BASE + P.C + cq * 16 : nullptr;   // endpoint feature rows (fp32, post-ReLU)
        if (!(P.exp & 1))
        tmem_chunks16<2>(acc, [&](int b, const uint32_t* v) {
          const float4* w = hw_res + b * 64 + cq * 16;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float4 t = __ldg(w + i);
            const float a = fmaxf(__uint_as_float(v[i]), 0.f);
            r0 = fmaf(a, t.x, r0); r1 = fmaf(a, t.y, r1); r2 = fmaf(a, t.z, r2);
          }
          if (g != nullptr) {
#pragma unroll
            for (int i = 0; i < 16; ++i) g[b * 64 + i] = fmaxf(__uint_as_float(v[i]), 0.f);
          }
        });
      }
      if (sem) {
        tmem_chunks16<2>(acc + 128, [&](int b, const uint32_t* v) { store16(v, H + b * CHUNK, ra, cq * 2); });
        fence_async_smem();
      }
      // partials of column quarters 1..3 -> scratch (H chunk 3 of this stream), quarter 0 sums in fixed order
      float4* scratch = reinterpret_cast<float4*>(smem + s * SM_STREAM + SM_SCRATCH);
      if (cq != 0) {
        float4* d = scratch + ((cq - 1) * TILE_M + row) * 2;
        d[0] = make_float4(h.sig, r0, r1, r2);
        d[1] = make_float4(h.alb[0], h.alb[1], h.alb[2], h.sh);
      }
      asm volatile("bar.sync 1, 512;" ::: "memory");
      if (cq == 0) {
        float sg = h.sig, al0 = h.alb[0], al1 = h.alb[1], al2 = h.alb[2], shv = h.sh;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const float4 p0 = scratch[(j * TILE_M + row) * 2], p1 = scratch[(j * TILE_M + row) * 2 + 1];
          sg += p0.x; r0 += p0.y; r1 += p0.z; r2 += p0.w;
          al0 += p1.x; al1 += p1.y; al2 += p1.z; shv += p1.w;
        }
        if (valid) {
          float r3[3], a3[3];
          r3[0] = sigmoid_(r0 + __ldg(P.bias + TCB_RES + 0));
          r3[1] = sigmoid_(r1 + __ldg(P.bias + TCB_RES + 1));
          r3[2] = sigmoid_(r2 + __ldg(P.bias + TCB_RES + 2));
          a3[0] = sigmoid_(al0 + __ldg(P.bias + TCB_ALB2 + 0));
          a3[1] = sigmoid_(al1 + __ldg(P.bias + TCB_ALB2 + 1));
          a3[2] = sigmoid_(al2 + __ldg(P.bias + TCB_ALB2 + 2));
          const float shd = sigmoid_(shv + __ldg(P.bias + TCB_SH2));
          const float sigma = sg + __ldg(P.bias + TCB_ALPHA_B);
#pragma unroll
          for (int i = 0; i < 3; ++i) grow[i] = __fadd_rn(__fmul_rn(a3[i], shd), r3[i]);
          grow[3] = sigma;
#pragma unroll
          for (int i = 0; i < 3; ++i) grow[4 + i] = a3[i];
          grow[7] = shd;
#pragma unroll
          for (int i = 0; i < 3; ++i) grow[8 + i] = r3[i];
        }
      }
      asm volatile("bar.sync 1, 512;" ::: "memory");     // scratch is H chunk 3: read before the next layer-0 epilogue
      if (!sem) ++it_of[s];
    } else {
      // ---- semantic logits: 32-column groups round-robin over the column quarters ---------------------------
      const int64_t tile = ((int64_t)it_of[s] * 2 + s) * gridDim.x + blockIdx.x;
      const int64_t m = tile * TILE_M + row;
      const bool valid = m < P.a.M;
      float* grow = P.a.raw + (valid ? m : 0) * P.out_ch;
      for (int c0 = cq * 32; c0 < P.C; c0 += 128) {
        uint32_t v[32];
        tmem_ld32(acc - cq * 16 + c0, v);
        tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c0 + i < P.C) grow[INRF_RAW_BASE + c0 + i] = __uint_as_float(v[i]) + __ldg(P.bias + TCB_SEM2 + c0 + i);
        }
      }
      ++it_of[s];
    }
    tc_fence_before();
    warp_arrive_leader(done, lane);
    PROF_END();
  });
#ifdef TC2_PROF
  { const int role = prof_role(); if (role >= 0) { g_prof[role * 64 + 30] = busy; g_prof[role * 64 + 29] = lat_wake; } }
#endif
}

// ------------------------------------------------------------------------------------------
// kernel
// ------------------------------------------------------------------------------------------
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1) k_mlp_tc2(const __grid_constant__ Params P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t smem_base = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SM_BAR + 8 * B_COUNT);
  const int rank = (int)cluster_ctarank();

  Sync sy;
  sy.bar0 = smem_base + SM_BAR;
  sy.dbg = P.dbg;
  sy.dead = false;
  sy.tile = -1;
  sy.phase = 0;
  if ((smem_base & 1023u) != 0) {                    // SWIZZLE_128B atoms need 1024 B alignment
    if (threadIdx.x == 0 && atomicCAS(P.dbg, 0, 2) == 0) P.dbg[1] = (int)smem_base;
    return;                                            // same for every CTA of the launch
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < NS; ++i) {
      mbar_init(sy.addr(B_WFULL + i), rank == 0 ? 2 : 1);     // leader: own tx + the peer's relay
      mbar_init(sy.addr(B_WEMPTY + i), 1);
    }
    for (int s = 0; s < 2; ++s) {
      const int sb = B_STREAM + 4 * s;
      mbar_init(sy.addr(sb + B_ACC_FULL), 1);
      mbar_init(sy.addr(sb + B_EPI_DONE), 32);
      mbar_init(sy.addr(sb + B_XREADY), 8);
      mbar_init(sy.addr(sb + B_XFREE), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == W_PROD) {
    // constant "ones" tile of the bias MMAs: 2 core matrices of 8 rows x 8 halves; K columns 0..2 are 1.0
    __half* ones = reinterpret_cast<__half*>(smem + SM_ONES);
    for (int i = lane; i < 128; i += 32) {
      const int k = (i >> 6) * 8 + (i & 7);          // element i = core*64 + row*8 + kk
      ones[i] = __float2half_rn(k < 3 ? 1.f : 0.f);
    }
    fence_async_smem();
  }
  if (warp == W_MMA) {                               // same warp in both CTAs: pair-wide TMEM allocation
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                  // the peer's barriers exist before any multicast commit / remote arrive
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  // "free"-type barriers start released: the first wait must pass on a fresh barrier
  sy.phase = (((1u << NS) - 1u) << B_WEMPTY) | (1u << (B_STREAM + B_EPI_DONE)) | (1u << (B_STREAM + 4 + B_EPI_DONE)) |
             (1u << (B_STREAM + B_XFREE)) | (1u << (B_STREAM + 4 + B_XFREE));

#ifdef TC2_PROF
  const long long t_start = clock64();
#endif
  // warp roles.  The scheduler favours the highest warp id of each sub-partition, so the MMA issuer (21)
  // and the weight producer (20) sit on top of theirs; both run converged and elect one lane to issue.
  if (warp == W_PROD) {
    if (!P.no_weights) producer(P, sy, smem_base, rank);
  } else if (warp == W_MMA) {
    if (rank == 0) issuer(P, sy, smem_base, tmem, smem);
    else if (!P.no_weights) relay(P, sy);
  } else if (warp >= W_FE0) {
    FrontEnd fe{P, sy, smem_base, (warp - W_FE0) * 32 + lane, lane, {}, {0u, 0u}};
    fe.ra.init(fe.row);
    fe.xready[0] = mapa(sy.addr(B_STREAM + B_XREADY), 0);
    fe.xready[1] = mapa(sy.addr(B_STREAM + 4 + B_XREADY), 0);
    fe.run();
  } else {
    epilogue(P, sy, smem, smem_base, tmem, warp & 3, warp >> 2, lane);
  }
#ifdef TC2_PROF
  { const int role = prof_role(); if (role >= 0) g_prof[role * 64 + 31] = clock64() - t_start; }
#endif
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                  // nobody leaves while the pair's MMAs / commits may still touch this CTA
  if (warp == W_MMA) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

}  // namespace tc2

int launch_mlp_tc2(const MlpArgs& a, cudaStream_t st) {
  if (a.M == 0) return INRF_OK;
  NetLayout L;
  int rc = make_layout(a.variant, a.n_classes, &L);
  if (rc) return rc;
  TcProgram prog;
  rc = make_tc_program(a.variant, a.n_classes, &prog);
  if (rc) return rc;
  tc2::Params P;
  memset(&P, 0, sizeof(P));
  P.a = a;
  const unsigned char* blob = static_cast<const unsigned char*>(a.packed);
  P.blocks = blob + L.tc_blocks;
  P.bias = reinterpret_cast<const float*>(blob + L.tc_bias);
  const bool sem = a.n_classes > 0;
  P.n_steps = sem ? 11 : 10;
  // ring fills in issue order: trunk 0..7, albedo1|shading1, views' [| sem1], [semantic logits]
  const int step_of[11] = {0, 1, 2, 3, 4, 5, 6, 7, TS_ALBSH, TS_VIEWS, TS_SEM2};
  int nf = 0;
  for (int k = 0; k < P.n_steps; ++k) {
    P.step_fill0[k] = nf;
    for (int f = 0; f < prog.n_fills; ++f) {
      if (prog.fill_step[f] != step_of[k]) continue;
      if (nf >= tc2::MAX_FILLS) { set_error("internal: too many ring fills"); return INRF_EINVAL; }
      const int bytes = prog.fill_bytes[f];
      if (step_of[k] == TS_SEM2) {
        // two K chunks of ceil16(C) rows each: every CTA takes its row half of both
        const int blk = bytes / 2;
        for (int r = 0; r < 2; ++r) {
          P.fill_src[r][nf] = prog.fill_off[f] + r * (blk / 2);
          P.fill_src2[r][nf] = prog.fill_off[f] + blk + r * (blk / 2);
        }
        P.fill_len[nf] = blk / 2;
        P.fill_len2[nf] = blk / 2;
      } else {
        for (int r = 0; r < 2; ++r) P.fill_src[r][nf] = prog.fill_off[f] + r * (bytes / 2);
        P.fill_len[nf] = bytes / 2;
      }
      if (P.fill_len[nf] + P.fill_len2[nf] > tc2::SLOT || (P.fill_len[nf] & 15) || (P.fill_len2[nf] & 15)) {
        set_error("internal: ring fill %d has %d+%d bytes", f, P.fill_len[nf], P.fill_len2[nf]);
        return INRF_EINVAL;
      }
      ++nf;
    }
    P.step_nfill[k] = nf - P.step_fill0[k];
  }
  P.out_ch = raw_channels(a.n_classes, a.endpoint);
  P.C = a.n_classes;
  P.sem_rows = (a.n_classes + 15) / 16 * 16;
  static const bool now_env = getenv("INRF_TC_NOWEIGHTS") != nullptr && getenv("INRF_TC_NOWEIGHTS")[0] == '1';
  P.no_weights = now_env ? 1 : 0;
  static const int exp_env = getenv("INRF_TC_EXP") ? atoi(getenv("INRF_TC_EXP")) : 0;
  P.exp = exp_env;
  int* dbg = nullptr;
  INRF_CUDA(cudaGetSymbolAddress((void**)&dbg, tc2::g_dbg));
  P.dbg = dbg;
  static const bool checked = getenv("INRF_TC_CHECK") != nullptr && getenv("INRF_TC_CHECK")[0] == '1';
  if (checked) INRF_CUDA(cudaMemsetAsync(dbg, 0, 16 * sizeof(int), st));
  int dev = 0, sms = 148;
  INRF_CUDA(cudaGetDevice(&dev));
  INRF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int64_t tiles = (a.M + tc2::TILE_M - 1) / tc2::TILE_M;
  int64_t want = (tiles + 1) / 2;              // two tiles per CTA and iteration
  want = (want + 1) / 2 * 2;                   // whole pairs
  int grid = (int)(want < sms ? want : sms / 2 * 2);
  P.n_iter = (int)((tiles + 2 * (int64_t)grid - 1) / (2 * (int64_t)grid));
  INRF_CUDA(cudaFuncSetAttribute(tc2::k_mlp_tc2, cudaFuncAttributeMaxDynamicSharedMemorySize, tc2::SM_TOTAL));
#ifdef TC2_PROF
  {
    long long* pp = nullptr;
    INRF_CUDA(cudaGetSymbolAddress((void**)&pp, tc2::g_prof));
    INRF_CUDA(cudaMemsetAsync(pp, 0, sizeof(long long) * 8 * 64, st));
  }
#endif
  tc2::k_mlp_tc2<<<grid, tc2::NUM_THREADS, tc2::SM_TOTAL, st>>>(P);
  INRF_LAUNCH_CHECK();
#ifdef TC2_PROF
  {
    static const char* roles[] = {"issuer", "producer0", "frontend0", "epilogue0.s0", "epilogue0.s1", "relay", "producer1", "epilogue1.s0"};
    static const char* bars[] = {"WFULL0", "WFULL1", "WFULL2", "WFULL3", "WEMPTY0", "WEMPTY1", "WEMPTY2", "WEMPTY3",
                                 "ACC_FULL.0", "EPI_DONE.0", "XREADY.0", "XFREE.0", "ACC_FULL.1", "EPI_DONE.1", "XREADY.1", "XFREE.1"};
    long long h[8 * 64];
    INRF_CUDA(cudaStreamSynchronize(st));
    INRF_CUDA(cudaMemcpyFromSymbol(h, tc2::g_prof, sizeof(h)));
    for (int r = 0; r < 8; ++r) {
      fprintf(stderr, "TC2PROF role=%s total=%lld busy=%lld lat=%lld n_iter=%d grid=%d\n", roles[r], h[r * 64 + 31], h[r * 64 + 30], h[r * 64 + 29], P.n_iter, grid);
      for (int b = 0; b < 16; ++b)
        if (h[r * 64 + 32 + b]) fprintf(stderr, "TC2PROF   %-10s waited %12lld cycles over %8lld waits\n", bars[b], h[r * 64 + b], h[r * 64 + 32 + b]);
    }
  }
#endif
  if (checked) {      // debug mode: synchronise and surface watchdog records as errors
    int h[16];
    INRF_CUDA(cudaStreamSynchronize(st));
    INRF_CUDA(cudaMemcpy(h, dbg, sizeof(h), cudaMemcpyDeviceToHost));
    if (h[0] == 1) { set_error("mlp_tc2 watchdog: barrier %d stuck (warp %d, tile %d, cta %d, parity %d)", h[1], h[2], h[3], h[4], h[5]); return INRF_ECUDA; }
    if (h[0] == 2) { set_error("mlp_tc2: dynamic shared memory base 0x%x is not 1024-byte aligned", h[1]); return INRF_ECUDA; }
  }
  return INRF_OK;
}

}  // namespace inrf
