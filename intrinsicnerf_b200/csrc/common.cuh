// Shared definitions for libinrf.so (sm_100a).  See include/inrf.h for the ABI.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "../../include/inrf.h"

namespace inrf {

// ---------------------------------------------------------------------------------
// error plumbing (thread-local message, integer codes across the C boundary)
// ---------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define INRF_CHECK_ARG(cond, msg)                                   \
  do {                                                              \
    if (!(cond)) { ::inrf::set_error("%s: %s", __func__, msg); return INRF_EINVAL; } \
  } while (0)
#define INRF_CHECK_SUPPORTED(cond, msg)                             \
  do {                                                              \
    if (!(cond)) { ::inrf::set_error("%s: unsupported: %s", __func__, msg); return INRF_EUNSUPPORTED; } \
  } while (0)
#define INRF_CUDA(call)                                             \
  do {                                                              \
    cudaError_t e__ = (call);                                       \
    if (e__ != cudaSuccess) return ::inrf::cuda_fail(e__, #call);   \
  } while (0)
void note_launch();                      // counts kernel launches (inrf_launch_count: the bench's gpu_launches claim)
#define INRF_LAUNCH_CHECK() do { ::inrf::note_launch(); INRF_CUDA(cudaGetLastError()); } while (0)

// ---------------------------------------------------------------------------------
// Deferred device status (status.cu).  Kernels cannot return codes, and no hot entry point may
// synchronise, so a kernel that trips its barrier watchdog or leaves the fp16 range writes a small
// record into a pinned, device-mapped host buffer (one per device, 64 bytes, allocated on first use);
// the host reads it - a plain memory load - at the entry of the next inrf_* call (status_poll) and
// turns it into an error code + message.  Records: [0] code, [1..6] details.
// ---------------------------------------------------------------------------------
enum DevStatusCode {
  DST_NONE = 0,
  DST_WATCHDOG = 1,      // a role of a tensor-core kernel waited > watchdog cycles on an mbarrier
  DST_SMEM_ALIGN = 2,    // dynamic shared memory base not 1024-byte aligned
  DST_F16_ACT = 3,       // a hidden activation reached the fp16 limit (saturated at 65504)
  DST_F16_WEIGHT = 4,    // a weight / bias does not fit fp16 (saturated at pack time)
  DST_F16_GRAD = 5,      // non-finite value in the tensor-core backward (fp16 gradient range)
};
int* status_flag_dev();                 // device-visible pointer of the current device's record (nullptr: error set)
int status_poll();                      // INRF_OK, or INRF_ECUDA / INRF_ERANGE with the message set; clears the record
unsigned long long* rng_epoch_dev();    // the current device's RNG epoch word (device memory, starts at 0; nullptr: error set)
int rng_epoch_set(int bump, cudaStream_t st);   // bump != 0: += 1 (a kernel, graph-capturable); 0: reset to 0
#ifdef __CUDACC__
// coarse sample depths (run_nerf.py:464-486; trainer.py:730-746) - shared by k_coarse_z and the fused kernel's front end
__device__ __forceinline__ float coarse_depth(float nearv, float farv, float t, int lindisp) {
  if (!lindisp) return __fadd_rn(__fmul_rn(nearv, __fsub_rn(1.f, t)), __fmul_rn(farv, t));
  float a = __fmul_rn(__fdiv_rn(1.f, nearv), __fsub_rn(1.f, t));
  float b = __fmul_rn(__fdiv_rn(1.f, farv), t);
  return __fdiv_rn(1.f, __fadd_rn(a, b));
}
// depth of sample s of a ray; t_rand_s: this sample's stratified-jitter draw (ignored when jitter == false)
__device__ __forceinline__ float coarse_z_sample(float nearv, float farv, const float* __restrict__ t_vals, int s, int S, int lindisp,
                                                 bool jitter, float t_rand_s) {
  float zc = coarse_depth(nearv, farv, t_vals[s], lindisp);
  if (jitter) {
    float zl = s > 0 ? coarse_depth(nearv, farv, t_vals[s - 1], lindisp) : zc;
    float zr = s < S - 1 ? coarse_depth(nearv, farv, t_vals[s + 1], lindisp) : zc;
    float lower = s > 0 ? __fmul_rn(0.5f, __fadd_rn(zc, zl)) : zc;
    float upper = s < S - 1 ? __fmul_rn(0.5f, __fadd_rn(zr, zc)) : zc;
    zc = __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), t_rand_s));
  }
  return zc;
}
// one writer per launch is elected by the caller (a device-side atomicCAS on its own claim word)
__device__ __forceinline__ void status_raise(int* flag, int code, int a = 0, int b = 0, int c = 0, int d = 0, int e = 0) {
  if (flag == nullptr) return;
  volatile int* f = flag;
  if (f[0] != 0) return;                  // an earlier record the host has not consumed yet wins
  f[1] = a; f[2] = b; f[3] = c; f[4] = d; f[5] = e;
  __threadfence_system();
  f[0] = code;
  __threadfence_system();
}
#endif

// ---------------------------------------------------------------------------------
// Network description (fixed architecture: D=8, W=256, skips=[4], PE L=10 / L=4)
// ---------------------------------------------------------------------------------
constexpr int W_HID = 256;
constexpr int PE_PTS = 63;    // 3 + 6*10
constexpr int PE_DIR = 27;    // 3 + 6*4
constexpr int MAX_CLASSES = 112;   // 11 + C + 128 endpoint must stay <= 256 raw channels

enum LayerId {
  L_T0 = 0, L_T1, L_T2, L_T3, L_T4, L_T5, L_T6, L_T7,
  L_ALPHA, L_FEAT, L_VIEWS, L_ALB1, L_ALB2, L_SH1, L_SH2, L_RES, L_SEM1, L_SEM2,
  L_COUNT
};

struct LayerDims { int K, N; };

__host__ __device__ inline LayerDims layer_dims(int l, int n_classes) {
  switch (l) {
    case L_T0: return {PE_PTS, W_HID};
    case L_T5: return {PE_PTS + W_HID, W_HID};
    case L_T1: case L_T2: case L_T3: case L_T4: case L_T6: case L_T7: return {W_HID, W_HID};
    case L_ALPHA: return {W_HID, 1};
    case L_FEAT: return {W_HID, W_HID};
    case L_VIEWS: return {W_HID + PE_DIR, 128};
    case L_ALB1: case L_SH1: case L_SEM1: return {W_HID, 128};
    case L_ALB2: case L_RES: return {128, 3};
    case L_SH2: return {128, 1};
    case L_SEM2: return {128, n_classes};
    default: return {0, 0};
  }
}

// Offsets (in floats) of each layer's weight / bias inside the canonical flat parameter
// vector, and inside the fp32 section of the packed blob.
struct NetLayout {
  int variant, n_classes, n_layers;
  int64_t flat_w[L_COUNT], flat_b[L_COUNT];   // canonical flat vector
  int64_t flat_count;
  // packed blob (byte offsets)
  int64_t f32_wt[L_COUNT];   // fp32 weights: transposed [K][N] for N>=128, else original [N][K]
  int64_t f32_b[L_COUNT];    // fp32 bias
  int64_t comp;              // fp32 composed views' weight [128][256] followed by its bias [128]
  int64_t tc_bias;           // fp32 bias table for the tensor-core kernel (see mlp_tc.cu)
  int64_t tc_blocks;         // fp16 pre-swizzled operand blocks
  int64_t tc_blocks_bytes;
  int64_t total_bytes;
};

int make_layout(int variant, int n_classes, NetLayout* out);   // returns INRF_* code

inline int raw_channels(int n_classes, int endpoint) { return INRF_RAW_BASE + n_classes + (endpoint ? 128 : 0); }
inline int rec_channels(int n_classes, int endpoint) { return INRF_REC_BASE + n_classes + (endpoint ? 128 : 0); }

// ---------------------------------------------------------------------------------
// Tensor-core operand block table (pack.cu builds the blob in this order, mlp_tc.cu
// streams it in the same order).  One block = rows x 64 fp16 (128 B per row), stored in
// the UMMA canonical K-major SWIZZLE_128B layout: 8-row atoms of 1024 B, the 16-byte
// unit index XORed with (row & 7).
// ---------------------------------------------------------------------------------
struct TcBlock {
  int16_t layer;     // source layer, or -1 for a composed matrix (views' = views o feature)
  int16_t rows;      // N rows in this block (128, 16 or 32)
  int16_t n0;        // first output row of the source matrix
  int16_t k0;        // first source K column (may be negative for padding layouts)
  int16_t kcols;     // valid source columns in this block (<=64), rest zero
  int16_t kind;      // 0: plain rows of `layer`; 1: views' (composed); 2: head-pair block-diag;
                     // 3: bias block (128 rows x K=16, no-swizzle core-matrix layout, cols 0..2 = hi/lo/lo2)
  int32_t byte_off;  // offset inside the tc_blocks section
  int32_t bytes;
};
constexpr int TC_MAX_BLOCKS = 128;
constexpr int TC_SLOT_BYTES = 32768;   // one ring stage: a 256-row x 64-K operand tile
constexpr int TC_MAX_FILLS = 64;

// The blob is a sequence of 128-row (or narrower) sub-blocks; consecutive sub-blocks that the MMA
// warp consumes as ONE operand tile (e.g. the two N halves of a 256-wide layer for one K chunk)
// are adjacent and streamed as one "fill" of the shared-memory ring.
// GEMM steps of one tile: trunk layers 0..7, then the tail
enum TcStep { TS_VIEWS = 8, TS_ALBSH = 9, TS_RES = 10, TS_ALB2SH2 = 11, TS_SEM2 = 12 };
constexpr int TC_BIAS_FLOATS = 8192;   // fp32 table behind the operand blocks (biases, sigma row, fp32 head weights)
struct TcProgram {
  int n_blocks;
  int bytes;
  int n_fills;
  int fill_off[TC_MAX_FILLS];
  int fill_bytes[TC_MAX_FILLS];
  int16_t fill_step[TC_MAX_FILLS];     // TcStep of the GEMM this fill feeds
  int16_t fill_block0[TC_MAX_FILLS];   // first sub-block of the fill
  TcBlock blk[TC_MAX_BLOCKS];
};
int make_tc_program(int variant, int n_classes, TcProgram* prog);

// Counter-based random numbers for the training-mode draws (stratified jitter run_nerf.py:472-486,
// u of sample_pdf run_nerf_helpers.py:414, sigma noise run_nerf.py:385-387): Philox4x32-10 keyed by the caller's seed,
// counter = (element index, stream id), so a draw is a pure function of (seed, which tensor, which element) - the
// backward pass regenerates the forward's noise instead of reading it back, and no generator kernel is launched.
// `epoch` (optional) points at a device word that is mixed into the key: a CUDA graph bakes `seed` into its kernel
// nodes, so a captured training step bumps the word with one tiny kernel at its start (inrf_rng_epoch_bump) and every
// replay draws fresh numbers while forward and backward of one replay still see the same ones.
struct Rng { unsigned long long seed; unsigned int stream; float scale; int on; const unsigned long long* epoch; };
enum { RNG_T_RAND = 1, RNG_U = 2, RNG_NOISE_COARSE = 3, RNG_NOISE_FINE = 4 };
#ifdef __CUDACC__
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned int hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const unsigned int hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u; k.y += 0xBB67AE85u;
  }
  return c;
}
__device__ __forceinline__ uint4 rng_bits(const Rng& g, long long i) {
  const unsigned long long key = g.seed + (g.epoch != nullptr ? __ldg(g.epoch) * 0x9E3779B97F4A7C15ull : 0ull);   // epoch 0: the seed itself
  return philox4x32_10(make_uint4((unsigned int)i, (unsigned int)((unsigned long long)i >> 32), g.stream, 0u),
                       make_uint2((unsigned int)key, (unsigned int)(key >> 32)));
}
// U[0,1) with 24 random bits (what torch.rand produces for float32)
__device__ __forceinline__ float rng_uniform(const Rng& g, long long i) { return (float)(rng_bits(g, i).x >> 8) * 5.9604644775390625e-08f; }
// N(0,1) * scale (Box-Muller on two of the four words)
__device__ __forceinline__ float rng_normal(const Rng& g, long long i) {
  const uint4 b = rng_bits(g, i);
  const float u1 = (float)((b.x >> 8) + 1u) * 5.9604644775390625e-08f;      // (0,1]
  const float u2 = (float)(b.y >> 8) * 5.9604644775390625e-08f;
  return g.scale * sqrtf(-2.f * logf(u1)) * cospif(2.f * u2);
}
#endif

// Pinhole camera of one frame (get_rays + render()'s packing, run_nerf_helpers.py:359-368 / run_nerf.py:100-128;
// SSR: rays.py:48-84): the ray record of pixel p is a pure function of it, so kernels can generate rays instead
// of reading a [N,11] table.
struct Cam { float fx, fy, cx, cy, m[12], nearv, farv; int opencv, euclidean; };
#ifdef __CUDACC__
// ray record o3 d3 near far viewdir3 of flat pixel p = row * W + column (the arithmetic of k_get_rays, stages.cu)
__device__ __forceinline__ void cam_ray(const Cam& c, int H, int W, int64_t p, float* o) {
  if (p < 0) p = 0;
  if (p >= (int64_t)H * W) p = (int64_t)H * W - 1;
  const int j = (int)(p / W), i = (int)(p - (int64_t)j * W);
  float dx = __fdiv_rn(__fsub_rn((float)i, c.cx), c.fx);
  float dy = __fdiv_rn(__fsub_rn((float)j, c.cy), c.fy);
  float dz = 1.f;
  if (!c.opencv) { dy = -dy; dz = -1.f; }          // OpenGL: x right, y up, camera looks along -z
  if (c.euclidean) {                                 // depth_type == "euclidean": unit camera-frame directions
    const float inv = __fdiv_rn(1.f, sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz))));
    dx = __fmul_rn(dx, inv); dy = __fmul_rn(dy, inv); dz = __fmul_rn(dz, inv);
  }
  float d[3];
#pragma unroll
  for (int r = 0; r < 3; ++r)      // torch.sum(dirs[..., None, :] * c2w[:3, :3], -1): ((x*m0 + y*m1) + z*m2)
    d[r] = __fadd_rn(__fadd_rn(__fmul_rn(dx, c.m[4 * r]), __fmul_rn(dy, c.m[4 * r + 1])), __fmul_rn(dz, c.m[4 * r + 2]));
  const float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])), __fmul_rn(d[2], d[2])));
  o[0] = c.m[3]; o[1] = c.m[7]; o[2] = c.m[11];
  o[3] = d[0]; o[4] = d[1]; o[5] = d[2];
  o[6] = c.nearv; o[7] = c.farv;
  o[8] = __fdiv_rn(d[0], nrm); o[9] = __fdiv_rn(d[1], nrm); o[10] = __fdiv_rn(d[2], nrm);
}
#endif

// ---------------------------------------------------------------------------------
// launchers implemented in the .cu files
// ---------------------------------------------------------------------------------
struct MlpArgs {
  const void* packed;
  int variant, n_classes, endpoint;
  float pe_scale;            // scalar_factor for points
  // addressing mode A: explicit points
  const float* pts;          // [M,3] or null
  const float* viewdirs;     // [M,3] or null
  // addressing mode C: rows already embedded (NeRF.forward's own input), [M,90] = gamma(x) | gamma(d)
  const float* emb;
  // addressing mode B': rays generated from a camera (fused tensor-core renderer only): ray n = pixel cam_pix0 + n
  int cam_on, cam_H, cam_W;
  int64_t cam_pix0;
  Cam cam;
  // addressing mode B: rays + depths
  const float* rays;         // [N,11] or null
  const float* z;            // [N,S]
  int S;
  int64_t M;                 // total rows (N*S in ray mode)
  float* raw;                // [M,out_ch]
  float* stash;              // optional [M, STASH_LD] fp32 activations for the backward pass (fp32 path only)
  unsigned char* stash_img;  // optional tensor-core training stash: IMG_STASH_SLOTS chunk images per 128-row tile
};

// ---------------------------------------------------------------------------------
// Tensor-core training path (mlp_tc.cu forward with stash, train_tc.cu backward).
// Every activation / gradient tile that a GEMM consumes lives in HBM as the exact image of the shared-memory
// operand chunk: 128 rows x 64 fp16 in the UMMA SWIZZLE_128B layout (8-row atoms of 1024 B, 16-byte unit index
// XOR (row & 7)), 16 KB.  The same image is a K-major operand (rows = M/N, columns = K: the dX GEMMs) and an
// MN-major operand (columns = M/N, rows = K: the dW GEMMs), so plain cp.async.bulk copies feed both.
// ---------------------------------------------------------------------------------
constexpr int IMG_BYTES = 16384;
// forward stash slots of one tile
constexpr int IS_PE = 0, IS_DIR = 1, IS_H = 2 /* + 4*layer + chunk */, IS_V = 34, IS_AS = 36, IS_S1 = 40;
// slots 42..44: the ReLU masks of image slots IS_H.. as bit words - for image slot s, 32-column half jj and row r the
// u32 at word ((s - IS_H) * 2 + jj) * 128 + r holds "column 32 jj + i of that chunk is > 0 in fp16" at bit i/2 (even i) or
// 16 + i/2 (odd i) (1 KB per image, 40 KB per tile).  The dX epilogue reads one coalesced word per thread and chunk instead of the 16 KB image.
constexpr int IS_MASK = 42, IMG_STASH_SLOTS = 45;
// backward workspace slots of one tile: G = head gradients (g_albedo[0:3] g_shading[3] g_residual[4:7] g_sigma[7]
// g_sem[8:8+C]), dZ of albedo1|shading1, views', sem1 and of the eight trunk layers
constexpr int IB_G = 0, IB_DAS = 2, IB_DV = 6, IB_DS1 = 8, IB_DZ = 10 /* + 4*layer + chunk */, IMG_BWD_SLOTS = 42;
// training stash (fp32, per sample row): post-ReLU outputs of the 8 trunk layers, the feature vector,
// relu(albedo1|shading1), relu(views) and relu(sem1)
constexpr int ST_H = 0, ST_FEAT = 2048, ST_AS = 2304, ST_V = 2560, ST_SEM1 = 2688, STASH_LD = 2816;

int launch_mlp_fp32(const MlpArgs& a, cudaStream_t st);
struct MlpBwdArgs {
  const float* flat;         // canonical flat parameters (fp32, nn.Linear layout)
  MlpArgs f;                 // forward addressing (pts/viewdirs | rays,z | emb), M, variant, ...; f.raw = forward output
  const float* stash;        // [M, STASH_LD] written by the training forward
  const float* grad_raw;     // [M, out_ch]
  float* grad_flat;          // [flat_count], accumulated into (+=)
};
int launch_mlp_bwd_fp32(const MlpBwdArgs& a, cudaStream_t st);
// In-kernel compositing / resampling of the fused renderer (mlp_tc.cu back-end warp; api.cu: inrf_render_fwd).
// With a FuseArgs the tensor-core kernel does not write raw rows to HBM: the rows of the tile in flight go through a
// small L2-resident ring and one warp composites them (raw2outputs, run_nerf.py:359-412) into per-ray records, and -
// coarse pass - resamples (sample_pdf + sort, run_nerf.py:499-503) into the fine pass's depths.
struct FuseArgs {
  int white_bkgd, lindisp;
  const float* t_vals;       // coarse pass, depths generated in-kernel (MlpArgs.z == nullptr): linspace(0,1,S)
  const float* t_rand;       // [N,S] stratified jitter or nullptr
  const float* noise;        // [N,S] sigma noise (already scaled) or nullptr
  float* rec;                // [N, 13 + C] per-ray records
  int n_importance;          // > 0: coarse pass also resamples (requires S == 64, n_importance == 128, deterministic u)
  const float* u_det;        // [n_importance] = linspace(0,1,n_importance)
  float* z_out;              // [N, S + n_importance] merged, ascending depths of the fine pass
  float* z_std;              // [N]
  float* ring;               // scratch: mlp_tc_ring_bytes(n_classes) (rows of two tiles per CTA)
};
int64_t mlp_tc_ring_bytes(int n_classes);
int launch_mlp_tc(const MlpArgs& a, cudaStream_t st, const FuseArgs* fuse = nullptr);
struct TcBwdArgs {
  const void* packed;        // packed blob (composed views' matrix, fp32 sections)
  const float* flat;         // canonical flat parameters
  int variant, n_classes, endpoint;
  int64_t M;
  const float* raw;          // [M, out_ch] forward output
  const float* grad_raw;     // [M, out_ch]
  const unsigned char* stash_img;   // forward stash (IMG_STASH_SLOTS images per tile)
  unsigned char* work;       // inrf_mlp_bwd_tc_workspace_bytes()
  float* grad_flat;          // accumulated into (+=)
};
int64_t tc_bwd_workspace_bytes(int variant, int n_classes, int64_t M);
int launch_mlp_bwd_tc(const TcBwdArgs& a, cudaStream_t st);

}  // namespace inrf
