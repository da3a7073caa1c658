/*
 * inrf.h - C ABI of libinrf.so: the B200 (sm_100a) implementation of IntrinsicNeRF's
 * volumetric ray-marching hot path and reflectance clustering.
 *
 * The reference (zju3dv/IntrinsicNeRF) is pure Python/PyTorch and has no FFI: its
 * "operator interface" for this path is a set of module-level Python functions.  Each
 * entry point below names the reference function (file:line under the reference
 * checkout) whose arithmetic it replaces; INTEGRATION.md shows the ctypes binding a
 * maintainer of the reference would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer on the current CUDA device unless marked host;
 *   - all tensors are fp32, row-major, contiguous, unless stated otherwise;
 *   - the caller owns every buffer (inputs, outputs, workspace, packed weights);
 *     the library allocates nothing persistent;
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); no entry point
 *     synchronises the device;
 *   - every entry returns 0 on success and a negative INRF_E* code on failure; the
 *     message is available from inrf_last_error_string() (thread-local).  Unsupported
 *     configurations are rejected - there is no fallback path.
 */
#ifndef INRF_H_
#define INRF_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define INRF_OK            0
#define INRF_EINVAL       -1   /* bad argument (null pointer, negative size, bad enum) */
#define INRF_EUNSUPPORTED -2   /* configuration outside what the kernels implement */
#define INRF_ECUDA        -3   /* CUDA runtime error (message carries cudaGetErrorString) */
#define INRF_EWORKSPACE   -4   /* workspace too small */
#define INRF_ERANGE       -5   /* a value left the fp16 range of the tensor-core path (weights at pack time,
                                  hidden activations or gradients at run time): results of that launch are
                                  saturated, not inf/NaN; rerun with INRF_PREC_FP32 */

/* Network variants.  Both are the 8x256 trunk with skip at layer 4, PE L=10 / L=4. */
#define INRF_NET_OBJECT 0      /* NeRF           object_level/run_nerf_helpers.py:247-325 */
#define INRF_NET_SSR    1      /* Semantic_NeRF  SSR/models/semantic_nerf.py:74-181       */

/* MLP arithmetic. */
#define INRF_PREC_TC   0       /* tcgen05 tensor cores: fp16 operands (RN), fp32 accumulate in
                                  TMEM, sigma head and all epilogues in fp32 */
#define INRF_PREC_FP32 1       /* CUDA-core fp32 FFMA everywhere (validation / strict mode) */

/* Channel layout of a per-sample "raw" row (reference: run_nerf_helpers.py:321,
 * semantic_nerf.py:171-181):  rgb[0:3] sigma[3] albedo[4:7] shading[7] residual[8:11]
 * sem_logits[11:11+C] endpoint_feature[.. +128].                                        */
#define INRF_RAW_BASE 11

/* Per-ray output record written by inrf_raw2outputs / inrf_render_fwd
 * (reference: run_nerf.py:398-412, model_utils.py:84-116):
 *   rgb[0:3] disp[3] acc[4] albedo[5:8] shading[8] residual[9:12] depth[12] sem[13:13+C]
 *   feat[13+C : 13+C+128] (only when endpoint_feat).                                     */
#define INRF_REC_BASE 13

const char* inrf_last_error_string(void);
int inrf_version(void);

/* Deferred device status.  No entry point synchronises the device, so a condition that only a running
 * kernel can detect - a tensor-core kernel whose barrier watchdog tripped (INRF_ECUDA), a weight /
 * activation / gradient outside the fp16 range of INRF_PREC_TC (INRF_ERANGE) - is recorded by the kernel in
 * a 64-byte pinned host buffer (the only allocation the library makes, one per device, on first use) and
 * reported by the NEXT inrf_pack_weights / inrf_mlp_* / inrf_render_fwd call on that device, which returns
 * the code without launching anything.  inrf_poll_status() checks explicitly (after the caller has
 * synchronised its stream, it reports everything enqueued so far); it clears the record.              */
int inrf_poll_status(void);

/* Number of CUDA kernels this library has launched in the calling process so far (all devices, all streams;
 * memsets and copies are not kernels).  bench.py reports the difference across its timed region.          */
int64_t inrf_launch_count(void);

/* ---------------------------------------------------------------------------------
 * Weights
 * --------------------------------------------------------------------------------- */
/* Number of fp32 values in the canonical flat parameter vector of one network:
 * for each layer in the order
 *   pts_linears.0..7, alpha, feature, views_linears.0, albedo1, albedo2, shading1,
 *   shading2, residual, [sem1, sem2 when n_classes>0]
 * the weight ([out,in] row-major, exactly nn.Linear.weight) followed by the bias.
 * (object fork: shading1/2 = test_linear1/2, residual = shading_linear;
 *  run_nerf_helpers.py:259-279.  SSR fork: semantic_nerf.py:96-118.)
 * Returns a negative code for unsupported (variant, n_classes).                        */
int64_t inrf_flat_param_count(int variant, int n_classes);

/* Bytes of the packed weight blob consumed by the kernels. */
int64_t inrf_packed_bytes(int variant, int n_classes);

/* flat_params (fp32, canonical order) -> packed blob (fp32 transposed copy for the
 * CUDA-core path, fp16 pre-swizzled UMMA operand blocks in MMA issue order for the
 * tensor-core path, fp32 biases).  Call again after every optimizer step.              */
int inrf_pack_weights(const float* flat_params, int variant, int n_classes,
                      void* packed, int64_t packed_bytes, void* stream);

/* ---------------------------------------------------------------------------------
 * Stage entry points (one per reference primitive; used by the drop-in Python layer
 * and by the parity tests)
 * --------------------------------------------------------------------------------- */
/* Embedder.embed  (run_nerf_helpers.py:195-225; semantic_nerf.py:14-65):
 * x[M,3] -> out[M,3+6L];  x is divided by `scalar_factor` first (1 for the object fork). */
int inrf_embed(const float* x, int64_t M, int n_freqs, float scalar_factor, float* out, void* stream);

/* run_network (run_nerf.py:42-56; model_utils.py:19-35) fused with the Embedder and the
 * network forward: pts[M,3], viewdirs[M,3] (one row per SAMPLE) -> raw[M,out_ch] with
 * out_ch = 11 + n_classes + (endpoint ? 128 : 0).                                       */
int inrf_mlp_fwd(const void* packed, int variant, int n_classes, int endpoint_feat,
                 float pe_scalar_factor, const float* pts, const float* viewdirs, int64_t M,
                 float* raw, int precision, void* stream);

/* NeRF.forward / Semantic_NeRF.forward on rows that are ALREADY embedded
 * (run_nerf_helpers.py:284, semantic_nerf.py:123): emb[M,90] = gamma(x)[63] | gamma(d)[27].  */
int inrf_mlp_fwd_embedded(const void* packed, int variant, int n_classes, int endpoint_feat,
                          const float* emb, int64_t M, float* raw, int precision, void* stream);

/* Same network evaluation addressed by rays: sample (n,s) sits at o_n + d_n * z[n,s]
 * (run_nerf.py:488,504) with view direction rays[n,8:11].  rays[N,11], z[N,S] -> raw[N,S,out_ch]. */
int inrf_mlp_fwd_rays(const void* packed, int variant, int n_classes, int endpoint_feat,
                      float pe_scalar_factor, const float* rays, const float* z, int64_t N, int S,
                      float* raw, int precision, void* stream);

/* Training forward / backward of the field network (fp32 CUDA cores).
 * Exactly one addressing mode is given: (pts, viewdirs) per sample row, or (rays, z, S), or emb
 * (already embedded rows); the others are NULL.  The forward additionally writes the activation
 * stash [M, inrf_stash_floats_per_row()] that the backward consumes.  The backward ACCUMULATES
 * dL/d(parameters) into grad_flat (canonical flat order of inrf_flat_param_count; the caller zeroes
 * it) given dL/d(raw) [M, out_ch].  Sample positions get no gradient (z_samples is detached in the
 * reference, run_nerf.py:501).  This is what makes `loss.backward()` of run_nerf.py:1018 /
 * trainer.py:990 work on the NeRF / Semantic_NeRF modules.                                        */
int64_t inrf_stash_floats_per_row(void);
int inrf_mlp_fwd_train(const void* packed, int variant, int n_classes, int endpoint_feat, float pe_scalar_factor,
                       const float* pts, const float* viewdirs, const float* rays, const float* z, int S,
                       const float* emb, int64_t M, float* raw, float* stash, void* stream);
int inrf_mlp_bwd(const float* flat_params, int variant, int n_classes, int endpoint_feat, float pe_scalar_factor,
                 const float* pts, const float* viewdirs, const float* rays, const float* z, int S,
                 const float* emb, int64_t M, const float* raw, const float* stash, const float* grad_raw,
                 float* grad_flat, void* stream);

/* Tensor-core training forward / backward (INRF_PREC_TC arithmetic: fp16 operands, fp32 accumulation).
 * The forward is the inference kernel with one addition: every activation tile is also written to
 * `stash_img` (inrf_mlp_stash_img_bytes(M) bytes) as 16 KB images of the tensor-core operand chunks (42 per 128-row
 * tile) followed by the ReLU decisions of those tiles as bit words (3 more 16 KB slots per tile).  The backward
 * walks those images with tcgen05 GEMMs (dX chain, then all dW in one launch) and ACCUMULATES dL/d(parameters)
 * into grad_flat like inrf_mlp_bwd; `workspace` needs inrf_mlp_bwd_tc_workspace_bytes(variant, n_classes, M) bytes.
 * Gradients are scaled on the device by a power of two derived from max|grad_raw| so that fp16 holds them. */
int64_t inrf_mlp_stash_img_bytes(int64_t M);
int64_t inrf_mlp_bwd_tc_workspace_bytes(int variant, int n_classes, int64_t M);
int inrf_mlp_fwd_train_tc(const void* packed, int variant, int n_classes, int endpoint_feat, float pe_scalar_factor,
                          const float* pts, const float* viewdirs, const float* rays, const float* z, int S,
                          const float* emb, int64_t M, float* raw, void* stash_img, void* stream);
int inrf_mlp_bwd_tc(const void* packed, const float* flat_params, int variant, int n_classes, int endpoint_feat,
                    int64_t M, const float* raw, const void* stash_img, const float* grad_raw, void* workspace,
                    int64_t workspace_bytes, float* grad_flat, void* stream);

/* raw2outputs (run_nerf.py:359-412; model_utils.py:39-116).
 * raw[N,S,ch], z[N,S], rays_d given as rays_d[N,ld] with row stride ld floats (ld=3 for a
 * packed [N,3] tensor, 11 to address columns 3:6 of a ray record - pass rays+3).
 * noise[N,S] (already multiplied by raw_noise_std) may be NULL.
 * Outputs: rec[N, 13 + n_classes + (endpoint?128:0)], weights[N,S] (may be NULL).        */
int inrf_raw2outputs(const float* raw, const float* z, const float* rays_d, int ld_rays_d,
                     const float* noise, int64_t N, int S, int n_classes, int endpoint_feat,
                     int white_bkgd, float* rec, float* weights, void* stream);

/* Backward of raw2outputs: given the forward inputs and dL/d(rec) [N, rec_ch] (and optionally
 * dL/d(weights) [N,S]), writes dL/d(raw) [N,S,ch].  z, rays_d and noise receive no gradient (the
 * reference never differentiates them: z_samples is detached, run_nerf.py:501).  S <= 256.
 * The disparity output contributes through d(1/max(1e-10, depth/acc)).                           */
int inrf_raw2outputs_bwd(const float* raw, const float* z, const float* rays_d, int ld_rays_d,
                         const float* noise, int64_t N, int S, int n_classes, int endpoint_feat,
                         int white_bkgd, const float* grad_rec, const float* grad_weights,
                         float* grad_raw, void* stream);

/* sample_pdf (run_nerf_helpers.py:402-445; SSR/models/rays.py:176-220).
 * bins[N,B], weights given with row stride ld_w and B-1 used entries per row
 * (pass weights_coarse+1 with ld_w=S to express weights[...,1:-1]).
 * u[N,n_samples] or NULL for det (u = linspace(0,1,n_samples), passed by the caller in
 * u_det[n_samples] so that it carries torch.linspace's exact values).
 * Outputs: samples[N,n_samples]; inds[N,n_samples] int64 (searchsorted(cdf,u,right=True),
 * may be NULL); cdf_out[N,B] (may be NULL).                                              */
int inrf_sample_pdf(const float* bins, const float* weights, int ld_w, const float* u,
                    const float* u_det, int64_t N, int B, int n_samples,
                    float* samples, int64_t* inds, float* cdf_out, void* stream);

/* Inversion only, for a caller-supplied cdf[N,B] (the exact-index contract of SURVEY 7.3). */
int inrf_invert_cdf(const float* bins, const float* cdf, const float* u, int64_t N, int B,
                    int n_samples, float* samples, int64_t* inds, void* stream);

/* sort(cat([z_vals, z_samples])) and std(z_samples, unbiased=False)
 * (run_nerf.py:503,519; trainer.py:766,800).  z_a[N,Sa], z_b[N,Sb] -> z_out[N,Sa+Sb]
 * ascending, z_std[N] over z_b (may be NULL).  Sa+Sb <= 1024.                            */
int inrf_merge_sorted(const float* z_a, const float* z_b, int64_t N, int Sa, int Sb,
                      float* z_out, float* z_std, void* stream);

/* Coarse sample depths (run_nerf.py:464-486; trainer.py:730-746): t_vals[S] is
 * torch.linspace(0,1,S); t_rand[N,S] or NULL (no stratified jitter). rays[N,11] -> z[N,S]. */
int inrf_coarse_z(const float* rays, const float* t_vals, const float* t_rand, int64_t N, int S,
                  int lindisp, float* z, void* stream);

/* Training-mode draws generated INSIDE the stage kernels.  The reference draws
 * t_rand = torch.rand(N,Sc) (run_nerf.py:478), u = torch.rand(N,Sf) (run_nerf_helpers.py:414) and
 * noise = torch.randn(N,S) * raw_noise_std (run_nerf.py:387) with three generator launches per pass; here a draw is
 * Philox4x32-10(key = seed, counter = (element index, tensor id)) evaluated where it is consumed: U[0,1) with 24 bits
 * for t_rand / u, Box-Muller N(0,1) * noise_std for the sigma noise.  The same (seed, element) always gives the same
 * number, so inrf_raw2outputs_bwd_rng regenerates the forward's noise.  `fine_pass` selects the noise tensor
 * (coarse / fine pass of one step use different streams of the same seed).  Everything else as the non-_rng entries.
 *
 * CUDA graphs: `seed` is a by-value argument and therefore baked into a captured launch.  The key of every draw is
 * seed + epoch * 0x9E3779B97F4A7C15, where `epoch` is a device-resident word (one per device, 0 until bumped).
 * inrf_rng_epoch_bump enqueues a one-thread kernel that increments it - capture it at the start of a training step and
 * every replay draws fresh jitter / noise while the backward of that replay regenerates exactly the forward's numbers;
 * inrf_rng_epoch_reset sets it back to 0 (the eager behaviour: key = seed). */
int inrf_rng_epoch_bump(void* stream);
int inrf_rng_epoch_reset(void* stream);
int inrf_coarse_z_rng(const float* rays, const float* t_vals, uint64_t seed, int64_t N, int S, int lindisp, float* z,
                      void* stream);
int inrf_sample_pdf_rng(const float* bins, const float* weights, int ld_w, uint64_t seed, int64_t N, int B,
                        int n_samples, float* samples, void* stream);
int inrf_raw2outputs_rng(const float* raw, const float* z, const float* rays_d, int ld_rays_d, float noise_std,
                         uint64_t seed, int fine_pass, int64_t N, int S, int n_classes, int endpoint_feat,
                         int white_bkgd, float* rec, float* weights, void* stream);
int inrf_raw2outputs_bwd_rng(const float* raw, const float* z, const float* rays_d, int ld_rays_d, float noise_std,
                             uint64_t seed, int fine_pass, int64_t N, int S, int n_classes, int endpoint_feat,
                             int white_bkgd, const float* grad_rec, const float* grad_weights, float* grad_raw,
                             void* stream);

/* Pinhole ray generation + packing of render() for a full image (run_nerf_helpers.py:359-368
 * get_rays, run_nerf.py:100-128 with use_viewdirs=True, ndc=False): pixel (i=column, j=row) ->
 * dir = ((i-cx)/fx, -(j-cy)/fy, -1), d = R dir, o = t, viewdir = d/|d|.
 * c2w[12] is the row-major 3x4 camera-to-world matrix (HOST pointer), rays[H*W,11] device.      */
int inrf_get_rays(int H, int W, float fx, float fy, float cx, float cy, const float* c2w_host,
                  float near, float far, float* rays, void* stream);

/* The same for selected pixels and for the SSR fork's create_rays (SSR/models/rays.py:48-76
 * get_rays_camera, :79-84 get_rays_world, :223-256 create_rays): pix[N] (device) holds flat pixel indices
 * row*W + column - e.g. sampling_index (rays.py:153-172) or the select_coords / neighbour pairs of
 * run_nerf.py:913-932; NULL means all H*W pixels in order (N must be H*W).  convention:
 * INRF_CAM_OPENGL dir = ((i-cx)/fx, -(j-cy)/fy, -1), INRF_CAM_OPENCV dir = ((i-cx)/fx, (j-cy)/fy, 1);
 * euclidean != 0 normalises the camera-frame direction first (depth_type "euclidean").  This replaces
 * the precomputed [n_images, H*W, 11] table (608 MB at the Replica config) by on-demand generation. */
#define INRF_CAM_OPENGL 0
#define INRF_CAM_OPENCV 1
int inrf_rays_from_pixels(const int64_t* pix, int64_t N, int H, int W, float fx, float fy, float cx,
                          float cy, const float* c2w_host, int convention, int euclidean, float near,
                          float far, float* rays, void* stream);

/* ---------------------------------------------------------------------------------
 * Fused renderer: render_rays (run_nerf.py:415-528) / SSRTrainer.volumetric_rendering
 * (SSR/training/trainer.py:717-808) for one chunk of rays.
 * --------------------------------------------------------------------------------- */
typedef struct InrfRenderCfg {
  int32_t variant;         /* INRF_NET_*                                                 */
  int32_t n_classes;       /* semantic classes C (0 = no semantic head)                  */
  int32_t n_samples;       /* coarse samples Sc (<= 256)                                 */
  int32_t n_importance;    /* fine samples Sf (0 = coarse pass only); Sc+Sf <= 1024      */
  int32_t lindisp;         /* sample linearly in inverse depth (object fork only)        */
  int32_t white_bkgd;
  int32_t endpoint_feat;   /* fine pass appends the 128-d endpoint feature (SSR)         */
  int32_t precision;       /* INRF_PREC_*                                                */
  float   pe_scalar_factor;/* 1 (object) or 10 (SSR points); directions always use 1     */
  int32_t reserved[7];
} InrfRenderCfg;

/* Workspace bytes needed by inrf_render_fwd for N rays (raw_* buffers included unless the
 * caller passes its own). */
int64_t inrf_render_workspace_bytes(const InrfRenderCfg* cfg, int64_t N);

/* rays[N,11] = o3 d3 near far viewdir3.  Randomness is injected: t_rand[N,Sc] (NULL = no
 * jitter), u[N,Sf] (NULL = det), noise_coarse[N,Sc] / noise_fine[N,Sc+Sf] already scaled
 * by raw_noise_std (NULL = none).  t_vals[Sc], u_det[Sf] are torch.linspace(0,1,.).
 * Outputs (each may be NULL when not wanted, except rec_fine / rec_coarse):
 *   rec_coarse[N,13+C], rec_fine[N,13+C(+128)], z_std[N],
 *   raw_coarse[N,Sc,11+C], raw_fine[N,Sc+Sf,11+C(+128)]  (taken from the workspace when NULL),
 *   z_fine[N,Sc+Sf] (merged sorted depths), weights_fine[N,Sc+Sf].                        */
int inrf_render_fwd(const float* rays, int64_t N, const void* packed_coarse, const void* packed_fine,
                    const InrfRenderCfg* cfg, const float* t_vals, const float* u_det,
                    const float* t_rand, const float* u, const float* noise_coarse,
                    const float* noise_fine, float* rec_coarse, float* rec_fine, float* z_std,
                    float* raw_coarse, float* raw_fine, float* z_fine, float* weights_fine,
                    void* workspace, int64_t workspace_bytes, void* stream);

/* The same renderer for a whole frame (or a pixel range of it) WITHOUT a ray table: render() called with c2w
 * (run_nerf.py:100-103 get_rays, :113-128 packing) and SSRTrainer.render_path over create_rays' rows generate the
 * rays of a frame from (H, W, K, c2w); here the kernels do that per pixel - ray n of the call is pixel pix0 + n
 * (row-major), bit-identical to inrf_get_rays / inrf_rays_from_pixels - so the 44 B/ray table is never written or
 * read.  Deterministic rendering only (no jitter / noise / random u: the frame drivers render with perturb = 0), and
 * only configurations the fused kernel covers (INRF_PREC_TC, sample counts multiples of 32, 64 + 128 when there is a
 * fine pass, no endpoint feature); anything else returns INRF_EUNSUPPORTED and the caller uses inrf_get_rays +
 * inrf_render_fwd.  Outputs as inrf_render_fwd.                                                                  */
typedef struct InrfCamera {
  int32_t H, W;
  float fx, fy, cx, cy;
  float c2w[12];           /* row-major 3x4 camera-to-world */
  float near_, far_;
  int32_t convention;      /* INRF_CAM_OPENGL / INRF_CAM_OPENCV */
  int32_t euclidean;       /* depth_type "euclidean": unit camera-frame directions */
  int32_t reserved[4];
} InrfCamera;
int inrf_render_fwd_camera(const InrfCamera* cam, int64_t pix0, int64_t N, const void* packed_coarse,
                           const void* packed_fine, const InrfRenderCfg* cfg, const float* t_vals,
                           const float* u_det, float* rec_coarse, float* rec_fine, float* z_std, float* z_fine,
                           void* workspace, int64_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------
 * Reflectance clustering (object_level/cluster.py, SSR/training/cluster.py)
 * --------------------------------------------------------------------------------- */
/* Cluster.mapping_color (cluster.py:266-275): rgb[P,3] -> d[P,3] = (I/3*f, g/I, b/I), I=r+g+b. */
int inrf_mapping_color(const float* rgb, int64_t P, float intensity_factor, float* out, void* stream);

/* Cluster.nearest_anchor over mapping_color(rgb) (cluster.py:217-252): for each pixel the
 * index of the anchor minimising |a|^2+|p|^2-2a.p (first minimum wins, as torch.argmin).
 * rgb[P,3], anchors[A,3] -> idx[P] int64.  map_color!=0 applies mapping_color first.       */
int inrf_nearest_anchor(const float* rgb, int64_t P, const float* anchors, int64_t A,
                        int map_color, float intensity_factor, int64_t* idx, void* stream);

/* Cluster.dest_color / dest_class (cluster.py:217-239): nearest anchor, then
 * out_rgb[P,3] = rgb_centers[links[idx]] and/or out_class[P] = links[idx].                 */
int inrf_dest_color(const float* rgb, int64_t P, const float* anchors, const int64_t* links,
                    int64_t A, const float* rgb_centers, int64_t K, float intensity_factor,
                    float* out_rgb, int64_t* out_class, void* stream);

/* Cluster.choose_anchors (cluster.py:150-176): voxelise mapped colours at leaf 0.01 into a
 * 100^3 grid; per occupied voxel keep the pixel closest to the voxel centre (ties: lowest
 * index - the deterministic statement of the reference's sort+last-write-wins scatter).
 * pixels[P,3] (already mapped), labels[P] int64; voxel_key[1e6] uint64 scratch.
 * Outputs anchors[<=1e6,3], links[<=1e6] int64 in ascending voxel order, *n_anchors (device). */
int inrf_choose_anchors(const float* pixels, const int64_t* labels, int64_t P,
                        unsigned long long* voxel_key, float* anchors, int64_t* links,
                        int32_t* n_anchors, void* stream);

/* One flat-kernel mean-shift sweep for a set of seeds (sklearn _mean_shift_single_seed as
 * called from cluster.py:139-140): iterate mean of points within `bandwidth` until the
 * shift is < 1e-3*bandwidth or max_iter.  points[P,3], seeds[Q,3] -> centers[Q,3],
 * n_within[Q] int32 (points inside the final window), n_iter[Q] int32.                      */
int inrf_meanshift_seeds(const float* points, int64_t P, const float* seeds, int64_t Q,
                         float bandwidth, int max_iter, float* centers, int32_t* n_within,
                         int32_t* n_iter, void* stream);

/* sklearn.estimate_bandwidth core: for each query (a subsample of the points) the distance
 * to its k-th nearest neighbour among points[P,3] (k counts the query itself, as
 * NearestNeighbors.kneighbors on the fitted data does).  queries[Q,3] -> kth_dist[Q].      */
int inrf_kth_neighbor_dist(const float* points, int64_t P, const float* queries, int64_t Q,
                           int k, float* kth_dist, void* stream);

/* ---------------------------------------------------------------------------------
 * Fused training losses: img2mse + compute_intrinsic_loss + cluster term
 * (object_level/run_nerf_helpers.py:11, 15-86; SSR/training/training_utils.py:124-207; composed at
 * run_nerf.py:975-1017 / trainer.py:913-990).  Maps are given as pointer + row stride in floats so
 * that the columns of a render record can be passed in place; rgb (and g_rgb) may be NULL.
 * gt_rgb[N,3]; label[N]: object mask (mode 0) or semantic label as float (mode 1);
 * target_albedo[N,3] (Cluster.dest_color) or NULL.
 * losses[8] (device) = { img, chroma, residual, reflect_sparsity, shading_smooth, far_reflect,
 *                        intensity, cluster }.
 * Backward: weights[8] (device) are the upstream gradients of those eight scalars (the loss weights
 * of the step); the gradient of their weighted sum is WRITTEN to g_* (same strides as the maps).
 * --------------------------------------------------------------------------------- */
int inrf_intrinsic_loss_fwd(const float* rgb, int ld_rgb, const float* albedo, int ld_alb,
                            const float* shading, int ld_sh, const float* residual, int ld_res,
                            const float* gt_rgb, const float* label, const float* target_albedo,
                            int64_t N, int mode, float* losses, void* stream);
int inrf_intrinsic_loss_bwd(const float* rgb, int ld_rgb, const float* albedo, int ld_alb,
                            const float* shading, int ld_sh, const float* residual, int ld_res,
                            const float* gt_rgb, const float* label, const float* target_albedo,
                            int64_t N, int mode, const float* weights, float* g_rgb, float* g_albedo,
                            float* g_shading, float* g_residual, void* stream);

/* ---------------------------------------------------------------------------------
 * Full-image driver outputs: what render_path does to every rendered frame on the
 * host after a 48 B/pixel .cpu().numpy() round trip (object_level/run_nerf.py:164-236,
 * SSR/training/trainer.py:1241-1441), done on the device from the per-ray record so that only 8-bit /
 * 16-bit planes (and the float maps the caller really returns) leave HBM.
 *
 * inrf_frame_finish: rec[H*W, rec_stride] (the INRF_REC layout above) -> any of the planes below
 * (every pointer nullable):
 *   to8b(x) = (uint8)(255 * clip(x, 0, 1)), truncating (run_nerf_helpers.py:13; trainer.py:1241-1242);
 *   rgb8[P,3] albedo8[P,3] shading8[P] residual8[P,3];
 *   label:  n_classes == 0 (object fork): label = acc > acc_threshold  (run_nerf.py:174 uses 10, i.e. always 0),
 *           label8 = to8b((float)label) is the 'acc###.png' plane (run_nerf.py:175, 211);
 *           n_classes  > 0 (SSR fork):    label = argmax_c sem_logits (first maximum; = argmax of the softmax,
 *           trainer.py:1243), label8 = (uint8)label;
 *   vis_label8[P,3] = colour_map[label] (uint8 [n_classes,3], trainer.py:1269/1291);
 *   entropy[P] = -sum softmax*log_softmax (trainer.py:1244), entropy8 = to8b(entropy);
 *   disp16 = (uint16)disp, depth_mm16 = (uint16)(depth*1000)  (trainer.py:1351-1352; truncation, values
 *           outside [0, 65535] wrap modulo 2^16 as numpy's C cast does on x86-64, non-finite -> 0);
 *   labels64[P] int64: the label plane in the dtype Cluster_Manager.dest_color consumes;
 *   every sub_step-th row and column (albedo[::2, ::2], label[::2, ::2]; run_nerf.py:183-186,
 *   trainer.py:1330-1334): sample_pixels[ceil(H/s)*ceil(W/s), 3] fp32, sample_labels[...] int64.
 * --------------------------------------------------------------------------------- */
typedef struct InrfFramePlanes {
  uint8_t* rgb8;        uint8_t* albedo8;     uint8_t* shading8;   uint8_t* residual8;
  uint8_t* label8;      uint8_t* vis_label8;  uint8_t* entropy8;   float*   entropy;
  uint16_t* disp16;     uint16_t* depth_mm16; int64_t* labels64;
  float*   sample_pixels;  int64_t* sample_labels;
  void* reserved[3];
} InrfFramePlanes;

int inrf_frame_finish(const float* rec, int32_t H, int32_t W, int32_t rec_stride, int32_t n_classes,
                      float acc_threshold, const uint8_t* colour_map, int32_t sub_step,
                      const InrfFramePlanes* out, void* stream);

/* Clustered-albedo and edit images of render_path (run_nerf.py:228-240, trainer.py:1425-1441):
 * cluster_rgb[P,3] = Cluster_Manager.dest_color(albedo, label);  c8 = to8b(cluster_rgb);
 * edit8 = to8b(cluster_rgb * shading + residual) with shading / residual read from rec.  c8, edit8 nullable. */
int inrf_edit_recompose(const float* cluster_rgb, const float* rec, int64_t P, int32_t rec_stride,
                        uint8_t* c8, uint8_t* edit8, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* INRF_H_ */
