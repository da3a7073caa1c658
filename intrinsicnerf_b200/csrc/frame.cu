// Full-image driver outputs.
//
// The reference's render_path (object_level/run_nerf.py:142-272, SSR/training/trainer.py:1221-1443) pulls six
// float maps per frame to the host (48 B/pixel), converts them to 8-bit with numpy, sub-samples the albedo for
// the cluster refresh, and later uploads the albedo again for Cluster_Manager.dest_color.  Here one launch reads
// the per-ray record once (52 + 4C B/pixel) and writes every plane the driver needs; a second small launch
// recomposes the clustered-albedo / edit images.  Both are HBM-streaming kernels: one thread per pixel, a warp
// touches 32 consecutive record rows (all sectors fully used), byte planes are written as consecutive bytes.
#include "common.cuh"

namespace inrf {

// to8b = lambda x: (255*np.clip(x,0,1)).astype(np.uint8)   (run_nerf_helpers.py:13, trainer.py:1241-1242)
// np.clip propagates NaN; the uint8 cast of NaN is 0 on x86-64 - cvt.rzi.u32.f32 also gives 0.
__device__ __forceinline__ uint8_t to8b(float x) {
  const float c = fminf(fmaxf(x, 0.f), 1.f);                       // fmaxf drops a NaN operand -> 0, the same final byte
  return (uint8_t)__float2uint_rz(__fmul_rn(255.f, c));
}

// ndarray.astype(np.uint16) of a float32: C cast through a 32-bit signed conversion (cvttss2si) - in-range values
// truncate, |v| < 2^31 wraps modulo 2^16, anything else (incl. NaN / inf) gives the "integer indefinite" 0x80000000 -> 0.
__device__ __forceinline__ uint16_t to_u16(float v) {
  if (!(fabsf(v) < 2147483648.f)) return 0;
  return (uint16_t)(uint32_t)__float2int_rz(v);
}

struct FrameArgs {
  const float* rec;
  int H, W, ld, C;
  float acc_thr;
  const uint8_t* cmap;
  int sub;
  InrfFramePlanes o;
};

__global__ void __launch_bounds__(256) k_frame_finish(const FrameArgs a) {
  const int64_t P = (int64_t)a.H * a.W;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (int64_t)gridDim.x * blockDim.x) {
    const float* r = a.rec + p * a.ld;
    const float rgb0 = r[0], rgb1 = r[1], rgb2 = r[2], disp = r[3], acc = r[4];
    const float al0 = r[5], al1 = r[6], al2 = r[7], sh = r[8], re0 = r[9], re1 = r[10], re2 = r[11], depth = r[12];
    if (a.o.rgb8) { uint8_t* d = a.o.rgb8 + 3 * p; d[0] = to8b(rgb0); d[1] = to8b(rgb1); d[2] = to8b(rgb2); }
    if (a.o.albedo8) { uint8_t* d = a.o.albedo8 + 3 * p; d[0] = to8b(al0); d[1] = to8b(al1); d[2] = to8b(al2); }
    if (a.o.shading8) a.o.shading8[p] = to8b(sh);
    if (a.o.residual8) { uint8_t* d = a.o.residual8 + 3 * p; d[0] = to8b(re0); d[1] = to8b(re1); d[2] = to8b(re2); }
    if (a.o.disp16) a.o.disp16[p] = to_u16(disp);
    if (a.o.depth_mm16) a.o.depth_mm16[p] = to_u16(__fmul_rn(depth, 1000.f));

    int label;
    if (a.C == 0) {
      label = acc > a.acc_thr ? 1 : 0;                               // run_nerf.py:174
      if (a.o.label8) a.o.label8[p] = to8b((float)label);            // accs.append(label.astype(float32)) -> to8b
    } else {
      const float* s = r + INRF_REC_BASE;
      float m = s[0];
      label = 0;
      for (int c = 1; c < a.C; ++c) { const float v = s[c]; if (v > m) { m = v; label = c; } }   // first maximum
      if (a.o.label8) a.o.label8[p] = (uint8_t)label;
      if (a.o.entropy || a.o.entropy8) {
        // logits_2_uncertainty (trainer.py:1244): sum_c -log_softmax(x)_c * softmax(x)_c
        float z = 0.f;
        for (int c = 0; c < a.C; ++c) z += expf(s[c] - m);
        const float lz = logf(z);
        float e = 0.f;
        for (int c = 0; c < a.C; ++c) {
          const float d = s[c] - m;
          e += (lz - d) * (expf(d) / z);
        }
        if (a.o.entropy) a.o.entropy[p] = e;
        if (a.o.entropy8) a.o.entropy8[p] = to8b(e);
      }
      if (a.o.vis_label8 && a.cmap) {
        uint8_t* d = a.o.vis_label8 + 3 * p;
        const uint8_t* c = a.cmap + 3 * label;
        d[0] = c[0]; d[1] = c[1]; d[2] = c[2];
      }
    }
    if (a.o.labels64) a.o.labels64[p] = label;
    if (a.sub > 0 && (a.o.sample_pixels || a.o.sample_labels)) {
      const int y = (int)(p / a.W), x = (int)(p - (int64_t)y * a.W);
      if (y % a.sub == 0 && x % a.sub == 0) {
        const int Ws = (a.W + a.sub - 1) / a.sub;
        const int64_t q = (int64_t)(y / a.sub) * Ws + x / a.sub;
        if (a.o.sample_pixels) { float* d = a.o.sample_pixels + 3 * q; d[0] = al0; d[1] = al1; d[2] = al2; }
        if (a.o.sample_labels) a.o.sample_labels[q] = label;
      }
    }
  }
}

// edit_img = result*shading + residual (run_nerf.py:237, trainer.py:1437): two roundings (mul, add), as numpy.
__global__ void __launch_bounds__(256) k_edit_recompose(const float* __restrict__ crgb, const float* __restrict__ rec,
                                                        int64_t P, int ld, uint8_t* c8, uint8_t* edit8) {
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (int64_t)gridDim.x * blockDim.x) {
    const float c0 = crgb[3 * p], c1 = crgb[3 * p + 1], c2 = crgb[3 * p + 2];
    if (c8) { uint8_t* d = c8 + 3 * p; d[0] = to8b(c0); d[1] = to8b(c1); d[2] = to8b(c2); }
    if (edit8) {
      const float* r = rec + p * ld;
      const float sh = r[8];
      uint8_t* d = edit8 + 3 * p;
      d[0] = to8b(__fadd_rn(__fmul_rn(c0, sh), r[9]));
      d[1] = to8b(__fadd_rn(__fmul_rn(c1, sh), r[10]));
      d[2] = to8b(__fadd_rn(__fmul_rn(c2, sh), r[11]));
    }
  }
}

static int grid_for(int64_t n, int threads) {
  int64_t b = (n + threads - 1) / threads;
  const int64_t cap = 148 * 8;                 // 8 resident 256-thread CTAs per SM, grid-stride beyond that
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace inrf

extern "C" {

int inrf_frame_finish(const float* rec, int32_t H, int32_t W, int32_t rec_stride, int32_t n_classes, float acc_threshold,
                      const uint8_t* colour_map, int32_t sub_step, const InrfFramePlanes* out, void* stream) {
  using namespace inrf;
  INRF_CHECK_ARG(out != nullptr, "null plane table");
  INRF_CHECK_ARG(H >= 0 && W >= 0, "negative image size");
  INRF_CHECK_ARG(n_classes >= 0 && n_classes <= MAX_CLASSES, "n_classes out of range");
  INRF_CHECK_ARG(rec_stride >= INRF_REC_BASE + n_classes, "record stride smaller than 13 + n_classes");
  INRF_CHECK_ARG(sub_step >= 0, "negative sub_step");
  INRF_CHECK_ARG(!(out->vis_label8 && !colour_map), "vis_label8 needs a colour map");
  INRF_CHECK_ARG(!((out->vis_label8 || out->entropy || out->entropy8) && n_classes == 0), "semantic planes need n_classes > 0");
  INRF_CHECK_ARG(!((out->sample_pixels || out->sample_labels) && sub_step == 0), "sample planes need sub_step > 0");
  if ((int64_t)H * W == 0) return INRF_OK;
  INRF_CHECK_ARG(rec != nullptr, "null record");
  FrameArgs a;
  a.rec = rec; a.H = H; a.W = W; a.ld = rec_stride; a.C = n_classes; a.acc_thr = acc_threshold; a.cmap = colour_map;
  a.sub = sub_step; a.o = *out;
  k_frame_finish<<<grid_for((int64_t)H * W, 256), 256, 0, (cudaStream_t)stream>>>(a);
  INRF_LAUNCH_CHECK();
  return INRF_OK;
}

int inrf_edit_recompose(const float* cluster_rgb, const float* rec, int64_t P, int32_t rec_stride, uint8_t* c8, uint8_t* edit8,
                        void* stream) {
  using namespace inrf;
  INRF_CHECK_ARG(P >= 0, "negative pixel count");
  if (P == 0) return INRF_OK;
  INRF_CHECK_ARG(cluster_rgb != nullptr, "null cluster_rgb");
  INRF_CHECK_ARG(!(edit8 && (!rec || rec_stride < INRF_REC_BASE)), "edit8 needs the render record");
  k_edit_recompose<<<grid_for(P, 256), 256, 0, (cudaStream_t)stream>>>(cluster_rgb, rec, P, rec_stride, c8, edit8);
  INRF_LAUNCH_CHECK();
  return INRF_OK;
}

}  // extern "C"
