// Deferred device status: a 64-byte pinned, device-mapped record per device (see common.cuh).
#include "common.cuh"
#include <atomic>
#include <mutex>

namespace inrf {

static std::atomic<long long> g_launches{0};
void note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
long long launch_count() { return g_launches.load(std::memory_order_relaxed); }

static std::mutex g_mu;
static int* g_host[64];
static int* g_dev[64];

static int slot(int* dev_out) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return -1;
  *dev_out = dev;
  return 0;
}

int* status_flag_dev() {
  int dev;
  if (slot(&dev)) { set_error("status: no current CUDA device"); return nullptr; }
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_dev[dev] == nullptr) {
    void* h = nullptr;
    cudaError_t e = cudaHostAlloc(&h, 64, cudaHostAllocMapped | cudaHostAllocPortable);
    if (e != cudaSuccess) { cuda_fail(e, "cudaHostAlloc(status record)"); return nullptr; }
    memset(h, 0, 64);
    void* d = nullptr;
    e = cudaHostGetDevicePointer(&d, h, 0);
    if (e != cudaSuccess) { cudaFreeHost(h); cuda_fail(e, "cudaHostGetDevicePointer(status record)"); return nullptr; }
    g_host[dev] = static_cast<int*>(h);
    g_dev[dev] = static_cast<int*>(d);
  }
  return g_dev[dev];
}

static unsigned long long* g_epoch[64];

unsigned long long* rng_epoch_dev() {
  int dev;
  if (slot(&dev)) { set_error("rng epoch: no current CUDA device"); return nullptr; }
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_epoch[dev] == nullptr) {
    void* d = nullptr;
    cudaError_t e = cudaMalloc(&d, 64);
    if (e != cudaSuccess) { cuda_fail(e, "cudaMalloc(rng epoch)"); return nullptr; }
    e = cudaMemset(d, 0, 64);
    if (e != cudaSuccess) { cudaFree(d); cuda_fail(e, "cudaMemset(rng epoch)"); return nullptr; }
    g_epoch[dev] = static_cast<unsigned long long*>(d);
  }
  return g_epoch[dev];
}

__global__ void k_rng_epoch(unsigned long long* w, int bump) { *w = bump ? *w + 1ull : 0ull; }

int rng_epoch_set(int bump, cudaStream_t st) {
  unsigned long long* w = rng_epoch_dev();
  if (w == nullptr) return INRF_ECUDA;
  k_rng_epoch<<<1, 1, 0, st>>>(w, bump);
  INRF_LAUNCH_CHECK();
  return INRF_OK;
}

int status_poll() {
  int dev;
  if (slot(&dev)) return INRF_OK;
  int* h;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    h = g_host[dev];
  }
  if (h == nullptr) return INRF_OK;
  volatile int* f = h;
  const int code = f[0];
  if (code == DST_NONE) return INRF_OK;
  const int a = f[1], b = f[2], c = f[3], d = f[4], e = f[5];
  for (int i = 0; i < 16; ++i) f[i] = 0;
  switch (code) {
    case DST_WATCHDOG:
      set_error("tensor-core kernel watchdog (kernel %d): barrier %d stuck (warp %d, tile %d, cta %d); the output of that launch "
                "is invalid", e, a, b, c, d);
      return INRF_ECUDA;
    case DST_SMEM_ALIGN:
      set_error("tensor-core kernel: dynamic shared memory base 0x%x is not 1024-byte aligned", a);
      return INRF_ECUDA;
    case DST_F16_ACT:
      set_error("INRF_PREC_TC: a hidden activation reached the fp16 limit 65504 (step %d, tile %d, cta %d) and was saturated; "
                "this network needs INRF_PREC_FP32", a, b, c);
      return INRF_ERANGE;
    case DST_F16_WEIGHT:
      if (b == 1)
        set_error("INRF_PREC_TC: more than a quarter of the non-zero weights of operand block %d lie below 2^-17, where fp16 keeps "
                  "fewer than 8 significant bits; this network needs INRF_PREC_FP32", a);
      else
        set_error("INRF_PREC_TC: a weight or bias exceeds the fp16 limit 65504 (operand block %d) and was saturated at pack time; "
                  "this network needs INRF_PREC_FP32", a);
      return INRF_ERANGE;
    case DST_F16_GRAD:
      set_error("INRF_PREC_TC backward: non-finite weight gradient (item %d): the back-propagated gradients left the fp16 range; "
                "use the fp32 training path", a);
      return INRF_ERANGE;
    default:
      set_error("unknown device status record %d", code);
      return INRF_ECUDA;
  }
}

}  // namespace inrf
