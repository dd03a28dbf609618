// Error plumbing, device queries and the FP32 peak micro-benchmark of libmdk.
#include "mdk_common.cuh"

#include <cstdarg>
#include <cstring>

namespace mdk {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
    return 148;
  return n;
}

// 16 independent FMA chains per thread; packed variant issues FFMA2 (two fp32 FMAs per
// lane per instruction, the only way to reach the 128 lanes/clk/SM FP32 rate on sm_100).
template <bool PACKED>
__global__ void __launch_bounds__(256) peak_fp32_kernel(float* out, int iters, float a, float b) {
  float2 acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f + i, blockIdx.x * 1e-3f - i);
  const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (PACKED) {
        acc[i] = __ffma2_rn(acc[i], a2, b2);
      } else {
        acc[i].x = fmaf(acc[i].x, a, b);
        acc[i].y = fmaf(acc[i].y, a, b);
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i].x + acc[i].y;
  if (s == 12345.678f) out[0] = s;  // keep the chains alive
}

// Result read-back without a copy engine: the SMs store a (small) device array into page-locked
// host memory that is mapped into the device address space.
__global__ void store_mapped_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst,
                                    long long n16, const unsigned char* __restrict__ src_tail,
                                    unsigned char* __restrict__ dst_tail, int tail) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n16;
       i += (long long)gridDim.x * blockDim.x)
    dst[i] = src[i];
  if (blockIdx.x == 0 && (int)threadIdx.x < tail) dst_tail[threadIdx.x] = src_tail[threadIdx.x];
}

}  // namespace mdk

using namespace mdk;

extern "C" int mdk_store_mapped(const void* src, void* dst_host_mapped, long long nbytes,
                                mdk_stream_t stream) {
  MDK_CHECK_ARG(nbytes >= 0, "store_mapped: negative size");
  if (nbytes == 0) return MDK_OK;
  MDK_CHECK_ARG(src && dst_host_mapped, "store_mapped: null pointer");
  MDK_CHECK_ARG(reinterpret_cast<uintptr_t>(src) % 16 == 0 &&
                    reinterpret_cast<uintptr_t>(dst_host_mapped) % 16 == 0,
                "store_mapped: pointers must be 16-byte aligned");
  const long long n16 = nbytes / 16;
  const int tail = (int)(nbytes - 16 * n16);
  long long blocks = (n16 + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > 4 * sm_count()) blocks = 4 * sm_count();
  const unsigned char* s8 = static_cast<const unsigned char*>(src) + 16 * n16;
  unsigned char* d8 = static_cast<unsigned char*>(dst_host_mapped) + 16 * n16;
  store_mapped_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(
      static_cast<const uint4*>(src), static_cast<uint4*>(dst_host_mapped), n16, s8, d8, tail);
  MDK_LAUNCH_CHECK();
  return MDK_OK;
}

extern "C" int mdk_version(void) { return MDK_VERSION; }
extern "C" const char* mdk_last_error(void) { return g_err; }
extern "C" int mdk_sm_count(void) { return sm_count(); }

extern "C" int mdk_peak_fp32(int packed, int iters, double* tflops) {
  MDK_CHECK_ARG(tflops && iters > 0, "peak_fp32: bad argument");
  float* d = nullptr;
  MDK_CUDA(cudaMalloc(&d, 16));
  const int grid = sm_count() * 8, block = 256;
  cudaEvent_t e0, e1;
  MDK_CUDA(cudaEventCreate(&e0));
  MDK_CUDA(cudaEventCreate(&e1));
  double best = 0.0;
  for (int rep = 0; rep < 4; ++rep) {
    MDK_CUDA(cudaEventRecord(e0));
    if (packed)
      peak_fp32_kernel<true><<<grid, block>>>(d, iters, 0.999f, 1e-3f);
    else
      peak_fp32_kernel<false><<<grid, block>>>(d, iters, 0.999f, 1e-3f);
    MDK_CUDA(cudaEventRecord(e1));
    MDK_CUDA(cudaEventSynchronize(e1));
    MDK_LAUNCH_CHECK();
    float ms = 0.f;
    MDK_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    const double flops = 2.0 * 16.0 * (double)iters * (double)grid * block;
    const double tf = flops / (ms * 1e-3) * 1e-12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  *tflops = best;
  return MDK_OK;
}
