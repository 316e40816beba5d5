// oxg_peak.cu — measures the roofline denominator of the batched k-NN / validity kernels:
// the issue rate of UNFUSED FP64 adds and multiplies (the reference never contracts a*b+c,
// so DFMA's two-flops-per-issue datasheet figure is not reachable by this workload).
#include <algorithm>

#include "oxg_internal.cuh"

namespace oxg {

constexpr int PEAK_CHAINS = 8;

__global__ void __launch_bounds__(256) fp64_peak_kernel(double* __restrict__ out, double a, double b,
                                                        int iters) {
  double x[PEAK_CHAINS];
#pragma unroll
  for (int c = 0; c < PEAK_CHAINS; ++c) x[c] = a * (double)(threadIdx.x + c + 1);
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < PEAK_CHAINS; ++c) {
      x[c] = __dmul_rn(x[c], b);
      x[c] = __dadd_rn(x[c], a);
    }
  }
  double s = 0.0;
#pragma unroll
  for (int c = 0; c < PEAK_CHAINS; ++c) s = __dadd_rn(s, x[c]);
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// FP32 pipe rate with the pre-filter's instruction mix: one packed subtract (FADD2) and one
// packed FMA (FFMA2) per pair of lanes, independent chains.
__global__ void __launch_bounds__(256) fp32_peak_kernel(float2* __restrict__ out, float a, float b,
                                                        int iters) {
  float2 x[PEAK_CHAINS];
#pragma unroll
  for (int c = 0; c < PEAK_CHAINS; ++c)
    x[c] = make_float2(a * (float)(threadIdx.x + c + 1), a * (float)(threadIdx.x + c + 2));
  const float2 bb = make_float2(b, b), aa = make_float2(a, a);
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < PEAK_CHAINS; ++c) {
      float2 d = __fadd2_rn(x[c], aa);
      x[c] = __ffma2_rn(d, bb, x[c]);
    }
  }
  float2 s = make_float2(0.f, 0.f);
#pragma unroll
  for (int c = 0; c < PEAK_CHAINS; ++c) s = __fadd2_rn(s, x[c]);
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// L2 read bandwidth: every CTA streams through the same L2-resident buffer with coalesced 16-byte
// loads that bypass L1 (ld.global.cg), the access shape of a warp reading a leaf of the
// bounding-box hierarchy.  The buffer (48 MB) fits the 126 MB L2 and is touched once before timing.
__global__ void __launch_bounds__(256) l2_peak_kernel(const float4* __restrict__ buf, size_t n_vec, int passes,
                                                      float4* __restrict__ out) {
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (int p = 0; p < passes; ++p) {
    // a different starting offset per pass and CTA, so that CTAs do not march in lock step
    size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x + (size_t)p * 7919u * blockDim.x) % n_vec;
    for (size_t t = 0; t < n_vec / stride; ++t) {
      const float4 v = __ldcg(buf + i);
      acc.x += v.x;
      acc.y += v.y;
      acc.z += v.z;
      acc.w += v.w;
      i += stride;
      if (i >= n_vec) i -= n_vec;
    }
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

}  // namespace oxg

using namespace oxg;

// Bytes per second read from L2 by all SMs (the store and its hierarchy are L2-resident by design,
// so this — not HBM — is the memory roofline of the batched neighbour kernels).
extern "C" int oxg_l2_peak(double* out_bytes_per_s, double* out_ms) {
  OXG_API_BEGIN
  if (!out_bytes_per_s) {
    set_error("out_bytes_per_s is NULL");
    return OXG_E_NULL;
  }
  const size_t bytes = 48ull << 20, n_vec = bytes / sizeof(float4);
  const int blocks = 148 * 8, threads = 256, passes = 128;
  float4 *d_buf = nullptr, *d_out = nullptr;
  OXG_CUDA_TRY(cudaMalloc((void**)&d_buf, bytes));
  if (cudaMalloc((void**)&d_out, sizeof(float4) * blocks * threads) != cudaSuccess) {
    cudaFree(d_buf);
    set_error("cudaMalloc failed");
    return OXG_E_CUDA;
  }
  cudaMemset(d_buf, 0, bytes);
  cudaEvent_t e0, e1;
  OXG_CUDA_TRY(cudaEventCreate(&e0));
  OXG_CUDA_TRY(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    OXG_CUDA_TRY(cudaEventRecord(e0, 0));
    l2_peak_kernel<<<blocks, threads>>>(d_buf, n_vec, passes, d_out);
    OXG_LAUNCH_CHECK();
    OXG_CUDA_TRY(cudaEventRecord(e1, 0));
    OXG_CUDA_TRY(cudaEventSynchronize(e1));
    float ms = 0.f;
    OXG_CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0) best = std::min(best, ms);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d_buf);
  cudaFree(d_out);
  const size_t per_thread = n_vec / ((size_t)blocks * threads);
  const double total = (double)per_thread * blocks * threads * passes * sizeof(float4);
  *out_bytes_per_s = total / (best * 1e-3);
  if (out_ms) *out_ms = best;
  return OXG_OK;
  OXG_API_END
}


// Lane-instructions per second of the FADD2+FFMA2 mix (a packed instruction counts as two
// lane-instructions; an FMA counts once).  Flop rate of the mix = 1.5 x this figure.
extern "C" int oxg_fp32_peak(double* out_ops_per_s, double* out_ms) {
  OXG_API_BEGIN
  if (!out_ops_per_s) {
    set_error("out_ops_per_s is NULL");
    return OXG_E_NULL;
  }
  const int blocks = 148 * 8, threads = 256, iters = 16384;
  float2* d_out = nullptr;
  OXG_CUDA_TRY(cudaMalloc((void**)&d_out, sizeof(float2) * blocks * threads));
  cudaEvent_t e0, e1;
  OXG_CUDA_TRY(cudaEventCreate(&e0));
  OXG_CUDA_TRY(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    OXG_CUDA_TRY(cudaEventRecord(e0, 0));
    fp32_peak_kernel<<<blocks, threads>>>(d_out, 1e-9f, 1.0000001f, iters);
    OXG_LAUNCH_CHECK();
    OXG_CUDA_TRY(cudaEventRecord(e1, 0));
    OXG_CUDA_TRY(cudaEventSynchronize(e1));
    float ms = 0.f;
    OXG_CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0) best = std::min(best, ms);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d_out);
  const double ops = (double)blocks * threads * (double)iters * PEAK_CHAINS * 2.0 * 2.0;
  *out_ops_per_s = ops / (best * 1e-3);
  if (out_ms) *out_ms = best;
  return OXG_OK;
  OXG_API_END
}

extern "C" int oxg_fp64_peak(double* out_ops_per_s, double* out_ms) {
  OXG_API_BEGIN
  if (!out_ops_per_s) {
    set_error("out_ops_per_s is NULL");
    return OXG_E_NULL;
  }
  const int blocks = 148 * 8, threads = 256, iters = 16384;
  double* d_out = nullptr;
  OXG_CUDA_TRY(cudaMalloc((void**)&d_out, sizeof(double) * blocks * threads));
  cudaEvent_t e0, e1;
  OXG_CUDA_TRY(cudaEventCreate(&e0));
  OXG_CUDA_TRY(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    OXG_CUDA_TRY(cudaEventRecord(e0, 0));
    fp64_peak_kernel<<<blocks, threads>>>(d_out, 1e-9, 1.0000001, iters);
    OXG_LAUNCH_CHECK();
    OXG_CUDA_TRY(cudaEventRecord(e1, 0));
    OXG_CUDA_TRY(cudaEventSynchronize(e1));
    float ms = 0.f;
    OXG_CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0) best = std::min(best, ms);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d_out);
  const double ops = (double)blocks * threads * (double)iters * PEAK_CHAINS * 2.0;
  *out_ops_per_s = ops / (best * 1e-3);
  if (out_ms) *out_ms = best;
  return OXG_OK;
  OXG_API_END
}
