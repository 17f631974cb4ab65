// Integer-issue probe: the roofline denominator for the NW and match-count kernels.
// MEASURED_PEAKS.json has HBM and bf16 tensor peaks only; neither bounds these kernels, so the INT32 / DPX issue
// rate is measured on the same device with dependency-free chains of the instructions the kernels are made of.
#include "common.cuh"

namespace dyna {
namespace {

constexpr int kChains = 8;       // independent dependency chains per thread
constexpr int kInner = 64;       // unrolled ops per chain per outer iteration
constexpr int kProbeThreads = 256;

template <int KIND>
__global__ void __launch_bounds__(kProbeThreads) probe_kernel(int iters, int a, int b, int* out) {
  int x[kChains];
#pragma unroll
  for (int c = 0; c < kChains; ++c) x[c] = threadIdx.x + c * 17 + a;
  uint32_t cnt = 0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < kInner; ++u) {
#pragma unroll
      for (int c = 0; c < kChains; ++c) {
        if (KIND == 0) {  // IADD3
          x[c] = x[c] + a + b;
          asm volatile("" : "+r"(x[c]));
        } else if (KIND == 1) {  // VIADDMNMX
          x[c] = __viaddmax_s32(x[c], a, b);
          asm volatile("" : "+r"(x[c]));
        } else if (KIND == 2) {  // VIMNMX3
          x[c] = __vimax3_s32(x[c], a, b + u);
          asm volatile("" : "+r"(x[c]));
        } else if (KIND == 3) {  // NW-like mix: add, max3, 2 viaddmax, compare+select
          const int m = x[c] + a;
          const int h = __vimax3_s32(m, b, x[(c + 1) % kChains]);
          x[c] = __viaddmax_s32(h, a, x[c]);
          cnt += (m == h) ? 1u : 0u;
        } else {  // ISETP + predicated add (the naive equality count)
          cnt += (x[c] == b + u) ? 1u : 0u;
          x[c] += a;
        }
      }
    }
  }
  int acc = (int)cnt;
#pragma unroll
  for (int c = 0; c < kChains; ++c) acc ^= x[c];
  if (acc == 0x7fffffff) out[0] = acc;  // keep the chains alive
}

template <int KIND>
int run_probe(double ops_per_inner, double* lane_ops_per_s, double* elapsed_ms, cudaStream_t st) {
  int* d_out = nullptr;
  DYNA_CUDA(cudaMalloc(&d_out, sizeof(int)));
  int sms = kNumSMsB200;
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = sms * 8;
  const int iters = 2000;
  cudaEvent_t e0, e1;
  DYNA_CUDA(cudaEventCreate(&e0));
  DYNA_CUDA(cudaEventCreate(&e1));
  probe_kernel<KIND><<<grid, kProbeThreads, 0, st>>>(50, 1, -3, d_out);  // warm-up
  DYNA_CUDA(cudaEventRecord(e0, st));
  probe_kernel<KIND><<<grid, kProbeThreads, 0, st>>>(iters, 1, -3, d_out);
  DYNA_CUDA(cudaEventRecord(e1, st));
  DYNA_CUDA(cudaEventSynchronize(e1));
  float ms = 0.f;
  DYNA_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  const double lane_ops = (double)grid * kProbeThreads * (double)iters * kInner * kChains * ops_per_inner;
  if (lane_ops_per_s) *lane_ops_per_s = lane_ops / (ms * 1e-3);
  if (elapsed_ms) *elapsed_ms = ms;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d_out);
  return DYNA_OK;
}

}  // namespace
}  // namespace dyna

extern "C" int dyna_probe_int_issue(int kind, double* lane_ops_per_s, double* elapsed_ms, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (kind) {
    case 0: return dyna::run_probe<0>(1.0, lane_ops_per_s, elapsed_ms, st);
    case 1: return dyna::run_probe<1>(1.0, lane_ops_per_s, elapsed_ms, st);
    case 2: return dyna::run_probe<2>(1.0, lane_ops_per_s, elapsed_ms, st);
    case 3: return dyna::run_probe<3>(5.0, lane_ops_per_s, elapsed_ms, st);  // IADD, VIMNMX3, VIADDMNMX, ISETP, IADD(pred)
    case 4: return dyna::run_probe<4>(3.0, lane_ops_per_s, elapsed_ms, st);  // ISETP, IADD(pred), IADD
    default: return dyna::fail(DYNA_ERR_INVALID, "dyna_probe_int_issue: unknown kind %d", kind);
  }
}
