// Integer-issue probe: the roofline denominator for the NW and match-count kernels.
// MEASURED_PEAKS.json has HBM and bf16 tensor peaks only; neither bounds these kernels, so the INT32 / DPX issue
// rate is measured on the same device with dependency-free chains of the instructions the kernels are made of.
#include <cuda_fp16.h>

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


// ---- per-opcode issue probe: 8 independent dependency chains per thread, each op pinned with inline PTX
enum OpKind { OP_IADD3 = 0, OP_VIADDMNMX, OP_VIMNMX3, OP_PRMT, OP_SELP, OP_SETP_SELP, OP_LOP3, OP_SHF, OP_IMAD,
              OP_HSET2, OP_HADD2, OP_VIADDMNMX_U16X2, OP_VIBMAX_S16X2, OP_VIMAX3_S16X2, OP_SETP_PADD, OP_SHFL,
              OP_LDS, OP_POPC, OP_VIBMAX_S32, OP_HSET2_ONLY, OP_HSET2_ISUB, OP_VIMNMX3_PRMT, OP_IMAD_PRMT,
              OP_VIADDMNMX_ISETP, OP_IADD3_3IN, OP_HSET2_IMAD, OP_PRMT_SEL, OP_VIMNMX3_IMAD, OP_PMOV,
              OP_IMAD_HI, OP_DPX16_IADD, OP_DPX16_2IADD, OP_DPX16_IMADHI, OP_DPX16_IMAD, OP_IADD_IMAD, OP_DPX16_LDS128, OP_COUNT };

template <int OP>
__global__ void __launch_bounds__(kProbeThreads) op_probe_kernel(int iters, uint32_t a, uint32_t b, uint32_t* out) {
  __shared__ __align__(16) uint32_t sm[kProbeThreads + 32];
  sm[threadIdx.x] = threadIdx.x * a;
  __syncthreads();
  uint32_t x[kChains];
#pragma unroll
  for (int c = 0; c < kChains; ++c) x[c] = threadIdx.x * 2654435761u + c * 40503u + a;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 32; ++u) {
#pragma unroll
      for (int c = 0; c < kChains; ++c) {
        uint32_t& v = x[c];
        if (OP == OP_IADD3) asm volatile("add.u32 %0, %0, %1;" : "+r"(v) : "r"(a));
        if (OP == OP_VIADDMNMX) { v = (uint32_t)__viaddmax_s32((int)v, (int)a, (int)b); asm volatile("" : "+r"(v)); }
        if (OP == OP_VIMNMX3) { v = (uint32_t)__vimax3_s32((int)v, (int)a, (int)b); asm volatile("" : "+r"(v)); }
        if (OP == OP_PRMT) asm volatile("prmt.b32 %0, %0, %1, 0x5140;" : "+r"(v) : "r"(b));
        if (OP == OP_SELP) asm volatile("{.reg .pred p; setp.ne.u32 p, %1, 0; selp.b32 %0, %0, %2, p;}" : "+r"(v) : "r"(a), "r"(b));
        if (OP == OP_SETP_SELP) asm volatile("{.reg .pred p; setp.ge.s32 p, %0, %1; selp.b32 %0, %1, %2, p;}" : "+r"(v) : "r"(a), "r"(b));
        if (OP == OP_LOP3) asm volatile("xor.b32 %0, %0, %1;" : "+r"(v) : "r"(a));
        if (OP == OP_SHF) asm volatile("shf.l.wrap.b32 %0, %0, %0, 7;" : "+r"(v));
        if (OP == OP_IMAD) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(v) : "r"(a), "r"(b));
        if (OP == OP_HSET2) { v = __heq2_mask(*reinterpret_cast<__half2*>(&v), *reinterpret_cast<const __half2*>(&b)) ^ a; asm volatile("" : "+r"(v)); }
        if (OP == OP_HADD2) { __half2 h = __hadd2(*reinterpret_cast<__half2*>(&v), *reinterpret_cast<const __half2*>(&b)); v = *reinterpret_cast<uint32_t*>(&h); asm volatile("" : "+r"(v)); }
        if (OP == OP_VIADDMNMX_U16X2) { v = __viaddmin_u16x2(v, a, b); asm volatile("" : "+r"(v)); }
        if (OP == OP_VIBMAX_S16X2) { bool p0, p1; v = __vibmax_s16x2(v, a, &p0, &p1); v += (p0 ? 1u : 0u); asm volatile("" : "+r"(v)); }
        if (OP == OP_VIMAX3_S16X2) { v = __vimax3_s16x2(v, a, b); asm volatile("" : "+r"(v)); }
        if (OP == OP_SETP_PADD) asm volatile("{.reg .pred p; setp.eq.u32 p, %0, %1; @p add.u32 %0, %0, 1;}" : "+r"(v) : "r"(b));
        if (OP == OP_SHFL) { v = __shfl_up_sync(0xFFFFFFFFu, v, 1); }
        if (OP == OP_LDS) { v = sm[(v & 31u) + (threadIdx.x & ~31u)]; }
        if (OP == OP_POPC) { v = __popc(v) + a; asm volatile("" : "+r"(v)); }
        if (OP == OP_HSET2_ONLY) { v = __heq2_mask(*reinterpret_cast<__half2*>(&v), *reinterpret_cast<const __half2*>(&b)); asm volatile("" : "+r"(v)); }
        if (OP == OP_HSET2_ISUB) { uint32_t m = __heq2_mask(*reinterpret_cast<__half2*>(&v), *reinterpret_cast<const __half2*>(&b)); asm volatile("sub.u32 %0, %0, %1;" : "+r"(v) : "r"(m)); }
        if (OP == OP_VIMNMX3_PRMT) { v = (uint32_t)__vimax3_s32((int)v, (int)a, (int)b); asm volatile("prmt.b32 %0, %0, %1, 0x5140;" : "+r"(v) : "r"(b)); }
        if (OP == OP_IMAD_PRMT) { asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(v) : "r"(a), "r"(b)); asm volatile("prmt.b32 %0, %0, %1, 0x5140;" : "+r"(v) : "r"(b)); }
        if (OP == OP_VIADDMNMX_ISETP) { v = (uint32_t)__viaddmax_s32((int)v, (int)a, (int)b); asm volatile("{.reg .pred p; setp.eq.u32 p, %0, %1; @p add.u32 %0, %0, 1;}" : "+r"(v) : "r"(b)); }
        if (OP == OP_IADD3_3IN) { v = v + a + x[(c + 1) % kChains]; asm volatile("" : "+r"(v)); }
        if (OP == OP_HSET2_IMAD) { v = __heq2_mask(*reinterpret_cast<__half2*>(&v), *reinterpret_cast<const __half2*>(&b)); asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(v) : "r"(a), "r"(b)); }
        if (OP == OP_PRMT_SEL) { asm volatile("prmt.b32 %0, %0, %1, 0x5140;" : "+r"(v) : "r"(b)); asm volatile("{.reg .pred p; setp.ne.u32 p, %1, 0; selp.b32 %0, %0, %2, p;}" : "+r"(v) : "r"(a), "r"(b)); }
        if (OP == OP_VIMNMX3_IMAD) { v = (uint32_t)__vimax3_s32((int)v, (int)a, (int)b); asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(v) : "r"(a), "r"(b)); }
        if (OP == OP_PMOV) { asm volatile("{.reg .pred p; setp.eq.u32 p, %0, %1; @p mov.u32 %0, %2;}" : "+r"(v) : "r"(b), "r"(a)); }
        if (OP == OP_IMAD_HI) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(v) : "r"(b), "r"(a));
        if (OP == OP_DPX16_IADD) { v = __viaddmax_s16x2(v, a, b); asm volatile("add.u32 %0, %0, %1;" : "+r"(v) : "r"(a)); }
        if (OP == OP_DPX16_2IADD) { v = __viaddmax_s16x2(v, a, b); asm volatile("add.u32 %0, %0, %1;" : "+r"(v) : "r"(a)); asm volatile("add.u32 %0, %0, %1;" : "+r"(v) : "r"(b)); }
        if (OP == OP_DPX16_IMADHI) { v = __viaddmax_s16x2(v, a, b); asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(v) : "r"(b), "r"(a)); }
        if (OP == OP_DPX16_IMAD) { v = __viaddmax_s16x2(v, a, b); asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(v) : "r"(a), "r"(b)); }
        if (OP == OP_IADD_IMAD) { asm volatile("add.u32 %0, %0, %1;" : "+r"(v) : "r"(a)); asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(v) : "r"(a), "r"(b)); }
        if (OP == OP_DPX16_LDS128) {  // four DPX ops per 128-bit shared load (conflict-free): does the load stream cost issue or ALU slots?
          const uint4 q = *reinterpret_cast<const uint4*>(&sm[((threadIdx.x * 4u) + (v & 0u)) & (kProbeThreads - 4)]);
          v = __viaddmax_s16x2(v, q.x, b); v = __viaddmax_s16x2(v, q.y, b); v = __viaddmax_s16x2(v, q.z, b); v = __viaddmax_s16x2(v, q.w, b);
          asm volatile("" : "+r"(v));
        }
        if (OP == OP_VIBMAX_S32) { bool p0; v = (uint32_t)__vibmax_s32((int)v, (int)a, &p0); v += (p0 ? 1u : 0u); asm volatile("" : "+r"(v)); }
      }
    }
  }
  uint32_t acc = 0;
#pragma unroll
  for (int c = 0; c < kChains; ++c) acc ^= x[c];
  if (acc == 0x12345u) out[0] = acc;
}

template <int OP>
int run_op_probe(double* chain_ops_per_s, cudaStream_t st) {
  uint32_t* d_out = nullptr;
  DYNA_CUDA(cudaMalloc(&d_out, sizeof(uint32_t)));
  int sms = kNumSMsB200, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = sms * 8, iters = 1500;
  cudaEvent_t e0, e1;
  DYNA_CUDA(cudaEventCreate(&e0));
  DYNA_CUDA(cudaEventCreate(&e1));
  op_probe_kernel<OP><<<grid, kProbeThreads, 0, st>>>(30, 3u, 0x3c003c00u, d_out);
  DYNA_CUDA(cudaEventRecord(e0, st));
  op_probe_kernel<OP><<<grid, kProbeThreads, 0, st>>>(iters, 3u, 0x3c003c00u, d_out);
  DYNA_CUDA(cudaEventRecord(e1, st));
  DYNA_CUDA(cudaEventSynchronize(e1));
  float ms = 0.f;
  DYNA_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  *chain_ops_per_s = (double)grid * kProbeThreads * (double)iters * 32 * kChains / (ms * 1e-3);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d_out);
  return DYNA_OK;
}

template <int OP>
int dispatch_op(int op, double* r, cudaStream_t st) {
  if (op == OP) return run_op_probe<OP>(r, st);
  if constexpr (OP + 1 < OP_COUNT) return dispatch_op<OP + 1>(op, r, st);
  return fail(DYNA_ERR_INVALID, "dyna_probe_op: unknown op %d", op);
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

// development probe: lane-level chain steps per second of one pinned opcode pattern (see OpKind in probe.cu)
extern "C" __attribute__((visibility("default"))) int dyna_probe_op(int op, double* chain_ops_per_s, void* stream) {
  return dyna::dispatch_op<0>(op, chain_ops_per_s, static_cast<cudaStream_t>(stream));
}
