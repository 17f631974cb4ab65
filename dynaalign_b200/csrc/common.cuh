// Shared host/device helpers for the dynaalign_b200 CUDA library (sm_100a only).
#pragma once

#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>

#include "dynaalign_b200.h"

namespace dyna {

// ---- thread-local error text, returned by dyna_last_error()
inline std::string& err_slot() {
  static thread_local std::string e;
  return e;
}
inline int& err_code_slot() {
  static thread_local int c = 0;
  return c;
}
inline int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  err_slot() = buf;
  err_code_slot() = code;
  return code;
}

#define DYNA_CUDA(call)                                                                              \
  do {                                                                                               \
    cudaError_t e__ = (call);                                                                        \
    if (e__ != cudaSuccess)                                                                          \
      return ::dyna::fail(DYNA_ERR_CUDA, "DynaAlign CUDA: %s failed: %s (%s:%d)", #call,             \
                          cudaGetErrorString(e__), __FILE__, __LINE__);                              \
  } while (0)

#define DYNA_TRY(expr)            \
  do {                            \
    int rc__ = (expr);            \
    if (rc__ != DYNA_OK) return rc__; \
  } while (0)

// Device memory comes from the stream-ordered allocator with the pool's release threshold lifted, so the buffers of
// one call are recycled by the next (clusterbreak invokes sim_fn once per recursion node; plain cudaMalloc/cudaFree of
// the multi-GB result buffers was measured at up to 1 s per call).  All allocation and release is ordered on the
// legacy default stream; the owners (plans, entry points) synchronise the stream their work ran on BEFORE releasing,
// so a block handed to the next allocation is never still in use, whatever kind of stream the caller passed; the first
// use of a fresh block on a non-blocking caller stream is ordered behind the allocation (WorkStreamGuard below).
inline void dev_pool_keep_cached(int device) {
  static bool done[64] = {false};
  if (device < 0 || device >= 64 || done[device]) return;
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
    unsigned long long threshold = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold);
  }
  done[device] = true;
}

// The stream the current entry point runs its work on (thread-local, set by WorkStreamGuard for the duration of the
// call).  Allocation is ordered on the LEGACY stream; a blocking or NULL caller stream is ordered with it by definition,
// a cudaStreamNonBlocking one is not -- a pool that has to map fresh memory does so in stream order, and work submitted to
// an unrelated stream can reach the pages first.  For such a stream every DevBuf::alloc records an event on the legacy
// stream behind the allocation and makes the work stream wait for it.  (Nothing changes for NULL / legacy callers.)
inline cudaStream_t& work_stream_slot() {
  static thread_local cudaStream_t s = nullptr;
  return s;
}
struct WorkStreamGuard {
  cudaStream_t prev;
  explicit WorkStreamGuard(cudaStream_t st) : prev(work_stream_slot()) { work_stream_slot() = st; }
  ~WorkStreamGuard() { work_stream_slot() = prev; }
  WorkStreamGuard(const WorkStreamGuard&) = delete;
  WorkStreamGuard& operator=(const WorkStreamGuard&) = delete;
};
inline void order_alloc_before_work_stream() {
  cudaStream_t st = work_stream_slot();
  if (st == nullptr || st == cudaStreamLegacy) return;
  struct Ev {
    cudaEvent_t ev = nullptr;
    int dev = -1;
    ~Ev() { if (ev) cudaEventDestroy(ev); }
  };
  static thread_local Ev e;
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess) return;
  if (!e.ev || e.dev != dev) {
    if (e.ev) cudaEventDestroy(e.ev);
    e.ev = nullptr;
    if (cudaEventCreateWithFlags(&e.ev, cudaEventDisableTiming) != cudaSuccess) { e.ev = nullptr; return; }
    e.dev = dev;
  }
  if (cudaEventRecord(e.ev, cudaStreamLegacy) == cudaSuccess) cudaStreamWaitEvent(st, e.ev, 0);
}

// RAII device buffer on the current device
template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  DevBuf() {}
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  int alloc(size_t count) {
    release();
    n = count;
    if (count == 0) count = 1;
    DYNA_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&p), count * sizeof(T), 0));
    order_alloc_before_work_stream();
    return DYNA_OK;
  }
  void release() {
    if (p) cudaFreeAsync(p, 0);
    p = nullptr;
    n = 0;
  }
};

// offsets[0..n] must start at 0 or above and never decrease (they come from a foreign caller: R, ctypes, ...)
inline int check_offsets(const int64_t* offsets, int64_t n, const char* who) {
  if (n < 0) return fail(DYNA_ERR_INVALID, "%s: negative sequence count", who);
  if (n > 0 && !offsets) return fail(DYNA_ERR_INVALID, "%s: null offsets", who);
  if (n > 0 && offsets[0] < 0) return fail(DYNA_ERR_INVALID, "%s: offsets[0] = %lld is negative", who, (long long)offsets[0]);
  for (int64_t i = 0; i < n; ++i)
    if (offsets[i + 1] < offsets[i])
      return fail(DYNA_ERR_INVALID, "%s: offsets must be non-decreasing (offsets[%lld] = %lld > offsets[%lld] = %lld)", who,
                  (long long)i, (long long)offsets[i], (long long)(i + 1), (long long)offsets[i + 1]);
  return DYNA_OK;
}

// Will `bytes` more device memory fit?  Counts what the driver reports free plus what the (never trimmed) default pool
// holds cached.  Lets the R-facing calls refuse an n x n double matrix that cannot exist with DYNA_ERR_UNSUPPORTED and
// the byte count instead of a raw out-of-memory error from the middle of the run.
inline int check_device_fits(double bytes, int device, const char* what) {
  size_t free_b = 0, total_b = 0;
  if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) {
    cudaGetLastError();
    return DYNA_OK;  // cannot tell: let the allocation itself decide
  }
  double avail = (double)free_b;
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
    unsigned long long reserved = 0, used = 0;
    if (cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &reserved) == cudaSuccess &&
        cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemCurrent, &used) == cudaSuccess && reserved > used)
      avail += (double)(reserved - used);
  }
  if (bytes > avail)
    return fail(DYNA_ERR_UNSUPPORTED,
                "%s needs %.0f bytes of device memory but only %.0f are available on device %d; use the row-range / "
                "edge-list entry points for inputs of this size",
                what, bytes, avail, device);
  return DYNA_OK;
}

inline int64_t tri_strict_index(int64_t n, int64_t i, int64_t j) { return i * n - i * (i + 1) / 2 + (j - i - 1); }
inline int64_t tri_strict_rows(int64_t n, int64_t r) { return r * n - r * (r + 1) / 2; }  // pairs in rows [0,r)
inline int64_t tri_diag_rows(int64_t n, int64_t r) { return r * n - r * (r - 1) / 2; }    // pairs incl. diagonal in rows [0,r)

constexpr int kNumSMsB200 = 148;

// weight of packed-triangle position k in the position-weighted checksums (gather.cu, mh_sparse.cu)
__host__ __device__ inline unsigned long long checksum_weight(unsigned long long k) {
  unsigned long long w = (k + 1ull) * 0x9E3779B97F4A7C15ull;
  return w ^ (w >> 31);
}

// exclusive scan of per-row counts (one block; rows <= a few hundred thousand) + grand total; defined in mh_kernels.cu
int launch_scan_rows(const unsigned long long* d_in, unsigned long long* d_out, int64_t rows, unsigned long long* d_total,
                     cudaStream_t st);

}  // namespace dyna
