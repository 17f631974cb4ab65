// MinHash match counts by joining on signature values instead of comparing every pair (sm_100a).
//
// similarityMH's match loop (src/minHash.cpp:164-177) counts, for every pair (i, j), the hash functions h with
// sig[i][h] == sig[j][h]: n(n-1)/2 * n_hash compares whatever the data.  The same counts are the number of hash rows in
// which i and j fall into the same GROUP of equal signature values, and the relabelling step of the 16-bit match path has
// already sorted every hash row by value (mh_kernels.cu, launch_mh_relabel).  So:
//   1. per hash row, walk the sorted row: an element at rank r >= 1 inside its group is paired with the r elements before
//      it (stable sort: they are the lower sequence indices)            -> one "incidence" (i, j) per matching (pair, h)
//   2. radix-sort the incidences by pair key i * n + j, run-length encode -> (pair, match count) for every pair with
//      count >= 1, in row-major pair order; all other pairs have count 0.
// Work and memory are O(n * n_hash + incidences) instead of O(n^2 * n_hash): on BASELINE config 4 (100,000 random
// 16-mers, 2.5e12 compares) there are ~6e6 incidences.  The result is exact -- the histogram, the type-7 quantile, the
// edge list, the checksum and (through mh_sparse_densify) the u16 triangle are the ones the all-pairs kernel gives --
// but the cost is data dependent: a few large groups (many identical sequences) make the incidence list quadratic, so the
// caller sets a cap and falls back to the all-pairs kernel beyond it.
#include "mh_sparse.cuh"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_run_length_encode.cuh>
#include <cub/device/device_select.cuh>
#include <cub/iterator/counting_input_iterator.cuh>

#include <algorithm>

namespace dyna {
namespace {

typedef unsigned long long u64;

// ---- block-wide scans over 1024 threads (warp shuffles + one shared exchange)
template <class T, class Op>
__device__ __forceinline__ T block_inclusive_scan(T v, Op op, T identity, T* warp_tot /* [32] shared */, T* block_total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const T o = __shfl_up_sync(0xFFFFFFFFu, v, d);
    if (lane >= d) v = op(o, v);
  }
  __syncthreads();  // warp_tot may still be read from the previous scan
  if (lane == 31) warp_tot[warp] = v;
  __syncthreads();
  if (warp == 0) {
    T w = warp_tot[lane];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const T o = __shfl_up_sync(0xFFFFFFFFu, w, d);
      if (lane >= d) w = op(o, w);
    }
    warp_tot[lane] = w;
  }
  __syncthreads();
  if (warp > 0) v = op(warp_tot[warp - 1], v);
  *block_total = warp_tot[31];
  return v;
}

struct MaxOp {
  __device__ __forceinline__ long long operator()(long long a, long long b) const { return a > b ? a : b; }
};
struct AddOp {
  __device__ __forceinline__ u64 operator()(u64 a, u64 b) const { return a + b; }
};

// One block per hash row of the sorted rows.  r(p) = p - (first position of p's group).
// WRITE = false: row_elems[h] = #elements with r >= 1, row_pairs[h] = sum of r   (the row's incidences)
// WRITE = true : the elements with r >= 1 in row order at row_elem_off[h]: (global position, r, offset of their first incidence)
template <bool WRITE>
__global__ void __launch_bounds__(1024)
mh_groups_kernel(const uint32_t* __restrict__ keys_sorted, int64_t n, int64_t npitch, u64* __restrict__ row_elems,
                 u64* __restrict__ row_pairs, const u64* __restrict__ row_elem_off, const u64* __restrict__ row_pair_off,
                 uint32_t* __restrict__ el_pos, uint32_t* __restrict__ el_r, u64* __restrict__ el_off) {
  __shared__ long long wt_max[32];
  __shared__ u64 wt_a[32], wt_b[32];
  const int h = blockIdx.x;
  const uint32_t* keys = keys_sorted + (int64_t)h * npitch;
  long long run_start = 0;  // start of the group that is open at the end of the previous chunk
  u64 run_cnt = 0, run_sum = 0;
  const u64 ebase = WRITE ? row_elem_off[h] : 0ull, pbase = WRITE ? row_pair_off[h] : 0ull;
  for (int64_t base = 0; base < n; base += 1024) {
    const int64_t p = base + threadIdx.x;
    const bool valid = p < n;
    const bool is_start = valid && (p == 0 || keys[p] != keys[p - 1]);
    long long tot_max;
    long long start = block_inclusive_scan<long long>(is_start ? (long long)p : -1ll, MaxOp(), -1ll, wt_max, &tot_max);
    if (start < 0) start = run_start;
    const u64 r = valid ? (u64)(p - start) : 0ull;
    const u64 nonfirst = r >= 1 ? 1ull : 0ull;
    u64 tot_c, tot_s;
    const u64 inc_c = block_inclusive_scan<u64>(nonfirst, AddOp(), 0ull, wt_a, &tot_c);
    const u64 inc_s = block_inclusive_scan<u64>(r, AddOp(), 0ull, wt_b, &tot_s);
    if (WRITE && nonfirst) {
      const u64 e = ebase + run_cnt + inc_c - 1ull;
      el_pos[e] = (uint32_t)((int64_t)h * npitch + p);
      el_r[e] = (uint32_t)r;
      el_off[e] = pbase + run_sum + inc_s - r;
    }
    run_cnt += tot_c;
    run_sum += tot_s;
    if (tot_max >= 0) run_start = tot_max;
  }
  if (!WRITE && threadIdx.x == 0) {
    row_elems[h] = run_cnt;
    row_pairs[h] = run_sum;
  }
}

// one thread per incidence t: element e = last with el_off[e] <= t, partner = the (t - el_off[e])-th element of e's group
__global__ void __launch_bounds__(256)
mh_emit_pairs_kernel(const uint32_t* __restrict__ idx_sorted, const uint32_t* __restrict__ el_pos, const uint32_t* __restrict__ el_r,
                     const u64* __restrict__ el_off, u64 n_elems, u64 n_pairs, u64 n, u64 row_begin, u64 row_end, u64 sentinel,
                     u64* __restrict__ keys_out) {
  for (u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x; t < n_pairs; t += (u64)gridDim.x * blockDim.x) {
    u64 lo = 0, hi = n_elems;  // el_off is non-decreasing, el_off[0] == 0
    while (hi - lo > 1) {
      const u64 mid = (lo + hi) >> 1;
      if (el_off[mid] <= t) lo = mid;
      else hi = mid;
    }
    const uint32_t pos = el_pos[lo], r = el_r[lo];
    const u64 u = t - el_off[lo];
    const u64 j = idx_sorted[pos], i = idx_sorted[pos - r + (uint32_t)u];  // stable sort: i < j
    keys_out[t] = (i >= row_begin && i < row_end) ? i * n + j : sentinel;
  }
}

__global__ void __launch_bounds__(256)
mh_runs_hist_kernel(const uint32_t* __restrict__ counts, u64 runs, int nbins, u64* __restrict__ hist) {
  __shared__ uint32_t sh[8192];
  const bool use_smem = nbins <= 8192;
  if (use_smem) {
    for (int b = threadIdx.x; b < nbins; b += blockDim.x) sh[b] = 0u;
    __syncthreads();
  }
  for (u64 q = (u64)blockIdx.x * blockDim.x + threadIdx.x; q < runs; q += (u64)gridDim.x * blockDim.x) {
    const uint32_t c = min(counts[q], (uint32_t)(nbins - 1));
    if (use_smem) atomicAdd(&sh[c], 1u);
    else atomicAdd(&hist[c], 1ull);
  }
  if (use_smem) {
    __syncthreads();
    for (int b = threadIdx.x; b < nbins; b += blockDim.x)
      if (sh[b]) atomicAdd(&hist[b], (u64)sh[b]);
  }
}

struct CountAtLeast {
  const uint32_t* counts;
  uint32_t min_count;
  __host__ __device__ __forceinline__ bool operator()(const uint32_t& q) const { return counts[q] >= min_count; }
};

__global__ void __launch_bounds__(256)
mh_gather_edges_kernel(const u64* __restrict__ keys, const uint32_t* __restrict__ counts, const uint32_t* __restrict__ sel, u64 n_sel,
                       u64 n, int32_t* __restrict__ ei, int32_t* __restrict__ ej, uint16_t* __restrict__ ec) {
  for (u64 q = (u64)blockIdx.x * blockDim.x + threadIdx.x; q < n_sel; q += (u64)gridDim.x * blockDim.x) {
    const uint32_t s = sel[q];
    const u64 key = keys[s];
    ei[q] = (int32_t)(key / n);
    ej[q] = (int32_t)(key % n);
    ec[q] = (uint16_t)counts[s];
  }
}

__global__ void __launch_bounds__(256)
mh_runs_checksum_kernel(const u64* __restrict__ keys, const uint32_t* __restrict__ counts, u64 runs, u64 n, u64* __restrict__ sum) {
  u64 acc = 0;
  for (u64 q = (u64)blockIdx.x * blockDim.x + threadIdx.x; q < runs; q += (u64)gridDim.x * blockDim.x) {
    const u64 key = keys[q], i = key / n, j = key % n;
    acc += (u64)counts[q] * checksum_weight(i * n - i * (i + 1) / 2 + (j - i - 1));  // global strict-triangle index
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, o);
  if ((threadIdx.x & 31) == 0 && acc) atomicAdd(sum, acc);
}

__global__ void __launch_bounds__(256)
mh_scatter_counts_kernel(const u64* __restrict__ keys, const uint32_t* __restrict__ counts, u64 runs, u64 n, u64 slab_base,
                         uint16_t* __restrict__ dense) {
  for (u64 q = (u64)blockIdx.x * blockDim.x + threadIdx.x; q < runs; q += (u64)gridDim.x * blockDim.x) {
    const u64 key = keys[q], i = key / n, j = key % n;
    dense[i * n - i * (i + 1) / 2 + (j - i - 1) - slab_base] = (uint16_t)counts[q];
  }
}

int grid_for(u64 items) { return (int)std::max<u64>(1, std::min<u64>((items + 255) / 256, (u64)kNumSMsB200 * 16)); }

}  // namespace

int mh_sparse_count_incidences(const uint32_t* d_keys_sorted, int64_t n, int n_hash, int64_t npitch, u64* d_row_elems,
                               u64* d_row_pairs, u64* d_row_elem_off, u64* d_row_pair_off, u64* d_totals /* [2] */,
                               cudaStream_t st) {
  mh_groups_kernel<false><<<n_hash, 1024, 0, st>>>(d_keys_sorted, n, npitch, d_row_elems, d_row_pairs, nullptr, nullptr, nullptr,
                                                    nullptr, nullptr);
  DYNA_CUDA(cudaGetLastError());
  DYNA_TRY(launch_scan_rows(d_row_elems, d_row_elem_off, n_hash, d_totals, st));
  DYNA_TRY(launch_scan_rows(d_row_pairs, d_row_pair_off, n_hash, d_totals + 1, st));
  return DYNA_OK;
}

int mh_sparse_emit(const uint32_t* d_keys_sorted, const uint32_t* d_idx_sorted, int64_t n, int n_hash, int64_t npitch,
                   const u64* d_row_elem_off, const u64* d_row_pair_off, uint32_t* d_el_pos, uint32_t* d_el_r, u64* d_el_off,
                   int64_t n_elems, int64_t n_pairs, int64_t row_begin, int64_t row_end, u64* d_pair_keys, cudaStream_t st) {
  mh_groups_kernel<true><<<n_hash, 1024, 0, st>>>(d_keys_sorted, n, npitch, nullptr, nullptr, d_row_elem_off, d_row_pair_off,
                                                   d_el_pos, d_el_r, d_el_off);
  DYNA_CUDA(cudaGetLastError());
  mh_emit_pairs_kernel<<<grid_for((u64)n_pairs), 256, 0, st>>>(d_idx_sorted, d_el_pos, d_el_r, d_el_off, (u64)n_elems, (u64)n_pairs,
                                                                (u64)n, (u64)row_begin, (u64)row_end, (u64)n * (u64)n, d_pair_keys);
  DYNA_CUDA(cudaGetLastError());
  return DYNA_OK;
}

static int key_bits(int64_t n) {
  const u64 top = (u64)n * (u64)n;  // the sentinel, the largest key
  int b = 1;
  while (b < 64 && (top >> b) != 0) ++b;
  return b;
}

size_t mh_sparse_sort_temp_bytes(int64_t n_pairs, int64_t n) {
  size_t a = 0, b = 0;
  cub::DeviceRadixSort::SortKeys(nullptr, a, (const u64*)nullptr, (u64*)nullptr, n_pairs, 0, key_bits(n));
  cub::DeviceRunLengthEncode::Encode(nullptr, b, (const u64*)nullptr, (u64*)nullptr, (uint32_t*)nullptr, (u64*)nullptr, n_pairs);
  return std::max(a, b) + 256;
}

int mh_sparse_sort_encode(u64* d_pair_keys, u64* d_sorted, int64_t n_pairs, int64_t n, void* d_temp, size_t temp_bytes,
                          u64* d_run_keys, uint32_t* d_run_counts, u64* d_num_runs, cudaStream_t st) {
  size_t bytes = temp_bytes;
  cudaError_t e = cub::DeviceRadixSort::SortKeys(d_temp, bytes, d_pair_keys, d_sorted, n_pairs, 0, key_bits(n), st);
  if (e != cudaSuccess) return fail(DYNA_ERR_CUDA, "DynaAlign CUDA: incidence sort failed: %s", cudaGetErrorString(e));
  bytes = temp_bytes;
  e = cub::DeviceRunLengthEncode::Encode(d_temp, bytes, d_sorted, d_run_keys, d_run_counts, d_num_runs, n_pairs, st);
  if (e != cudaSuccess) return fail(DYNA_ERR_CUDA, "DynaAlign CUDA: run-length encode failed: %s", cudaGetErrorString(e));
  return DYNA_OK;
}

int mh_sparse_histogram(const uint32_t* d_run_counts, int64_t runs, int n_hash, u64* d_hist, cudaStream_t st) {
  DYNA_CUDA(cudaMemsetAsync(d_hist, 0, sizeof(u64) * (size_t)(n_hash + 1), st));
  if (runs <= 0) return DYNA_OK;
  mh_runs_hist_kernel<<<grid_for((u64)runs), 256, 0, st>>>(d_run_counts, (u64)runs, n_hash + 1, d_hist);
  DYNA_CUDA(cudaGetLastError());
  return DYNA_OK;
}

size_t mh_sparse_select_temp_bytes(int64_t runs) {
  size_t a = 0;
  cub::CountingInputIterator<uint32_t> it(0u);
  cub::DeviceSelect::If(nullptr, a, it, (uint32_t*)nullptr, (u64*)nullptr, runs, CountAtLeast{nullptr, 0u});
  return a + 256;
}

int mh_sparse_select(const uint32_t* d_run_counts, int64_t runs, uint32_t min_count, void* d_temp, size_t temp_bytes,
                     uint32_t* d_sel, u64* d_num_sel, cudaStream_t st) {
  size_t bytes = temp_bytes;
  cub::CountingInputIterator<uint32_t> it(0u);
  cudaError_t e = cub::DeviceSelect::If(d_temp, bytes, it, d_sel, d_num_sel, runs, CountAtLeast{d_run_counts, min_count}, st);
  if (e != cudaSuccess) return fail(DYNA_ERR_CUDA, "DynaAlign CUDA: edge selection failed: %s", cudaGetErrorString(e));
  return DYNA_OK;
}

int mh_sparse_gather_edges(const u64* d_run_keys, const uint32_t* d_run_counts, const uint32_t* d_sel, int64_t n_sel, int64_t n,
                           int32_t* d_i, int32_t* d_j, uint16_t* d_c, cudaStream_t st) {
  if (n_sel <= 0) return DYNA_OK;
  mh_gather_edges_kernel<<<grid_for((u64)n_sel), 256, 0, st>>>(d_run_keys, d_run_counts, d_sel, (u64)n_sel, (u64)n, d_i, d_j, d_c);
  DYNA_CUDA(cudaGetLastError());
  return DYNA_OK;
}

int mh_sparse_checksum(const u64* d_run_keys, const uint32_t* d_run_counts, int64_t runs, int64_t n, u64* d_sum, cudaStream_t st) {
  DYNA_CUDA(cudaMemsetAsync(d_sum, 0, sizeof(u64), st));
  if (runs <= 0) return DYNA_OK;
  mh_runs_checksum_kernel<<<grid_for((u64)runs), 256, 0, st>>>(d_run_keys, d_run_counts, (u64)runs, (u64)n, d_sum);
  DYNA_CUDA(cudaGetLastError());
  return DYNA_OK;
}

int mh_sparse_densify(const u64* d_run_keys, const uint32_t* d_run_counts, int64_t runs, int64_t n, int64_t slab_base,
                      int64_t slab_pairs, uint16_t* d_dense, cudaStream_t st) {
  DYNA_CUDA(cudaMemsetAsync(d_dense, 0, sizeof(uint16_t) * (size_t)slab_pairs, st));
  if (runs <= 0) return DYNA_OK;
  mh_scatter_counts_kernel<<<grid_for((u64)runs), 256, 0, st>>>(d_run_keys, d_run_counts, (u64)runs, (u64)n, (u64)slab_base, d_dense);
  DYNA_CUDA(cudaGetLastError());
  return DYNA_OK;
}

}  // namespace dyna
