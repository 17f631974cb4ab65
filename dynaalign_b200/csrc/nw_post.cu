// Threshold + sparsify for the NW path (sm_100a): what clusterbreak does right after sim_fn returns
// (R/clusterbreak.R:217-221)
//     pep.sim   <- sim_fn(pep)                                   # similarityNW: matches / alignment_length
//     threshold <- quantile(pep.sim[upper.tri(pep.sim)], thresh_p)
//     pep.sim[pep.sim < threshold] <- 0
// without the dense n x n double matrix.  Every identity is a ratio of two small integers, so the type-7 quantile is
// exactly computable from the histogram of (matches, length) and the thresholding is the IEEE comparison
// (double)matches / (double)length >= threshold evaluated per pair on the device (correctly rounded division: the
// same double the reference produces at src/pairwiseSeqAlign.cpp:311).
//
// A recursion node (R/clusterbreak.R:250-254 calls sim_fn again on every oversized cluster) is a subset of the root's
// sequences in their original order; its similarity matrix is the corresponding sub-matrix of the root triangle, so the
// node kernels index the root's slab through the member list instead of re-aligning anything.
#include "nw_post.cuh"

#include <algorithm>

namespace dyna {
namespace {

constexpr int kNwHistSmemBins = 8192;

__device__ __forceinline__ int64_t node_seq(const NwNode& nd, int64_t a) { return nd.members ? (int64_t)nd.members[a] : a; }
// slab offset of pair (i, j = i): the row's entries follow at + (j - i)
__device__ __forceinline__ int64_t row_base(const NwNode& nd, int64_t i) { return i * nd.n - i * (i - 1) / 2 - i - nd.slab_base; }

__device__ __forceinline__ bool keep_pair(uint32_t m, uint32_t l, double threshold) {
  return m > 0u && __ddiv_rn((double)m, (double)l) >= threshold;
}

// one warp per node row; lanes over the node's columns
__global__ void __launch_bounds__(256)
nw_stat_hist_kernel(NwNode nd, int64_t mdim, int64_t ldim, unsigned long long* __restrict__ hist) {
  __shared__ uint32_t sh[kNwHistSmemBins];
  const int64_t bins = mdim * ldim;
  const bool use_smem = bins <= kNwHistSmemBins;
  if (use_smem) {
    for (int b = threadIdx.x; b < (int)bins; b += blockDim.x) sh[b] = 0u;
    __syncthreads();
  }
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t a = warp; a < nd.n_node; a += nwarps) {
    const int64_t i = node_seq(nd, a);
    if (i < nd.row_begin || i >= nd.row_end) continue;
    const int64_t base = row_base(nd, i);
    for (int64_t b = a + 1 + lane; b < nd.n_node; b += 32) {
      const int64_t j = node_seq(nd, b);
      const uint32_t m = nd.matches[base + j], l = nd.length[base + j];
      if (m < (uint64_t)mdim && l < (uint64_t)ldim) {  // always true for a computed slab (matches <= max_len, length <= 2 max_len)
        const int64_t bin = (int64_t)m * ldim + l;
        if (use_smem) atomicAdd(&sh[bin], 1u);
        else atomicAdd(&hist[bin], 1ull);
      }
    }
  }
  if (use_smem) {
    __syncthreads();
    for (int b = threadIdx.x; b < (int)bins; b += blockDim.x)
      if (sh[b]) atomicAdd(&hist[b], (unsigned long long)sh[b]);
  }
}

// FILL = false counts the kept pairs per node row, FILL = true writes them in column order at the row's offset:
// the edge list is deterministic (row-major) and needs no global atomics
template <bool FILL>
__global__ void __launch_bounds__(256)
nw_edges_kernel(NwNode nd, double threshold, unsigned long long* __restrict__ row_counts,
                const unsigned long long* __restrict__ row_offsets, int32_t* __restrict__ ei, int32_t* __restrict__ ej,
                uint32_t* __restrict__ em, uint32_t* __restrict__ el) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t a = warp; a < nd.n_node; a += nwarps) {
    const int64_t i = node_seq(nd, a);
    const bool mine = i >= nd.row_begin && i < nd.row_end;
    unsigned long long kept = 0;
    if (mine) {
      const int64_t base = row_base(nd, i);
      const unsigned long long off = FILL ? row_offsets[a] : 0ull;
      for (int64_t b0 = a + 1; b0 < nd.n_node; b0 += 32) {
        const int64_t b = b0 + lane;
        uint32_t m = 0u, l = 1u;
        if (b < nd.n_node) {
          const int64_t j = node_seq(nd, b);
          m = nd.matches[base + j];
          l = nd.length[base + j];
        }
        const bool keep = b < nd.n_node && keep_pair(m, l, threshold);
        const unsigned mask = __ballot_sync(0xFFFFFFFFu, keep);
        if (FILL && keep) {
          const unsigned long long pos = off + kept + __popc(mask & ((1u << lane) - 1u));
          ei[pos] = (int32_t)a;
          ej[pos] = (int32_t)b;
          em[pos] = m;
          el[pos] = l;
        }
        kept += __popc(mask);
      }
    }
    if (!FILL && lane == 0) row_counts[a] = kept;
  }
}

// the node's diagonal (self-alignments): what graph_from_adjacency_matrix(mode = "upper") reads as self-loop weights
__global__ void __launch_bounds__(256)
nw_diag_kernel(NwNode nd, uint32_t* __restrict__ dm, uint32_t* __restrict__ dl) {
  for (int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; a < nd.n_node; a += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = node_seq(nd, a);
    const bool mine = i >= nd.row_begin && i < nd.row_end;
    const int64_t slot = row_base(nd, i) + i;
    dm[a] = mine ? nd.matches[slot] : 0u;
    dl[a] = mine ? nd.length[slot] : 0u;
  }
}

int node_grid(const NwNode& nd) {
  return (int)std::max<int64_t>(1, std::min<int64_t>((nd.n_node + 7) / 8, (int64_t)kNumSMsB200 * 16));
}

}  // namespace

int launch_nw_stat_hist(const NwNode& nd, int64_t mdim, int64_t ldim, unsigned long long* d_hist, cudaStream_t st) {
  DYNA_CUDA(cudaMemsetAsync(d_hist, 0, sizeof(unsigned long long) * (size_t)(mdim * ldim), st));
  if (nd.n_node < 2 || nd.row_end <= nd.row_begin) return DYNA_OK;
  nw_stat_hist_kernel<<<node_grid(nd), 256, 0, st>>>(nd, mdim, ldim, d_hist);
  DYNA_CUDA(cudaGetLastError());
  return DYNA_OK;
}

int launch_nw_diag(const NwNode& nd, uint32_t* d_m, uint32_t* d_l, cudaStream_t st) {
  if (nd.n_node <= 0) return DYNA_OK;
  nw_diag_kernel<<<(int)std::min<int64_t>((nd.n_node + 255) / 256, (int64_t)kNumSMsB200 * 8), 256, 0, st>>>(nd, d_m, d_l);
  DYNA_CUDA(cudaGetLastError());
  return DYNA_OK;
}

int launch_nw_edges_count(const NwNode& nd, double threshold, unsigned long long* d_row_counts, unsigned long long* d_row_offsets,
                          unsigned long long* d_total, cudaStream_t st) {
  if (nd.n_node <= 0) return DYNA_OK;
  nw_edges_kernel<false><<<node_grid(nd), 256, 0, st>>>(nd, threshold, d_row_counts, nullptr, nullptr, nullptr, nullptr, nullptr);
  DYNA_CUDA(cudaGetLastError());
  return launch_scan_rows(d_row_counts, d_row_offsets, nd.n_node, d_total, st);
}

int launch_nw_edges_fill(const NwNode& nd, double threshold, const unsigned long long* d_row_offsets, int32_t* d_i, int32_t* d_j,
                         uint32_t* d_m, uint32_t* d_l, cudaStream_t st) {
  if (nd.n_node <= 0) return DYNA_OK;
  nw_edges_kernel<true><<<node_grid(nd), 256, 0, st>>>(nd, threshold, nullptr, d_row_offsets, d_i, d_j, d_m, d_l);
  DYNA_CUDA(cudaGetLastError());
  return DYNA_OK;
}

}  // namespace dyna
