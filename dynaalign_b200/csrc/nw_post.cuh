// The step after the NW hot path in clusterbreak (R/clusterbreak.R:217-221) for sim_fn = similarityNW, on the device:
// histogram of (matches, alignment length) over the strict upper triangle -> exact type-7 quantile of the identities on
// the host -> the pairs that survive `sim[sim < threshold] <- 0` as an edge list.  See nw_post.cu.
#pragma once
#include "common.cuh"

namespace dyna {

// A clusterbreak recursion node over a computed NW triangle: the sequences `members[0..n_node)` of the plan (strictly
// increasing plan indices; nullptr = all n sequences).  NW pair results do not depend on the other sequences, so a
// sub-cluster's similarity matrix is a sub-matrix of the root's and nothing is re-aligned.
struct NwNode {
  const uint32_t* matches;  // the plan's slab: packed upper triangle, diagonal included, rows [row_begin, row_end)
  const uint32_t* length;
  int64_t n;                // sequences of the plan
  int64_t row_begin, row_end;
  int64_t slab_base;        // packed index of the slab's first pair
  const int32_t* members;   // device pointer or nullptr
  int64_t n_node;           // number of node sequences (n when members == nullptr)
};

// hist[matches * ldim + length] += 1 for every node pair a < b whose row lies in the slab; d_hist is zeroed first
int launch_nw_stat_hist(const NwNode& nd, int64_t mdim, int64_t ldim, unsigned long long* d_hist, cudaStream_t st);
// (matches, length) of the node's self-alignments (zeros for rows outside the slab)
int launch_nw_diag(const NwNode& nd, uint32_t* d_m, uint32_t* d_l, cudaStream_t st);
// kept pairs: matches > 0 and (double)matches / (double)length >= threshold.  Per node row counts -> exclusive offsets + total
int launch_nw_edges_count(const NwNode& nd, double threshold, unsigned long long* d_row_counts, unsigned long long* d_row_offsets,
                          unsigned long long* d_total, cudaStream_t st);
// node-local indices a < b (equal to the plan indices when members == nullptr), row-major order
int launch_nw_edges_fill(const NwNode& nd, double threshold, const unsigned long long* d_row_offsets, int32_t* d_i, int32_t* d_j,
                         uint32_t* d_m, uint32_t* d_l, cudaStream_t st);

}  // namespace dyna
