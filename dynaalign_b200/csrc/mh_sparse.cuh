// MinHash match counts by joining the sorted hash rows on equal signature values (sm_100a).  See mh_sparse.cu.
#pragma once
#include "common.cuh"

namespace dyna {

// Step 1a: per hash row of the sorted rows (keys_sorted[h][0..n), pitch npitch) the number of elements that have a
// partner before them and the number of incidences; exclusive offsets over the rows; d_totals = {elements, incidences}
int mh_sparse_count_incidences(const uint32_t* d_keys_sorted, int64_t n, int n_hash, int64_t npitch, unsigned long long* d_row_elems,
                               unsigned long long* d_row_pairs, unsigned long long* d_row_elem_off,
                               unsigned long long* d_row_pair_off, unsigned long long* d_totals, cudaStream_t st);
// Step 1b: the incidences as pair keys i * n + j (i < j); pairs whose row i is outside [row_begin, row_end) become n * n
int mh_sparse_emit(const uint32_t* d_keys_sorted, const uint32_t* d_idx_sorted, int64_t n, int n_hash, int64_t npitch,
                   const unsigned long long* d_row_elem_off, const unsigned long long* d_row_pair_off, uint32_t* d_el_pos,
                   uint32_t* d_el_r, unsigned long long* d_el_off, int64_t n_elems, int64_t n_pairs, int64_t row_begin,
                   int64_t row_end, unsigned long long* d_pair_keys, cudaStream_t st);
// Step 2: sort + run-length encode -> (pair key, match count) in row-major pair order; *d_num_runs counts the runs
// (the last one is the n * n sentinel if any pair was outside the row range)
size_t mh_sparse_sort_temp_bytes(int64_t n_pairs, int64_t n);
int mh_sparse_sort_encode(unsigned long long* d_pair_keys, unsigned long long* d_sorted, int64_t n_pairs, int64_t n, void* d_temp,
                          size_t temp_bytes, unsigned long long* d_run_keys, uint32_t* d_run_counts,
                          unsigned long long* d_num_runs, cudaStream_t st);
// consumers of the runs: histogram of the counts (bin 0 is left 0: the caller knows how many pairs have no run), the runs
// with count >= min_count as an edge list, the position-weighted checksum, and the dense u16 triangle
int mh_sparse_histogram(const uint32_t* d_run_counts, int64_t runs, int n_hash, unsigned long long* d_hist, cudaStream_t st);
size_t mh_sparse_select_temp_bytes(int64_t runs);
int mh_sparse_select(const uint32_t* d_run_counts, int64_t runs, uint32_t min_count, void* d_temp, size_t temp_bytes,
                     uint32_t* d_sel, unsigned long long* d_num_sel, cudaStream_t st);
int mh_sparse_gather_edges(const unsigned long long* d_run_keys, const uint32_t* d_run_counts, const uint32_t* d_sel, int64_t n_sel,
                           int64_t n, int32_t* d_i, int32_t* d_j, uint16_t* d_c, cudaStream_t st);
int mh_sparse_checksum(const unsigned long long* d_run_keys, const uint32_t* d_run_counts, int64_t runs, int64_t n,
                       unsigned long long* d_sum, cudaStream_t st);
int mh_sparse_densify(const unsigned long long* d_run_keys, const uint32_t* d_run_counts, int64_t runs, int64_t n, int64_t slab_base,
                      int64_t slab_pairs, uint16_t* d_dense, cudaStream_t st);

}  // namespace dyna
