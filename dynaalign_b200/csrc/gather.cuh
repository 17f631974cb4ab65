// Result assembly for sm_100a: expansion of the packed-triangle results into the reference's column-major n x n
// double matrix (src/pairwiseSeqAlign.cpp:311,349-350,356-362; src/minHash.cpp:161,174-176), position-weighted
// checksums of the slabs, and the narrow host-returnable forms.  See gather.cu.
#pragma once
#include "common.cuh"

namespace dyna {

constexpr int kMaxSlabs = 8;

// The packed upper triangle cut into consecutive row blocks ("slabs"), each living on the device that computed it.
// With peer access enabled (NVLink / NVSwitch) the expansion kernel of any device reads all of them directly.
struct TriSlabs {
  int nslabs;
  int64_t row_begin[kMaxSlabs + 1];  // slab g owns rows [row_begin[g], row_begin[g+1])
  const void* a[kMaxSlabs];          // NW: matches (u32) | MinHash: counts (u16)
  const void* b[kMaxSlabs];          // NW: length (u32)  | MinHash: unused
};

// Column block [col_begin, col_end) of the n x n matrix -> out_block (column-major, leading dimension n, so the block
// is one contiguous range of the caller's matrix).  Entries on and below the diagonal of a column are read along the
// owning row (coalesced); entries above it are read along THEIR rows and transposed through shared memory.
int launch_nw_expand_block(const TriSlabs& s, int64_t n, int64_t col_begin, int64_t col_end, double* d_out_block, cudaStream_t st);
// MinHash: value table[count] (n_hash + 1 doubles on this device), `diag` on the diagonal
int launch_mh_expand_block(const TriSlabs& s, int64_t n, int64_t col_begin, int64_t col_end, const double* d_table, double diag,
                           double* d_out_block, cudaStream_t st);

// sum_k value[k] * mix(first_index + k)  (mod 2^64): additive over any partition of the triangle into slabs
int launch_checksum_u32(const uint32_t* d_v, int64_t count, int64_t first_index, unsigned long long* d_sum, cudaStream_t st);
int launch_checksum_u16(const uint16_t* d_v, int64_t count, int64_t first_index, unsigned long long* d_sum, cudaStream_t st);

// (matches, length) u32 -> u8 each (caller guarantees every value <= 255)
int launch_nw_pack8(const uint32_t* d_m, const uint32_t* d_l, int64_t count, uint8_t* d_m8, uint8_t* d_l8, cudaStream_t st);
// counts u16 -> u8 saturated at 255; every count >= 255 is also appended to the escape list (global pair index, count)
int launch_mh_narrow8(const uint16_t* d_counts, int64_t count, int64_t first_index, uint8_t* d_out8, int64_t esc_capacity,
                      long long* d_esc_index, uint16_t* d_esc_count, unsigned long long* d_esc_n, cudaStream_t st);

}  // namespace dyna
