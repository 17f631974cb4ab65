// MinHash device kernels (sm_100a): launch wrappers.  See mh_kernels.cu for the kernels themselves.
#pragma once
#include "common.cuh"

namespace dyna {

// Device layout of a signature set used by the match kernel: hash-major ("sigT"), i.e. row h holds the
// h-th signature component of every sequence, padded so every tile read is in bounds:
//   sigT[h * npitch + i],  h < hrows (= n_hash rounded up to kMatchBK), i < npitch (>= n + 256, multiple of 128)
constexpr int kMatchBM = 128;  // pairs tile: rows
constexpr int kMatchBN = 128;  // pairs tile: cols
constexpr int kMatchBK = 16;   // hash components per pipeline stage

inline int64_t mh_npitch(int64_t n) { return ((n + 127) / 128) * 128 + 256; }
inline int mh_hrows(int n_hash) { return ((n_hash + kMatchBK - 1) / kMatchBK) * kMatchBK; }
// rows of the packed 16-bit layout sigP[hrows2][npitch]: word (hp, i) = code(2hp, i) | code(2hp+1, i) << 16
inline int mh_hrows2(int n_hash) { return (((n_hash + 1) / 2 + kMatchBK - 1) / kMatchBK) * kMatchBK; }

// device workspace of the relabelling step (owned by the plan)
struct MhRelabelWork {
  void* temp = nullptr;          // CUB temporary storage
  size_t temp_bytes = 0;
  uint32_t* keys_out = nullptr;  // [hrows][npitch] sorted values
  uint32_t* vals_in = nullptr;   // [hrows][npitch] sequence indices 0..npitch-1 per row (constant)
  uint32_t* vals_out = nullptr;  // [hrows][npitch] indices in sorted order
  int* seg_begin = nullptr;      // [hrows] h * npitch
  int* seg_end = nullptr;        // [hrows] h * npitch + n
  uint32_t* sigP = nullptr;      // [hrows2][npitch] packed codes
  int* overflow = nullptr;       // set to 1 if some hash row has more than 61,440 distinct values
};

// K1: signatures from raw residues with MurmurHash3_x86_32 (src/minHash.cpp:21-64,140-157)
int launch_mh_signature_murmur3(const uint8_t* d_res, const int64_t* d_off, int64_t n, int64_t max_len, int k,
                                const uint32_t* d_seeds, int n_hash, uint32_t* d_sig, cudaStream_t st);
// K2: signatures from vocabulary ranks with (a*x+b) mod m (R/minHash.R:104-106,126-143)
int launch_mh_signature_linear(const int32_t* d_ranks, const int64_t* d_roff, int64_t n, const int64_t* d_a,
                               const int64_t* d_b, int64_t m, int n_hash, uint32_t* d_sig, cudaStream_t st);
// row-major sig[n][n_hash] -> hash-major sigT[hrows][npitch], zero padding
int launch_mh_transpose(const uint32_t* d_sig, int64_t n, int n_hash, uint32_t* d_sigT, int64_t npitch, int hrows,
                        cudaStream_t st);
// K3: match counts for rows [row_begin,row_end) into the packed strict-upper-triangle slab
// d_sigP == nullptr: 32-bit path only.  Otherwise the 16-bit kernel runs unless *d_overflow != 0, in which case the
// 32-bit kernel (enqueued right behind it) does the work.
int launch_mh_match(const uint32_t* d_sigT, int64_t npitch, int hrows, int n_hash, int64_t n, int64_t row_begin,
                    int64_t row_end, uint16_t* d_counts, const uint32_t* d_sigP, const int* d_overflow, cudaStream_t st,
                    int* launches);
// exact 16-bit relabelling of every hash row (see mh_kernels.cu)
size_t mh_relabel_temp_bytes(int64_t npitch, int hrows);
int launch_mh_iota(uint32_t* d_vals, int64_t npitch, int rows, cudaStream_t st);
int launch_mh_relabel(const uint32_t* d_sigT, int64_t n, int n_hash, int64_t npitch, int hrows, const MhRelabelWork& w,
                      int code_row_begin, int code_row_end, cudaStream_t st, int* launches);
int launch_mh_gather_rows(const uint32_t* d_sig, const int64_t* d_idx, int64_t m, int n_hash, uint32_t* d_out, cudaStream_t st);
// threshold + sparsify (R/clusterbreak.R:219-221) on the counts slab
int launch_mh_count_hist(const uint16_t* d_counts, int64_t total, int n_hash, unsigned long long* d_hist, cudaStream_t st);
int launch_mh_edges_count(const uint16_t* d_counts, int64_t n, int64_t row_begin, int64_t row_end, uint32_t min_count,
                          unsigned long long* d_row_counts, unsigned long long* d_row_offsets, unsigned long long* d_total,
                          cudaStream_t st);
int launch_mh_edges_fill(const uint16_t* d_counts, int64_t n, int64_t row_begin, int64_t row_end, uint32_t min_count,
                         const unsigned long long* d_row_offsets, int32_t* d_i, int32_t* d_j, uint16_t* d_c, cudaStream_t st);

}  // namespace dyna
