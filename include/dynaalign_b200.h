/* dynaalign_b200 -- C ABI of the B200-native (sm_100a) all-pairs similarity hot path of DynaAlign.
 *
 * This header is the drop-in boundary.  Every entry point takes plain pointers and sizes (no R, Rcpp,
 * torch or CUDA types) and is what the reference's Rcpp layer binds instead of its own C++ loops:
 *
 *   reference (paths under the DynaAlign source tree)            replaced by
 *   -----------------------------------------------------------  ---------------------------------
 *   similarityMH()            src/minHash.cpp:119-188             dyna_similarityMH
 *     HashFamily              src/minHash.cpp:67-89               dyna_hashfamily_seeds (host, mt19937)
 *     signature loop          src/minHash.cpp:140-157             dyna_mh_signatures_murmur3
 *     murmur3_32              src/minHash.cpp:21-64               (device code of the above)
 *     match loop              src/minHash.cpp:160-178             dyna_mh_match_counts / dyna_mh_match_matrix
 *   similarityNW()            src/pairwiseSeqAlign.cpp:331-365    dyna_similarityNW
 *     calculate_similarity    src/pairwiseSeqAlign.cpp:209-313    dyna_nw_pair_stats (matches, length per pair)
 *     getSubstitutionMatrix   src/pairwiseSeqAlign.cpp:190-206    dyna_substitution_matrix
 *     aa_to_index             src/pairwiseSeqAlign.cpp:15-21      dyna_aa_index_table
 *   minhash() (pure R)        R/minHash.R:206-221
 *     compute_signature_matrix R/minHash.R:126-143                dyna_mh_signatures_linear
 *     compute_distance_matrix  R/minHash.R:166-182                dyna_mh_match_matrix(kind=DYNA_MH_DISTANCE)
 *   .Call entry points        src/RcppExports.cpp:15,28           unchanged; see INTEGRATION.md
 *
 * Conventions
 *   - Sequences are passed flattened: `residues` holds all bytes back to back, `offsets[n+1]` delimits them
 *     (sequence i = residues[offsets[i] .. offsets[i+1])).  Bytes are raw R string bytes (no terminator).
 *   - Square results are column-major doubles, both triangles and the diagonal filled, exactly as the
 *     reference's NumericMatrix (consumed by clusterbreak/netcluster: R/clusterbreak.R:217-222).
 *   - "tri" results are the row-major packed upper triangle.  MinHash: strict (j>i), pair (i,j) at
 *     i*n - i*(i+1)/2 + (j-i-1).  NW: diagonal included (j>=i), pair (i,j) at i*n - i*(i-1)/2 + (j-i).
 *     A row range [row_begin,row_end) owns one contiguous slab starting at its first pair; this is the unit
 *     of multi-GPU sharding (one slab per GPU, no data-path collective).
 *   - Every function returns DYNA_OK (0) or an error code; dyna_last_error() then holds the message.  For
 *     argument errors the message is byte-identical to the reference's Rcpp::stop() text so the R side can
 *     re-raise it unchanged.  There is NO CPU fallback: without a usable CUDA device every compute entry
 *     point fails with DYNA_ERR_CUDA.
 *   - n_gpus: number of devices one call may drive from this process (row blocks balanced by work, one host thread
 *     per device).  The n x n matrix is then assembled in column blocks: every device expands its share of the
 *     columns, reading the result slabs of the other devices by peer loads over NVLink, and copies that contiguous
 *     block to the caller's matrix.  Values <= 0 mean "all visible".  Multi-process launches (one rank per GPU) use the row-range entry
 *     points with dyna_partition_rows instead.
 *   - Thread safety: calls are serialised per device internally; the API is re-entrant across devices.
 */
#ifndef DYNAALIGN_B200_H
#define DYNAALIGN_B200_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define DYNA_API __attribute__((visibility("default")))
#else
#define DYNA_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define DYNA_OK 0
#define DYNA_ERR_INVALID 1     /* argument / input validation (reference error strings) */
#define DYNA_ERR_CUDA 2        /* CUDA runtime failure or no device */
#define DYNA_ERR_UNSUPPORTED 3 /* outside the implemented domain (documented limits) */

#define DYNA_MH_SIMILARITY 0 /* matches/n_hash, diagonal 1.0   (src/minHash.cpp:161,174) */
#define DYNA_MH_DISTANCE 1   /* 1 - mean(==),   diagonal 0.0   (R/minHash.R:171-179)     */

/* ---------------------------------------------------------------- misc */
DYNA_API const char* dyna_last_error(void);
DYNA_API int dyna_version(void);
DYNA_API int dyna_device_count(void); /* 0 when no usable device; never fails */
DYNA_API int dyna_set_device(int device); /* device used by single-GPU entry points of this thread (default 0) */
/* Device buffers of finished calls stay cached in the CUDA memory pool for the next call (clusterbreak calls sim_fn
 * once per recursion node); this returns the cached memory of `device` to the driver. */
DYNA_API int dyna_release_cached_memory(int device);

/* Balanced row partition of the upper triangle.  weights==NULL: every pair costs 1
 * (MinHash); otherwise pair (i,j) costs weights[i]*weights[j] (NW: sequence lengths -> DP cells).
 * include_diagonal: 1 for NW, 0 for MinHash.  bounds_out[nshards+1], bounds[0]=0, bounds[nshards]=n. */
DYNA_API int dyna_partition_rows(int64_t n, const int64_t* weights, int include_diagonal, int nshards, int64_t* bounds_out);

/* ---------------------------------------------------------------- MinHash, C++ flavour (src/minHash.cpp) */
/* HashFamily(n_hash, seed): seeds[i] = i-th output of std::mt19937(seed) (src/minHash.cpp:73-80). Host only. */
DYNA_API int dyna_hashfamily_seeds(uint32_t seed, int n_hash, uint32_t* seeds_out);
/* What the reference uses when no seed is given: std::random_device{}() (src/minHash.cpp:73). */
DYNA_API uint32_t dyna_random_seed(void);

/* sig_out[n*n_hash] row-major uint32: min over all k-byte windows of murmur3_32(window, k, seeds[h]);
 * sequences shorter than k keep UINT32_MAX (src/minHash.cpp:140-157). */
DYNA_API int dyna_mh_signatures_murmur3(const uint8_t* residues, const int64_t* offsets, int64_t n, int k,
                               const uint32_t* seeds, int n_hash, uint32_t* sig_out);

/* counts_tri_out: #{h : sig[i][h]==sig[j][h]} for rows [row_begin,row_end), strict upper triangle slab
 * (src/minHash.cpp:167-173).  n_hash <= 65535. */
DYNA_API int dyna_mh_match_counts(const uint32_t* sig, int64_t n, int n_hash, int64_t row_begin, int64_t row_end,
                         uint16_t* counts_tri_out);

/* Full n x n double matrix from signatures; kind selects similarity (C++) or distance (R pipeline). */
DYNA_API int dyna_mh_match_matrix(const uint32_t* sig, int64_t n, int n_hash, int kind, double* out_colmajor, int n_gpus);

/* Drop-in for similarityMH(sequences, k, n_hash).  seeds==NULL: HashFamily seeded from dyna_random_seed()
 * (the reference's behaviour); tests inject seeds. */
DYNA_API int dyna_similarityMH(const uint8_t* residues, const int64_t* offsets, int64_t n, int k, int n_hash,
                      const uint32_t* seeds, double* out_colmajor, int n_gpus);

/* ---------------------------------------------------------------- MinHash, R flavour (R/minHash.R) */
/* sig_out[n*n_hash] row-major (== R's n_hash x n column-major sig_matrix): min over the document's shingle
 * ranks x (1-based vocabulary rank) of (a[h]*x + b[h]) mod m, computed in 64-bit (R/minHash.R:104-106,126-143).
 * Requires 1 <= m < 2^31, 0 <= a,b < 2^31, 1 <= ranks <= 2^31-1; documents with no shingles keep UINT32_MAX. */
DYNA_API int dyna_mh_signatures_linear(const int32_t* ranks, const int64_t* rank_offsets, int64_t n, const int64_t* a,
                              const int64_t* b, int64_t m, int n_hash, uint32_t* sig_out);

/* Device front end of the R pipeline: create_vocab (R/minHash.R:38-41) and, instead of the dense V x N characteristic
 * matrix of create_char_matrix (:60-66), the 1-based vocabulary rank of every k-shingle in document order
 * (ranks_out[rank_offsets_out[d] .. rank_offsets_out[d+1]) for document d) -- exactly the input of
 * dyna_mh_signatures_linear.  Vocabulary entries are the k bytes packed big-endian into a uint64 (sorted ascending ==
 * byte-wise string order).  k <= 8.  Any sequence shorter than k raises shingle()'s error text (:15-16).
 * vocab_keys_out may be NULL; ranks_out must hold sum(len_d - k + 1) entries, rank_offsets_out n+1. */
DYNA_API int dyna_minhash_vocab_ranks(const uint8_t* residues, const int64_t* offsets, int64_t n, int k, uint64_t* vocab_keys_out,
                             int64_t vocab_capacity, int64_t* vocab_size_out, int32_t* ranks_out, int64_t* rank_offsets_out);

/* ---------------------------------------------------------------- Needleman-Wunsch (src/pairwiseSeqAlign.cpp) */
DYNA_API int dyna_substitution_matrix(const char* name, int8_t* out576); /* 24x24 row-major; unknown name -> reference error */
DYNA_API void dyna_aa_index_table(int8_t* out256);                       /* byte -> 0..23, -1 = not in the alphabet */

/* (matches, alignment_length) of the reference's traceback for every pair i<=j of rows [row_begin,row_end),
 * lower index on rows.  similarity = (double)matches/length (0/0 -> NaN for two empty strings).
 * Sequence length limit: 65535 residues. */
DYNA_API int dyna_nw_pair_stats(const uint8_t* residues, const int64_t* offsets, int64_t n, const char* matrix_name,
                       int gap_open, int gap_ext, int64_t row_begin, int64_t row_end, uint32_t* matches_tri_out,
                       uint32_t* length_tri_out);

/* dyna_nw_pair_stats in the narrow form of dyna_nw_plan_fetch_packed8 (short peptides: 2 bytes per pair). */
DYNA_API int dyna_nw_pair_stats8(const uint8_t* residues, const int64_t* offsets, int64_t n, const char* matrix_name,
                        int gap_open, int gap_ext, int64_t row_begin, int64_t row_end, uint8_t* matches8_tri_out,
                        uint8_t* length8_tri_out);

/* Drop-in for similarityNW(sequences, matrixName, gapOpen, gapExt). n==0 -> OK, nothing written (0x0 matrix). */
DYNA_API int dyna_similarityNW(const uint8_t* residues, const int64_t* offsets, int64_t n, const char* matrix_name,
                      int gap_open, int gap_ext, double* out_colmajor, int n_gpus);

/* ---------------------------------------------------------------- device-resident plans
 * A plan owns device buffers on one GPU for one row range.  `run` only enqueues kernels on `stream`
 * (a cudaStream_t passed as void*, NULL = default stream) so callers can time it with their own events;
 * `fetch` copies results to host memory (synchronous on `stream`).  Used by bench.py for the
 * inputs-resident-in-HBM figure and by the host entry points above.  Device memory comes from the stream-ordered
 * allocator on the legacy default stream and stays cached in the pool between calls; work enqueued on a non-blocking
 * stream must have completed before the plan is destroyed. */
typedef struct dyna_mh_plan dyna_mh_plan;
DYNA_API dyna_mh_plan* dyna_mh_plan_create(int64_t n, int n_hash, int64_t row_begin, int64_t row_end, int device);
DYNA_API int dyna_mh_plan_upload_sequences(dyna_mh_plan*, const uint8_t* residues, const int64_t* offsets, int k,
                                  const uint32_t* seeds, void* stream);
DYNA_API int dyna_mh_plan_upload_signatures(dyna_mh_plan*, const uint32_t* sig, void* stream);
DYNA_API int dyna_mh_plan_run_signatures(dyna_mh_plan*, void* stream); /* K1 + layout transform */
DYNA_API int dyna_mh_plan_run_match(dyna_mh_plan*, void* stream);      /* K3 over the plan's row range */
/* Multi-rank form of run_signatures (the all-gather variant of SURVEY.md section 8(e); the reference has no
 * counterpart, src/minHash.cpp:143-157 recomputes nothing because it is one process): K1 and the layout transform
 * run for every hash row, the exact 16-bit relabelling only for packed code rows [code_row_begin, code_row_end)
 * out of dyna_mh_plan_code_rows() (0 = this plan matches on the 32-bit signatures and needs no exchange).  Before
 * run_match the caller all-gathers the code table -- rows of dyna_mh_plan_code_row_bytes() bytes at
 * dyna_mh_plan_codes_device_ptr() -- and max-reduces the int at dyna_mh_plan_overflow_device_ptr() across ranks
 * (NCCL over NVLink in bench.py). */
DYNA_API int dyna_mh_plan_run_signatures_shard(dyna_mh_plan*, int code_row_begin, int code_row_end, void* stream);
DYNA_API int dyna_mh_plan_code_rows(const dyna_mh_plan*);
DYNA_API int64_t dyna_mh_plan_code_row_bytes(const dyna_mh_plan*);
DYNA_API void* dyna_mh_plan_codes_device_ptr(dyna_mh_plan*);
DYNA_API void* dyna_mh_plan_overflow_device_ptr(dyna_mh_plan*);
/* run_match + fetch_counts fused: row chunks are copied to the host (pinned memory recommended) while the next chunks
 * are still being matched */
DYNA_API int dyna_mh_plan_run_match_fetch(dyna_mh_plan*, uint16_t* counts_tri_out, void* stream);
DYNA_API int dyna_mh_plan_fetch_signatures(dyna_mh_plan*, uint32_t* sig_out, void* stream);
DYNA_API int dyna_mh_plan_fetch_counts(dyna_mh_plan*, uint16_t* counts_tri_out, void* stream);
/* Child plan over a subset of the parent's sequences (clusterbreak's recursion re-invokes sim_fn on each oversized
 * cluster, R/clusterbreak.R:250-254): gathers the parent's device-resident signature rows, nothing is re-hashed.
 * indices[m] are 0-based positions in the parent; the child is independent of the parent afterwards. */
DYNA_API dyna_mh_plan* dyna_mh_plan_create_subset(dyna_mh_plan* parent, const int64_t* indices, int64_t m, int64_t row_begin,
                                         int64_t row_end);
/* The step after the hot path in clusterbreak (R/clusterbreak.R:219-221), on the plan's device-resident counts:
 *   threshold <- quantile(sim[upper.tri(sim)], thresh_p);  sim[sim < threshold] <- 0;  graph from the upper triangle.
 * Similarities are count/n_hash, so the type-7 quantile is exact from the histogram of counts (sum the histograms of
 * all row ranges first when the triangle is sharded), and the thresholded matrix is the edge list of pairs with
 * count >= min_count (zero counts are never edges), emitted in row-major (i, then j) order. */
DYNA_API int dyna_mh_plan_count_histogram(dyna_mh_plan*, uint64_t* hist_out /* n_hash+1 */, void* stream);
DYNA_API int dyna_quantile_type7_counts(const uint64_t* hist, int n_hash, double prob, double* threshold_out, int* min_count_out);
DYNA_API int dyna_mh_plan_threshold_edges(dyna_mh_plan*, int min_count, int64_t max_edges, int32_t* i_out, int32_t* j_out,
                                 uint16_t* count_out, int64_t* n_edges_out, void* stream);
/* run_match_fetch in the narrow host form (1 byte per pair instead of 2): counts8_out holds min(count, 255) and every
 * pair with count >= 255 is also reported exactly as (packed pair index in the FULL strict triangle, count) -- lossless.
 * Escapes arrive unordered; more than esc_capacity of them fails with DYNA_ERR_INVALID and *n_esc_out = the number. */
DYNA_API int dyna_mh_plan_run_match_fetch8(dyna_mh_plan*, uint8_t* counts8_out, int64_t esc_capacity, int64_t* esc_index_out,
                                  uint16_t* esc_count_out, int64_t* n_esc_out, void* stream);
/* sum over the plan's slab of count[k] * w(global pair index k), w(k) = x ^ (x >> 31), x = (k+1)*0x9E3779B97F4A7C15,
 * in wrap-around 64-bit arithmetic.  Additive over the slabs of any row partition: summed over all ranks it equals
 * the single-device value exactly when the slabs tile the triangle (bench.py's multi-GPU parity check). */
/* The same match counts by a JOIN on equal signature values instead of all-pairs compares (csrc/mh_sparse.cu): the hash
 * rows are already sorted by the 16-bit relabelling; every group of equal values contributes its pairs, the pair list is
 * sorted and run-length encoded.  Work and memory are O(n * n_hash + matches), not O(n^2 * n_hash) -- BASELINE config 4
 * (2.5e12 compares) has ~6e6 matches -- and the plan needs no dense triangle, so n is not bound by n^2 memory.
 * Exact, but data dependent: *n_incidences_out reports the number of (pair, hash function) matches and *done_out = 0
 * means the join was NOT run (more than max_incidences matches; 0 = a cap derived from free device memory; or the plan
 * has no sorted rows: n < 2048, or a sharded relabelling) and the caller should call dyna_mh_plan_run_match instead.
 * After done = 1, count_histogram / threshold_edges / checksum read the join's result directly; fetch_counts and
 * counts_device_ptr first scatter it into the dense u16 triangle. */
DYNA_API int dyna_mh_plan_run_match_sparse(dyna_mh_plan*, int64_t max_incidences, int64_t* n_incidences_out, int* done_out,
                                  void* stream);
DYNA_API int dyna_mh_plan_checksum(dyna_mh_plan*, uint64_t* sum_out /* 1 */, void* stream);
DYNA_API int64_t dyna_mh_plan_pairs(const dyna_mh_plan*);
DYNA_API int dyna_mh_plan_launches(const dyna_mh_plan*); /* kernels enqueued by the last run_* call */
DYNA_API void* dyna_mh_plan_counts_device_ptr(dyna_mh_plan*);
DYNA_API void dyna_mh_plan_destroy(dyna_mh_plan*);

typedef struct dyna_nw_plan dyna_nw_plan;
DYNA_API dyna_nw_plan* dyna_nw_plan_create(const uint8_t* residues, const int64_t* offsets, int64_t n, const char* matrix_name,
                                  int gap_open, int gap_ext, int64_t row_begin, int64_t row_end, int device);
DYNA_API int dyna_nw_plan_run(dyna_nw_plan*, void* stream);
DYNA_API int dyna_nw_plan_fetch(dyna_nw_plan*, uint32_t* matches_tri_out, uint32_t* length_tri_out, void* stream);
/* Narrow host form: one byte each for matches and alignment length (2 B/pair instead of 8).  Only when every
 * alignment length fits a byte (2 * longest sequence <= 255); DYNA_ERR_UNSUPPORTED otherwise. */
DYNA_API int dyna_nw_plan_fetch_packed8(dyna_nw_plan*, uint8_t* matches8_tri_out, uint8_t* length8_tri_out, void* stream);
/* Same weighting as dyna_mh_plan_checksum over the NW triangle (diagonal included): sum_out[0] matches, [1] length. */
DYNA_API int dyna_nw_plan_checksum(dyna_nw_plan*, uint64_t* sum_out /* 2 */, void* stream);
/* Threshold + sparsify for sim_fn = similarityNW, the step right after the hot path in clusterbreak
 * (R/clusterbreak.R:217-221: threshold <- quantile(sim[upper.tri(sim)], thresh_p); sim[sim < threshold] <- 0), without
 * the dense matrix.  A recursion node is the subset `members[0..n_members)` of the plan's sequences (strictly
 * increasing 0-based indices; NULL = all sequences): sim_fn(sub-cluster) (R/clusterbreak.R:250-254) is a sub-matrix of
 * the root's, so nothing is re-aligned.  Results cover the node pairs whose row lies in the plan's row range; ranks add
 * their histograms and concatenate their edge lists.
 *   stat_histogram : hist_out[(matches) * (2*max_len+1) + length] over pairs a < b, (max_len+1) x (2*max_len+1) counters
 *   quantile       : exact type-7 quantile of (double)matches/length; two empty sequences (NaN) -> R's quantile error
 *   threshold_edges: pairs with matches > 0 and (double)matches/length >= threshold, row-major, node-local indices */
DYNA_API int dyna_nw_plan_max_len(const dyna_nw_plan*); /* longest sequence of the plan (>= 1) */
DYNA_API int dyna_nw_plan_stat_histogram(dyna_nw_plan*, const int32_t* members, int64_t n_members, uint64_t* hist_out, void* stream);
DYNA_API int dyna_nw_plan_fetch_diagonal(dyna_nw_plan*, const int32_t* members, int64_t n_members, uint32_t* matches_out,
                                uint32_t* length_out, void* stream); /* self-alignments = self-loop weights of netcluster */
DYNA_API int dyna_quantile_type7_identities(const uint64_t* hist, int64_t mdim, int64_t ldim, double prob, double* threshold_out);
DYNA_API int dyna_nw_plan_threshold_edges(dyna_nw_plan*, const int32_t* members, int64_t n_members, double threshold,
                                 int64_t max_edges, int32_t* i_out, int32_t* j_out, uint32_t* matches_out,
                                 uint32_t* length_out, int64_t* n_edges_out, void* stream);
/* The planner's work units without touching a device (host logic only): which kernel family and strip height every
 * (row | row pair) x column block of the triangle goes to.  The returned plan supports only unit_count / export_units /
 * pairs / cells / destroy.  export_units writes 6 int32 per unit: kind, R, row, second row (-1 for single-row kinds),
 * first column, column count. */
DYNA_API dyna_nw_plan* dyna_nw_plan_layout(const uint8_t* residues, const int64_t* offsets, int64_t n, const char* matrix_name,
                                  int gap_open, int gap_ext, int64_t row_begin, int64_t row_end);
DYNA_API int64_t dyna_nw_plan_unit_count(const dyna_nw_plan*);
DYNA_API int dyna_nw_plan_export_units(const dyna_nw_plan*, int32_t* out6);
DYNA_API int64_t dyna_nw_plan_pairs(const dyna_nw_plan*);
DYNA_API int64_t dyna_nw_plan_cells(const dyna_nw_plan*); /* sum of len_i*len_j over the plan's pairs */
DYNA_API int dyna_nw_plan_launches(const dyna_nw_plan*);
DYNA_API void dyna_nw_plan_destroy(dyna_nw_plan*);

/* ---------------------------------------------------------------- measurement helpers
 * Dependency-free integer-issue probe used as the NW / match-count roofline denominator:
 * returns lane-operations per second of the chosen instruction mix on the current device.
 * kind: 0 = IADD3, 1 = VIADDMNMX (DPX), 2 = VIMNMX3 (DPX), 3 = mixed NW cell mix, 4 = ISETP+predicated add */
DYNA_API int dyna_probe_int_issue(int kind, double* lane_ops_per_s, double* elapsed_ms, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DYNAALIGN_B200_H */
