/* TEST INFRASTRUCTURE ONLY -- not part of the product.
 * Plain-C restatement of DynaAlign's all-pairs similarity hot path (see dyna_oracle.c). */
#ifndef DYNA_ORACLE_H
#define DYNA_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
const char* orc_last_error(void);
uint32_t orc_murmur3_32(const uint8_t* key, uint64_t len, uint32_t seed);
void     orc_hashfamily_seeds(uint32_t seed, int n_hash, uint32_t* seeds_out);
int      orc_mh_signatures(const uint8_t* residues, const int64_t* offsets, int64_t n, int k,
                           const uint32_t* seeds, int n_hash, uint32_t* sig_rowmajor);
void     orc_mh_match_counts(const uint32_t* sig, int64_t n, int n_hash, int64_t row_begin, int64_t row_end,
                             uint16_t* counts_tri);
int      orc_similarityMH(const uint8_t* residues, const int64_t* offsets, int64_t n, int k, int n_hash,
                          uint32_t seed, double* out_colmajor);
int      orc_substitution_matrix(const char* name, int8_t* out576);
int      orc_nw_pair(const uint8_t* a, int64_t m, const uint8_t* b, int64_t n, const int8_t* sub576,
                     int gap_open, int gap_ext, int32_t* matches, int32_t* aln_len);
int      orc_nw_pair_forward(const uint8_t* a, int64_t m, const uint8_t* b, int64_t n, const int8_t* sub576,
                             int gap_open, int gap_ext, int32_t* matches, int32_t* aln_len);
int      orc_nw_pair_stats(const uint8_t* residues, const int64_t* offsets, int64_t n, const char* matrix_name,
                           int gap_open, int gap_ext, int64_t row_begin, int64_t row_end,
                           uint32_t* matches_tri, uint32_t* len_tri);
int      orc_similarityNW(const uint8_t* residues, const int64_t* offsets, int64_t n, const char* matrix_name,
                          int gap_open, int gap_ext, double* out_colmajor);
int      orc_mh_signatures_linear(const int32_t* ranks, const int64_t* rank_offsets, int64_t n,
                                  const int64_t* a, const int64_t* b, int64_t m, int n_hash, uint32_t* sig_rowmajor);
void     orc_mh_distance_matrix(const uint32_t* sig, int64_t n, int n_hash, double* out_colmajor);
#ifdef __cplusplus
}
#endif
#endif
