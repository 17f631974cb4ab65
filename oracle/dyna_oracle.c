/* TEST INFRASTRUCTURE ONLY -- not part of the product; nothing under dynaalign_b200/ links or
 * loads this file.  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs use it.
 *
 * Plain-C restatement of the reference's algorithms for the all-pairs similarity hot path, each
 * function citing the reference lines it follows (paths relative to /root/reference):
 *
 *   orc_murmur3_32            src/minHash.cpp:21-64      MurmurHash3_x86_32
 *   orc_hashfamily_seeds      src/minHash.cpp:67-81      seeds[i] = i-th std::mt19937(seed) output
 *   orc_mh_signatures         src/minHash.cpp:92-105,140-157
 *   orc_mh_match_counts       src/minHash.cpp:160-173
 *   orc_similarityMH          src/minHash.cpp:119-188
 *   orc_nw_pair               src/pairwiseSeqAlign.cpp:209-313 (full matrices + backtrack, as written)
 *   orc_nw_pair_forward       same result, O(n) memory: (matches, diag steps) carried forward
 *   orc_similarityNW          src/pairwiseSeqAlign.cpp:331-365
 *   orc_mh_signatures_linear  R/minHash.R:104-106,126-143 ((a*x+b) %% m on 1-based vocabulary ranks)
 *   orc_mh_distance_matrix    R/minHash.R:166-182
 *
 * Parity status: PINNED.  tests/test_oracle_vs_reference.py checks every function here against
 * oracle/_ref/libdynaref.so (the reference's own C++ compiled unmodified) and against the committed
 * fixtures in tests/golden/ that were generated from it (tests/golden/make_golden.py).
 */
#include "dyna_oracle.h"

#include <limits.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "oracle_tables.h"

static _Thread_local char g_err[256];
const char* orc_last_error(void) { return g_err; }

/* ------------------------------------------------------------------ MurmurHash3_x86_32 */
static uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

uint32_t orc_murmur3_32(const uint8_t* key, uint64_t len, uint32_t seed) {
  /* src/minHash.cpp:21-64.  Blocks are read little-endian (the reference casts the byte pointer
   * to uint32_t* on x86). */
  uint32_t h = seed;
  uint64_t nblocks = len / 4;
  for (uint64_t i = 0; i < nblocks; ++i) {
    const uint8_t* p = key + 4 * i;
    uint32_t k = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
    k *= 0xcc9e2d51u;
    k = rotl32(k, 15);
    k *= 0x1b873593u;
    h ^= k;
    h = rotl32(h, 13) * 5u + 0xe6546b64u;
  }
  const uint8_t* tail = key + 4 * nblocks;
  uint32_t k1 = 0;
  unsigned rem = (unsigned)(len & 3u);
  if (rem) {
    if (rem >= 3) k1 ^= (uint32_t)tail[2] << 16;
    if (rem >= 2) k1 ^= (uint32_t)tail[1] << 8;
    k1 ^= (uint32_t)tail[0];
    k1 *= 0xcc9e2d51u;
    k1 = rotl32(k1, 15);
    k1 *= 0x1b873593u;
    h ^= k1;
  }
  h ^= (uint32_t)len;
  h ^= h >> 16;
  h *= 0x85ebca6bu;
  h ^= h >> 13;
  h *= 0xc2b2ae35u;
  h ^= h >> 16;
  return h;
}

/* ------------------------------------------------------------------ std::mt19937 (32-bit MT) */
typedef struct { uint32_t s[624]; int idx; } mt19937_t;

static void mt_seed(mt19937_t* g, uint32_t seed) {
  g->s[0] = seed;
  for (int i = 1; i < 624; ++i) g->s[i] = 1812433253u * (g->s[i - 1] ^ (g->s[i - 1] >> 30)) + (uint32_t)i;
  g->idx = 624;
}
static uint32_t mt_next(mt19937_t* g) {
  if (g->idx >= 624) {
    for (int i = 0; i < 624; ++i) {
      uint32_t y = (g->s[i] & 0x80000000u) | (g->s[(i + 1) % 624] & 0x7fffffffu);
      uint32_t v = g->s[(i + 397) % 624] ^ (y >> 1);
      if (y & 1u) v ^= 0x9908b0dfu;
      g->s[i] = v;
    }
    g->idx = 0;
  }
  uint32_t y = g->s[g->idx++];
  y ^= y >> 11;
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= y >> 18;
  return y;
}

void orc_hashfamily_seeds(uint32_t seed, int n_hash, uint32_t* seeds_out) {
  /* src/minHash.cpp:73-80: mt19937 gen(seed); uniform_int_distribution<uint32_t> dis; seeds[i]=dis(gen).
   * A full-range uint32 distribution over a 32-bit engine passes the raw stream through (libstdc++). */
  mt19937_t g;
  mt_seed(&g, seed);
  for (int i = 0; i < n_hash; ++i) seeds_out[i] = mt_next(&g);
}

/* ------------------------------------------------------------------ similarityMH */
int orc_mh_signatures(const uint8_t* residues, const int64_t* offsets, int64_t n, int k,
                      const uint32_t* seeds, int n_hash, uint32_t* sig) {
  /* src/minHash.cpp:140-157 with generate_kmers (:92-105): every k-byte window, duplicates included;
   * L < k leaves the row at UINT32_MAX. */
  if (k <= 0 || n_hash <= 0) { snprintf(g_err, sizeof g_err, "bad k/n_hash"); return 1; }
#pragma omp parallel for schedule(dynamic, 16)
  for (int64_t i = 0; i < n; ++i) {
    const uint8_t* s = residues + offsets[i];
    int64_t L = offsets[i + 1] - offsets[i];
    uint32_t* row = sig + i * (int64_t)n_hash;
    for (int h = 0; h < n_hash; ++h) row[h] = UINT32_MAX;
    for (int64_t p = 0; p + k <= L; ++p)
      for (int h = 0; h < n_hash; ++h) {
        uint32_t v = orc_murmur3_32(s + p, (uint64_t)k, seeds[h]);
        if (v < row[h]) row[h] = v;
      }
  }
  return 0;
}

static int64_t tri_strict_index(int64_t n, int64_t i, int64_t j) { /* i<j, row-major strict upper triangle */
  return i * n - i * (i + 1) / 2 + (j - i - 1);
}

void orc_mh_match_counts(const uint32_t* sig, int64_t n, int n_hash, int64_t row_begin, int64_t row_end,
                         uint16_t* counts_tri) {
  /* src/minHash.cpp:160-173; counts for rows [row_begin,row_end) written as one contiguous slab of the
   * packed strict upper triangle (slab origin = index of (row_begin,row_begin+1)). */
  int64_t base = row_begin < n - 1 ? tri_strict_index(n, row_begin, row_begin + 1) : 0;
#pragma omp parallel for schedule(dynamic, 4)
  for (int64_t i = row_begin; i < row_end; ++i)
    for (int64_t j = i + 1; j < n; ++j) {
      int matches = 0;
      const uint32_t* a = sig + i * (int64_t)n_hash;
      const uint32_t* b = sig + j * (int64_t)n_hash;
      for (int h = 0; h < n_hash; ++h) matches += (a[h] == b[h]);
      counts_tri[tri_strict_index(n, i, j) - base] = (uint16_t)matches;
    }
}

int orc_similarityMH(const uint8_t* residues, const int64_t* offsets, int64_t n, int k, int n_hash,
                     uint32_t seed, double* out) {
  /* src/minHash.cpp:119-188 (error strings :122,:126,:130) */
  if (n == 0) { snprintf(g_err, sizeof g_err, "Input sequences vector cannot be empty"); return 1; }
  if (k <= 0) { snprintf(g_err, sizeof g_err, "'k' must be a positive integer"); return 1; }
  if (n_hash <= 0) { snprintf(g_err, sizeof g_err, "Number of hash functions must be positive"); return 1; }
  uint32_t* seeds = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)n_hash);
  uint32_t* sig = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)n_hash * (size_t)n);
  orc_hashfamily_seeds(seed, n_hash, seeds);
  orc_mh_signatures(residues, offsets, n, k, seeds, n_hash, sig);
  for (int64_t i = 0; i < n; ++i) {
    out[i + i * n] = 1.0;
#pragma omp parallel for
    for (int64_t j = i + 1; j < n; ++j) {
      int matches = 0;
      for (int h = 0; h < n_hash; ++h) matches += (sig[i * (int64_t)n_hash + h] == sig[j * (int64_t)n_hash + h]);
      double s = (double)matches / n_hash;
      out[i + j * n] = s;
      out[j + i * n] = s;
    }
  }
  free(sig);
  free(seeds);
  return 0;
}

/* ------------------------------------------------------------------ similarityNW */
static int aa_index(uint8_t c) { /* src/pairwiseSeqAlign.cpp:15-21 */
  for (int i = 0; i < 24; ++i) if ((uint8_t)orc_alphabet[i] == c) return i;
  return -1;
}

int orc_substitution_matrix(const char* name, int8_t* out576) { /* :190-206 */
  for (int t = 0; t < ORC_NUM_TABLES; ++t)
    if (strcmp(name, orc_table_names[t]) == 0) {
      memcpy(out576, orc_tables[t], 576);
      return 0;
    }
  snprintf(g_err, sizeof g_err, "Invalid substitution matrix name: %s", name);
  return 1;
}

#define ORC_NEG (INT_MIN / 2)
static int32_t wsub(int32_t a, int32_t b) { return (int32_t)((uint32_t)a - (uint32_t)b); } /* int32 wrap, no UB */
static int32_t wadd(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }
static int32_t wmul(int32_t a, int32_t b) { return (int32_t)((uint32_t)a * (uint32_t)b); }
static int32_t max2(int32_t a, int32_t b) { return a > b ? a : b; }

int orc_nw_pair(const uint8_t* a, int64_t m, const uint8_t* b, int64_t n, const int8_t* S,
                int go, int ge, int32_t* matches_out, int32_t* len_out) {
  /* src/pairwiseSeqAlign.cpp:209-313, as written: three (m+1)x(n+1) int matrices + pointer matrix,
   * border init :222-235, fill :238-281 (tie-break D >= U >= L, M overwritten by the winner),
   * backtrack :284-308.  Residue validation happens lazily inside the fill loops (:239-250). */
  size_t W = (size_t)n + 1, cells = ((size_t)m + 1) * W;
  int32_t* M = (int32_t*)malloc(cells * sizeof(int32_t));
  int32_t* Ix = (int32_t*)malloc(cells * sizeof(int32_t));
  int32_t* Iy = (int32_t*)malloc(cells * sizeof(int32_t));
  char* tb = (char*)malloc(cells);
  if (!M || !Ix || !Iy || !tb) { free(M); free(Ix); free(Iy); free(tb); snprintf(g_err, sizeof g_err, "oom"); return 2; }
  for (size_t c = 0; c < cells; ++c) { M[c] = Ix[c] = Iy[c] = ORC_NEG; tb[c] = '0'; }
  M[0] = 0;
  for (int64_t i = 1; i <= m; ++i) {
    M[i * W] = ORC_NEG;
    Ix[i * W] = wsub(-go, wmul((int32_t)(i - 1), ge));
    Iy[i * W] = ORC_NEG;
    tb[i * W] = 'U';
  }
  for (int64_t j = 1; j <= n; ++j) {
    M[j] = ORC_NEG;
    Ix[j] = ORC_NEG;
    Iy[j] = wsub(-go, wmul((int32_t)(j - 1), ge));
    tb[j] = 'L';
  }
  int rc = 0;
  for (int64_t i = 1; i <= m && !rc; ++i) {
    int i1 = aa_index(a[i - 1]);
    if (i1 < 0) { snprintf(g_err, sizeof g_err, "Invalid amino acid in sequence1: %c", a[i - 1]); rc = 1; break; }
    for (int64_t j = 1; j <= n; ++j) {
      int i2 = aa_index(b[j - 1]);
      if (i2 < 0) { snprintf(g_err, sizeof g_err, "Invalid amino acid in sequence2: %c", b[j - 1]); rc = 1; break; }
      int32_t s = S[i1 * 24 + i2];
      size_t c = (size_t)i * W + (size_t)j, up = c - W, lf = c - 1, dg = c - W - 1;
      int32_t ix = max2(wsub(M[up], wadd(go, ge)), wsub(Ix[up], ge));
      int32_t iy = max2(wsub(M[lf], wadd(go, ge)), wsub(Iy[lf], ge));
      int32_t mm = max2(max2(wadd(M[dg], s), wadd(Ix[dg], s)), wadd(Iy[dg], s));
      Ix[c] = ix; Iy[c] = iy;
      if (mm >= ix && mm >= iy) { tb[c] = 'D'; M[c] = mm; }
      else if (ix >= iy)        { tb[c] = 'U'; M[c] = ix; }
      else                      { tb[c] = 'L'; M[c] = iy; }
    }
  }
  if (!rc) {
    int32_t matches = 0, len = 0;
    int64_t i = m, j = n;
    while (i > 0 || j > 0) {
      char t = tb[(size_t)i * W + (size_t)j];
      if (t == 'D') { if (a[i - 1] == b[j - 1]) ++matches; --i; --j; }
      else if (t == 'U') --i;
      else --j;
      ++len;
    }
    *matches_out = matches;
    *len_out = len;
  }
  free(M); free(Ix); free(Iy); free(tb);
  return rc;
}

int orc_nw_pair_forward(const uint8_t* a, int64_t m, const uint8_t* b, int64_t n, const int8_t* S,
                        int go, int ge, int32_t* matches_out, int32_t* len_out) {
  /* Same function as orc_nw_pair, restated without the pointer matrix: each
   * cell has exactly one predecessor, so (matches, #diagonal steps) can be carried forward along
   * the chosen pointer; alignment length = m + n - #diagonal steps.  Rolling rows, O(n) memory.
   * Used to cross-check the formulation the CUDA kernels implement; validated against orc_nw_pair
   * and the compiled reference in tests/. */
  size_t W = (size_t)n + 1;
  int32_t* H = (int32_t*)malloc(W * sizeof(int32_t));   /* max(M,Ix,Iy) of previous row (diag role) */
  int32_t* Mo = (int32_t*)malloc(W * sizeof(int32_t));  /* M of previous row as an open-source      */
  int32_t* X = (int32_t*)malloc(W * sizeof(int32_t));   /* Ix of previous row                       */
  int32_t* mt = (int32_t*)malloc(W * sizeof(int32_t));
  int32_t* dg = (int32_t*)malloc(W * sizeof(int32_t));
  H[0] = 0; Mo[0] = 0; X[0] = ORC_NEG; mt[0] = 0; dg[0] = 0;
  for (int64_t j = 1; j <= n; ++j) {
    H[j] = wsub(-go, wmul((int32_t)(j - 1), ge));       /* max(NEG, NEG, Iy[0][j]) */
    Mo[j] = ORC_NEG; X[j] = ORC_NEG; mt[j] = 0; dg[j] = 0;
  }
  int rc = 0;
  for (int64_t i = 1; i <= m && !rc; ++i) {
    int i1 = aa_index(a[i - 1]);
    if (i1 < 0) { snprintf(g_err, sizeof g_err, "Invalid amino acid in sequence1: %c", a[i - 1]); rc = 1; break; }
    int32_t Hd = H[0], mtd = mt[0], dgd = dg[0];                     /* (i-1, j-1) */
    int32_t bord = wsub(-go, wmul((int32_t)(i - 1), ge));           /* Ix[i][0] */
    int32_t Hl = bord, Ml = ORC_NEG, Yl = ORC_NEG, mtl = 0, dgl = 0; /* (i, j-1) */
    H[0] = bord; Mo[0] = ORC_NEG; X[0] = bord; mt[0] = 0; dg[0] = 0;
    for (int64_t j = 1; j <= n; ++j) {
      int i2 = aa_index(b[j - 1]);
      if (i2 < 0) { snprintf(g_err, sizeof g_err, "Invalid amino acid in sequence2: %c", b[j - 1]); rc = 1; break; }
      int32_t s = S[i1 * 24 + i2];
      int32_t ix = max2(wsub(Mo[j], wadd(go, ge)), wsub(X[j], ge));
      int32_t iy = max2(wsub(Ml, wadd(go, ge)), wsub(Yl, ge));
      int32_t mm = wadd(Hd, s);
      int32_t h, nm, nd;
      if (mm >= ix && mm >= iy) { h = mm; nm = mtd + (a[i - 1] == b[j - 1]); nd = dgd + 1; }
      else if (ix >= iy)        { h = ix; nm = mt[j]; nd = dg[j]; }
      else                      { h = iy; nm = mtl;   nd = dgl; }
      Hd = H[j]; mtd = mt[j]; dgd = dg[j];
      H[j] = h; Mo[j] = h; X[j] = ix; mt[j] = nm; dg[j] = nd;
      Hl = h; Ml = h; Yl = iy; mtl = nm; dgl = nd;
    }
    (void)Hl;
  }
  if (!rc) { *matches_out = mt[n]; *len_out = (int32_t)(m + n) - dg[n]; }
  free(H); free(Mo); free(X); free(mt); free(dg);
  return rc;
}

static int64_t tri_diag_index(int64_t n, int64_t i, int64_t j) { /* i<=j, row-major upper triangle incl. diagonal */
  return i * n - i * (i - 1) / 2 + (j - i);
}

int orc_nw_pair_stats(const uint8_t* residues, const int64_t* offsets, int64_t n, const char* matrix_name,
                      int go, int ge, int64_t row_begin, int64_t row_end,
                      uint32_t* matches_tri, uint32_t* len_tri) {
  /* driver loop of src/pairwiseSeqAlign.cpp:340-352 restricted to rows [row_begin,row_end): all j>=i,
   * diagonal included, lower index on rows.  Output slab origin = index of (row_begin,row_begin). */
  int8_t S[576];
  if (orc_substitution_matrix(matrix_name, S)) return 1;
  int64_t base = tri_diag_index(n, row_begin, row_begin);
  int rc = 0;
  char err[256] = {0};
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t i = row_begin; i < row_end; ++i)
    for (int64_t j = i; j < n; ++j) {
      int32_t mt = 0, ln = 0;
      int r = orc_nw_pair(residues + offsets[i], offsets[i + 1] - offsets[i], residues + offsets[j],
                          offsets[j + 1] - offsets[j], S, go, ge, &mt, &ln);
      if (r) {
#pragma omp critical
        { if (!rc) { rc = r; memcpy(err, g_err, sizeof err); } }
        continue;
      }
      matches_tri[tri_diag_index(n, i, j) - base] = (uint32_t)mt;
      len_tri[tri_diag_index(n, i, j) - base] = (uint32_t)ln;
    }
  if (rc) memcpy(g_err, err, sizeof err);
  return rc;
}

int orc_similarityNW(const uint8_t* residues, const int64_t* offsets, int64_t n, const char* matrix_name,
                     int go, int ge, double* out) {
  /* src/pairwiseSeqAlign.cpp:331-365, single-threaded and in the reference's pair order so that the
   * first error raised is the reference's. */
  int8_t S[576];
  if (orc_substitution_matrix(matrix_name, S)) return 1;
  for (int64_t i = 0; i < n; ++i)
    for (int64_t j = i; j < n; ++j) {
      int32_t mt = 0, ln = 0;
      if (orc_nw_pair(residues + offsets[i], offsets[i + 1] - offsets[i], residues + offsets[j],
                      offsets[j + 1] - offsets[j], S, go, ge, &mt, &ln)) return 1;
      double s = (double)mt / ln; /* 0/0 -> NaN for two empty strings, as :311 */
      out[i + j * n] = s;
      out[j + i * n] = s;
    }
  return 0;
}

/* ------------------------------------------------------------------ R pipeline (R/minHash.R) */
int orc_mh_signatures_linear(const int32_t* ranks, const int64_t* rank_offsets, int64_t n,
                             const int64_t* a, const int64_t* b, int64_t m, int n_hash, uint32_t* sig) {
  /* R/minHash.R:126-143 with apply_hash (:104-106): sig[h, doc] = min over the doc's shingle ranks x
   * (1-based) of (a_h*x + b_h) %% m.  The reference iterates vocabulary rows and pmin()s; the minimum
   * over the set of present rows is the same value.  64-bit arithmetic (R would give NA once a*x+b
   * exceeds 2^31-1 with integer inputs; callers stay below that for bit-exact comparisons). */
  if (m <= 0) { snprintf(g_err, sizeof g_err, "bad modulus"); return 1; }
  for (int64_t d = 0; d < n; ++d) {
    uint32_t* row = sig + d * (int64_t)n_hash;
    for (int h = 0; h < n_hash; ++h) {
      uint64_t best = UINT64_MAX;
      for (int64_t p = rank_offsets[d]; p < rank_offsets[d + 1]; ++p) {
        uint64_t v = ((uint64_t)a[h] * (uint64_t)ranks[p] + (uint64_t)b[h]) % (uint64_t)m;
        if (v < best) best = v;
      }
      row[h] = best == UINT64_MAX ? UINT32_MAX : (uint32_t)best;
    }
  }
  return 0;
}

void orc_mh_distance_matrix(const uint32_t* sig, int64_t n, int n_hash, double* out) {
  /* R/minHash.R:166-182: d[i,j] = 1 - mean(sig[,i]==sig[,j]), diagonal 0.  R's mean() of a logical
   * vector accumulates in long double and divides in long double before rounding to double. */
  for (int64_t i = 0; i < n; ++i) {
    out[i + i * n] = 0.0;
    for (int64_t j = i + 1; j < n; ++j) {
      int cnt = 0;
      for (int h = 0; h < n_hash; ++h) cnt += (sig[i * (int64_t)n_hash + h] == sig[j * (int64_t)n_hash + h]);
      double sim = (double)((long double)cnt / (long double)n_hash);
      out[i + j * n] = 1.0 - sim;
      out[j + i * n] = 1.0 - sim;
    }
  }
}
