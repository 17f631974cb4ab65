// Needleman-Wunsch identity kernels (sm_100a): launch interface.  See nw_kernels.cu.
#pragma once
#include "common.cuh"

namespace dyna {

// A unit of work: one row sequence (the lower-index sequence, on the DP rows) against a run of
// consecutive column sequences j_begin .. j_begin + j_count - 1 (all >= row).
struct NwUnit {
  int32_t row;
  int32_t j_begin;
  int32_t j_count;  // two-rows kernels: column count in the low 16 bits, (second row - row) in the high 16 bits
};
// two-rows units: the second row sequence is row + delta, delta >= 1 (columns below it are computed for `row` alone)
inline int32_t nw_pack_count(int64_t count, int64_t row_delta) { return (int32_t)(count | (row_delta << 16)); }

struct NwDeviceData {
  const uint8_t* codes;  // residues encoded 0..23, all sequences back to back
  const int32_t* off;    // n+1 offsets into codes
  const int8_t* sub;     // 24x24 substitution table, row-major [row residue][column residue]
  int64_t n;
  int64_t slab_base;     // packed-triangle (diagonal included) index of the plan's first pair
  uint32_t* matches;     // outputs, slab order
  uint32_t* length;
  int gap_open, gap_ext;
  uint32_t one;          // always 1; passed as data so the compiler keeps it in a register (see strip_column)
  uint32_t zero;         // always 0; an opaque addend that lives in a uniform register / constant operand (see stat_select)
  uint32_t bias16;       // nw_rows2_kernel: 0 = signed 16-bit lanes; otherwise the offset of its unsigned domain (strip_column4)
};

constexpr int kNwThreadMaxRows = 32;  // rows handled by the thread-per-pair kernel
constexpr int kNwWarpMaxR = 24;       // rows per lane of the warp-per-pair kernel (32*24 = 768 rows per pass)
constexpr int kNwWarp2MaxCols = 1024;  // column-sequence length limit of the packed warp kernel (shared-memory staging)
constexpr int kNwWarp2MaxR = 20;      // strip height limit of the two-pairs-per-warp 16-bit kernel (register budget)
constexpr int kNwWarpUnitPairs = 32;  // pairs per unit (8 warps x 4) of the 32-bit warp kernel
// column sequences per unit of the packed single-pass warp kernel: 64 by default, kNwWarp2UnitColsMax for large inputs
// (chosen in cabi.cu); DYNA_NW_UNITCOLS overrides it for measurements
constexpr int kNwWarp2UnitCols = 64;
constexpr int kNwWarp2UnitColsMax = 256;
constexpr int kNwThreadUnitPairs = 512;
constexpr int kNwMultiPassGrid = 148 * 2;

// Strip height for a row sequence of length m (m >= 1): <= 32 -> thread kernel (R = m rounded up to 4),
// otherwise the warp kernel with R = ceil(m/32) capped at kNwWarpMaxR (longer rows take several passes).
inline bool nw_use_thread_kernel(int m) { return m <= kNwThreadMaxRows; }
inline int nw_thread_R(int m) { return ((m + 3) / 4) * 4; }
inline int nw_warp_R(int m) { int r = (m + 31) / 32; return r > kNwWarpMaxR ? kNwWarpMaxR : r; }

int launch_nw_empty_rows(const NwDeviceData& d, const NwUnit* d_units, int num_units, cudaStream_t st);
int launch_nw_thread(int R, bool slant, const NwDeviceData& d, const NwUnit* d_units, int num_units, cudaStream_t st);
// scratch: only for multipass (rows longer than 32*R): kNwMultiPassGrid * 8 warps * 3 * max_cols ints
int launch_nw_warp(int R, bool slant, bool multipass, const NwDeviceData& d, const NwUnit* d_units, int num_units, int32_t* d_scratch,
                   int max_cols, cudaStream_t st);
// two pairs per warp in 16-bit lanes (host guarantees the value range); R in 2..kNwWarp2MaxR
int launch_nw_warp2(int R, const NwDeviceData& d, const NwUnit* d_units, int num_units, cudaStream_t st);
int launch_nw_thread2(int R, const NwDeviceData& d, const NwUnit* d_units, int num_units, cudaStream_t st);
// two rows (units[].row and row+1, both <= 32 residues) per thread against one column sequence each; units of up to
// 2 * kNwThreadUnitPairs columns
int launch_nw_thread_rows2(int R, const NwDeviceData& d, const NwUnit* d_units, int num_units, cudaStream_t st);
// packed kernel, several passes of 32*R rows (R in 7..12); scratch: kNwMultiPassGrid * 32 pair-sets * kNwWarp2MpMaxCols * 16 bytes
int launch_nw_warp2mp(int R, const NwDeviceData& d, const NwUnit* d_units, int num_units, void* d_scratch, cudaStream_t st);
// cooperative packed kernel: two warps share one 64-lane wavefront (rows 385..768, R = ceil(m/64) in 7..12, so that the
// second warp always owns rows); a unit is one row against up to kNwCoUnitCols column sequences
constexpr int kNwCoUnitCols = 128;
constexpr int kNwCoMinRows = 32 * 12 + 1;
constexpr int kNwCoMaxRows = 64 * 12;
inline int nw_co_R(int m) { return (m + 63) / 64; }
int launch_nw_warp2co(int R, const NwDeviceData& d, const NwUnit* d_units, int num_units, cudaStream_t st);
// two-rows packed kernel: rows i and i+1 (33..384 residues each, strips R <= 12) against the same column sequences;
// a unit is the row pair against up to kNwRows2UnitCols column sequences (units[].row is the first row)
constexpr int kNwRows2UnitCols = 256;
constexpr int kNwRows2MaxCols = 2048;  // its column-sequence length limit (one staged sequence per warp)
int launch_nw_rows2(int R, const NwDeviceData& d, const NwUnit* d_units, int num_units, cudaStream_t st);
// two-rows multi-pass kernel: row pairs of 769..kNwRows2MpMaxRows residues, both ending in the same (last) pass of 32*R rows,
// R = nw_rows2mp_R(longer row) in 7..12; persistent grid of kNwRows2MpGrid CTAs of 16 warps; scratch: one line of
// kNwRows2MaxCols 16-byte entries per warp (it fits the scratch of launch_nw_warp2mp and shares it: same stream)
constexpr int kNwRows2MpMaxRows = 32 * 12 * 8;
constexpr int kNwRows2MpGrid = 148;
inline int nw_rows2mp_passes(int m) { return (m + 32 * 12 - 1) / (32 * 12); }
inline int nw_rows2mp_R(int m) { const int np = nw_rows2mp_passes(m); const int r = (m + 32 * np - 1) / (32 * np); return r < 7 ? 7 : r; }
int launch_nw_rows2mp(int R, const NwDeviceData& d, const NwUnit* d_units, int num_units, void* d_scratch, cudaStream_t st);
// cooperative two-rows kernel: row pairs of 385..768 residues (R = ceil(max/64) <= 12); units of up to kNwCoUnitCols columns
constexpr int kNwRows2CoMaxRows = 64 * 12;
// its column-sequence length limit: the staging buffers share the SM's 227 KB with 172 KB of tables from R = 11 on
constexpr int nw_rows2co_max_cols(int R) { return R <= 10 ? 2048 : 1024; }
int launch_nw_rows2co(int R, const NwDeviceData& d, const NwUnit* d_units, int num_units, cudaStream_t st);
constexpr int kNwWarp2MpMaxRows = 32 * 12 * 8;  // 8 passes at most
constexpr int kNwWarp2MpMaxCols = 2048;        // its column-sequence limit (staging buffer)
}  // namespace dyna
