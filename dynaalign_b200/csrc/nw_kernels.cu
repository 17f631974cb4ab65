// Needleman-Wunsch identity kernels for sm_100a.
//
// What is computed (reference: calculate_similarity, src/pairwiseSeqAlign.cpp:209-313): a three-state affine-gap
// global alignment whose result is NOT the score but (matches, alignment_length) of the traceback path, with the
// reference's exact quirks -- M overwritten by the winning state (:273-278), tie-break D >= U >= L (:271-279),
// border gaps one extension cheaper than interior gaps (:226,:233 vs :255,:260), raw-character equality for
// matches (:294), row sequence = lower index (:340-346).
//
// How (no traceback matrix): every cell has exactly one predecessor, so the pair (matches, #diagonal steps) is
// carried forward with the scores; alignment_length = m + n - #diagonal steps.  Per cell, with
//   F = Ix[i][j]  (carried down the column, already holding max(H_up - (go+ge), F_up - ge)),
//   E = Iy[i][j]  (carried along the row, same form), diag = max(M,Ix,Iy)[i-1][j-1]:
//   Mraw = diag + s                       IADD
//   H    = max3(Mraw, F, E)               VIMNMX3            (DPX)
//   D    = (Mraw == H); U = (F >= E)      2 x ISETP
//   stat = D ? stat_diag + inc : (U ? stat_up : stat_left)   IADD + 2 x SEL   (inc = 1 | eq << 16)
//   E'   = max(H - (go+ge), E - ge)       IADD + VIADDMNMX   (DPX)
//   F'   = max(H - (go+ge), F - ge)       IADD + VIADDMNMX   (DPX)
// Storing E'/F' (the value the NEXT cell needs) instead of Iy/Ix makes the reference's asymmetric borders pure
// initial values: the border's "open source" M[i][0] = M[0][j] = NEG never has to coexist with the border's
// "diagonal source" max(M,Ix,Iy) = -go-(k-1)ge in one register.
//
// The substitution scores come from a query profile of the ROW sequence staged in shared memory:
// prof[c][row] = (int8 S[a_row][c], uint8 a_row == c) for the 24 residue classes c, so one step of a lane needs a
// single contiguous read of R 16-bit entries, decoded with two PRMT per cell (sign-extended score, stat increment).
//
//   K5 nw_warp_kernel<R>    one warp per pair: lane l owns rows [l*R, l*R+R) in registers and walks the columns;
//                           lane l is one column behind lane l-1 (anti-diagonal wavefront) and receives the bottom
//                           row of its upper neighbour (H, F, stat) through __shfl_up_sync.  A CTA shares one row
//                           sequence (one profile) across its 8 warps.  Rows longer than 32*R take several passes,
//                           the last lane spilling its bottom row to a per-warp global scratch line.
//   K4 nw_thread_kernel<R>  one thread per pair for short rows (<= 32 residues): the whole column strip lives in
//                           the thread's registers, no shuffles; lanes of a warp share the row sequence, so
//                           profile reads are bank-conflict free (distinct residue class -> distinct bank).
#include "nw_kernels.cuh"

#include <type_traits>

#include <algorithm>
#include <climits>
#include <cstdlib>

namespace dyna {
namespace {

constexpr int kNeg = INT_MIN / 2;  // the reference's "minus infinity" (src/pairwiseSeqAlign.cpp:216)

__device__ __forceinline__ int wadd(int a, int b) { return (int)((unsigned)a + (unsigned)b); }
__device__ __forceinline__ int wsub(int a, int b) { return (int)((unsigned)a - (unsigned)b); }
__device__ __forceinline__ int wmul(int a, int b) { return (int)((unsigned)a * (unsigned)b); }

// border diagonal source: max(M,Ix,Iy) at (k,0) or (0,k):  Bd(0) = 0, Bd(k) = -go - (k-1)*ge
__device__ __forceinline__ int border_diag(int k, int go, int ge) { return k == 0 ? 0 : wsub(-go, wmul(k - 1, ge)); }

template <int R>
struct Strip {
  static constexpr int RP = R + (R & 1);      // 16-bit entries per lane strip, padded to even
  static constexpr int RW = RP / 2;           // 32-bit words actually read per step
  static constexpr int RWS = (RP / 2) | 1;    // word stride between strips: odd -> conflict-free across lanes
  // stride for 64-bit profile loads: 8-byte aligned and == 2 (mod 4) words, so the 16 lanes of a half-warp wavefront
  // hit 16 distinct bank pairs (2, 6, 10 words)
  static constexpr int RWS64 = ((RW + 1) / 4) * 4 + 2;
};

// PTX prmt with full selector semantics (bit 3 of a selector nibble replicates the sign of the selected byte;
// the __byte_perm intrinsic masks that bit away).
template <uint32_t SEL>
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "n"(SEL));
  return d;
}

// Gap-extension slanting.  With SLANT every DP value of cell (i,j) is stored as value + (i+j)*ge.  All
// comparisons inside a cell share the offset, the extension term cancels from both gap recurrences
//   E'(i,j+1) = max(H - go, E)      F'(i+1,j) = max(H - go, F)        (one VIADDMNMX each, no IADD)
// and the diagonal step picks up + 2*ge, which is folded into the profile's score byte.  Exact as long as
// nothing overflows (checked on the host); without SLANT the textbook form with explicit "- ge" is used.
template <bool SLANT>
struct Gap {
  int neg_open;  // added to H when a gap opens:  SLANT ? -go : -(go+ge)
  int ext;       // subtracted when a gap extends: SLANT ? 0 : ge
  int sl;        // slant per unit of (i+j):       SLANT ? ge : 0
};

// One column of a strip of R rows: reads the column to the left (Ho, So), writes this column (Hn, Sn); E is
// updated in place.  Ping-ponging H and S between two register sets keeps the loop free of register moves.
template <int R, bool SLANT>
__device__ __forceinline__ void strip_column(const int (&Ho)[R], const uint32_t (&So)[R], int (&Hn)[R], uint32_t (&Sn)[R],
                                             int (&El)[R], const uint32_t (&pw)[Strip<R>::RW], int diagH, uint32_t diagS,
                                             int F, uint32_t upS, const Gap<SLANT>& g, uint32_t one, int& outF) {
#pragma unroll
  for (int k = 0; k < R; ++k) {
    const uint32_t w = pw[k >> 1];
    // entry layout (16 bit): low byte = int8 score (+ 2*ge when slanted), high byte = 1 if residues are equal
    const int s = (k & 1) ? (int)prmt<0xAAA2>(w, 0u) : (int)prmt<0x8880>(w, 0u);
    const uint32_t inc = (k & 1) ? prmt<0x5354>(w, one) : prmt<0x5154>(w, one);  // 1 | eq << 16
    const int E = El[k];
    const int Mraw = diagH + s;
    const int H = __vimax3_s32(Mraw, F, E);
    const bool pD = (Mraw == H);  // Mraw >= F && Mraw >= E
    const bool pU = (F >= E);
    uint32_t S = pU ? upS : So[k];
    if (pD) S = diagS + inc;
    diagH = Ho[k];
    diagS = So[k];
    Hn[k] = H;
    Sn[k] = S;
    if (SLANT) {
      El[k] = __viaddmax_s32(H, g.neg_open, E);
      F = __viaddmax_s32(H, g.neg_open, F);
    } else {
      El[k] = __viaddmax_s32(H, g.neg_open, E - g.ext);
      F = __viaddmax_s32(H, g.neg_open, F - g.ext);
    }
    upS = S;
  }
  outF = F;
}

// profile of rows [row0, row0 + LANES*R) of the row sequence (length m) into shared memory, strip-major per class
template <int R, int LANES, int STRIDE_WORDS = Strip<R>::RWS>
__device__ __forceinline__ void build_profile(uint32_t* prof, const uint8_t* __restrict__ a, int m, int row0,
                                              const int8_t* __restrict__ sub, int bias, int tid, int nthreads) {
  using S = Strip<R>;
  uint16_t* p16 = reinterpret_cast<uint16_t*>(prof);
  constexpr int per_class = LANES * STRIDE_WORDS * 2;  // 16-bit entries per residue class (incl. padding)
  for (int idx = tid; idx < 24 * LANES * S::RP; idx += nthreads) {
    const int c = idx / (LANES * S::RP);
    const int rem = idx - c * (LANES * S::RP);
    const int lane = rem / S::RP, k = rem - lane * S::RP;
    const int r = row0 + lane * R + k;
    uint16_t e = 0;
    if (k < R && r < m) {
      const int ar = a[r];
      e = (uint16_t)((uint8_t)(sub[ar * 24 + c] + bias)) | (uint16_t)((ar == c) ? 0x0100 : 0);
    }
    p16[c * per_class + lane * (STRIDE_WORDS * 2) + k] = e;
  }
}

__device__ __forceinline__ int64_t pair_slot(int64_t n, int64_t i, int64_t j, int64_t slab_base) {
  return i * n - i * (i - 1) / 2 + (j - i) - slab_base;
}

template <bool SLANT>
__device__ __forceinline__ Gap<SLANT> make_gap(int go, int ge) {
  Gap<SLANT> g;
  g.neg_open = SLANT ? wsub(0, go) : wsub(0, wadd(go, ge));
  g.ext = SLANT ? 0 : ge;
  g.sl = SLANT ? ge : 0;
  return g;
}
// Ix[1][j] = Iy[i][1] = max(NEG-(go+ge), NEG-ge)  (src/pairwiseSeqAlign.cpp:255-262 applied to the border)
__device__ __forceinline__ int neg_init_value(int go, int ge) { return max(wsub(kNeg, wadd(go, ge)), wsub(kNeg, ge)); }

// ------------------------------------------------------------------------------------------------
// K5: one warp per pair
// ------------------------------------------------------------------------------------------------
constexpr int kWarpThreads = 256;

template <int R, bool SLANT, bool MULTIPASS>
__global__ void __launch_bounds__(kWarpThreads)
nw_warp_kernel(NwDeviceData d, const NwUnit* __restrict__ units, int num_units, int32_t* __restrict__ scratch,
               int max_cols) {
  using S = Strip<R>;
  __shared__ uint32_t prof[24 * 32 * S::RWS];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int nwarps = kWarpThreads / 32;
  const int go = d.gap_open, ge = d.gap_ext;
  const Gap<SLANT> g = make_gap<SLANT>(go, ge);
  const int neg_init = neg_init_value(go, ge);
  const uint32_t one = d.one;  // opaque 1 (kept in a register so PRMT can take an immediate selector)
  const unsigned full = 0xFFFFFFFFu;

  for (int u = blockIdx.x; u < num_units; u += gridDim.x) {
    const NwUnit un = units[u];
    const int row = un.row;
    const int m = d.off[row + 1] - d.off[row];
    const uint8_t* a = d.codes + d.off[row];
    const int npass = MULTIPASS ? (m + 32 * R - 1) / (32 * R) : 1;
    // multipass units carry at most one pair per warp, so one scratch line per warp persists across passes
    int32_t* wscr = MULTIPASS ? scratch + ((int64_t)blockIdx.x * nwarps + warp) * 3 * (int64_t)max_cols : nullptr;

    for (int pass = 0; pass < npass; ++pass) {
      const int row0 = pass * 32 * R;
      __syncthreads();  // previous profile no longer in use
      build_profile<R, 32>(prof, a, m, row0, d.sub, 2 * g.sl, tid, kWarpThreads);
      __syncthreads();
      const bool last_pass = (pass == npass - 1);
      const int lm = last_pass ? (m - 1 - row0) / R : 31;  // last lane with real rows in this pass
      const int km = (m - 1 - row0) - lm * R;              // row of (m, .) inside lane lm (last pass only)
      const int r0 = row0 + lane * R;                      // 0-based first row of this lane's strip

      for (int jj = warp; jj < un.j_count; jj += nwarps) {
        const int j = un.j_begin + jj;
        const int n = d.off[j + 1] - d.off[j];
        const uint8_t* __restrict__ b = d.codes + d.off[j];

        // two register sets for (H, S): step t reads set (t&1), writes set ((t+1)&1); both start as the border
        int H0[R], H1[R], El[R];
        uint32_t S0[R], S1[R];
#pragma unroll
        for (int k = 0; k < R; ++k) {
          const int i = r0 + k + 1;                                     // 1-based DP row
          H0[k] = H1[k] = wadd(wsub(-go, wmul(i - 1, ge)), wmul(i, g.sl));  // Bd(i) at column 0
          El[k] = wadd(neg_init, wmul(i + 1, g.sl));                    // Iy[i][1]
          S0[k] = S1[k] = 0u;
        }
        int prevUpH = wadd(border_diag(r0, go, ge), wmul(r0, g.sl));  // diagonal source (r0, 0)
        uint32_t prevUpS = 0u;
        int outH = 0, outF = 0;
        uint32_t outS = 0u;
        const int T = n + lm;
        for (int t0 = 0; t0 < T; t0 += 2) {
#pragma unroll
          for (int ph = 0; ph < 2; ++ph) {
            const int t = t0 + ph;
            const int jc = t - lane;  // 0-based column handled by this lane in this step
            int rH = __shfl_up_sync(full, outH, 1);
            int rF = __shfl_up_sync(full, outF, 1);
            uint32_t rS = __shfl_up_sync(full, outS, 1);
            if (lane == 0) {
              if (!MULTIPASS || pass == 0) {
                rH = wadd(wsub(-go, wmul(t, ge)), wmul(t + 1, g.sl));  // Bd(t+1) at (0, t+1)
                rF = wadd(neg_init, wmul(t + 2, g.sl));                // Ix[1][t+1]
                rS = 0u;
              } else if (t < n) {
                rH = wscr[3 * t + 0];
                rF = wscr[3 * t + 1];
                rS = (uint32_t)wscr[3 * t + 2];
              }
            }
            if (jc >= 0 && jc < n && lane <= lm) {
              const int c = b[jc];
              uint32_t pw[S::RW];
              const uint32_t* pp = prof + c * (32 * S::RWS) + lane * S::RWS;
#pragma unroll
              for (int w = 0; w < S::RW; ++w) pw[w] = pp[w];
              // parity of the step as seen by this lane: its first column always reads set (lane&1)... use t
              if (((t0 + ph) & 1) == 0) {
                strip_column<R, SLANT>(H0, S0, H1, S1, El, pw, prevUpH, prevUpS, rF, rS, g, one, outF);
                outH = H1[R - 1];
                outS = S1[R - 1];
              } else {
                strip_column<R, SLANT>(H1, S1, H0, S0, El, pw, prevUpH, prevUpS, rF, rS, g, one, outF);
                outH = H0[R - 1];
                outS = S0[R - 1];
              }
              prevUpH = rH;
              prevUpS = rS;
              if (MULTIPASS && !last_pass && lane == 31) {  // bottom row of this pass feeds lane 0 of the next
                wscr[3 * jc + 0] = outH;
                wscr[3 * jc + 1] = outF;
                wscr[3 * jc + 2] = (int32_t)outS;
              }
            }
          }
        }
        if (last_pass) {
          // lane lm finished its last column at step t = lm + n - 1, which wrote set ((lm + n) & 1)
          const bool in1 = (((lm + n) & 1) != 0);
          uint32_t res = 0u;
#pragma unroll
          for (int k = 0; k < R; ++k)
            if (k == km) res = in1 ? S1[k] : S0[k];
          res = __shfl_sync(full, res, lm);
          if (lane == 0) {
            const int64_t slot = pair_slot(d.n, row, j, d.slab_base);
            d.matches[slot] = res >> 16;
            d.length[slot] = (uint32_t)(m + n) - (res & 0xFFFFu);
          }
        }
        if (MULTIPASS) __syncwarp();
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K5x2: one warp per TWO pairs, 16-bit lanes (s16x2 DPX)
//
// The 32-bit kernel is bound by the ALU pipe (8 of its 10 instructions per cell issue there at 16 lanes/clk per
// sub-partition).  When every DP value provably fits 16 bits (host check: slanted values live in
// [smin - 3go + 2ge, smax*min(m,n) + (m+n)*ge], see cabi.cu), two column sequences share a warp: the low half of
// every score register belongs to pair A = (row, jA), the high half to pair B = (row, jB), and VIADDMNMX.S16x2 /
// VIMNMX.S16x2 (which returns both ">=" predicates) update two cells per instruction:
//     sP   = PRMT(wA, wB)                 packed sign-extended scores                     ALU
//     Mraw = VIADDMNMX.S16x2(diag, sP, MIN)                                               ALU
//     g    = VIMNMX.S16x2(F, E)  -> pU_A, pU_B                                            ALU
//     H    = VIMNMX.S16x2(Mraw, g) -> pD_A, pD_B                                          ALU
//     E'   = VIADDMNMX.S16x2(H, -go, E);  F' = VIADDMNMX.S16x2(H, -go, F)                 ALU x2
//     per pair: inc = PRMT, S = SEL(pU), @pD S = diagS + inc                              ALU x2, other x1
// = 10 ALU-pipe instructions per two cells instead of 16.  The traceback statistics stay 32-bit per pair.
// The shorter of the two column sequences is padded with residue class 24 (zero profile) and its result is
// captured when its last real column has been processed; the values computed past its end are never used.
// "Minus infinity" is a flat sentinel (-30000) in the slanted domain: it only has to lose every comparison once.
// ------------------------------------------------------------------------------------------------
constexpr int kSentinel16 = -30000;

__device__ __forceinline__ uint32_t pack16(int v) { return ((uint32_t)v & 0xFFFFu) | ((uint32_t)v << 16); }

// A unit field read again where it is needed (opaque load) instead of being carried in a register across the step loop:
// the two-rows kernels sit at the 128-register limit of 512-thread CTAs, and one more live value cost 6 % (measured when
// NwUnit::row2 replaced "row + 1").
__device__ __forceinline__ int reload_i32(const int32_t* p) {
  int v;
  asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}

// stat update without touching the ALU pipe: S = left + cn; if (up) S = upS + cn; if (diag) S = diagS + inc,
// written as one plain and two predicated 2-input adds (a SEL or a MOV would issue on the ALU pipe, which is the
// pipe this kernel saturates; 2-input adds issue on the other integer pipes).  cn is the per-pair constant added on
// a non-diagonal step (0, or the "non-diagonal step" counter increment of variant 2).
__device__ __forceinline__ uint32_t stat_select(uint32_t left, uint32_t up, uint32_t dsum_a, uint32_t dsum_b, bool pu, bool pd,
                                                uint32_t cn) {
  uint32_t S;
  asm("{\n\t.reg .pred pu, pd;\n\t"
      "setp.ne.u32 pu, %5, 0;\n\t"
      "setp.ne.u32 pd, %6, 0;\n\t"
      "add.u32 %0, %1, %7;\n\t"
      "@pu add.u32 %0, %2, %7;\n\t"
      "@pd add.u32 %0, %3, %4;\n\t}"
      : "=&r"(S)
      : "r"(left), "r"(up), "r"(dsum_a), "r"(dsum_b), "r"((uint32_t)pu), "r"((uint32_t)pd), "r"(cn));
  return S;
}

// the same update with one select (ALU pipe) and one predicated add: one instruction less, one more on the ALU pipe.
// nw_rows2_kernel is bound by issue slots with the ALU pipe at ~70 %, so a FEW rows per column can afford the trade
// (rows chosen by the bit masks DYNA_ROWS2_SELMASK_A / _B).
__device__ __forceinline__ uint32_t stat_select_sel(uint32_t left, uint32_t up, uint32_t dsum_a, uint32_t dsum_b, bool pu, bool pd) {
  uint32_t S;
  asm("{\n\t.reg .pred pu, pd;\n\t"
      "setp.ne.u32 pu, %5, 0;\n\t"
      "setp.ne.u32 pd, %6, 0;\n\t"
      "selp.b32 %0, %2, %1, pu;\n\t"
      "@pd add.u32 %0, %3, %4;\n\t}"
      : "=&r"(S)
      : "r"(left), "r"(up), "r"(dsum_a), "r"(dsum_b), "r"((uint32_t)pu), "r"((uint32_t)pd));
  return S;
}

// Variants of the per-pair statistics update (same results):
//   VAR 1: increments (1 | eq << 16) permuted out of the profile word (two PRMT per two cells); select by three adds
//   VAR 2: increments read ready-made from a shared-memory table, 128 bits per four rows (strips R <= 12)
// (Measured and dropped: SEL + predicated add, 2.51 vs 2.64 TCUPS; both increments from one permute through
//  IMAD / IMAD.HI, 2.38 TCUPS.)
struct Stat2Consts {
  uint32_t one, zero;
};

// Shared-memory loads through explicit 32-bit shared-window addresses (the generic-pointer form made ptxas rebuild
// the window base with four uniform-datapath instructions in every step of the hot loop).
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint2 lds_v2(uint32_t addr) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint4 lds_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}

template <int R, int VAR>
__device__ __forceinline__ void strip_column2(const uint32_t (&Ho)[R], uint32_t (&Hn)[R], uint32_t (&El)[R],
                                              const uint32_t (&SAo)[R], uint32_t (&SAn)[R], const uint32_t (&SBo)[R],
                                              uint32_t (&SBn)[R], const uint32_t (&pwA)[Strip<R>::RW],
                                              const uint32_t (&pwB)[Strip<R>::RW], uint32_t diagH, uint32_t dSA,
                                              uint32_t dSB, uint32_t F, uint32_t upSA, uint32_t upSB, uint32_t ngo2,
                                              const Stat2Consts& c, uint32_t& outF, uint32_t incA_sh = 0u,
                                              uint32_t incB_sh = 0u) {
  // VAR 2: the stat increments (1 | eq << 16) are read ready-made from a shared-memory table, four rows per 128-bit
  // load, instead of being permuted out of the profile word: two fewer ALU-pipe instructions per two cells
  uint4 qa = make_uint4(0, 0, 0, 0), qb = make_uint4(0, 0, 0, 0);
#pragma unroll
  for (int k = 0; k < R; ++k) {
    const uint32_t wA = pwA[k >> 1], wB = pwB[k >> 1];
    const uint32_t sP = (k & 1) ? prmt<0xE6A2>(wA, wB) : prmt<0xC480>(wA, wB);  // [sext16(sA) | sext16(sB) << 16]
    uint32_t incA, incB;
    if (VAR == 2) {
      if ((k & 3) == 0) {
        qa = lds_v4(incA_sh + 4u * (unsigned)k);  // rows k..k+3 (constant offset after unrolling)
        qb = lds_v4(incB_sh + 4u * (unsigned)k);
      }
      incA = (k & 3) == 0 ? qa.x : (k & 3) == 1 ? qa.y : (k & 3) == 2 ? qa.z : qa.w;
      incB = (k & 3) == 0 ? qb.x : (k & 3) == 1 ? qb.y : (k & 3) == 2 ? qb.z : qb.w;
    } else {
      incA = (k & 1) ? prmt<0x5354>(wA, c.one) : prmt<0x5154>(wA, c.one);  // 1 | eq << 16
      incB = (k & 1) ? prmt<0x5354>(wB, c.one) : prmt<0x5154>(wB, c.one);
    }
    const uint32_t E = El[k];
    const uint32_t Mraw = __viaddmax_s16x2(diagH, sP, 0x80008000u);
    bool puB, puA, pdB, pdA;
    const uint32_t g = __vibmax_s16x2(F, E, &puB, &puA);   // pred_hi -> pair B, pred_lo -> pair A
    const uint32_t H = __vibmax_s16x2(Mraw, g, &pdB, &pdA);
    const uint32_t SA = stat_select(SAo[k], upSA, dSA, incA, puA, pdA, c.zero);
    const uint32_t SB = stat_select(SBo[k], upSB, dSB, incB, puB, pdB, c.zero);
    diagH = Ho[k];
    dSA = SAo[k];
    dSB = SBo[k];
    Hn[k] = H;
    SAn[k] = SA;
    SBn[k] = SB;
    El[k] = __viaddmax_s16x2(H, ngo2, E);
    F = __viaddmax_s16x2(H, ngo2, F);
    upSA = SA;
    upSB = SB;
  }
  outF = F;
}

// VAR 3: ONE table.  Per residue class and lane a record of RW score-profile words (four int8 scores each) followed
// by R increment words,
// padded to a stride of 4*odd words, read only with 128-bit loads (conflict-free: the 8 lanes of a quarter-warp
// wavefront hit 8 distinct 16-byte bank groups whatever class each lane reads, because the class stride is a multiple
// of 128 bytes).  Against VAR 2 a column costs one address computation per sequence instead of two and
// ceil((RW+R)/4) loads instead of RW + ceil(R/4); ncu on the first cooperative kernel showed the 32-bit tail loads of
// VAR 2 (9th increment row, 5th profile word) paying 4 and 2 wavefronts each at strides 12 and 6.
template <int R>
struct Rec {
  static constexpr int RW = (R + 3) / 4;       // profile words: FOUR int8 scores per word (the equality flag of the
                                               // 16-bit entries is not needed here -- the increments carry it)
  static constexpr int kWords = RW + R;
  static constexpr int NQ = (kWords + 3) / 4;  // 128-bit loads per record
  static constexpr int kStride = (NQ | 1) * 4; // words per lane record: a multiple of 4, an odd multiple
  // first strip row that uses a word of quad q (profile word w serves rows 4w .. 4w+3; increment word RW+k row k)
  __host__ __device__ static constexpr int first_needed(int q) {
    int best = R;
    for (int i = 4 * q; i < 4 * q + 4 && i < kWords; ++i) {
      const int row = i < RW ? 4 * i : i - RW;
      if (row < best) best = row;
    }
    return best;
  }
  __host__ __device__ static constexpr int load_row(int q) { return first_needed(q) >= 2 ? first_needed(q) - 2 : 0; }  // two rows ahead
};

template <int R, int LANES>
__device__ __forceinline__ void build_records(uint32_t* rec, const uint8_t* __restrict__ a, int m, int row0,
                                              const int8_t* __restrict__ sub, int bias, int tid, int nthreads) {
  using RC = Rec<R>;
  for (int idx = tid; idx < 25 * LANES * RC::kStride; idx += nthreads) {
    const int cls = idx / (LANES * RC::kStride);
    const int rem = idx - cls * (LANES * RC::kStride);
    const int ln = rem / RC::kStride, w = rem - ln * RC::kStride;
    uint32_t v = 0u;
    if (w < RC::RW) {  // four int8 scores (+ bias): rows 4w .. 4w+3 of the lane's strip
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        const int k = 4 * w + h, r = row0 + ln * R + k;
        if (k < R && r < m && cls < 24) v |= (uint32_t)(uint8_t)(sub[a[r] * 24 + cls] + bias) << (8 * h);
      }
    } else if (w < RC::kWords) {  // 1 | (row residue == class) << 16; padding rows and class 24 count steps only
      const int k = w - RC::RW, r = row0 + ln * R + k;
      v = 1u | ((r < m && cls < 24 && a[r] == cls) ? 0x10000u : 0u);
    }
    rec[idx] = v;
  }
}

template <int R>
__device__ __forceinline__ void strip_column3(const uint32_t (&Ho)[R], uint32_t (&Hn)[R], uint32_t (&El)[R],
                                              const uint32_t (&SAo)[R], uint32_t (&SAn)[R], const uint32_t (&SBo)[R],
                                              uint32_t (&SBn)[R], uint32_t recA, uint32_t recB, uint32_t diagH,
                                              uint32_t dSA, uint32_t dSB, uint32_t F, uint32_t upSA, uint32_t upSB,
                                              uint32_t ngo2, const Stat2Consts& c, uint32_t& outF) {
  using RC = Rec<R>;
  uint32_t wa[RC::NQ * 4], wb[RC::NQ * 4];
#pragma unroll
  for (int k = 0; k < R; ++k) {
#pragma unroll
    for (int q = 0; q < RC::NQ; ++q) {
      if (RC::load_row(q) == k) {
        const uint4 va = lds_v4(recA + 16u * (unsigned)q), vb = lds_v4(recB + 16u * (unsigned)q);
        wa[4 * q + 0] = va.x; wa[4 * q + 1] = va.y; wa[4 * q + 2] = va.z; wa[4 * q + 3] = va.w;
        wb[4 * q + 0] = vb.x; wb[4 * q + 1] = vb.y; wb[4 * q + 2] = vb.z; wb[4 * q + 3] = vb.w;
      }
    }
    const uint32_t wA = wa[k >> 2], wB = wb[k >> 2];
    // [sext16(sA) | sext16(sB) << 16] from byte k&3 of both words (selector bit 3 replicates the sign)
    const uint32_t sP = (k & 3) == 0 ? prmt<0xC480>(wA, wB) : (k & 3) == 1 ? prmt<0xD591>(wA, wB)
                      : (k & 3) == 2 ? prmt<0xE6A2>(wA, wB) : prmt<0xF7B3>(wA, wB);
    const uint32_t incA = wa[RC::RW + k], incB = wb[RC::RW + k];
    const uint32_t E = El[k];
    const uint32_t Mraw = __viaddmax_s16x2(diagH, sP, 0x80008000u);
    bool puB, puA, pdB, pdA;
    const uint32_t g = __vibmax_s16x2(F, E, &puB, &puA);
    const uint32_t H = __vibmax_s16x2(Mraw, g, &pdB, &pdA);
    const uint32_t SA = stat_select(SAo[k], upSA, dSA, incA, puA, pdA, c.zero);
    const uint32_t SB = stat_select(SBo[k], upSB, dSB, incB, puB, pdB, c.zero);
    diagH = Ho[k];
    dSA = SAo[k];
    dSB = SBo[k];
    Hn[k] = H;
    SAn[k] = SA;
    SBn[k] = SB;
    El[k] = __viaddmax_s16x2(H, ngo2, E);
    F = __viaddmax_s16x2(H, ngo2, F);
    upSA = SA;
    upSB = SB;
  }
  outF = F;
}

// INPLACE: one register set for (H, S) instead of the ping-pong pair -- 4R instead of 7R state registers at the price
// of three register moves per row (they issue as IMAD.MOV, off the ALU pipe).  Used for tall strips (R >= 13), where
// the ping-pong version drops to one CTA per SM.
// Per-step overhead matters (an 11-row strip is only ~150 instructions per step), so the step loop avoids ALU-pipe work
// that is not the recurrence: both column sequences are staged per warp in shared memory (the shorter one padded with
// class 24, so no bounds test and no 64-bit address arithmetic per step), the active test is one unsigned compare, and
// the border row reaches lane 0 through a ROTATING shuffle from lane 31, which holds the border constants whenever the
// strip layout leaves it idle (m <= 31*R) -- no per-step select for lane 0.
constexpr int kNwStageCols = 1024;

// dynamic shared memory layout of nw_warp2_kernel
template <int R, int VAR, int THREADS, int STAGE = kNwStageCols>
struct Warp2Smem {
  // increment-table stride per lane strip: a multiple of 4 words (128-bit loads) and an ODD multiple (conflict-free
  // across the 8 lanes of a quarter warp): 4, 12 or 20
  static constexpr int kIncStride = R <= 4 ? 4 : (R <= 12 ? 12 : 20);
  // score-profile stride per lane strip (odd word count: conflict-free 32-bit loads).  A 128-bit-load layout of the
  // score profile was measured slower (2.76 vs 2.94 TCUPS) and is not kept.
  // VAR 2 reads it with 64-bit loads (stride == 2 mod 4 words: conflict-free for the 16 lanes of a wavefront).
  static constexpr bool kProf64 = false;  // measured on the SASS: the register pairs of 64-bit loads cost more moves than the loads save (R = 11: 422 vs 394 instructions per two steps)
  static constexpr int kProfStride = kProf64 ? Strip<R>::RWS64 : Strip<R>::RWS;
  static constexpr int kProfBytes = VAR == 3 ? 0 : 25 * 32 * kProfStride * 4;
  static constexpr int kIncBytes = VAR == 2 ? 25 * 32 * kIncStride * 4 : (VAR == 3 ? 25 * 32 * Rec<R>::kStride * 4 : 0);
  static constexpr int kStageBytes = (THREADS / 32) * 2 * (STAGE + 8);
  static constexpr int kIncOff = 0;                                  // 16-byte aligned first
  static constexpr int kProfOff = kIncOff + kIncBytes;
  static constexpr int kStageOff = kProfOff + kProfBytes;
  static constexpr int kTotal = kStageOff + kStageBytes;
};  // column-sequence length limit of the packed warp kernel (host-checked)

template <int R, int VAR, bool INPLACE, int THREADS>
__global__ void __launch_bounds__(THREADS, THREADS >= 256 ? 2 : 3)
nw_warp2_kernel(NwDeviceData d, const NwUnit* __restrict__ units, int num_units) {
  using S = Strip<R>;
  using L = Warp2Smem<R, VAR, THREADS>;
  constexpr int nwarps = THREADS / 32;
  // VAR 2 needs ~77 KB (dynamic shared memory); the other variants stay within the 48 KB static limit, which also
  // gives the compiler constant shared addresses (measured 2.5 % faster than the dynamic form)
  uint32_t* prof;        // score profile, class 24 = padding residue (all-zero entries)
  uint32_t* incT;        // VAR 2: [class][lane][kIncStride] stat increments
  uint8_t* stage_base;   // per-warp staged column sequences
  if constexpr (VAR == 2 || VAR == 3) {
    extern __shared__ __align__(16) unsigned char smem_dyn[];
    prof = reinterpret_cast<uint32_t*>(smem_dyn + L::kProfOff);
    incT = reinterpret_cast<uint32_t*>(smem_dyn + L::kIncOff);  // VAR 3: the record table (profile words + increments)
    stage_base = smem_dyn + L::kStageOff;
  } else {
    __shared__ uint32_t prof_s[25 * 32 * L::kProfStride];
    __shared__ uint8_t stage_s[L::kStageBytes];
    prof = prof_s;
    incT = prof_s;
    stage_base = stage_s;
  }
  // results of one unit (<= 64 consecutive pair slots), written back as two coalesced 256-byte rows instead of one
  // 4-byte store per pair (ncu: the scattered stores cost 30 % extra DRAM write traffic plus read-modify-write reads)
  __shared__ uint32_t res_m[kNwWarp2UnitColsMax], res_l[kNwWarp2UnitColsMax];
  // the unit's column sequences ordered by decreasing length: the two sequences that share a warp then differ by a
  // residue or two instead of the ~11 of a random pairing (the shorter one idles for the difference)
  __shared__ int col_len[kNwWarp2UnitColsMax];
  __shared__ uint8_t col_ord[kNwWarp2UnitColsMax];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int go = d.gap_open, ge = d.gap_ext;
  const uint32_t ngo2 = pack16(-go);
  const uint32_t sent2 = pack16(kSentinel16);
  const uint32_t bord2 = pack16(ge - go);  // slanted border: Bd(k) + k*ge = -go + ge for every k >= 1
  Stat2Consts c;  // opaque constants (from a kernel parameter) so the compiler keeps them in registers
  c.one = d.one;
  c.zero = d.one - 1u;
  const unsigned full = 0xFFFFFFFFu;
  const int src_lane = (lane + 31) & 31;  // rotating "shuffle up": lane 0 reads lane 31
  uint8_t* sA = stage_base + (warp * 2 + 0) * (kNwStageCols + 8);
  uint8_t* sB = stage_base + (warp * 2 + 1) * (kNwStageCols + 8);
  const uint32_t sA_sh = (uint32_t)__cvta_generic_to_shared(sA), sB_sh = (uint32_t)__cvta_generic_to_shared(sB);

  for (int u = blockIdx.x; u < num_units; u += gridDim.x) {
    const NwUnit un = units[u];
    const int row = un.row;
    const int m = d.off[row + 1] - d.off[row];
    __syncthreads();
    for (int q = tid; q < un.j_count; q += THREADS) col_len[q] = d.off[un.j_begin + q + 1] - d.off[un.j_begin + q];
    if constexpr (VAR == 3) {
      build_records<R, 32>(incT, d.codes + d.off[row], m, 0, d.sub, 2 * ge, tid, THREADS);
    } else {
      build_profile<R, 32, L::kProfStride>(prof, d.codes + d.off[row], m, 0, d.sub, 2 * ge, tid, THREADS);
      for (int idx = tid; idx < 32 * L::kProfStride; idx += THREADS) prof[24 * 32 * L::kProfStride + idx] = 0u;
    }
    if (VAR == 2) {  // increment table: 1 | (row residue == class) << 16; padding rows and class 24 count steps only
      const uint8_t* __restrict__ a = d.codes + d.off[row];
      for (int idx = tid; idx < 25 * 32 * L::kIncStride; idx += THREADS) {
        const int cls = idx / (32 * L::kIncStride);
        const int rem = idx - cls * (32 * L::kIncStride);
        const int ln = rem / L::kIncStride, k = rem - ln * L::kIncStride;
        const int r = ln * R + k;
        incT[idx] = 1u | ((k < R && r < m && cls < 24 && a[r] == cls) ? 0x10000u : 0u);
      }
    }
    __syncthreads();
    for (int me = tid; me < un.j_count; me += THREADS) {  // rank by (length descending, index ascending): a permutation
      const int mine = col_len[me];
      int rank = 0;
      for (int q = 0; q < un.j_count; ++q) {
        const int other = col_len[q];
        rank += (other > mine || (other == mine && q < me)) ? 1 : 0;
      }
      col_ord[rank] = (uint8_t)me;
    }
    __syncthreads();
    const int lm = (m - 1) / R;
    const int km = (m - 1) - lm * R;
    const int r0 = lane * R;
    const bool rot = (lm < 31);  // lane 31 idle: it can hold the border row for lane 0
    const int npairs2 = (un.j_count + 1) >> 1;
    const uint32_t plane_sh = (uint32_t)__cvta_generic_to_shared(prof + lane * L::kProfStride);
    const uint32_t ilane_sh = (uint32_t)__cvta_generic_to_shared(incT + lane * (VAR == 3 ? Rec<R>::kStride : L::kIncStride));

    for (int pp = warp; pp < npairs2; pp += nwarps) {
      const bool hasB = (2 * pp + 1 < un.j_count);
      int jA = un.j_begin + col_ord[2 * pp];
      int jB = hasB ? un.j_begin + col_ord[2 * pp + 1] : jA;
      int nA = d.off[jA + 1] - d.off[jA], nB = d.off[jB + 1] - d.off[jB];
      if (nB > nA) {  // A is the longer column sequence
        int tj = jA; jA = jB; jB = tj;
        int tn = nA; nA = nB; nB = tn;
      }
      {  // stage both column sequences (B padded with the zero-profile class up to nA)
        const uint8_t* __restrict__ bA = d.codes + d.off[jA];
        const uint8_t* __restrict__ bB = d.codes + d.off[jB];
        __syncwarp();
        for (int q = lane; q < nA; q += 32) {
          sA[q] = bA[q];
          sB[q] = (q < nB) ? bB[q] : (uint8_t)24;
        }
        __syncwarp();
      }

      constexpr int R2 = INPLACE ? 1 : R;  // second register set only for the ping-pong version
      uint32_t H0[R], H1[R2], El[R], SA0[R], SA1[R2], SB0[R], SB1[R2];
#pragma unroll
      for (int k = 0; k < R; ++k) {
        H0[k] = bord2;
        El[k] = sent2;
        SA0[k] = SB0[k] = 0u;  // border column: no diagonal step yet
        if (!INPLACE) {
          H1[k] = bord2;
          SA1[k] = SB1[k] = 0u;
        }
      }
      uint32_t prevUpH = (r0 == 0) ? 0u : bord2;
      uint32_t prevUpSA = 0u, prevUpSB = 0u;
      // lane 31 (idle when rot) carries the border row: H_diag = -go+ge (slanted), F = sentinel, stats 0
      // what the lane below reads is always the bottom row written in the previous step, i.e. H/SA/SB[R-1] of the set
      // this phase reads -- shuffled straight out of the strip registers (no copies); only F needs its own register.
      // An idle lane 31 never writes its strip, so it keeps offering the border constants it was initialised with.
      uint32_t outF = sent2;
      uint32_t resB = 0u;
      const unsigned n_act = (lane <= lm) ? (unsigned)nA : 0u;  // columns this lane processes
      const int capB = (lane == lm) ? nB - 1 : -1;              // column at which pair B's result is final
      const int T = nA + lm;
      // the step loop exists twice, specialised on ROT (warp-uniform): with the rotation the border row costs nothing,
      // without it lane 0 overrides four registers per step -- selects that would otherwise sit on the ALU pipe of
      // every step of every unit
      auto run_steps = [&](auto rot_c) {
      constexpr bool ROT = decltype(rot_c)::value;
      for (int t0 = 0; t0 < T; t0 += 2) {
#pragma unroll
        for (int ph = 0; ph < 2; ++ph) {
          const int jc = t0 + ph - lane;
          const bool from1 = !INPLACE && ph == 1;  // the register set the previous step wrote
          uint32_t rH = __shfl_sync(full, from1 ? H1[INPLACE ? 0 : R - 1] : H0[R - 1], src_lane);
          uint32_t rF = __shfl_sync(full, outF, src_lane);
          uint32_t rSA = __shfl_sync(full, from1 ? SA1[INPLACE ? 0 : R - 1] : SA0[R - 1], src_lane);
          uint32_t rSB = __shfl_sync(full, from1 ? SB1[INPLACE ? 0 : R - 1] : SB0[R - 1], src_lane);
          if (!ROT) {  // all 32 lanes own rows: lane 0 takes the border row explicitly
            if (lane == 0) {
              rH = bord2;
              rF = sent2;
              rSA = 0u;
              rSB = 0u;
            }
          }
          if ((unsigned)jc < n_act) {
            const uint32_t cA = lds_u8(sA_sh + (uint32_t)jc), cB = lds_u8(sB_sh + (uint32_t)jc);
            if constexpr (VAR == 3) {
              const uint32_t ra = ilane_sh + cA * (32u * Rec<R>::kStride * 4u);
              const uint32_t rb = ilane_sh + cB * (32u * Rec<R>::kStride * 4u);
              if (ph == 0) {
                strip_column3<R>(H0, H1, El, SA0, SA1, SB0, SB1, ra, rb, prevUpH, prevUpSA, prevUpSB, rF, rSA, rSB, ngo2, c, outF);
              } else {
                strip_column3<R>(H1, H0, El, SA1, SA0, SB1, SB0, ra, rb, prevUpH, prevUpSA, prevUpSB, rF, rSA, rSB, ngo2, c, outF);
              }
            } else {
              uint32_t pwA[S::RW], pwB[S::RW];
              const uint32_t pa = plane_sh + cA * (32u * L::kProfStride * 4u);
              const uint32_t pb = plane_sh + cB * (32u * L::kProfStride * 4u);
              const uint32_t ia = ilane_sh + cA * (32u * L::kIncStride * 4u);
              const uint32_t ib = ilane_sh + cB * (32u * L::kIncStride * 4u);
              if constexpr (L::kProf64) {
  #pragma unroll
                for (int w = 0; w < S::RW; w += 2) {  // an odd RW reads one padding word of the lane's stride
                  const uint2 va = lds_v2(pa + 4u * (unsigned)w), vb = lds_v2(pb + 4u * (unsigned)w);
                  pwA[w] = va.x;
                  pwB[w] = vb.x;
                  if (w + 1 < S::RW) {
                    pwA[w + 1] = va.y;
                    pwB[w + 1] = vb.y;
                  }
                }
              } else {
  #pragma unroll
                for (int w = 0; w < S::RW; ++w) {
                  pwA[w] = lds_u32(pa + 4u * (unsigned)w);
                  pwB[w] = lds_u32(pb + 4u * (unsigned)w);
                }
              }
              if constexpr (INPLACE) {
                strip_column2<R, VAR>(H0, H0, El, SA0, SA0, SB0, SB0, pwA, pwB, prevUpH, prevUpSA, prevUpSB, rF, rSA, rSB,
                                      ngo2, c, outF, ia, ib);
              } else if (ph == 0) {
                strip_column2<R, VAR>(H0, H1, El, SA0, SA1, SB0, SB1, pwA, pwB, prevUpH, prevUpSA, prevUpSB, rF, rSA, rSB,
                                      ngo2, c, outF, ia, ib);
              } else {
                strip_column2<R, VAR>(H1, H0, El, SA1, SA0, SB1, SB0, pwA, pwB, prevUpH, prevUpSA, prevUpSB, rF, rSA, rSB,
                                      ngo2, c, outF, ia, ib);
              }
            }
            prevUpH = rH;
            prevUpSA = rSA;
            prevUpSB = rSB;
            if (jc == capB) {  // the shorter sequence ends here: capture its result
#pragma unroll
              for (int k = 0; k < R; ++k)
                if (k == km) resB = (!INPLACE && ph == 0) ? SB1[INPLACE ? 0 : k] : SB0[k];
            }
          }
        }
      }
      };
      if (rot) run_steps(std::true_type{});
      else run_steps(std::false_type{});
      const bool in1 = (((lm + nA) & 1) != 0);
      uint32_t resA = 0u;
#pragma unroll
      for (int k = 0; k < R; ++k)
        if (k == km) resA = (!INPLACE && in1) ? SA1[INPLACE ? 0 : k] : SA0[k];
      resA = __shfl_sync(full, resA, lm);
      resB = __shfl_sync(full, resB, lm);
      if (lane == 0) {
        // stat word: matches << 16 | diag steps;  length = m + n - diag
        res_m[jA - un.j_begin] = resA >> 16;
        res_l[jA - un.j_begin] = (uint32_t)(m + nA) - (resA & 0xFFFFu);
        if (hasB) {
          res_m[jB - un.j_begin] = resB >> 16;
          res_l[jB - un.j_begin] = (uint32_t)(m + nB) - (resB & 0xFFFFu);
        }
      }
    }
    __syncthreads();
    {  // the unit's pairs occupy consecutive slots of the packed triangle
      const int64_t slot0 = pair_slot(d.n, row, un.j_begin, d.slab_base);
      for (int q = tid; q < 2 * un.j_count; q += THREADS) {
        if (q < un.j_count) d.matches[slot0 + q] = res_m[q];
        else d.length[slot0 + q - un.j_count] = res_l[q - un.j_count];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K5x2 "two rows": the two 16-bit halves of every score register belong to two ROW sequences -- rows i and i+1 of the
// triangle -- aligned against the SAME column sequence, instead of one row sequence against two column sequences.
// Both halves then see the same residue class c in every column, so
//   * the packed score operand [s(a2_k, c) : s(a1_k, c)] is ONE precomputed table word: the PRMT that merged the scores
//     of two classes (an ALU-pipe instruction per row, the pipe this kernel loads most) is gone;
//   * there is one staged column sequence, one residue load, one table address and one record stream per step;
//   * the shorter-column padding (class 24, result capture at the shorter sequence's last column, pairing the columns
//     by length) disappears: both alignments end in the same column, only in different rows.
// Record per (class, lane): for each of the R strip rows two words -- packed scores and the increment both pairs share
// (Rec2 below; three words with one increment per pair until the statistics word was re-laid: 3826 -> see DESIGN.md) --
// read with 128-bit loads a few rows ahead of their use (stride 4*odd words: conflict-free whatever class each lane
// reads).  2R words per lane and class is ~86 KB for R = 11; with the staged sequences a CTA owns a whole SM: 16
// warps, one unit = the two rows against up to 256 column sequences (16 per warp).
// The two rows have different lengths m1, m2: the strip layout follows the longer one, the rows the shorter one lacks
// are zero-score padding below its last row and are never read back (each result is taken from its own last row).
// Column j = i (the first row against itself) has no counterpart for row i+1; that half is computed and dropped.
// ------------------------------------------------------------------------------------------------
#ifndef DYNA_ROWS2_THREADS
#define DYNA_ROWS2_THREADS 512
#endif
#ifndef DYNA_TROWS2_SELMASK_A  // the same trade in nw_thread_rows2_kernel (short probes)
#define DYNA_TROWS2_SELMASK_A 0xFFFF
#endif
#ifndef DYNA_TROWS2_SELMASK_B
#define DYNA_TROWS2_SELMASK_B 0
#endif
#ifndef DYNA_TROWS2_SELMASK_UA  // unsigned domain
#define DYNA_TROWS2_SELMASK_UA 0xFFFF
#endif
#ifndef DYNA_TROWS2_SELMASK_UB
#define DYNA_TROWS2_SELMASK_UB 0x00FF
#endif
#ifndef DYNA_ROWS2_SELMASK_A  // signed lanes (five DPX instructions per row on the ALU pipe): 7 rows of the first pair
#define DYNA_ROWS2_SELMASK_A 0x7F
#endif
#ifndef DYNA_ROWS2_SELMASK_B
#define DYNA_ROWS2_SELMASK_B 0
#endif
#ifndef DYNA_ROWS2_SELMASK_UA  // unsigned domain (four DPX instructions per row): all rows of the first pair, 7 of the second
#define DYNA_ROWS2_SELMASK_UA 0xFFF
#endif
#ifndef DYNA_ROWS2_SELMASK_UB
#define DYNA_ROWS2_SELMASK_UB 0x7F
#endif
#ifndef DYNA_ROWS2_LOOKAHEAD
#define DYNA_ROWS2_LOOKAHEAD 2
#endif
constexpr int kRows2Threads = DYNA_ROWS2_THREADS;

// Statistics word of the two-rows kernels: bits 0..9 matches of pair 1 (row i), bits 10..19 matches of pair 2 (row i+1),
// bits 20..31 diagonal steps.  With this layout ONE increment word per strip row and residue class,
//     1 << 20 | (a2_k == c) << 10 | (a1_k == c),
// serves both pairs: pair 1's statistics also count pair 2's matches along pair 1's path (and vice versa) in the field
// they do not own, which is never read back.  A field grows by at most one per diagonal step and a path has at most
// min(m, n) <= 768 of them, so no field carries into its neighbour.  Two words per row (scores, increment) instead of
// three: the record stream -- the shared-memory pipe was at 85-90 % with three -- shrinks by a third.
constexpr uint32_t kStat2DiagOne = 1u << 20;
static_assert(kNwRows2CoMaxRows <= 1023 && 32 * 12 <= 1023 && kNwThreadMaxRows <= 1023, "10-bit match fields of the two-rows statistics word");
__host__ __device__ __forceinline__ uint32_t inc2_word(bool eq1, bool eq2) {
  return kStat2DiagOne | (eq2 ? 1u << 10 : 0u) | (eq1 ? 1u : 0u);
}
__device__ __forceinline__ uint32_t stat2_matches1(uint32_t s) { return s & 0x3FFu; }
__device__ __forceinline__ uint32_t stat2_matches2(uint32_t s) { return (s >> 10) & 0x3FFu; }
__device__ __forceinline__ uint32_t stat2_diag(uint32_t s) { return s >> 20; }

template <int R>
struct Rec2 {
  static constexpr int kWords = 2 * R;
  static constexpr int NQ = (kWords + 3) / 4;   // 128-bit loads per record
  static constexpr int kStride = (NQ | 1) * 4;  // words per lane record: a multiple of 4, an odd multiple
  static constexpr int kTableBytes = 24 * 32 * kStride * 4;
  static constexpr int kStageBytes = (kRows2Threads / 32) * (kNwRows2MaxCols + 8);  // two records' words leave room for 2048 columns
  static constexpr int kTotal = kTableBytes + kStageBytes;
  __host__ __device__ static constexpr int first_needed(int q) { return 2 * q; }  // row of the quad's first word
  __host__ __device__ static constexpr int load_row(int q) {
    return first_needed(q) >= DYNA_ROWS2_LOOKAHEAD ? first_needed(q) - DYNA_ROWS2_LOOKAHEAD : 0;
  }
};

// U (unsigned domain, strip_column4): the score word is the INTEGER sum s2 * 65536 + s1, not the concatenation of two
// 16-bit two's-complement halves -- the diagonal step is a plain 32-bit add there, and a negative s1 written as a
// 16-bit half would carry into the upper half; as an integer sum the borrow is already in the word (equal for scores >= 0).
__host__ __device__ __forceinline__ uint32_t score2_word(int s1, int s2, bool u) {
  return u ? (uint32_t)(s2 * 65536 + s1) : (((uint32_t)s1 & 0xFFFFu) | ((uint32_t)s2 << 16));
}

template <int R, bool U>
__device__ __forceinline__ void build_records2(uint32_t* rec, const uint8_t* __restrict__ a1, int m1,
                                               const uint8_t* __restrict__ a2, int m2, const int8_t* __restrict__ sub, int bias,
                                               int tid, int nthreads) {
  using RC = Rec2<R>;
  for (int idx = tid; idx < 24 * 32 * RC::kStride; idx += nthreads) {
    const int cls = idx / (32 * RC::kStride);
    const int rem = idx - cls * (32 * RC::kStride);
    const int ln = rem / RC::kStride, w = rem - ln * RC::kStride;
    uint32_t v = 0u;
    if (w < RC::kWords) {
      const int k = w >> 1, r = ln * R + k;
      if ((w & 1) == 0) {  // packed scores, sign-extended to 16 bits each; rows beyond a sequence's end score 0
        const int s1 = r < m1 ? (int)(int8_t)(sub[a1[r] * 24 + cls] + bias) : 0;
        const int s2 = r < m2 ? (int)(int8_t)(sub[a2[r] * 24 + cls] + bias) : 0;
        v = score2_word(s1, s2, U);
      } else {
        v = inc2_word(r < m1 && a1[r] == cls, r < m2 && a2[r] == cls);
      }
    }
    rec[idx] = v;
  }
}

// U: the UNSIGNED domain.  Every DP value is stored + d.bias16 so that it is a positive 16-bit number and "minus
// infinity" is 0.  The diagonal step diagH + sP is then a plain 32-bit add -- off the ALU pipe that the four remaining DPX
// instructions and the selects share -- because both halves of the sum stay inside [0, 65535] (the table word is the
// integer sum of the two scores, see score2_word: negative scores included), and the compares are the .U16x2 forms.
// Same predicates, same statistics.
template <int R, bool U = false>
__device__ __forceinline__ void strip_column4(const uint32_t (&Ho)[R], uint32_t (&Hn)[R], uint32_t (&El)[R],
                                              const uint32_t (&SAo)[R], uint32_t (&SAn)[R], const uint32_t (&SBo)[R],
                                              uint32_t (&SBn)[R], uint32_t rec_sh, uint32_t diagH, uint32_t dSA, uint32_t dSB,
                                              uint32_t F, uint32_t upSA, uint32_t upSB, uint32_t ngo2, const Stat2Consts& c,
                                              uint32_t& outF) {
  using RC = Rec2<R>;
  uint32_t w[RC::NQ * 4];
#pragma unroll
  for (int k = 0; k < R; ++k) {
#pragma unroll
    for (int q = 0; q < RC::NQ; ++q) {
      if (RC::load_row(q) == k) {
        const uint4 v = lds_v4(rec_sh + 16u * (unsigned)q);
        w[4 * q + 0] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
      }
    }
    const uint32_t sP = w[2 * k], incA = w[2 * k + 1], incB = incA;
    const uint32_t E = El[k];
    const uint32_t Mraw = U ? diagH + sP : __viaddmax_s16x2(diagH, sP, 0x80008000u);
    bool puB, puA, pdB, pdA;
    // pred_hi -> pair 2 (row i+1), pred_lo -> pair 1 (row i)
    const uint32_t g = U ? __vibmax_u16x2(F, E, &puB, &puA) : __vibmax_s16x2(F, E, &puB, &puA);
    const uint32_t H = U ? __vibmax_u16x2(Mraw, g, &pdB, &pdA) : __vibmax_s16x2(Mraw, g, &pdB, &pdA);
    const uint32_t SA = (((U ? DYNA_ROWS2_SELMASK_UA : DYNA_ROWS2_SELMASK_A) >> k) & 1) ? stat_select_sel(SAo[k], upSA, dSA, incA, puA, pdA)
                                                 : stat_select(SAo[k], upSA, dSA, incA, puA, pdA, c.zero);
    const uint32_t SB = (((U ? DYNA_ROWS2_SELMASK_UB : DYNA_ROWS2_SELMASK_B) >> k) & 1) ? stat_select_sel(SBo[k], upSB, dSB, incB, puB, pdB)
                                                     : stat_select(SBo[k], upSB, dSB, incB, puB, pdB, c.zero);
    diagH = Ho[k];
    dSA = SAo[k];
    dSB = SBo[k];
    Hn[k] = H;
    SAn[k] = SA;
    SBn[k] = SB;
    El[k] = U ? __viaddmax_u16x2(H, ngo2, E) : __viaddmax_s16x2(H, ngo2, E);
    F = U ? __viaddmax_u16x2(H, ngo2, F) : __viaddmax_s16x2(H, ngo2, F);
    upSA = SA;
    upSB = SB;
  }
  outF = F;
}

template <int R, bool U>
__global__ void __launch_bounds__(kRows2Threads, 1)
nw_rows2_kernel(NwDeviceData d, const NwUnit* __restrict__ units, int num_units) {
  using RC = Rec2<R>;
  constexpr int nwarps = kRows2Threads / 32;
  extern __shared__ __align__(16) unsigned char smem_dyn[];
  uint32_t* rec = reinterpret_cast<uint32_t*>(smem_dyn);
  uint8_t* stage_base = smem_dyn + RC::kTableBytes;
  __shared__ uint32_t res_m1[kNwRows2UnitCols], res_l1[kNwRows2UnitCols], res_m2[kNwRows2UnitCols], res_l2[kNwRows2UnitCols];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int go = d.gap_open, ge = d.gap_ext;
  const uint32_t ngo2 = pack16(-go);
  Stat2Consts c;
  c.one = d.one;
  c.zero = d.zero;
  // register operands (see nw_thread2_kernel); U: everything + bias16, the sentinel is 0
  const uint32_t sent2 = (U ? 0u : pack16(kSentinel16)) + c.zero;
  const uint32_t bord2 = pack16(ge - go + (U ? (int)d.bias16 : 0));  // slanted border: -go + ge for every k >= 1
  const uint32_t corner2 = U ? pack16((int)d.bias16) : 0u;
  const unsigned full = 0xFFFFFFFFu;
  const int src_lane = (lane + 31) & 31;  // rotating "shuffle up": lane 0 reads lane 31
  uint8_t* sC = stage_base + warp * (kNwRows2MaxCols + 8);
  const uint32_t sC_sh = (uint32_t)__cvta_generic_to_shared(sC);
  const uint32_t rlane_sh = (uint32_t)__cvta_generic_to_shared(rec + lane * RC::kStride);

  for (int u = blockIdx.x; u < num_units; u += gridDim.x) {
    NwUnit un = units[u];
    const int row = un.row;
    int m2;
    const uint8_t* a2p;
    {
      const int row2 = row + (un.j_count >> 16);
      un.j_count &= 0xFFFF;
      m2 = d.off[row2 + 1] - d.off[row2];
      a2p = d.codes + d.off[row2];
    }
    const int m1 = d.off[row + 1] - d.off[row];
    __syncthreads();
    build_records2<R, U>(rec, d.codes + d.off[row], m1, a2p, m2, d.sub, 2 * ge, tid, kRows2Threads);
    __syncthreads();
    const int lmA = (m1 - 1) / R, kmA = (m1 - 1) - lmA * R;
    const int lmB = (m2 - 1) / R, kmB = (m2 - 1) - lmB * R;
    const int lm = max(lmA, lmB);
    const int r0 = lane * R;
    const bool rot = (lm < 31);  // lane 31 idle: it holds the border row for lane 0

    for (int pp = warp; pp < un.j_count; pp += nwarps) {
      const int j = un.j_begin + pp;
      const int n = d.off[j + 1] - d.off[j];
      {
        const uint8_t* __restrict__ b = d.codes + d.off[j];
        __syncwarp();
        for (int q = lane; q < n; q += 32) sC[q] = b[q];
        __syncwarp();
      }
      uint32_t H0[R], H1[R], El[R], SA0[R], SA1[R], SB0[R], SB1[R];
#pragma unroll
      for (int k = 0; k < R; ++k) {
        H0[k] = H1[k] = bord2;
        El[k] = sent2;
        SA0[k] = SA1[k] = SB0[k] = SB1[k] = 0u;
      }
      uint32_t prevUpH = (r0 == 0) ? corner2 : bord2;
      uint32_t prevUpSA = 0u, prevUpSB = 0u;
      uint32_t outF = sent2;
      const unsigned n_act = (lane <= lm) ? (unsigned)n : 0u;
      const int T = n + lm;
      auto run_steps = [&](auto rot_c) {
        constexpr bool ROT = decltype(rot_c)::value;
        for (int t0 = 0; t0 < T; t0 += 2) {
#pragma unroll
          for (int ph = 0; ph < 2; ++ph) {
            const int jc = t0 + ph - lane;
            uint32_t rH = __shfl_sync(full, ph == 1 ? H1[R - 1] : H0[R - 1], src_lane);
            uint32_t rF = __shfl_sync(full, outF, src_lane);
            uint32_t rSA = __shfl_sync(full, ph == 1 ? SA1[R - 1] : SA0[R - 1], src_lane);
            uint32_t rSB = __shfl_sync(full, ph == 1 ? SB1[R - 1] : SB0[R - 1], src_lane);
            if (!ROT) {  // all 32 lanes own rows: lane 0 takes the border row explicitly
              if (lane == 0) {
                rH = bord2;
                rF = sent2;
                rSA = 0u;
                rSB = 0u;
              }
            }
            if ((unsigned)jc < n_act) {
              const uint32_t cc = lds_u8(sC_sh + (uint32_t)jc);
              const uint32_t ra = rlane_sh + cc * (32u * RC::kStride * 4u);
              if (ph == 0) {
                strip_column4<R, U>(H0, H1, El, SA0, SA1, SB0, SB1, ra, prevUpH, prevUpSA, prevUpSB, rF, rSA, rSB, ngo2, c, outF);
              } else {
                strip_column4<R, U>(H1, H0, El, SA1, SA0, SB1, SB0, ra, prevUpH, prevUpSA, prevUpSB, rF, rSA, rSB, ngo2, c, outF);
              }
            }
            // Unconditional: what the upper neighbour delivered in this step is the diagonal source of the next one.  A
            // lane that has not started yet receives its neighbour's initial (= border) registers, which is exactly the
            // diagonal of its first column; lane 0 starts at step 0 with the corner.  Outside the branch the two phases
            // just alternate register names -- inside it these were three predicated moves per step.
            prevUpH = rH;
            prevUpSA = rSA;
            prevUpSB = rSB;
          }
        }
      };
      if (rot) run_steps(std::true_type{});
      else run_steps(std::false_type{});
      // every lane finished its last column at step lane + n - 1, which wrote register set ((lane + n) & 1)
      const bool in1 = (((lane + n) & 1) != 0);
      uint32_t resA = 0u, resB = 0u;
#pragma unroll
      for (int k = 0; k < R; ++k) {
        if (k == kmA) resA = in1 ? SA1[k] : SA0[k];
        if (k == kmB) resB = in1 ? SB1[k] : SB0[k];
      }
      resA = __shfl_sync(full, resA, lmA);
      resB = __shfl_sync(full, resB, lmB);
      if (lane == 0) {  // length = m + n - diagonal steps
        res_m1[pp] = stat2_matches1(resA);
        res_l1[pp] = (uint32_t)(m1 + n) - stat2_diag(resA);
        res_m2[pp] = stat2_matches2(resB);
        res_l2[pp] = (uint32_t)(m2 + n) - stat2_diag(resB);
      }
    }
    __syncthreads();
    {  // both rows' pairs occupy consecutive slots of the packed triangle; the columns below row2 exist for the first row only
      const int row2 = row + (reload_i32(&units[u].j_count) >> 16);
      const int64_t slot1 = pair_slot(d.n, row, un.j_begin, d.slab_base);
      const int skip = un.j_begin < row2 ? row2 - un.j_begin : 0;
      const int64_t slot2 = pair_slot(d.n, row2, un.j_begin + skip, d.slab_base) - skip;
      for (int q = tid; q < un.j_count; q += kRows2Threads) {
        d.matches[slot1 + q] = res_m1[q];
        d.length[slot1 + q] = res_l1[q];
        if (q >= skip) {
          d.matches[slot2 + q] = res_m2[q];
          d.length[slot2 + q] = res_l2[q];
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K5x2 "two rows", multi-pass: row pairs longer than the cooperative form holds (769..3072 residues).  The rows are walked
// in passes of 32*R rows with the strip code of nw_rows2_kernel; lane 31's bottom row (H, F', both statistics: 16 bytes per
// column) goes through a per-warp global scratch line (in place: lane 31 is 31 columns behind lane 0).  Lane 0 of the next
// pass does not load it from global memory step by step -- an L2 round trip is longer than a step of this kernel's four
// warps per scheduler, the first version ran at 2.67 TCUPS with the warp waiting for lane 0 in every step of passes 1.. --
// the warp copies it 32 entries at a time with cp.async into a 64-entry shared-memory ring, one block (32 steps) ahead, and
// lane 0 overwrites its four shuffled registers with one predicated 128-bit shared load (pass 0: a ring of border entries).  The record table belongs to one pass and to the whole CTA, so the 16
// warps move through the passes together: in every ROUND each warp takes ONE column sequence through all passes (the
// table is rebuilt 16 times more often than with the passes outermost, ~2 % of the work, and the scratch is 32 KB per warp
// instead of one line per column sequence of the unit).
// Records: three words per row (scores, increment of pair 1, increment of pair 2) and statistics words matches << 16 |
// diagonal steps -- the shared 10+10+12-bit increment of the single-pass kernels stops at 1023 diagonal steps.
// Host: both rows end in the last pass, 4 * shorter >= 3 * longer, columns <= kNwRows2MaxCols, unsigned domain.
// ------------------------------------------------------------------------------------------------
template <int R>
struct Rec3 {
  static constexpr int kWords = 3 * R;
  static constexpr int NQ = (kWords + 3) / 4;
  static constexpr int kStride = (NQ | 1) * 4;
  static constexpr int kTableBytes = 24 * 32 * kStride * 4;
  static constexpr int kStageBytes = (kRows2Threads / 32) * (kNwRows2MaxCols + 8);
  static constexpr int kRingBytes = (kRows2Threads / 32) * 64 * 16;  // per warp: 64 boundary entries of 16 bytes
  static constexpr int kTotal = kTableBytes + kStageBytes + kRingBytes;
  __host__ __device__ static constexpr int first_needed(int q) { return (4 * q) / 3; }
  __host__ __device__ static constexpr int load_row(int q) {
    return first_needed(q) >= DYNA_ROWS2_LOOKAHEAD ? first_needed(q) - DYNA_ROWS2_LOOKAHEAD : 0;
  }
};

// rows [0, 32R) of the pass: a1 / a2 point at the pass's first row, m1 / m2 are what is left of the sequences (<= 0: none)
template <int R, bool U>
__device__ __forceinline__ void build_records3(uint32_t* rec, const uint8_t* __restrict__ a1, int m1,
                                               const uint8_t* __restrict__ a2, int m2, const int8_t* __restrict__ sub, int bias,
                                               int tid, int nthreads) {
  using RC = Rec3<R>;
  for (int idx = tid; idx < 24 * 32 * RC::kStride; idx += nthreads) {
    const int cls = idx / (32 * RC::kStride);
    const int rem = idx - cls * (32 * RC::kStride);
    const int ln = rem / RC::kStride, w = rem - ln * RC::kStride;
    uint32_t v = 0u;
    if (w < RC::kWords) {
      const int k = w / 3, t = w - 3 * k, r = ln * R + k;
      if (t == 0) {
        const int s1 = r < m1 ? (int)(int8_t)(sub[a1[r] * 24 + cls] + bias) : 0;
        const int s2 = r < m2 ? (int)(int8_t)(sub[a2[r] * 24 + cls] + bias) : 0;
        v = score2_word(s1, s2, U);
      } else if (t == 1) {
        v = 1u | ((r < m1 && a1[r] == cls) ? 0x10000u : 0u);
      } else {
        v = 1u | ((r < m2 && a2[r] == cls) ? 0x10000u : 0u);
      }
    }
    rec[idx] = v;
  }
}

// strip_column4 on three-word records (one increment per pair)
template <int R, bool U>
__device__ __forceinline__ void strip_column4x(const uint32_t (&Ho)[R], uint32_t (&Hn)[R], uint32_t (&El)[R],
                                               const uint32_t (&SAo)[R], uint32_t (&SAn)[R], const uint32_t (&SBo)[R],
                                               uint32_t (&SBn)[R], uint32_t rec_sh, uint32_t diagH, uint32_t dSA, uint32_t dSB,
                                               uint32_t F, uint32_t upSA, uint32_t upSB, uint32_t ngo2, const Stat2Consts& c,
                                               uint32_t& outF) {
  using RC = Rec3<R>;
  uint32_t w[RC::NQ * 4];
#pragma unroll
  for (int k = 0; k < R; ++k) {
#pragma unroll
    for (int q = 0; q < RC::NQ; ++q) {
      if (RC::load_row(q) == k) {
        const uint4 v = lds_v4(rec_sh + 16u * (unsigned)q);
        w[4 * q + 0] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
      }
    }
    const uint32_t sP = w[3 * k], incA = w[3 * k + 1], incB = w[3 * k + 2];
    const uint32_t E = El[k];
    const uint32_t Mraw = U ? diagH + sP : __viaddmax_s16x2(diagH, sP, 0x80008000u);
    bool puB, puA, pdB, pdA;
    const uint32_t g = U ? __vibmax_u16x2(F, E, &puB, &puA) : __vibmax_s16x2(F, E, &puB, &puA);
    const uint32_t H = U ? __vibmax_u16x2(Mraw, g, &pdB, &pdA) : __vibmax_s16x2(Mraw, g, &pdB, &pdA);
    const uint32_t SA = ((DYNA_ROWS2_SELMASK_UA >> k) & 1) ? stat_select_sel(SAo[k], upSA, dSA, incA, puA, pdA)
                                                           : stat_select(SAo[k], upSA, dSA, incA, puA, pdA, c.zero);
    const uint32_t SB = ((DYNA_ROWS2_SELMASK_UB >> k) & 1) ? stat_select_sel(SBo[k], upSB, dSB, incB, puB, pdB)
                                                           : stat_select(SBo[k], upSB, dSB, incB, puB, pdB, c.zero);
    diagH = Ho[k];
    dSA = SAo[k];
    dSB = SBo[k];
    Hn[k] = H;
    SAn[k] = SA;
    SBn[k] = SB;
    El[k] = U ? __viaddmax_u16x2(H, ngo2, E) : __viaddmax_s16x2(H, ngo2, E);
    F = U ? __viaddmax_u16x2(H, ngo2, F) : __viaddmax_s16x2(H, ngo2, F);
    upSA = SA;
    upSB = SB;
  }
  outF = F;
}

template <int R, bool U>
__global__ void __launch_bounds__(kRows2Threads, 1)
nw_rows2mp_kernel(NwDeviceData d, const NwUnit* __restrict__ units, int num_units, uint4* __restrict__ scratch) {
  using RC = Rec3<R>;
  constexpr int nwarps = kRows2Threads / 32;
  extern __shared__ __align__(16) unsigned char smem_dyn[];
  uint32_t* rec = reinterpret_cast<uint32_t*>(smem_dyn);
  uint8_t* stage_base = smem_dyn + RC::kTableBytes;
  __shared__ uint32_t res_m1[kNwRows2UnitCols], res_l1[kNwRows2UnitCols], res_m2[kNwRows2UnitCols], res_l2[kNwRows2UnitCols];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int go = d.gap_open, ge = d.gap_ext;
  const uint32_t ngo2 = pack16(-go);
  Stat2Consts c;
  c.one = d.one;
  c.zero = d.zero;
  const uint32_t sent2 = (U ? 0u : pack16(kSentinel16)) + c.zero;
  const uint32_t bord2 = pack16(ge - go + (U ? (int)d.bias16 : 0));
  const uint32_t corner2 = U ? pack16((int)d.bias16) : 0u;
  const unsigned full = 0xFFFFFFFFu;
  uint8_t* sC = stage_base + warp * (kNwRows2MaxCols + 8);
  const uint32_t sC_sh = (uint32_t)__cvta_generic_to_shared(sC);
  const uint32_t rlane_sh = (uint32_t)__cvta_generic_to_shared(rec + lane * RC::kStride);
  uint4* scr = scratch + ((int64_t)blockIdx.x * nwarps + warp) * kNwRows2MaxCols;
  __shared__ int col_len[kNwRows2UnitCols];
  __shared__ uint8_t col_ord[kNwRows2UnitCols];
  static_assert(kNwRows2UnitCols <= 256 && kNwRows2UnitCols <= kRows2Threads, "column order: one byte, one thread per column");
  static_assert((RC::kTableBytes + RC::kStageBytes) % 16 == 0, "ring alignment");
  uint4* ring = reinterpret_cast<uint4*>(smem_dyn + RC::kTableBytes + RC::kStageBytes) + warp * 64;
  const uint32_t ring_sh = (uint32_t)__cvta_generic_to_shared(ring);

  for (int u = blockIdx.x; u < num_units; u += gridDim.x) {
    NwUnit un = units[u];
    const int row = un.row, row2 = row + (un.j_count >> 16);
    un.j_count &= 0xFFFF;
    const int m1 = d.off[row + 1] - d.off[row], m2 = d.off[row2 + 1] - d.off[row2];
    const uint8_t* __restrict__ a1 = d.codes + d.off[row];
    const uint8_t* __restrict__ a2 = d.codes + d.off[row2];
    const int npass = (max(m1, m2) + 32 * R - 1) / (32 * R);
    const int rounds = (un.j_count + nwarps - 1) / nwarps;
    // the warps of a round wait for each other at every pass: the unit's column sequences are taken in order of length,
    // so that a round holds 16 of about the same length (a proteome-like mix ran at 40 % efficiency in input order)
    __syncthreads();
    if (tid < un.j_count) col_len[tid] = d.off[un.j_begin + tid + 1] - d.off[un.j_begin + tid];
    __syncthreads();
    if (tid < un.j_count) {
      const int mine = col_len[tid];
      int rank = 0;
      for (int q = 0; q < un.j_count; ++q) {
        const int other = col_len[q];
        rank += (other > mine || (other == mine && q < tid)) ? 1 : 0;
      }
      col_ord[rank] = (uint8_t)tid;
    }
    __syncthreads();

    for (int rd = 0; rd < rounds; ++rd) {
      const bool has = rd * nwarps + warp < un.j_count;
      const int pp = has ? (int)col_ord[rd * nwarps + warp] : 0;
      int n = 0;
      if (has) {
        const int j = un.j_begin + pp;
        n = d.off[j + 1] - d.off[j];
        const uint8_t* __restrict__ b = d.codes + d.off[j];
        __syncwarp();
        for (int q = lane; q < n; q += 32) sC[q] = b[q];
        __syncwarp();
      }
      for (int pass = 0; pass < npass; ++pass) {
        const int row0 = pass * 32 * R;
        const bool last = (pass == npass - 1);
        __syncthreads();  // nobody reads the previous pass's table any more
        build_records3<R, U>(rec, a1 + row0, m1 - row0, a2 + row0, m2 - row0, d.sub, 2 * ge, tid, kRows2Threads);
        __syncthreads();
        if (!has) continue;
        const int lmA = last ? (m1 - 1 - row0) / R : 31, kmA = (m1 - 1 - row0) - lmA * R;
        const int lmB = last ? (m2 - 1 - row0) / R : 31, kmB = (m2 - 1 - row0) - lmB * R;
        const int lm = max(lmA, lmB);
        uint32_t H0[R], H1[R], El[R], SA0[R], SA1[R], SB0[R], SB1[R];
#pragma unroll
        for (int k = 0; k < R; ++k) {
          H0[k] = H1[k] = bord2;
          El[k] = sent2;
          SA0[k] = SA1[k] = SB0[k] = SB1[k] = 0u;
        }
        uint32_t prevUpH = (row0 == 0 && lane == 0) ? corner2 : bord2;
        uint32_t prevUpSA = 0u, prevUpSB = 0u;
        uint32_t outF = sent2;
        const unsigned n_act = (lane <= lm) ? (unsigned)n : 0u;
        const int T = n + lm;
        // what lane 0 sees above its first row, per column: ring entry (column & 63)
        auto request = [&](int first) {  // the warp asks for scratch entries first .. first + 31 (those that exist)
          const int idx = first + lane;
          if (idx < n)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(ring_sh + (uint32_t)(idx & 63) * 16u), "l"(scr + idx));
          asm volatile("cp.async.commit_group;");
        };
        __syncwarp();
        if (pass == 0) {
          ring[lane] = make_uint4(bord2, sent2, 0u, 0u);
          ring[lane + 32] = make_uint4(bord2, sent2, 0u, 0u);
        } else {
          request(0);
          request(32);
          asm volatile("cp.async.wait_all;" ::: "memory");
        }
        __syncwarp();
        for (int t0 = 0; t0 < T; t0 += 2) {
          if (pass > 0 && (t0 & 31) == 0 && t0 > 0) {  // lane 0 enters block t0 / 32, requested 32 steps ago; ask for the next one
            asm volatile("cp.async.wait_all;" ::: "memory");
            __syncwarp();
            request(t0 + 32);
          }
#pragma unroll
          for (int ph = 0; ph < 2; ++ph) {
            const int jc = t0 + ph - lane;
            uint32_t rH = __shfl_up_sync(full, ph == 1 ? H1[R - 1] : H0[R - 1], 1);
            uint32_t rF = __shfl_up_sync(full, outF, 1);
            uint32_t rSA = __shfl_up_sync(full, ph == 1 ? SA1[R - 1] : SA0[R - 1], 1);
            uint32_t rSB = __shfl_up_sync(full, ph == 1 ? SB1[R - 1] : SB0[R - 1], 1);
            asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, %4, 0;\n\t@p ld.shared.v4.u32 {%0, %1, %2, %3}, [%5];\n\t}"
                         : "+r"(rH), "+r"(rF), "+r"(rSA), "+r"(rSB)
                         : "r"((uint32_t)lane), "r"(ring_sh + (uint32_t)((t0 + ph) & 63) * 16u)
                         : "memory");
            if ((unsigned)jc < n_act) {
              const uint32_t cc = lds_u8(sC_sh + (uint32_t)jc);
              const uint32_t ra = rlane_sh + cc * (32u * RC::kStride * 4u);
              if (ph == 0) {
                strip_column4x<R, U>(H0, H1, El, SA0, SA1, SB0, SB1, ra, prevUpH, prevUpSA, prevUpSB, rF, rSA, rSB, ngo2, c, outF);
              } else {
                strip_column4x<R, U>(H1, H0, El, SA1, SA0, SB1, SB0, ra, prevUpH, prevUpSA, prevUpSB, rF, rSA, rSB, ngo2, c, outF);
              }
              if (!last && lane == 31)  // the bottom row just written (the other register set)
                scr[jc] = ph == 0 ? make_uint4(H1[R - 1], outF, SA1[R - 1], SB1[R - 1])
                                  : make_uint4(H0[R - 1], outF, SA0[R - 1], SB0[R - 1]);
            }
            prevUpH = rH;
            prevUpSA = rSA;
            prevUpSB = rSB;
          }
        }
        if (last) {
          const bool in1 = (((lane + n) & 1) != 0);
          uint32_t resA = 0u, resB = 0u;
#pragma unroll
          for (int k = 0; k < R; ++k) {
            if (k == kmA) resA = in1 ? SA1[k] : SA0[k];
            if (k == kmB) resB = in1 ? SB1[k] : SB0[k];
          }
          resA = __shfl_sync(full, resA, lmA);
          resB = __shfl_sync(full, resB, lmB);
          if (lane == 0) {  // stat word: matches << 16 | diagonal steps;  length = m + n - diagonal steps
            res_m1[pp] = resA >> 16;
            res_l1[pp] = (uint32_t)(m1 + n) - (resA & 0xFFFFu);
            res_m2[pp] = resB >> 16;
            res_l2[pp] = (uint32_t)(m2 + n) - (resB & 0xFFFFu);
          }
        }
        __syncwarp();  // lane 31's scratch entries of this pass before lane 0 reads them in the next
      }
    }
    __syncthreads();
    {
      const int64_t slot1 = pair_slot(d.n, row, un.j_begin, d.slab_base);
      const int skip = un.j_begin < row2 ? row2 - un.j_begin : 0;
      const int64_t slot2 = pair_slot(d.n, row2, un.j_begin + skip, d.slab_base) - skip;
      for (int q = tid; q < un.j_count; q += kRows2Threads) {
        d.matches[slot1 + q] = res_m1[q];
        d.length[slot1 + q] = res_l1[q];
        if (q >= skip) {
          d.matches[slot2 + q] = res_m2[q];
          d.length[slot2 + q] = res_l2[q];
        }
      }
    }
    __syncthreads();  // res_* are rewritten by the next unit
  }
}

// ------------------------------------------------------------------------------------------------
// K5x2 multi-pass: the packed two-pairs kernel for rows longer than one pass of 32*R rows (R <= 12 keeps the fast
// ping-pong / increment-table configuration).  The CTA walks the row sequence in passes of 32*R rows; within a pass
// every warp processes its pair-sets exactly like nw_warp2_kernel, except that
//   * lane 0 of pass p > 0 takes its upper neighbour (H, F, statA, statB per column) from a global scratch line that
//     lane 31 of pass p-1 wrote (in place: lane 31 writes column t-31 while lane 0 reads column t), and
//   * results are captured in the last pass only.
// Scratch: one 16-byte entry per column, per pair-set, per resident CTA (persistent grid).
// ------------------------------------------------------------------------------------------------
constexpr int kNwMpPairSets = 32;  // pair-sets (= 64 pairs) per unit at most
constexpr int kNwMpStageCols = 2048;  // column-sequence length limit of the multi-pass packed kernel

template <int R>
__global__ void __launch_bounds__(kWarpThreads, 2)
nw_warp2mp_kernel(NwDeviceData d, const NwUnit* __restrict__ units, int num_units, uint4* __restrict__ scratch) {
  using S = Strip<R>;
  using L = Warp2Smem<R, 2, kWarpThreads, kNwMpStageCols>;
  constexpr int nwarps = kWarpThreads / 32;
  extern __shared__ __align__(16) unsigned char smem_dyn[];
  uint32_t* prof = reinterpret_cast<uint32_t*>(smem_dyn + L::kProfOff);
  uint32_t* incT = reinterpret_cast<uint32_t*>(smem_dyn + L::kIncOff);
  uint8_t* stage_base = smem_dyn + L::kStageOff;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int go = d.gap_open, ge = d.gap_ext;
  const uint32_t ngo2 = pack16(-go);
  Stat2Consts c;
  c.one = d.one;
  c.zero = d.one - 1u;
  const uint32_t sent2 = pack16(kSentinel16) + c.zero;  // keep it a register operand (see nw_thread2_kernel)
  const uint32_t bord2 = pack16(ge - go);
  const unsigned full = 0xFFFFFFFFu;
  uint8_t* sA = stage_base + (warp * 2 + 0) * (kNwMpStageCols + 8);
  uint8_t* sB = stage_base + (warp * 2 + 1) * (kNwMpStageCols + 8);
  const uint32_t sA_sh = (uint32_t)__cvta_generic_to_shared(sA), sB_sh = (uint32_t)__cvta_generic_to_shared(sB);
  const uint32_t plane_sh = (uint32_t)__cvta_generic_to_shared(prof + lane * L::kProfStride);
  const uint32_t ilane_sh = (uint32_t)__cvta_generic_to_shared(incT + lane * L::kIncStride);
  __shared__ int col_len[2 * kNwMpPairSets];
  __shared__ uint8_t col_ord[2 * kNwMpPairSets];

  for (int u = blockIdx.x; u < num_units; u += gridDim.x) {
    const NwUnit un = units[u];
    const int row = un.row;
    const int m = d.off[row + 1] - d.off[row];
    const uint8_t* __restrict__ a = d.codes + d.off[row];
    const int npass = (m + 32 * R - 1) / (32 * R);
    const int npairs2 = (un.j_count + 1) >> 1;
    // column sequences paired by length, as in nw_warp2_kernel (long proteins differ by ~100 residues when paired at
    // random); the order is fixed for the unit, so the per-pair-set scratch lines stay valid across passes
    __syncthreads();
    if (tid < un.j_count) col_len[tid] = d.off[un.j_begin + tid + 1] - d.off[un.j_begin + tid];
    __syncthreads();
    if (tid < un.j_count) {
      const int mine = col_len[tid];
      int rank = 0;
      for (int q = 0; q < un.j_count; ++q) {
        const int other = col_len[q];
        rank += (other > mine || (other == mine && q < tid)) ? 1 : 0;
      }
      col_ord[rank] = (uint8_t)tid;
    }

    for (int pass = 0; pass < npass; ++pass) {
      const int row0 = pass * 32 * R;
      const bool last_pass = (pass == npass - 1);
      __syncthreads();  // previous pass (or unit) no longer reads the tables
      build_profile<R, 32, L::kProfStride>(prof, a, m, row0, d.sub, 2 * ge, tid, kWarpThreads);
      for (int idx = tid; idx < 32 * L::kProfStride; idx += kWarpThreads) prof[24 * 32 * L::kProfStride + idx] = 0u;
      for (int idx = tid; idx < 25 * 32 * L::kIncStride; idx += kWarpThreads) {
        const int cls = idx / (32 * L::kIncStride);
        const int rem = idx - cls * (32 * L::kIncStride);
        const int ln = rem / L::kIncStride, k = rem - ln * L::kIncStride;
        const int r = row0 + ln * R + k;
        incT[idx] = 1u | ((k < R && r < m && cls < 24 && a[r] == cls) ? 0x10000u : 0u);
      }
      __syncthreads();
      const int lm = last_pass ? (m - 1 - row0) / R : 31;
      const int km = (m - 1 - row0) - lm * R;
      const int r0 = row0 + lane * R;

      for (int pp = warp; pp < npairs2; pp += nwarps) {
        const bool hasB = (2 * pp + 1 < un.j_count);
        int jA = un.j_begin + col_ord[2 * pp];
        int jB = hasB ? un.j_begin + col_ord[2 * pp + 1] : jA;
        int nA = d.off[jA + 1] - d.off[jA], nB = d.off[jB + 1] - d.off[jB];
        if (nB > nA) {
          int tj = jA; jA = jB; jB = tj;
          int tn = nA; nA = nB; nB = tn;
        }
        uint4* __restrict__ scr = scratch + ((int64_t)blockIdx.x * kNwMpPairSets + pp) * kNwMpStageCols;
        {
          const uint8_t* __restrict__ bA = d.codes + d.off[jA];
          const uint8_t* __restrict__ bB = d.codes + d.off[jB];
          __syncwarp();
          for (int q = lane; q < nA; q += 32) {
            sA[q] = bA[q];
            sB[q] = (q < nB) ? bB[q] : (uint8_t)24;
          }
          __syncwarp();
        }
        uint32_t H0[R], H1[R], El[R], SA0[R], SA1[R], SB0[R], SB1[R];
#pragma unroll
        for (int k = 0; k < R; ++k) {
          H0[k] = H1[k] = bord2;  // border column (slanted): the same for every row >= 1
          El[k] = sent2;
          SA0[k] = SA1[k] = SB0[k] = SB1[k] = 0u;
        }
        uint32_t prevUpH = (r0 == 0) ? 0u : bord2;  // diagonal source (r0, 0): corner only for the very first row
        uint32_t prevUpSA = 0u, prevUpSB = 0u;
        uint32_t outF = 0u;  // H and the stats of the bottom row are shuffled straight out of the strip registers
        uint32_t resB = 0u;
        const unsigned n_act = (lane <= lm) ? (unsigned)nA : 0u;
        const int capB = (last_pass && lane == lm) ? nB - 1 : -1;
        const int T = nA + lm;
        // lane 0 of a later pass reads its upper neighbour from scratch, one column ahead of its use (the load
        // latency hides behind the current column's work)
        const bool from_scr = (lane == 0 && pass > 0);
        uint4 nxt = make_uint4(bord2, sent2, 0u, 0u);
        if (from_scr && nA > 0) nxt = scr[0];
        for (int t0 = 0; t0 < T; t0 += 2) {
#pragma unroll
          for (int ph = 0; ph < 2; ++ph) {
            const int jc = t0 + ph - lane;
            // the previous step wrote the register set this phase reads
            uint32_t rH = __shfl_up_sync(full, ph == 1 ? H1[R - 1] : H0[R - 1], 1);
            uint32_t rF = __shfl_up_sync(full, outF, 1);
            uint32_t rSA = __shfl_up_sync(full, ph == 1 ? SA1[R - 1] : SA0[R - 1], 1);
            uint32_t rSB = __shfl_up_sync(full, ph == 1 ? SB1[R - 1] : SB0[R - 1], 1);
            if (lane == 0) {  // pass 0: border row (nxt stays at the border constants); later passes: scratch
              rH = nxt.x;
              rF = nxt.y;
              rSA = nxt.z;
              rSB = nxt.w;
              if (from_scr && jc + 1 < nA) nxt = scr[jc + 1];
            }
            if ((unsigned)jc < n_act) {
              const uint32_t cA = lds_u8(sA_sh + (uint32_t)jc), cB = lds_u8(sB_sh + (uint32_t)jc);
              uint32_t pwA[S::RW], pwB[S::RW];
              const uint32_t pa = plane_sh + cA * (32u * L::kProfStride * 4u);
              const uint32_t pb = plane_sh + cB * (32u * L::kProfStride * 4u);
              const uint32_t ia = ilane_sh + cA * (32u * L::kIncStride * 4u);
              const uint32_t ib = ilane_sh + cB * (32u * L::kIncStride * 4u);
#pragma unroll
              for (int w = 0; w < S::RW; ++w) {
                pwA[w] = lds_u32(pa + 4u * (unsigned)w);
                pwB[w] = lds_u32(pb + 4u * (unsigned)w);
              }
              if (ph == 0) {
                strip_column2<R, 2>(H0, H1, El, SA0, SA1, SB0, SB1, pwA, pwB, prevUpH, prevUpSA, prevUpSB, rF, rSA, rSB,
                                    ngo2, c, outF, ia, ib);
              } else {
                strip_column2<R, 2>(H1, H0, El, SA1, SA0, SB1, SB0, pwA, pwB, prevUpH, prevUpSA, prevUpSB, rF, rSA, rSB,
                                    ngo2, c, outF, ia, ib);
              }
              prevUpH = rH;
              prevUpSA = rSA;
              prevUpSB = rSB;
              if (!last_pass && lane == 31)  // the bottom row just written (the other register set)
                scr[jc] = ph == 0 ? make_uint4(H1[R - 1], outF, SA1[R - 1], SB1[R - 1])
                                  : make_uint4(H0[R - 1], outF, SA0[R - 1], SB0[R - 1]);
              if (jc == capB) {
#pragma unroll
                for (int k = 0; k < R; ++k)
                  if (k == km) resB = (ph == 0) ? SB1[k] : SB0[k];
              }
            }
          }
        }
        if (last_pass) {
          const bool in1 = (((lm + nA) & 1) != 0);
          uint32_t resA = 0u;
#pragma unroll
          for (int k = 0; k < R; ++k)
            if (k == km) resA = in1 ? SA1[k] : SA0[k];
          resA = __shfl_sync(full, resA, lm);
          resB = __shfl_sync(full, resB, lm);
          if (lane == 0) {
            const int64_t slotA = pair_slot(d.n, row, jA, d.slab_base);
            d.matches[slotA] = resA >> 16;
            d.length[slotA] = (uint32_t)(m + nA) - (resA & 0xFFFFu);
            if (hasB) {
              const int64_t slotB = pair_slot(d.n, row, jB, d.slab_base);
              d.matches[slotB] = resB >> 16;
              d.length[slotB] = (uint32_t)(m + nB) - (resB & 0xFFFFu);
            }
          }
        }
        __syncwarp();
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K5x2 cooperative: K warps share one 32*K-lane wavefront (rows 385..768: the 566-residue HA sequences of BASELINE
// config 2).  A strip taller than 12 rows per lane loses the ping-pong register sets and the increment table
// (nw_warp2_kernel<R >= 13>: one register set, PRMT increments, 2.6 TCUPS against 3.2 for R <= 12); here the row
// sequence is cut into K consecutive blocks of 32*R rows, R <= 12, and warp w of a group owns block w with exactly
// the strip code of the fast kernel.  Warp w+1's lane 0 needs, per column, the bottom row (H, F', statA, statB) of
// warp w's lane 31.  It travels through a 128-entry shared-memory ring:
//   * lane 0 of the PRODUCER receives lane 31's bottom row one step later through the rotating shuffle every lane
//     uses anyway (src = lane-1 mod 32), so the four values already sit in four fresh registers: one predicated
//     128-bit store, no register moves, no divergent lane-31 code;
//   * lane 0 of the CONSUMER overwrites the four shuffled registers with one predicated 128-bit load; the first
//     warp of a group loads the constant border entry the same way (no per-step selects for the border row);
//   * the two warps are NOT in lock step: each runs its own stream of (n + 31)-step jobs, the consumer trailing by
//     at least 32 columns.  Progress counters (entries produced / consumed, published every 16 steps with
//     st.release / ld.acquire at CTA scope) keep the consumer behind the producer and the producer less than one
//     ring ahead; they are polled only when the cached copy says "not yet".
// Warp-steps per pair-set are K*(n+31), the same as K sequential passes, and nothing leaves the SM (the multi-pass
// kernel pays a global scratch line and a table rebuild per pass).  One CTA of 16 warps per SM: both warps' tables
// (K*38 KB increments, K*16..22 KB scores) are resident, 8 groups share them.
// ------------------------------------------------------------------------------------------------
constexpr int kCoThreads = 512;
constexpr int kCoRing = 128;   // ring entries (16 bytes each) between two neighbouring warps of a group
constexpr int kCoChunk = 16;   // flow control and progress publication once per kCoChunk steps
constexpr int kCoGap = 32;     // ring indices skipped between pair-sets (see the unconditional store below)

__device__ __forceinline__ void st_release_shared(uint32_t addr, uint32_t v) {
  asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_shared(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}

template <int R, int K, int VAR>
struct CoSmem {
  static constexpr int kLanes = 32 * K;
  static constexpr int kWarps = kCoThreads / 32;
  static constexpr int kGroups = kWarps / K;
  static constexpr int kIncStride = R <= 4 ? 4 : 12;  // multiple of 4 words, odd multiple: conflict-free 128-bit loads
  static constexpr int kProfStride = Strip<R>::RWS64; // score profile read with 64-bit loads
  static constexpr int kRingBytes = kWarps * kCoRing * 16;  // one outgoing ring per warp (the last warp's is a sink)
  static constexpr int kRecStride = Rec<R>::kStride;  // VAR 3: one record table instead of the two tables
  static constexpr int kIncBytes = VAR == 3 ? 25 * kLanes * kRecStride * 4 : 25 * kLanes * kIncStride * 4;
  static constexpr int kProfBytes = VAR == 3 ? 0 : 25 * kLanes * kProfStride * 4;
  static constexpr int kStageBytes = kWarps * 2 * (kNwStageCols + 8);  // every warp stages its own copy
  // offsets from the first 2048-byte aligned address of the dynamic window (ring addresses are formed with OR)
  static constexpr int kRingOff = 0;
  static constexpr int kIncOff = kRingOff + kRingBytes;
  static constexpr int kProfOff = kIncOff + kIncBytes;
  static constexpr int kStageOff = kProfOff + kProfBytes;
  static constexpr int kTotal = kStageOff + kStageBytes + 2048;
  static_assert(R <= 12, "increment table stride");
  static_assert(kWarps % K == 0, "groups");
  static_assert(kCoRing * 16 == 2048, "ring addresses: (offset & 0x7F0) | base");
};

template <int R, int K, int VAR>
__global__ void __launch_bounds__(kCoThreads, 1)
nw_warp2co_kernel(NwDeviceData d, const NwUnit* __restrict__ units, int num_units) {
  using S = Strip<R>;
  using L = CoSmem<R, K, VAR>;
  extern __shared__ __align__(16) unsigned char smem_dyn[];
  __shared__ uint32_t res_m[kNwCoUnitCols], res_l[kNwCoUnitCols];
  __shared__ int col_len[kNwCoUnitCols];
  __shared__ uint8_t col_ord[kNwCoUnitCols];
  __shared__ uint32_t prod_cnt[L::kWarps], cons_cnt[L::kWarps];  // per warp: ring entries it has written / read
  const uint32_t dyn_sh = ((uint32_t)__cvta_generic_to_shared(smem_dyn) + 2047u) & ~2047u;
  unsigned char* dyn = smem_dyn + (dyn_sh - (uint32_t)__cvta_generic_to_shared(smem_dyn));
  uint32_t* incT = reinterpret_cast<uint32_t*>(dyn + L::kIncOff);
  uint32_t* prof = reinterpret_cast<uint32_t*>(dyn + L::kProfOff);
  uint8_t* stage_base = dyn + L::kStageOff;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int group = warp / K, role = warp - group * K;
  const int go = d.gap_open, ge = d.gap_ext;
  const uint32_t ngo2 = pack16(-go);
  Stat2Consts c;
  c.one = d.one;
  c.zero = d.one - 1u;
  const uint32_t sent2 = pack16(kSentinel16) + c.zero;  // register operand (see nw_thread2_kernel)
  const uint32_t bord2 = pack16(ge - go);
  const unsigned full = 0xFFFFFFFFu;
  const int src_lane = (lane + 31) & 31;
  const bool producer = role < K - 1, consumer = role > 0;
  uint8_t* sA = stage_base + (warp * 2 + 0) * (kNwStageCols + 8);
  uint8_t* sB = stage_base + (warp * 2 + 1) * (kNwStageCols + 8);
  const uint32_t sA_sh = (uint32_t)__cvta_generic_to_shared(sA), sB_sh = (uint32_t)__cvta_generic_to_shared(sB);
  const uint32_t plane_sh = (uint32_t)__cvta_generic_to_shared(prof + (role * 32 + lane) * L::kProfStride);
  const uint32_t ilane_sh = (uint32_t)__cvta_generic_to_shared(incT + (role * 32 + lane) * (VAR == 3 ? L::kRecStride : L::kIncStride));
  // Every warp writes its bottom row into its own ring; warp w > 0 of a group reads warp w-1's ring.  The first warp of
  // a group "reads" entry 0 of the sink ring of the group's last warp, which holds the border constants for the whole
  // unit (offset mask 0); the last warp's own stores go to entries 1.. of that sink ring and are never read.
  const uint32_t out_ring_sh = dyn_sh + L::kRingOff + (uint32_t)warp * (kCoRing * 16);
  const uint32_t in_ring_sh = dyn_sh + L::kRingOff + (uint32_t)(consumer ? warp - 1 : group * K + K - 1) * (kCoRing * 16);
  const uint32_t in_mask16 = consumer ? (uint32_t)(kCoRing * 16 - 16) : 0u;
  const uint32_t out_mask16 = (uint32_t)(kCoRing * 16 - 16);
  const uint32_t out_fix = producer ? 0u : 16u;  // keeps the sink's stores away from entry 0 (see above)
  const uint32_t my_prod_sh = (uint32_t)__cvta_generic_to_shared(&prod_cnt[warp]);
  const uint32_t my_cons_sh = (uint32_t)__cvta_generic_to_shared(&cons_cnt[warp]);
  const uint32_t up_prod_sh = (uint32_t)__cvta_generic_to_shared(&prod_cnt[consumer ? warp - 1 : warp]);
  const uint32_t dn_cons_sh = (uint32_t)__cvta_generic_to_shared(&cons_cnt[producer ? warp + 1 : warp]);

  for (int u = blockIdx.x; u < num_units; u += gridDim.x) {
    const NwUnit un = units[u];
    const int row = un.row;
    const int m = d.off[row + 1] - d.off[row];
    const uint8_t* __restrict__ a = d.codes + d.off[row];
    __syncthreads();
    if (tid < un.j_count) col_len[tid] = d.off[un.j_begin + tid + 1] - d.off[un.j_begin + tid];
    if (tid < L::kWarps) prod_cnt[tid] = cons_cnt[tid] = 0u;
    if (tid < L::kGroups)  // border entry of each group: H_diag = -go+ge (slanted), F = sentinel, no statistics yet
      *reinterpret_cast<uint4*>(dyn + L::kRingOff + (tid * K + K - 1) * (kCoRing * 16)) = make_uint4(bord2, sent2, 0u, 0u);
    if constexpr (VAR == 3) {
      build_records<R, L::kLanes>(incT, a, m, 0, d.sub, 2 * ge, tid, kCoThreads);
    } else {
      build_profile<R, L::kLanes, L::kProfStride>(prof, a, m, 0, d.sub, 2 * ge, tid, kCoThreads);
      for (int idx = tid; idx < L::kLanes * L::kProfStride; idx += kCoThreads) prof[24 * L::kLanes * L::kProfStride + idx] = 0u;
      for (int idx = tid; idx < 25 * L::kLanes * L::kIncStride; idx += kCoThreads) {
        const int cls = idx / (L::kLanes * L::kIncStride);
        const int rem = idx - cls * (L::kLanes * L::kIncStride);
        const int ln = rem / L::kIncStride, k = rem - ln * L::kIncStride;
        const int r = ln * R + k;
        incT[idx] = 1u | ((k < R && r < m && cls < 24 && a[r] == cls) ? 0x10000u : 0u);
      }
    }
    __syncthreads();
    if (tid < un.j_count) {  // rank by (length descending, index ascending)
      const int mine = col_len[tid];
      int rank = 0;
      for (int q = 0; q < un.j_count; ++q) {
        const int other = col_len[q];
        rank += (other > mine || (other == mine && q < tid)) ? 1 : 0;
      }
      col_ord[rank] = (uint8_t)tid;
    }
    __syncthreads();
    const int Lm = (m - 1) / R;              // last lane (0 .. 32K-1) that owns rows
    const int km = (m - 1) - Lm * R;
    const int last_role = Lm >> 5;           // host guarantees last_role == K-1 (m > 32*R*(K-1))
    const int lm = (role < last_role) ? 31 : (Lm & 31);
    const bool owns_result = (role == last_role);
    const int r0 = (role * 32 + lane) * R;
    const int npairs2 = (un.j_count + 1) >> 1;
    // Ring index space: a pair-set with nA columns owns indices [base, base + nA) followed by a gap of kCoGap unused
    // indices.  The gap absorbs the stores lane 0 issues before lane 31 has produced anything (steps 0..31 of the
    // NEXT pair-set write indices base-32 .. base-1), so the store needs no condition at all.
    int base_p = 0, base_c = 0;       // first ring index of the current pair-set (written / read side)
    int seen_prod = 0, seen_cons = 0; // cached copies of the neighbours' counters

    for (int pp = group; pp < npairs2; pp += L::kGroups) {
      const bool hasB = (2 * pp + 1 < un.j_count);
      int jA = un.j_begin + col_ord[2 * pp];
      int jB = hasB ? un.j_begin + col_ord[2 * pp + 1] : jA;
      int nA = d.off[jA + 1] - d.off[jA], nB = d.off[jB + 1] - d.off[jB];
      if (nB > nA) {
        int tj = jA; jA = jB; jB = tj;
        int tn = nA; nA = nB; nB = tn;
      }
      {
        const uint8_t* __restrict__ bA = d.codes + d.off[jA];
        const uint8_t* __restrict__ bB = d.codes + d.off[jB];
        __syncwarp();
        for (int q = lane; q < nA; q += 32) {
          sA[q] = bA[q];
          sB[q] = (q < nB) ? bB[q] : (uint8_t)24;
        }
        __syncwarp();
      }
      uint32_t H0[R], H1[R], El[R], SA0[R], SA1[R], SB0[R], SB1[R];
#pragma unroll
      for (int k = 0; k < R; ++k) {
        H0[k] = H1[k] = bord2;
        El[k] = sent2;
        SA0[k] = SA1[k] = SB0[k] = SB1[k] = 0u;
      }
      uint32_t prevUpH = (r0 == 0) ? 0u : bord2;
      uint32_t prevUpSA = 0u, prevUpSB = 0u;
      uint32_t outF = sent2;
      uint32_t resB = 0u;
      const unsigned n_act = (lane <= lm) ? (unsigned)nA : 0u;
      const int capB = (owns_result && lane == lm) ? nB - 1 : -1;
      // a producer runs one extra (idle) step so that lane 0 also receives and stores lane 31's last column
      const int T = nA + lm + (producer ? 1 : 0);
      // running ring addresses: step t stores index base_p + t - 32 and loads index base_c + t
      uint32_t out_addr = out_ring_sh | (((uint32_t)(base_p - 32) * 16u) & out_mask16) | out_fix;
      uint32_t in_addr = in_ring_sh | (((uint32_t)base_c * 16u) & in_mask16);
      for (int tc = 0; tc < T; tc += kCoChunk) {
        const int tend = min(tc + kCoChunk, T);
        // ---- flow control once per chunk, warp-uniform (all lanes hold the same counters)
        if (consumer) {
          const int need = base_c + min(tend, nA);  // ring indices below this are read in this chunk
          while (seen_prod < need) seen_prod = (int)ld_acquire_shared(up_prod_sh);
        }
        if (producer) {
          const int top = base_p + tend + 1 - 32;   // ring indices below this are written in this chunk
          while (top - seen_cons > kCoRing) seen_cons = (int)ld_acquire_shared(dn_cons_sh);
        }
        if (lane == 0) {
          if (producer) st_release_shared(my_prod_sh, (uint32_t)(base_p + max(tc - 32, 0)));
          if (consumer) st_release_shared(my_cons_sh, (uint32_t)(base_c + min(tc, nA)));
        }
        for (int t0 = tc; t0 < tend; t0 += 2) {
#pragma unroll
          for (int ph = 0; ph < 2; ++ph) {
            const int jc = t0 + ph - lane;
            uint32_t rH = __shfl_sync(full, ph == 1 ? H1[R - 1] : H0[R - 1], src_lane);
            uint32_t rF = __shfl_sync(full, outF, src_lane);
            uint32_t rSA = __shfl_sync(full, ph == 1 ? SA1[R - 1] : SA0[R - 1], src_lane);
            uint32_t rSB = __shfl_sync(full, ph == 1 ? SB1[R - 1] : SB0[R - 1], src_lane);
            // lane 0: what arrived from lane 31 is this warp's bottom row at column t - 32 -> ring; its own upper
            // neighbour comes from the ring of the warp above (or the border entry)
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "setp.eq.u32 p, %6, 0;\n\t"
                "@p st.shared.v4.u32 [%4], {%0, %1, %2, %3};\n\t"
                "@p ld.shared.v4.u32 {%0, %1, %2, %3}, [%5];\n\t}"
                : "+r"(rH), "+r"(rF), "+r"(rSA), "+r"(rSB)
                : "r"(out_addr), "r"(in_addr), "r"(lane)
                : "memory");
            out_addr = out_ring_sh | ((out_addr + 16u) & out_mask16) | out_fix;
            in_addr = in_ring_sh | ((in_addr + 16u) & in_mask16);
            if ((unsigned)jc < n_act) {
              const uint32_t cA = lds_u8(sA_sh + (uint32_t)jc), cB = lds_u8(sB_sh + (uint32_t)jc);
              if constexpr (VAR == 3) {
                const uint32_t ra = ilane_sh + cA * (uint32_t)(L::kLanes * L::kRecStride * 4);
                const uint32_t rb = ilane_sh + cB * (uint32_t)(L::kLanes * L::kRecStride * 4);
                if (ph == 0) {
                  strip_column3<R>(H0, H1, El, SA0, SA1, SB0, SB1, ra, rb, prevUpH, prevUpSA, prevUpSB, rF, rSA, rSB, ngo2, c, outF);
                } else {
                  strip_column3<R>(H1, H0, El, SA1, SA0, SB1, SB0, ra, rb, prevUpH, prevUpSA, prevUpSB, rF, rSA, rSB, ngo2, c, outF);
                }
              } else {
                uint32_t pwA[S::RW], pwB[S::RW];
                const uint32_t pa = plane_sh + cA * (uint32_t)(L::kLanes * L::kProfStride * 4);
                const uint32_t pb = plane_sh + cB * (uint32_t)(L::kLanes * L::kProfStride * 4);
                const uint32_t ia = ilane_sh + cA * (uint32_t)(L::kLanes * L::kIncStride * 4);
                const uint32_t ib = ilane_sh + cB * (uint32_t)(L::kLanes * L::kIncStride * 4);
#pragma unroll
                for (int w = 0; w < S::RW; w += 2) {  // an odd RW reads one padding word of the lane's stride
                  const uint2 va = lds_v2(pa + 4u * (unsigned)w), vb = lds_v2(pb + 4u * (unsigned)w);
                  pwA[w] = va.x;
                  pwB[w] = vb.x;
                  if (w + 1 < S::RW) {
                    pwA[w + 1] = va.y;
                    pwB[w + 1] = vb.y;
                  }
                }
                if (ph == 0) {
                  strip_column2<R, 2>(H0, H1, El, SA0, SA1, SB0, SB1, pwA, pwB, prevUpH, prevUpSA, prevUpSB, rF, rSA, rSB,
                                      ngo2, c, outF, ia, ib);
                } else {
                  strip_column2<R, 2>(H1, H0, El, SA1, SA0, SB1, SB0, pwA, pwB, prevUpH, prevUpSA, prevUpSB, rF, rSA, rSB,
                                      ngo2, c, outF, ia, ib);
                }
              }
              prevUpH = rH;
              prevUpSA = rSA;
              prevUpSB = rSB;
              if (jc == capB) {
#pragma unroll
                for (int k = 0; k < R; ++k)
                  if (k == km) resB = (ph == 0) ? SB1[k] : SB0[k];
              }
            }
          }
        }
      }
      // the pair-set is complete on this warp: publish, so that a neighbour waiting on the tail can proceed
      base_p += nA + kCoGap;
      base_c += nA + kCoGap;
      __syncwarp();
      if (lane == 0) {
        if (producer) st_release_shared(my_prod_sh, (uint32_t)base_p);
        if (consumer) st_release_shared(my_cons_sh, (uint32_t)base_c);
      }
      if (owns_result) {
        const bool in1 = (((lm + nA) & 1) != 0);
        uint32_t resA = 0u;
#pragma unroll
        for (int k = 0; k < R; ++k)
          if (k == km) resA = in1 ? SA1[k] : SA0[k];
        resA = __shfl_sync(full, resA, lm);
        resB = __shfl_sync(full, resB, lm);
        if (lane == 0) {
          res_m[jA - un.j_begin] = resA >> 16;
          res_l[jA - un.j_begin] = (uint32_t)(m + nA) - (resA & 0xFFFFu);
          if (hasB) {
            res_m[jB - un.j_begin] = resB >> 16;
            res_l[jB - un.j_begin] = (uint32_t)(m + nB) - (resB & 0xFFFFu);
          }
        }
      }
    }
    __syncthreads();
    {
      const int64_t slot0 = pair_slot(d.n, row, un.j_begin, d.slab_base);
      for (int q = tid; q < 2 * un.j_count; q += kCoThreads) {
        if (q < un.j_count) d.matches[slot0 + q] = res_m[q];
        else d.length[slot0 + q - un.j_count] = res_l[q - un.j_count];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K5x2 cooperative + two rows: the cooperative wavefront of nw_warp2co_kernel (K warps, ring between neighbours) with
// the operand layout of nw_rows2_kernel (rows i and i+1 in the two 16-bit halves, one column sequence, one record
// stream, no score PRMT).  2R words per lane and class over 32*K lanes is 172 KB for R = 12, K = 2 -- it fits next to
// the rings and the staged sequences -- so it serves row pairs of 385..768 residues, among them the 566-residue HA
// sequences of BASELINE config 2 (with one increment word per pair, 3R words, the tables stopped at R = 9 = 576 rows).  Rings: one per producer warp, one shared sink for the last warp of every group (its bottom row
// is never read), one constant border entry.
// ------------------------------------------------------------------------------------------------
template <int R, int K>
struct Rows2CoSmem {
  static constexpr int kLanes = 32 * K;
  static constexpr int kWarps = kCoThreads / 32;
  static constexpr int kGroups = kWarps / K;
  static constexpr int kStride = Rec2<R>::kStride;
  static constexpr int kRings = kGroups * (K - 1) + 1;            // producers' rings + the sink
  static constexpr int kRingBytes = kRings * kCoRing * 16;
  static constexpr int kTableBytes = 24 * kLanes * kStride * 4;
  static constexpr int kStageCols = nw_rows2co_max_cols(R);       // 2048 where the tables leave room (R <= 10)
  static constexpr int kStageBytes = kWarps * (kStageCols + 8);
  static constexpr int kRingOff = 0;                              // from the first 2048-byte aligned address
  static constexpr int kTableOff = kRingOff + kRingBytes;
  static constexpr int kStageOff = kTableOff + kTableBytes;
  static constexpr int kTotal = kStageOff + kStageBytes + 2048;
};

template <int R, int K, bool U>
__global__ void __launch_bounds__(kCoThreads, 1)
nw_rows2co_kernel(NwDeviceData d, const NwUnit* __restrict__ units, int num_units) {
  using RC = Rec2<R>;
  using L = Rows2CoSmem<R, K>;
  extern __shared__ __align__(16) unsigned char smem_dyn[];
  __shared__ uint32_t res_m1[kNwCoUnitCols], res_l1[kNwCoUnitCols], res_m2[kNwCoUnitCols], res_l2[kNwCoUnitCols];
  __shared__ uint32_t prod_cnt[L::kWarps], cons_cnt[L::kWarps];
  __shared__ __align__(16) uint4 border_entry;
  const uint32_t dyn_sh = ((uint32_t)__cvta_generic_to_shared(smem_dyn) + 2047u) & ~2047u;
  unsigned char* dyn = smem_dyn + (dyn_sh - (uint32_t)__cvta_generic_to_shared(smem_dyn));
  uint32_t* rec = reinterpret_cast<uint32_t*>(dyn + L::kTableOff);
  uint8_t* stage_base = dyn + L::kStageOff;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int group = warp / K, role = warp - group * K;
  const int go = d.gap_open, ge = d.gap_ext;
  const uint32_t ngo2 = pack16(-go);
  Stat2Consts c;
  c.one = d.one;
  c.zero = d.zero;
  const uint32_t sent2 = (U ? 0u : pack16(kSentinel16)) + c.zero;     // U: the unsigned domain of strip_column4
  const uint32_t bord2 = pack16(ge - go + (U ? (int)d.bias16 : 0));
  const uint32_t corner2 = U ? pack16((int)d.bias16) : 0u;
  const unsigned full = 0xFFFFFFFFu;
  const int src_lane = (lane + 31) & 31;
  const bool producer = role < K - 1, consumer = role > 0;
  uint8_t* sC = stage_base + warp * (L::kStageCols + 8);
  const uint32_t sC_sh = (uint32_t)__cvta_generic_to_shared(sC);
  const uint32_t rlane_sh = (uint32_t)__cvta_generic_to_shared(rec + (role * 32 + lane) * L::kStride);
  // ring this warp writes: its own when another warp of the group consumes it, the shared sink otherwise
  const uint32_t out_ring_sh = dyn_sh + L::kRingOff + (uint32_t)(producer ? group * (K - 1) + role : L::kRings - 1) * (kCoRing * 16);
  const uint32_t in_ring_sh = consumer ? dyn_sh + L::kRingOff + (uint32_t)(group * (K - 1) + role - 1) * (kCoRing * 16)
                                       : (uint32_t)__cvta_generic_to_shared(&border_entry);
  const uint32_t in_mask16 = consumer ? (uint32_t)(kCoRing * 16 - 16) : 0u;
  const uint32_t out_mask16 = (uint32_t)(kCoRing * 16 - 16);
  const uint32_t my_prod_sh = (uint32_t)__cvta_generic_to_shared(&prod_cnt[warp]);
  const uint32_t my_cons_sh = (uint32_t)__cvta_generic_to_shared(&cons_cnt[warp]);
  const uint32_t up_prod_sh = (uint32_t)__cvta_generic_to_shared(&prod_cnt[consumer ? warp - 1 : warp]);
  const uint32_t dn_cons_sh = (uint32_t)__cvta_generic_to_shared(&cons_cnt[producer ? warp + 1 : warp]);

  for (int u = blockIdx.x; u < num_units; u += gridDim.x) {
    NwUnit un = units[u];
    const int row = un.row;
    const int row2_top = row + (un.j_count >> 16);
    un.j_count &= 0xFFFF;
    const int m1 = d.off[row + 1] - d.off[row], m2 = d.off[row2_top + 1] - d.off[row2_top];
    __syncthreads();
    if (tid < L::kWarps) prod_cnt[tid] = cons_cnt[tid] = 0u;
    if (tid == 0) border_entry = make_uint4(bord2, sent2, 0u, 0u);
    {  // records of all 32*K lanes
      const uint8_t* __restrict__ a1 = d.codes + d.off[row];
      const uint8_t* __restrict__ a2 = d.codes + d.off[row2_top];
      for (int idx = tid; idx < 24 * L::kLanes * L::kStride; idx += kCoThreads) {
        const int cls = idx / (L::kLanes * L::kStride);
        const int rem = idx - cls * (L::kLanes * L::kStride);
        const int ln = rem / L::kStride, w = rem - ln * L::kStride;
        uint32_t v = 0u;
        if (w < RC::kWords) {
          const int k = w >> 1, r = ln * R + k;
          if ((w & 1) == 0) {
            const int s1 = r < m1 ? (int)(int8_t)(d.sub[a1[r] * 24 + cls] + 2 * ge) : 0;
            const int s2 = r < m2 ? (int)(int8_t)(d.sub[a2[r] * 24 + cls] + 2 * ge) : 0;
            v = score2_word(s1, s2, U);
          } else {
            v = inc2_word(r < m1 && a1[r] == cls, r < m2 && a2[r] == cls);
          }
        }
        rec[idx] = v;
      }
    }
    __syncthreads();
    const int LmA = (m1 - 1) / R, kmA = (m1 - 1) - LmA * R;  // last lane (0 .. 32K-1) and strip row of each sequence
    const int LmB = (m2 - 1) / R, kmB = (m2 - 1) - LmB * R;
    const int Lm = max(LmA, LmB);
    const int last_role = Lm >> 5;  // host guarantees last_role == K-1
    const int lm = (role < last_role) ? 31 : (Lm & 31);
    const bool ownA = (role == (LmA >> 5)), ownB = (role == (LmB >> 5));
    const int r0 = (role * 32 + lane) * R;
    int base_p = 0, base_c = 0;
    int seen_prod = 0, seen_cons = 0;

    for (int pp = group; pp < un.j_count; pp += L::kGroups) {
      const int j = un.j_begin + pp;
      const int n = d.off[j + 1] - d.off[j];
      {
        const uint8_t* __restrict__ b = d.codes + d.off[j];
        __syncwarp();
        for (int q = lane; q < n; q += 32) sC[q] = b[q];
        __syncwarp();
      }
      uint32_t H0[R], H1[R], El[R], SA0[R], SA1[R], SB0[R], SB1[R];
#pragma unroll
      for (int k = 0; k < R; ++k) {
        H0[k] = H1[k] = bord2;
        El[k] = sent2;
        SA0[k] = SA1[k] = SB0[k] = SB1[k] = 0u;
      }
      uint32_t prevUpH = (r0 == 0) ? corner2 : bord2;
      uint32_t prevUpSA = 0u, prevUpSB = 0u;
      uint32_t outF = sent2;
      const unsigned n_act = (lane <= lm) ? (unsigned)n : 0u;
      const int T = n + lm + (producer ? 1 : 0);
      uint32_t out_addr = out_ring_sh | (((uint32_t)(base_p - 32) * 16u) & out_mask16);
      uint32_t in_addr = in_ring_sh | (((uint32_t)base_c * 16u) & in_mask16);
      for (int tc = 0; tc < T; tc += kCoChunk) {
        const int tend = min(tc + kCoChunk, T);
        if (consumer) {
          const int need = base_c + min(tend, n);
          while (seen_prod < need) seen_prod = (int)ld_acquire_shared(up_prod_sh);
        }
        if (producer) {
          const int top = base_p + tend + 1 - 32;
          while (top - seen_cons > kCoRing) seen_cons = (int)ld_acquire_shared(dn_cons_sh);
        }
        if (lane == 0) {
          if (producer) st_release_shared(my_prod_sh, (uint32_t)(base_p + max(tc - 32, 0)));
          if (consumer) st_release_shared(my_cons_sh, (uint32_t)(base_c + min(tc, n)));
        }
        for (int t0 = tc; t0 < tend; t0 += 2) {
#pragma unroll
          for (int ph = 0; ph < 2; ++ph) {
            const int jc = t0 + ph - lane;
            uint32_t rH = __shfl_sync(full, ph == 1 ? H1[R - 1] : H0[R - 1], src_lane);
            uint32_t rF = __shfl_sync(full, outF, src_lane);
            uint32_t rSA = __shfl_sync(full, ph == 1 ? SA1[R - 1] : SA0[R - 1], src_lane);
            uint32_t rSB = __shfl_sync(full, ph == 1 ? SB1[R - 1] : SB0[R - 1], src_lane);
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "setp.eq.u32 p, %6, 0;\n\t"
                "@p st.shared.v4.u32 [%4], {%0, %1, %2, %3};\n\t"
                "@p ld.shared.v4.u32 {%0, %1, %2, %3}, [%5];\n\t}"
                : "+r"(rH), "+r"(rF), "+r"(rSA), "+r"(rSB)
                : "r"(out_addr), "r"(in_addr), "r"(lane)
                : "memory");
            out_addr = out_ring_sh | ((out_addr + 16u) & out_mask16);
            in_addr = in_ring_sh | ((in_addr + 16u) & in_mask16);
            if ((unsigned)jc < n_act) {
              const uint32_t cc = lds_u8(sC_sh + (uint32_t)jc);
              const uint32_t ra = rlane_sh + cc * (uint32_t)(L::kLanes * L::kStride * 4);
              if (ph == 0) {
                strip_column4<R, U>(H0, H1, El, SA0, SA1, SB0, SB1, ra, prevUpH, prevUpSA, prevUpSB, rF, rSA, rSB, ngo2, c, outF);
              } else {
                strip_column4<R, U>(H1, H0, El, SA1, SA0, SB1, SB0, ra, prevUpH, prevUpSA, prevUpSB, rF, rSA, rSB, ngo2, c, outF);
              }
              prevUpH = rH;
              prevUpSA = rSA;
              prevUpSB = rSB;
            }
          }
        }
      }
      base_p += n + kCoGap;
      base_c += n + kCoGap;
      __syncwarp();
      if (lane == 0) {
        if (producer) st_release_shared(my_prod_sh, (uint32_t)base_p);
        if (consumer) st_release_shared(my_cons_sh, (uint32_t)base_c);
      }
      if (ownA || ownB) {
        const bool in1 = (((lane + n) & 1) != 0);  // the set this lane's last column (step lane + n - 1) wrote
        uint32_t resA = 0u, resB = 0u;
#pragma unroll
        for (int k = 0; k < R; ++k) {
          if (k == kmA) resA = in1 ? SA1[k] : SA0[k];
          if (k == kmB) resB = in1 ? SB1[k] : SB0[k];
        }
        resA = __shfl_sync(full, resA, LmA & 31);
        resB = __shfl_sync(full, resB, LmB & 31);
        if (lane == 0) {
          if (ownA) {
            res_m1[pp] = stat2_matches1(resA);
            res_l1[pp] = (uint32_t)(m1 + n) - stat2_diag(resA);
          }
          if (ownB) {
            res_m2[pp] = stat2_matches2(resB);
            res_l2[pp] = (uint32_t)(m2 + n) - stat2_diag(resB);
          }
        }
      }
    }
    __syncthreads();
    {
      const int row2 = row + (reload_i32(&units[u].j_count) >> 16);
      const int64_t slot1 = pair_slot(d.n, row, un.j_begin, d.slab_base);
      const int skip = un.j_begin < row2 ? row2 - un.j_begin : 0;
      const int64_t slot2 = pair_slot(d.n, row2, un.j_begin + skip, d.slab_base) - skip;
      for (int q = tid; q < un.j_count; q += kCoThreads) {
        d.matches[slot1 + q] = res_m1[q];
        d.length[slot1 + q] = res_l1[q];
        if (q >= skip) {
          d.matches[slot2 + q] = res_m2[q];
          d.length[slot2 + q] = res_l2[q];
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K4: one thread per pair (rows <= R <= 32)
// ------------------------------------------------------------------------------------------------
constexpr int kThreadThreads = 128;

template <int R, bool SLANT>
__global__ void __launch_bounds__(kThreadThreads)
nw_thread_kernel(NwDeviceData d, const NwUnit* __restrict__ units, int num_units) {
  using S = Strip<R>;
  __shared__ uint32_t prof[24 * S::RWS];
  const int tid = threadIdx.x;
  const int go = d.gap_open, ge = d.gap_ext;
  const Gap<SLANT> g = make_gap<SLANT>(go, ge);
  const int neg_init = neg_init_value(go, ge);
  const uint32_t one = d.one;

  for (int u = blockIdx.x; u < num_units; u += gridDim.x) {
    const NwUnit un = units[u];
    const int row = un.row;
    const int m = d.off[row + 1] - d.off[row];
    __syncthreads();
    build_profile<R, 1>(prof, d.codes + d.off[row], m, 0, d.sub, 2 * g.sl, tid, kThreadThreads);
    __syncthreads();
    for (int jj = tid; jj < un.j_count; jj += kThreadThreads) {
      const int j = un.j_begin + jj;
      const int n = d.off[j + 1] - d.off[j];
      const uint8_t* __restrict__ b = d.codes + d.off[j];
      int H0[R], H1[R], El[R];
      uint32_t S0[R], S1[R];
#pragma unroll
      for (int k = 0; k < R; ++k) {
        const int i = k + 1;
        H0[k] = H1[k] = wadd(wsub(-go, wmul(i - 1, ge)), wmul(i, g.sl));
        El[k] = wadd(neg_init, wmul(i + 1, g.sl));
        S0[k] = S1[k] = 0u;
      }
      int diagH = 0;  // Bd(0) at (0,0), slant 0
      int outF;
      for (int t0 = 0; t0 < n; t0 += 2) {
#pragma unroll
        for (int ph = 0; ph < 2; ++ph) {
          const int t = t0 + ph;
          if (t < n) {
            const int c = b[t];
            uint32_t pw[S::RW];
#pragma unroll
            for (int w = 0; w < S::RW; ++w) pw[w] = prof[c * S::RWS + w];
            const int rF = wadd(neg_init, wmul(t + 2, g.sl));  // Ix[1][t+1]
            if (ph == 0)
              strip_column<R, SLANT>(H0, S0, H1, S1, El, pw, diagH, 0u, rF, 0u, g, one, outF);
            else
              strip_column<R, SLANT>(H1, S1, H0, S0, El, pw, diagH, 0u, rF, 0u, g, one, outF);
            diagH = wadd(wsub(-go, wmul(t, ge)), wmul(t + 1, g.sl));  // Bd(t+1) at (0, t+1) for the next column
          }
        }
      }
      const bool in1 = ((n & 1) != 0);
      uint32_t res = 0u;
#pragma unroll
      for (int k = 0; k < R; ++k)
        if (k == m - 1) res = in1 ? S1[k] : S0[k];
      const int64_t slot = pair_slot(d.n, row, j, d.slab_base);
      d.matches[slot] = res >> 16;
      d.length[slot] = (uint32_t)(m + n) - (res & 0xFFFFu);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K4x2: one thread per TWO pairs (rows <= 32), 16-bit lanes -- the short-probe kernel with the same s16x2 packing
// as nw_warp2_kernel: a thread walks two column sequences against the CTA's row sequence, pair A in the low halves,
// pair B in the high halves.  Stat increments stay on the PRMT form (VAR 1): every lane reads its own residue class
// here, so the 128-bit increment-table loads of VAR 2 cost four shared-memory wavefronts each and the kernel turns
// shared-memory bound (measured 2.73 vs 3.09 TCUPS on 50,000 16-mers).  A single register set (the INPLACE form of
// the warp kernel) at 96 registers / 5 CTAs per SM was also measured and dropped: 2.55 vs 3.20 TCUPS.
// ------------------------------------------------------------------------------------------------
// Strips up to 16 rows are held to 128 registers (4 CTAs per SM instead of 3; a few spilled words at R = 16):
// measured 3.20 vs 3.09 TCUPS on 16-mers.  Taller strips keep the full register budget.
template <int R>
__global__ void __launch_bounds__(kThreadThreads, R <= 16 ? 4 : 1)
nw_thread2_kernel(NwDeviceData d, const NwUnit* __restrict__ units, int num_units) {
  using S = Strip<R>;
  __shared__ uint32_t prof[25 * S::RWS];  // class 24 = padding residue
  const int tid = threadIdx.x;
  const int go = d.gap_open, ge = d.gap_ext;
  const uint32_t ngo2 = pack16(-go);
  Stat2Consts c;
  c.one = d.one;
  c.zero = d.one - 1u;
  // The sentinel must live in a register the compiler cannot see through: with an immediate operand ptxas commutes
  // VIMNMX.S16x2 and the ">=" predicates it returns come back with the wrong sense (observed with CUDA 12.9:
  // the first DP row then takes 'U' where it must take 'L').  "+ opaque zero" keeps it a plain register operand.
  const uint32_t sent2 = pack16(kSentinel16) + c.zero;
  const uint32_t bord2 = pack16(ge - go);

  for (int u = blockIdx.x; u < num_units; u += gridDim.x) {
    const NwUnit un = units[u];
    const int row = un.row;
    const int m = d.off[row + 1] - d.off[row];
    __syncthreads();
    build_profile<R, 1>(prof, d.codes + d.off[row], m, 0, d.sub, 2 * ge, tid, kThreadThreads);
    for (int idx = tid; idx < S::RWS; idx += kThreadThreads) prof[24 * S::RWS + idx] = 0u;
    __syncthreads();
    const int npairs2 = (un.j_count + 1) >> 1;
    for (int pp = tid; pp < npairs2; pp += kThreadThreads) {
      int jA = un.j_begin + 2 * pp;
      int jB = jA + 1;
      const bool hasB = (jB < un.j_begin + un.j_count);
      if (!hasB) jB = jA;
      int nA = d.off[jA + 1] - d.off[jA], nB = d.off[jB + 1] - d.off[jB];
      if (nB > nA) {
        int tj = jA; jA = jB; jB = tj;
        int tn = nA; nA = nB; nB = tn;
      }
      const uint8_t* __restrict__ bA = d.codes + d.off[jA];
      const uint8_t* __restrict__ bB = d.codes + d.off[jB];
      uint32_t H0[R], H1[R], El[R], SA0[R], SA1[R], SB0[R], SB1[R];
#pragma unroll
      for (int k = 0; k < R; ++k) {
        H0[k] = H1[k] = bord2;
        El[k] = sent2;
        SA0[k] = SA1[k] = SB0[k] = SB1[k] = 0u;
      }
      uint32_t diagH = 0u;  // corner (0,0)
      uint32_t resB = 0u, outF;
      for (int t0 = 0; t0 < nA; t0 += 2) {
#pragma unroll
        for (int ph = 0; ph < 2; ++ph) {
          const int t = t0 + ph;
          if (t < nA) {
            const int cA = bA[t];
            const int cB = (t < nB) ? (int)bB[t] : 24;
            uint32_t pwA[S::RW], pwB[S::RW];
#pragma unroll
            for (int w = 0; w < S::RW; ++w) {
              pwA[w] = prof[cA * S::RWS + w];
              pwB[w] = prof[cB * S::RWS + w];
            }
            if (ph == 0)
              strip_column2<R, 1>(H0, H1, El, SA0, SA1, SB0, SB1, pwA, pwB, diagH, 0u, 0u, sent2, 0u, 0u, ngo2, c, outF);
            else
              strip_column2<R, 1>(H1, H0, El, SA1, SA0, SB1, SB0, pwA, pwB, diagH, 0u, 0u, sent2, 0u, 0u, ngo2, c, outF);
            diagH = bord2;  // border row, slanted: -go + ge for every column >= 1
            if (t == nB - 1) {
#pragma unroll
              for (int k = 0; k < R; ++k)
                if (k == m - 1) resB = (ph == 0) ? SB1[k] : SB0[k];
            }
          }
        }
      }
      const bool in1 = ((nA & 1) != 0);
      uint32_t resA = 0u;
#pragma unroll
      for (int k = 0; k < R; ++k)
        if (k == m - 1) resA = in1 ? SA1[k] : SA0[k];
      const int64_t slotA = pair_slot(d.n, row, jA, d.slab_base);
      d.matches[slotA] = resA >> 16;
      d.length[slotA] = (uint32_t)(m + nA) - (resA & 0xFFFFu);
      if (hasB) {
        const int64_t slotB = pair_slot(d.n, row, jB, d.slab_base);
        d.matches[slotB] = resB >> 16;
        d.length[slotB] = (uint32_t)(m + nB) - (resB & 0xFFFFu);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K4x2 "two rows": one thread per column sequence, the two 16-bit halves carry rows i and i+1 (both <= 32 residues) --
// the short-probe counterpart of nw_rows2_kernel.  Every thread of the CTA aligns the same two row sequences, so the
// table has no lane dimension: per residue class and strip row two words (packed scores, the increment both pairs
// share: Rec2), class stride odd, so that the 32 lanes of a warp -- each at its own column residue -- read
// conflict-free 32-bit words (equal classes broadcast).  Against nw_thread2_kernel (one row, two column sequences) a
// row costs two shared loads instead of one load and three PRMT: 5 instead of 8 ALU-pipe instructions per row, the
// pipe that kernel saturates (ncu r01d: ALU 78 %).
// ------------------------------------------------------------------------------------------------
// one column of the two-rows thread kernel: rows 0..R-1 of both row sequences against column residue class `base`
template <int R, bool U>
__device__ __forceinline__ void thread_rows2_column(const uint32_t (&Ho)[R], uint32_t (&Hn)[R], uint32_t (&El)[R],
                                                    const uint32_t (&SAo)[R], uint32_t (&SAn)[R], const uint32_t (&SBo)[R],
                                                    uint32_t (&SBn)[R], uint32_t base, uint32_t diagH, uint32_t F, uint32_t ngo2,
                                                    uint32_t pgo2, const Stat2Consts& c) {
  uint32_t dSA = 0u, dSB = 0u, upSA = 0u, upSB = 0u;
#pragma unroll
  for (int k = 0; k < R; ++k) {
    const uint32_t sP = lds_u32(base + 8u * (unsigned)k);
    const uint32_t incA = lds_u32(base + 8u * (unsigned)k + 4u), incB = incA;  // one increment word for both pairs (Rec2)
    const uint32_t E = El[k];
    const uint32_t Mraw = U ? diagH + sP : __viaddmax_s16x2(diagH, sP, 0x80008000u);  // U: see strip_column4
    if (U && k == 0) {
      // First row, unsigned domain: the vertical gap from the border row is "minus infinity" = 0, so max(F, E) = E, the
      // "came from above" predicate can only hold where E is the sentinel too (column 0), and there the statistics above
      // and to the left are both 0: the select has two cases, and F' = max(H - go, 0) is a plain subtraction (every real
      // value exceeds go in both halves, see bias16 in cabi.cu).  Also keeps the compiler from expanding the DPX compare
      // of a loop-invariant F into seven scalar instructions per column, which it did.
      bool pdB, pdA;
      const uint32_t H = __vibmax_u16x2(Mraw, E, &pdB, &pdA);
      uint32_t SA, SB;
      asm("{\n\t.reg .pred pd;\n\tsetp.ne.u32 pd, %3, 0;\n\tadd.u32 %0, %1, %4;\n\t@pd add.u32 %0, %2, %4;\n\t}"
          : "=&r"(SA) : "r"(SAo[0]), "r"(incA), "r"((uint32_t)pdA), "r"(c.zero));
      asm("{\n\t.reg .pred pd;\n\tsetp.ne.u32 pd, %3, 0;\n\tadd.u32 %0, %1, %4;\n\t@pd add.u32 %0, %2, %4;\n\t}"
          : "=&r"(SB) : "r"(SBo[0]), "r"(incB), "r"((uint32_t)pdB), "r"(c.zero));
      diagH = Ho[0];
      dSA = SAo[0];
      dSB = SBo[0];
      Hn[0] = H;
      SAn[0] = SA;
      SBn[0] = SB;
      El[0] = __viaddmax_u16x2(H, ngo2, E);
      F = H - pgo2;  // per-half H - go, pgo2 = [go : go]: no borrow, both halves exceed go
      upSA = SA;
      upSB = SB;
      continue;
    }
    bool puB, puA, pdB, pdA;
    const uint32_t g = U ? __vibmax_u16x2(F, E, &puB, &puA) : __vibmax_s16x2(F, E, &puB, &puA);
    const uint32_t H = U ? __vibmax_u16x2(Mraw, g, &pdB, &pdA) : __vibmax_s16x2(Mraw, g, &pdB, &pdA);
    const uint32_t SA = (((U ? DYNA_TROWS2_SELMASK_UA : DYNA_TROWS2_SELMASK_A) >> k) & 1) ? stat_select_sel(SAo[k], upSA, dSA, incA, puA, pdA)
                                                          : stat_select(SAo[k], upSA, dSA, incA, puA, pdA, c.zero);
    const uint32_t SB = (((U ? DYNA_TROWS2_SELMASK_UB : DYNA_TROWS2_SELMASK_B) >> k) & 1) ? stat_select_sel(SBo[k], upSB, dSB, incB, puB, pdB)
                                                          : stat_select(SBo[k], upSB, dSB, incB, puB, pdB, c.zero);
    diagH = Ho[k];
    dSA = SAo[k];
    dSB = SBo[k];
    Hn[k] = H;
    SAn[k] = SA;
    SBn[k] = SB;
    El[k] = U ? __viaddmax_u16x2(H, ngo2, E) : __viaddmax_s16x2(H, ngo2, E);
    F = U ? __viaddmax_u16x2(H, ngo2, F) : __viaddmax_s16x2(H, ngo2, F);
    upSA = SA;
    upSB = SB;
  }
}

template <int R, bool U>
__global__ void __launch_bounds__(kThreadThreads, R <= 16 ? 4 : 1)
nw_thread_rows2_kernel(NwDeviceData d, const NwUnit* __restrict__ units, int num_units) {
  constexpr int TS = (2 * R) | 1;  // words per residue class (odd: conflict-free across classes)
  __shared__ uint32_t tab[24 * TS];
  const int tid = threadIdx.x;
  const int go = d.gap_open, ge = d.gap_ext;
  const uint32_t ngo2 = pack16(-go), pgo2 = pack16(go);
  Stat2Consts c;
  c.one = d.one;
  c.zero = d.zero;
  const uint32_t sent2 = (U ? 0u : pack16(kSentinel16)) + c.zero;  // register operand (see nw_thread2_kernel)
  const uint32_t bord2 = pack16(ge - go + (U ? (int)d.bias16 : 0));
  const uint32_t corner2 = U ? pack16((int)d.bias16) : 0u;
  uint32_t tab_sh;  // opaque: otherwise the shared-window base is rebuilt (S2R + MOV + LEA) in every column
  asm volatile("mov.u32 %0, %1;" : "=r"(tab_sh) : "r"((uint32_t)__cvta_generic_to_shared(tab)));

  for (int u = blockIdx.x; u < num_units; u += gridDim.x) {
    NwUnit un = units[u];
    const int row = un.row, row2 = un.row + (un.j_count >> 16);
    un.j_count &= 0xFFFF;
    const int m1 = d.off[row + 1] - d.off[row], m2 = d.off[row2 + 1] - d.off[row2];
    const uint8_t* __restrict__ a1 = d.codes + d.off[row];
    const uint8_t* __restrict__ a2 = d.codes + d.off[row2];
    __syncthreads();
    for (int idx = tid; idx < 24 * TS; idx += kThreadThreads) {
      const int cls = idx / TS, w = idx - cls * TS;
      uint32_t v = 0u;
      if (w < 2 * R) {
        const int k = w >> 1;
        if ((w & 1) == 0) {
          const int s1 = k < m1 ? (int)(int8_t)(d.sub[a1[k] * 24 + cls] + 2 * ge) : 0;
          const int s2 = k < m2 ? (int)(int8_t)(d.sub[a2[k] * 24 + cls] + 2 * ge) : 0;
          v = score2_word(s1, s2, U);
        } else {
          v = inc2_word(k < m1 && a1[k] == cls, k < m2 && a2[k] == cls);
        }
      }
      tab[idx] = v;
    }
    __syncthreads();
    for (int jj = tid; jj < un.j_count; jj += kThreadThreads) {
      const int j = un.j_begin + jj;
      const int n = d.off[j + 1] - d.off[j];
      const uint8_t* __restrict__ b = d.codes + d.off[j];
      uint32_t H0[R], H1[R], El[R], SA0[R], SA1[R], SB0[R], SB1[R];
#pragma unroll
      for (int k = 0; k < R; ++k) {
        H0[k] = H1[k] = bord2;
        El[k] = sent2;
        SA0[k] = SA1[k] = SB0[k] = SB1[k] = 0u;
      }
      uint32_t diag0 = corner2;  // corner (0,0); the border row (slanted: -go + ge) for every later column
      const uint8_t* bp = b;     // walked as a pointer: b[t] cost a 64-bit add (three ALU-pipe instructions) per column
      for (int t0 = 0; t0 < n; t0 += 2, bp += 2) {
#pragma unroll
        for (int ph = 0; ph < 2; ++ph) {
          const int t = t0 + ph;
          if (t < n) {
            uint32_t cres;
            if (ph == 0) asm volatile("ld.global.nc.u8 %0, [%1];" : "=r"(cres) : "l"(bp));
            else asm volatile("ld.global.nc.u8 %0, [%1+1];" : "=r"(cres) : "l"(bp));
            const uint32_t base = tab_sh + cres * (uint32_t)(TS * 4);
            if (ph == 0) thread_rows2_column<R, U>(H0, H1, El, SA0, SA1, SB0, SB1, base, diag0, sent2, ngo2, pgo2, c);
            else thread_rows2_column<R, U>(H1, H0, El, SA1, SA0, SB1, SB0, base, diag0, sent2, ngo2, pgo2, c);
            diag0 = bord2;
          }
        }
      }
      const bool in1 = ((n & 1) != 0);
      uint32_t resA = 0u, resB = 0u;
#pragma unroll
      for (int k = 0; k < R; ++k) {
        if (k == m1 - 1) resA = in1 ? SA1[k] : SA0[k];
        if (k == m2 - 1) resB = in1 ? SB1[k] : SB0[k];
      }
      const int64_t slot1 = pair_slot(d.n, row, j, d.slab_base);
      d.matches[slot1] = stat2_matches1(resA);
      d.length[slot1] = (uint32_t)(m1 + n) - stat2_diag(resA);
      if (j >= row2) {  // column `row` exists for the first row only
        const int64_t slot2 = pair_slot(d.n, row2, j, d.slab_base);
        d.matches[slot2] = stat2_matches2(resB);
        d.length[slot2] = (uint32_t)(m2 + n) - stat2_diag(resB);
      }
    }
  }
}

// rows of length 0: no DP; the path is n left moves -> matches 0, length n (0/0 -> NaN handled at the division)
__global__ void nw_empty_rows_kernel(NwDeviceData d, const NwUnit* __restrict__ units, int num_units) {
  for (int u = blockIdx.x; u < num_units; u += gridDim.x) {
    const NwUnit un = units[u];
    for (int jj = threadIdx.x; jj < un.j_count; jj += blockDim.x) {
      const int j = un.j_begin + jj;
      const int64_t slot = pair_slot(d.n, un.row, j, d.slab_base);
      d.matches[slot] = 0u;
      d.length[slot] = (uint32_t)(d.off[j + 1] - d.off[j]);
    }
  }
}

template <int R>
int launch_warp_R(bool slant, const NwDeviceData& d, const NwUnit* d_units, int num_units, cudaStream_t st) {
  if (slant) nw_warp_kernel<R, true, false><<<num_units, kWarpThreads, 0, st>>>(d, d_units, num_units, nullptr, 0);
  else       nw_warp_kernel<R, false, false><<<num_units, kWarpThreads, 0, st>>>(d, d_units, num_units, nullptr, 0);
  DYNA_CUDA(cudaGetLastError());
  return DYNA_OK;
}

template <int R, int VAR, bool INPLACE, int THREADS>
int launch_warp2_inst(const NwDeviceData& d, const NwUnit* d_units, int num_units, cudaStream_t st) {
  using L = Warp2Smem<R, VAR, THREADS>;
  if constexpr (VAR == 2 || VAR == 3) {
    DYNA_CUDA(cudaFuncSetAttribute(nw_warp2_kernel<R, VAR, INPLACE, THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
    nw_warp2_kernel<R, VAR, INPLACE, THREADS><<<num_units, THREADS, L::kTotal, st>>>(d, d_units, num_units);
  } else {
    nw_warp2_kernel<R, VAR, INPLACE, THREADS><<<num_units, THREADS, 0, st>>>(d, d_units, num_units);
  }
  DYNA_CUDA(cudaGetLastError());
  return DYNA_OK;
}

template <int R>
int launch_warp2_R(const NwDeviceData& d, const NwUnit* d_units, int num_units, cudaStream_t st) {
  // measured on B200 (TCUPS, config 5 / config 2):
  //   strips R <= 12: ping-pong register sets, 8-warp CTAs, record table (VAR 3: int8 scores and increments in one record
  //                   per lane and class, 128-bit loads only): 3.38 against 3.26 for VAR 2 on the same box (round 2)
  //   history of VAR 2 (increments from their own shared-memory table): 2.94
  //                   (VAR 1, increments by PRMT: 2.72 on the same box; one register set: 2.58; with the final step
  //                   loop: table for all rows 3.17, for the first 8 rows 3.12, first 4 rows 3.08, PRMT only 3.00)
  //   strips R >= 13: one register set, 4-warp CTAs, increments by PRMT (VAR 1): 2.60 (VAR 2: 2.48; ping-pong: 2.27;
  //                   VAR 2 with one register set in 8-warp CTAs at 128 registers, 16 warps/SM, R <= 18: 2.56)
  if constexpr (R >= 13) return launch_warp2_inst<R, 1, true, 128>(d, d_units, num_units, st);
  else {
    const int var = getenv("DYNA_NW2_VARIANT") ? atoi(getenv("DYNA_NW2_VARIANT")) : 3;
    if (var == 1) return launch_warp2_inst<R, 1, false, kWarpThreads>(d, d_units, num_units, st);
    if (var == 3) return launch_warp2_inst<R, 3, false, kWarpThreads>(d, d_units, num_units, st);
    return launch_warp2_inst<R, 2, false, kWarpThreads>(d, d_units, num_units, st);
  }
}

int launch_warp_multipass(bool slant, const NwDeviceData& d, const NwUnit* d_units, int num_units, int32_t* d_scratch,
                          int max_cols, cudaStream_t st) {
  const int grid = std::min(num_units, kNwMultiPassGrid);
  if (slant)
    nw_warp_kernel<kNwWarpMaxR, true, true><<<grid, kWarpThreads, 0, st>>>(d, d_units, num_units, d_scratch, max_cols);
  else
    nw_warp_kernel<kNwWarpMaxR, false, true><<<grid, kWarpThreads, 0, st>>>(d, d_units, num_units, d_scratch, max_cols);
  DYNA_CUDA(cudaGetLastError());
  return DYNA_OK;
}

}  // namespace

int launch_nw_empty_rows(const NwDeviceData& d, const NwUnit* d_units, int num_units, cudaStream_t st) {
  if (num_units == 0) return DYNA_OK;
  nw_empty_rows_kernel<<<std::min(num_units, kNumSMsB200 * 8), 256, 0, st>>>(d, d_units, num_units);
  DYNA_CUDA(cudaGetLastError());
  return DYNA_OK;
}

int launch_nw_thread(int R, bool slant, const NwDeviceData& d, const NwUnit* d_units, int num_units, cudaStream_t st) {
  if (num_units == 0) return DYNA_OK;
  switch (R) {
#define DYNA_CASE(RR)                                                                                        \
  case RR:                                                                                                   \
    if (slant) nw_thread_kernel<RR, true><<<num_units, kThreadThreads, 0, st>>>(d, d_units, num_units);      \
    else       nw_thread_kernel<RR, false><<<num_units, kThreadThreads, 0, st>>>(d, d_units, num_units);     \
    break;
    DYNA_CASE(4) DYNA_CASE(8) DYNA_CASE(12) DYNA_CASE(16) DYNA_CASE(20) DYNA_CASE(24) DYNA_CASE(28) DYNA_CASE(32)
#undef DYNA_CASE
    default:
      return fail(DYNA_ERR_UNSUPPORTED, "nw thread kernel: unsupported strip height %d", R);
  }
  DYNA_CUDA(cudaGetLastError());
  return DYNA_OK;
}

int launch_nw_warp(int R, bool slant, bool multipass, const NwDeviceData& d, const NwUnit* d_units, int num_units, int32_t* d_scratch,
                   int max_cols, cudaStream_t st) {
  if (num_units == 0) return DYNA_OK;
  if (multipass) {
    if (R != kNwWarpMaxR) return fail(DYNA_ERR_UNSUPPORTED, "nw multipass requires R=%d", kNwWarpMaxR);
    return launch_warp_multipass(slant, d, d_units, num_units, d_scratch, max_cols, st);
  }
  switch (R) {
#define DYNA_CASE(RR) \
  case RR:            \
    return launch_warp_R<RR>(slant, d, d_units, num_units, st);
    DYNA_CASE(1) DYNA_CASE(2) DYNA_CASE(3) DYNA_CASE(4) DYNA_CASE(5) DYNA_CASE(6) DYNA_CASE(7) DYNA_CASE(8)
    DYNA_CASE(9) DYNA_CASE(10) DYNA_CASE(11) DYNA_CASE(12) DYNA_CASE(13) DYNA_CASE(14) DYNA_CASE(15) DYNA_CASE(16)
    DYNA_CASE(17) DYNA_CASE(18) DYNA_CASE(19) DYNA_CASE(20) DYNA_CASE(21) DYNA_CASE(22) DYNA_CASE(23) DYNA_CASE(24)
#undef DYNA_CASE
    default:
      return fail(DYNA_ERR_UNSUPPORTED, "nw warp kernel: unsupported strip height %d", R);
  }
}

template <int R>
int launch_warp2mp_R(const NwDeviceData& d, const NwUnit* d_units, int num_units, uint4* d_scratch, cudaStream_t st) {
  using L = Warp2Smem<R, 2, kWarpThreads, kNwMpStageCols>;
  DYNA_CUDA(cudaFuncSetAttribute(nw_warp2mp_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
  const int grid = std::min(num_units, kNwMultiPassGrid);
  nw_warp2mp_kernel<R><<<grid, kWarpThreads, L::kTotal, st>>>(d, d_units, num_units, d_scratch);
  DYNA_CUDA(cudaGetLastError());
  return DYNA_OK;
}

int launch_nw_warp2mp(int R, const NwDeviceData& d, const NwUnit* d_units, int num_units, void* d_scratch, cudaStream_t st) {
  if (num_units == 0) return DYNA_OK;
  uint4* scr = static_cast<uint4*>(d_scratch);
  switch (R) {
#define DYNA_CASE(RR) \
  case RR:            \
    return launch_warp2mp_R<RR>(d, d_units, num_units, scr, st);
    DYNA_CASE(7) DYNA_CASE(8) DYNA_CASE(9) DYNA_CASE(10) DYNA_CASE(11) DYNA_CASE(12)
#undef DYNA_CASE
    default:
      return fail(DYNA_ERR_UNSUPPORTED, "nw warp2 multipass kernel: unsupported strip height %d", R);
  }
}

template <int R, int VAR>
int launch_warp2co_inst(const NwDeviceData& d, const NwUnit* d_units, int num_units, cudaStream_t st) {
  using L = CoSmem<R, 2, VAR>;
  DYNA_CUDA(cudaFuncSetAttribute(nw_warp2co_kernel<R, 2, VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
  nw_warp2co_kernel<R, 2, VAR><<<num_units, kCoThreads, L::kTotal, st>>>(d, d_units, num_units);
  DYNA_CUDA(cudaGetLastError());
  return DYNA_OK;
}
template <int R>
int launch_warp2co_R(const NwDeviceData& d, const NwUnit* d_units, int num_units, cudaStream_t st) {
  const int var = getenv("DYNA_NW_CO_VARIANT") ? atoi(getenv("DYNA_NW_CO_VARIANT")) : 3;
  if (var == 2) return launch_warp2co_inst<R, 2>(d, d_units, num_units, st);
  return launch_warp2co_inst<R, 3>(d, d_units, num_units, st);
}

int launch_nw_warp2co(int R, const NwDeviceData& d, const NwUnit* d_units, int num_units, cudaStream_t st) {
  if (num_units == 0) return DYNA_OK;
  switch (R) {
#define DYNA_CASE(RR) \
  case RR:            \
    return launch_warp2co_R<RR>(d, d_units, num_units, st);
    DYNA_CASE(7) DYNA_CASE(8) DYNA_CASE(9) DYNA_CASE(10) DYNA_CASE(11) DYNA_CASE(12)
#undef DYNA_CASE
    default:
      return fail(DYNA_ERR_UNSUPPORTED, "nw cooperative warp2 kernel: unsupported strip height %d", R);
  }
}

template <int R>
int launch_rows2_R(const NwDeviceData& d, const NwUnit* d_units, int num_units, cudaStream_t st) {
  using RC = Rec2<R>;
  if (d.bias16 != 0u) {  // unsigned domain (host: every table score + 2*ge is non-negative, values + bias16 fit 16 bits)
    DYNA_CUDA(cudaFuncSetAttribute(nw_rows2_kernel<R, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, RC::kTotal));
    nw_rows2_kernel<R, true><<<num_units, kRows2Threads, RC::kTotal, st>>>(d, d_units, num_units);
  } else {
    DYNA_CUDA(cudaFuncSetAttribute(nw_rows2_kernel<R, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, RC::kTotal));
    nw_rows2_kernel<R, false><<<num_units, kRows2Threads, RC::kTotal, st>>>(d, d_units, num_units);
  }
  DYNA_CUDA(cudaGetLastError());
  return DYNA_OK;
}

template <int R>
int launch_rows2mp_R(const NwDeviceData& d, const NwUnit* d_units, int num_units, uint4* scr, cudaStream_t st) {
  using RC = Rec3<R>;
  DYNA_CUDA(cudaFuncSetAttribute(nw_rows2mp_kernel<R, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, RC::kTotal));
  nw_rows2mp_kernel<R, true><<<std::min(num_units, kNwRows2MpGrid), kRows2Threads, RC::kTotal, st>>>(d, d_units, num_units, scr);
  DYNA_CUDA(cudaGetLastError());
  return DYNA_OK;
}

int launch_nw_rows2mp(int R, const NwDeviceData& d, const NwUnit* d_units, int num_units, void* d_scratch, cudaStream_t st) {
  if (num_units == 0) return DYNA_OK;
  if (d.bias16 == 0u) return fail(DYNA_ERR_UNSUPPORTED, "nw two-rows multi-pass kernel: needs the unsigned 16-bit domain");
  uint4* scr = static_cast<uint4*>(d_scratch);
  switch (R) {
#define DYNA_CASE(RR) \
  case RR:            \
    return launch_rows2mp_R<RR>(d, d_units, num_units, scr, st);
    DYNA_CASE(7) DYNA_CASE(8) DYNA_CASE(9) DYNA_CASE(10) DYNA_CASE(11) DYNA_CASE(12)
#undef DYNA_CASE
    default:
      return fail(DYNA_ERR_UNSUPPORTED, "nw two-rows multi-pass kernel: unsupported strip height %d", R);
  }
}

int launch_nw_rows2(int R, const NwDeviceData& d, const NwUnit* d_units, int num_units, cudaStream_t st) {
  if (num_units == 0) return DYNA_OK;
  switch (R) {
#define DYNA_CASE(RR) \
  case RR:            \
    return launch_rows2_R<RR>(d, d_units, num_units, st);
    DYNA_CASE(2) DYNA_CASE(3) DYNA_CASE(4) DYNA_CASE(5) DYNA_CASE(6) DYNA_CASE(7) DYNA_CASE(8) DYNA_CASE(9)
    DYNA_CASE(10) DYNA_CASE(11) DYNA_CASE(12)
#undef DYNA_CASE
    default:
      return fail(DYNA_ERR_UNSUPPORTED, "nw two-rows kernel: unsupported strip height %d", R);
  }
}

template <int R>
int launch_rows2co_R(const NwDeviceData& d, const NwUnit* d_units, int num_units, cudaStream_t st) {
  using L = Rows2CoSmem<R, 2>;
  if (d.bias16 != 0u) {
    DYNA_CUDA(cudaFuncSetAttribute(nw_rows2co_kernel<R, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
    nw_rows2co_kernel<R, 2, true><<<num_units, kCoThreads, L::kTotal, st>>>(d, d_units, num_units);
  } else {
    DYNA_CUDA(cudaFuncSetAttribute(nw_rows2co_kernel<R, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
    nw_rows2co_kernel<R, 2, false><<<num_units, kCoThreads, L::kTotal, st>>>(d, d_units, num_units);
  }
  DYNA_CUDA(cudaGetLastError());
  return DYNA_OK;
}

int launch_nw_rows2co(int R, const NwDeviceData& d, const NwUnit* d_units, int num_units, cudaStream_t st) {
  if (num_units == 0) return DYNA_OK;
  switch (R) {
#define DYNA_CASE(RR) \
  case RR:            \
    return launch_rows2co_R<RR>(d, d_units, num_units, st);
    DYNA_CASE(7) DYNA_CASE(8) DYNA_CASE(9) DYNA_CASE(10) DYNA_CASE(11) DYNA_CASE(12)
#undef DYNA_CASE
    default:
      return fail(DYNA_ERR_UNSUPPORTED, "nw cooperative two-rows kernel: unsupported strip height %d", R);
  }
}

int launch_nw_thread2(int R, const NwDeviceData& d, const NwUnit* d_units, int num_units, cudaStream_t st) {
  if (num_units == 0) return DYNA_OK;
  switch (R) {
#define DYNA_CASE(RR)                                                                  \
  case RR:                                                                             \
    nw_thread2_kernel<RR><<<num_units, kThreadThreads, 0, st>>>(d, d_units, num_units); \
    break;
    DYNA_CASE(4) DYNA_CASE(8) DYNA_CASE(12) DYNA_CASE(16) DYNA_CASE(20) DYNA_CASE(24) DYNA_CASE(28) DYNA_CASE(32)
#undef DYNA_CASE
    default:
      return fail(DYNA_ERR_UNSUPPORTED, "nw thread2 kernel: unsupported strip height %d", R);
  }
  DYNA_CUDA(cudaGetLastError());
  return DYNA_OK;
}

int launch_nw_thread_rows2(int R, const NwDeviceData& d, const NwUnit* d_units, int num_units, cudaStream_t st) {
  if (num_units == 0) return DYNA_OK;
  switch (R) {
#define DYNA_CASE(RR)                                                                        \
  case RR:                                                                                   \
    if (d.bias16 != 0u) nw_thread_rows2_kernel<RR, true><<<num_units, kThreadThreads, 0, st>>>(d, d_units, num_units); \
    else nw_thread_rows2_kernel<RR, false><<<num_units, kThreadThreads, 0, st>>>(d, d_units, num_units);               \
    break;
    DYNA_CASE(4) DYNA_CASE(8) DYNA_CASE(12) DYNA_CASE(16) DYNA_CASE(20) DYNA_CASE(24) DYNA_CASE(28) DYNA_CASE(32)
#undef DYNA_CASE
    default:
      return fail(DYNA_ERR_UNSUPPORTED, "nw thread two-rows kernel: unsupported strip height %d", R);
  }
  DYNA_CUDA(cudaGetLastError());
  return DYNA_OK;
}

int launch_nw_warp2(int R, const NwDeviceData& d, const NwUnit* d_units, int num_units, cudaStream_t st) {
  if (num_units == 0) return DYNA_OK;
  switch (R) {
#define DYNA_CASE(RR) \
  case RR:            \
    return launch_warp2_R<RR>(d, d_units, num_units, st);
    DYNA_CASE(2) DYNA_CASE(3) DYNA_CASE(4) DYNA_CASE(5) DYNA_CASE(6) DYNA_CASE(7) DYNA_CASE(8) DYNA_CASE(9)
    DYNA_CASE(10) DYNA_CASE(11) DYNA_CASE(12) DYNA_CASE(13) DYNA_CASE(14) DYNA_CASE(15) DYNA_CASE(16) DYNA_CASE(17)
    DYNA_CASE(18) DYNA_CASE(19) DYNA_CASE(20)
#undef DYNA_CASE
    default:
      return fail(DYNA_ERR_UNSUPPORTED, "nw warp2 kernel: unsupported strip height %d", R);
  }
}

}  // namespace dyna
