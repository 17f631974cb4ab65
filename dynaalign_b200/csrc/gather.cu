// Result assembly kernels (sm_100a).
//
// The kernels of nw_kernels.cu / mh_kernels.cu leave (matches, length) or match counts in the row-major packed upper
// triangle, one contiguous slab per row block -- and, when one process drives several GPUs, one slab per GPU.  The
// reference hands R ONE column-major n x n double matrix with both triangles filled (src/pairwiseSeqAlign.cpp:349-350,
// src/minHash.cpp:174-176).  The expansion below produces that matrix in COLUMN BLOCKS: device g writes columns
// [c_g, c_g+1) -- a contiguous range of the caller's matrix, so its device-to-host copy is one plain copy that runs
// in parallel with the other devices' -- and reads whatever slabs those columns need straight out of the other
// devices' memory (peer loads over NVLink / NVSwitch; the slabs are never staged through the host or copied whole).
//   column c, rows r >= c : pair (c, r), consecutive r are consecutive slots of row c         -> coalesced as is
//   column c, rows r <  c : pair (r, c), consecutive c are consecutive slots of row r         -> read tiles along
//                           the rows, transpose through shared memory, write along the column
#include "gather.cuh"

namespace dyna {
namespace {

__device__ __forceinline__ int slab_of_row(const TriSlabs& s, int64_t row) {
  int g = 0;
  while (g + 1 < s.nslabs && row >= s.row_begin[g + 1]) ++g;
  return g;
}

// value of pair (i, j), i <= j, from the slab that owns row i
struct NwValue {
  int64_t n;
  __device__ __forceinline__ double operator()(const TriSlabs& s, int64_t i, int64_t j) const {
    const int g = slab_of_row(s, i);
    const int64_t rb = s.row_begin[g];
    const int64_t slot = (i * n - i * (i - 1) / 2 + (j - i)) - (rb * n - rb * (rb - 1) / 2);
    const uint32_t m = static_cast<const uint32_t*>(s.a[g])[slot];
    const uint32_t l = static_cast<const uint32_t*>(s.b[g])[slot];
    // static_cast<double>(matches) / alignment_length (src/pairwiseSeqAlign.cpp:311): IEEE double divide, 0/0 = NaN
    return __ddiv_rn((double)m, (double)l);
  }
};
struct MhValue {
  int64_t n;
  const double* table;
  double diag;
  __device__ __forceinline__ double operator()(const TriSlabs& s, int64_t i, int64_t j) const {
    if (i == j) return diag;  // similarityMatrix(i, i) = 1.0 (src/minHash.cpp:161) / dist 0 (R/minHash.R:171)
    const int g = slab_of_row(s, i);
    const int64_t rb = s.row_begin[g];
    const int64_t slot = (i * n - i * (i + 1) / 2 + (j - i - 1)) - (rb * n - rb * (rb + 1) / 2);
    return table[static_cast<const uint16_t*>(s.a[g])[slot]];
  }
};

constexpr int kTile = 32;

template <class Value>
__global__ void __launch_bounds__(kTile * 8)
tri_expand_block_kernel(TriSlabs s, Value val, int64_t n, int64_t col_begin, int64_t col_end, double* __restrict__ out) {
  __shared__ double tile[kTile][kTile + 1];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int64_t row_tiles = (n + kTile - 1) / kTile;
  for (int64_t t = blockIdx.x; t < row_tiles * ((col_end - col_begin + kTile - 1) / kTile); t += gridDim.x) {
    const int64_t R0 = (t % row_tiles) * kTile, C0 = col_begin + (t / row_tiles) * kTile;
    if (R0 >= C0 + kTile - 1) {
      // on / below the diagonal everywhere: pair (c, r), r runs along the slab row
#pragma unroll
      for (int k = 0; k < kTile; k += 8) {
        const int64_t c = C0 + ty + k, r = R0 + tx;
        if (c < col_end && r < n) out[(c - col_begin) * n + r] = val(s, c, r);
      }
    } else if (R0 + kTile - 1 <= C0) {
      // on / above the diagonal everywhere: pair (r, c), c runs along the slab row -> transpose
      __syncthreads();
#pragma unroll
      for (int k = 0; k < kTile; k += 8) {
        const int64_t r = R0 + ty + k, c = C0 + tx;
        if (c < col_end && r < n) tile[ty + k][tx] = val(s, r, c);
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < kTile; k += 8) {
        const int64_t c = C0 + ty + k, r = R0 + tx;
        if (c < col_end && r < n) out[(c - col_begin) * n + r] = tile[tx][ty + k];
      }
    } else {
      // the diagonal crosses this tile
#pragma unroll
      for (int k = 0; k < kTile; k += 8) {
        const int64_t c = C0 + ty + k, r = R0 + tx;
        if (c < col_end && r < n) out[(c - col_begin) * n + r] = r >= c ? val(s, c, r) : val(s, r, c);
      }
    }
  }
}

template <class Value>
int launch_expand(const TriSlabs& s, const Value& v, int64_t n, int64_t c0, int64_t c1, double* d_out, cudaStream_t st) {
  if (c1 <= c0 || n <= 0) return DYNA_OK;
  const int64_t tiles = ((n + kTile - 1) / kTile) * ((c1 - c0 + kTile - 1) / kTile);
  const unsigned grid = (unsigned)std::min<int64_t>(tiles, (int64_t)kNumSMsB200 * 64);
  tri_expand_block_kernel<Value><<<grid, dim3(kTile, 8), 0, st>>>(s, v, n, c0, c1, d_out);
  DYNA_CUDA(cudaGetLastError());
  return DYNA_OK;
}

// ---- checksums: sum of value[k] * mix(global index) in wrap-around 64-bit arithmetic.  The weight depends on the
// pair's GLOBAL position in the packed triangle, so the sum over all ranks' slabs equals the single-device sum only
// if the slabs tile the triangle exactly (an off-by-one row in a partition changes it).
__device__ __forceinline__ unsigned long long mix_index(unsigned long long k) { return checksum_weight(k); }

template <class T>
__global__ void __launch_bounds__(256)
checksum_kernel(const T* __restrict__ v, int64_t count, int64_t first, unsigned long long* __restrict__ sum) {
  unsigned long long acc = 0;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < count; k += (int64_t)gridDim.x * blockDim.x)
    acc += (unsigned long long)v[k] * mix_index((unsigned long long)(first + k));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, o);
  __shared__ unsigned long long part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long b = 0;
    for (int w = 0; w < 8; ++w) b += part[w];
    atomicAdd(sum, b);
  }
}

template <class T>
int launch_checksum(const T* d_v, int64_t count, int64_t first, unsigned long long* d_sum, cudaStream_t st) {
  DYNA_CUDA(cudaMemsetAsync(d_sum, 0, sizeof(unsigned long long), st));
  if (count <= 0) return DYNA_OK;
  const unsigned grid = (unsigned)std::min<int64_t>((count + 255) / 256, (int64_t)kNumSMsB200 * 16);
  checksum_kernel<T><<<grid, 256, 0, st>>>(d_v, count, first, d_sum);
  DYNA_CUDA(cudaGetLastError());
  return DYNA_OK;
}

__global__ void __launch_bounds__(256)
nw_pack8_kernel(const uint32_t* __restrict__ m, const uint32_t* __restrict__ l, int64_t count, uint8_t* __restrict__ m8,
                uint8_t* __restrict__ l8) {
  // four pairs per thread: 128-bit loads, 32-bit stores (the slabs and the outputs are 16-byte aligned allocations)
  const int64_t quads = count >> 2;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < quads; q += (int64_t)gridDim.x * blockDim.x) {
    const uint4 a = reinterpret_cast<const uint4*>(m)[q], b = reinterpret_cast<const uint4*>(l)[q];
    reinterpret_cast<uint32_t*>(m8)[q] = a.x | (a.y << 8) | (a.z << 16) | (a.w << 24);
    reinterpret_cast<uint32_t*>(l8)[q] = b.x | (b.y << 8) | (b.z << 16) | (b.w << 24);
  }
  if (blockIdx.x == 0 && threadIdx.x < (count & 3)) {
    const int64_t k = (quads << 2) + threadIdx.x;
    m8[k] = (uint8_t)m[k];
    l8[k] = (uint8_t)l[k];
  }
}

__global__ void __launch_bounds__(256)
mh_narrow8_kernel(const uint16_t* __restrict__ counts, int64_t count, int64_t first, uint8_t* __restrict__ out8,
                  int64_t esc_capacity, long long* __restrict__ esc_index, uint16_t* __restrict__ esc_count,
                  unsigned long long* __restrict__ esc_n) {
  // eight counts per thread (128-bit load, 64-bit store) while the slab offset keeps them aligned; scalar otherwise
  const bool aligned = ((reinterpret_cast<uintptr_t>(counts) & 15) == 0) && ((reinterpret_cast<uintptr_t>(out8) & 7) == 0);
  const int64_t octs = aligned ? (count >> 3) : 0;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < octs; q += (int64_t)gridDim.x * blockDim.x) {
    const uint4 v = reinterpret_cast<const uint4*>(counts)[q];
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t lo = 0, hi = 0;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const uint32_t c = (w[e >> 1] >> (16 * (e & 1))) & 0xFFFFu;
      const uint32_t b = c < 255u ? c : 255u;
      if (e < 4) lo |= b << (8 * e);
      else hi |= b << (8 * (e - 4));
      if (c >= 255u) {
        const unsigned long long pos = atomicAdd(esc_n, 1ull);
        if ((int64_t)pos < esc_capacity) {
          esc_index[pos] = first + (q << 3) + e;
          esc_count[pos] = (uint16_t)c;
        }
      }
    }
    reinterpret_cast<uint2*>(out8)[q] = make_uint2(lo, hi);
  }
  for (int64_t k = (octs << 3) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < count; k += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t c = counts[k];
    out8[k] = (uint8_t)(c < 255u ? c : 255u);
    if (c >= 255u) {
      const unsigned long long pos = atomicAdd(esc_n, 1ull);
      if ((int64_t)pos < esc_capacity) {
        esc_index[pos] = first + k;
        esc_count[pos] = (uint16_t)c;
      }
    }
  }
}

}  // namespace

int launch_nw_expand_block(const TriSlabs& s, int64_t n, int64_t col_begin, int64_t col_end, double* d_out_block, cudaStream_t st) {
  return launch_expand(s, NwValue{n}, n, col_begin, col_end, d_out_block, st);
}
int launch_mh_expand_block(const TriSlabs& s, int64_t n, int64_t col_begin, int64_t col_end, const double* d_table, double diag,
                           double* d_out_block, cudaStream_t st) {
  return launch_expand(s, MhValue{n, d_table, diag}, n, col_begin, col_end, d_out_block, st);
}
int launch_checksum_u32(const uint32_t* d_v, int64_t count, int64_t first_index, unsigned long long* d_sum, cudaStream_t st) {
  return launch_checksum<uint32_t>(d_v, count, first_index, d_sum, st);
}
int launch_checksum_u16(const uint16_t* d_v, int64_t count, int64_t first_index, unsigned long long* d_sum, cudaStream_t st) {
  return launch_checksum<uint16_t>(d_v, count, first_index, d_sum, st);
}
int launch_nw_pack8(const uint32_t* d_m, const uint32_t* d_l, int64_t count, uint8_t* d_m8, uint8_t* d_l8, cudaStream_t st) {
  if (count <= 0) return DYNA_OK;
  const unsigned grid = (unsigned)std::min<int64_t>((count / 4 + 255) / 256 + 1, (int64_t)kNumSMsB200 * 32);
  nw_pack8_kernel<<<grid, 256, 0, st>>>(d_m, d_l, count, d_m8, d_l8);
  DYNA_CUDA(cudaGetLastError());
  return DYNA_OK;
}
int launch_mh_narrow8(const uint16_t* d_counts, int64_t count, int64_t first_index, uint8_t* d_out8, int64_t esc_capacity,
                      long long* d_esc_index, uint16_t* d_esc_count, unsigned long long* d_esc_n, cudaStream_t st) {
  if (count <= 0) return DYNA_OK;
  const unsigned grid = (unsigned)std::min<int64_t>((count / 8 + 255) / 256 + 1, (int64_t)kNumSMsB200 * 32);
  mh_narrow8_kernel<<<grid, 256, 0, st>>>(d_counts, count, first_index, d_out8, esc_capacity, d_esc_index, d_esc_count, d_esc_n);
  DYNA_CUDA(cudaGetLastError());
  return DYNA_OK;
}

}  // namespace dyna
