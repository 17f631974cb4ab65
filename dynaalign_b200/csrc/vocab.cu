// Device front end of the pure-R MinHash pipeline: create_vocab (R/minHash.R:38-41) and the
// vocabulary ranks that stand in for create_char_matrix (:60-66) -- without the dense V x N matrix.
//   1. every k-shingle (k <= 8 bytes) is packed big-endian into a 64-bit key, so integer order == byte-wise string order
//      (R's sort() collation for the upper-case residue alphabets)
//   2. sort + unique (CUB radix sort / select: library helpers around the hand-written pack and rank kernels)
//   3. rank of every shingle = 1 + position in the vocabulary, by binary search
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>

#include <algorithm>
#include <vector>

#include "common.cuh"

namespace dyna {
namespace {

__global__ void vocab_pack_kernel(const uint8_t* __restrict__ res, const int64_t* __restrict__ off,
                                  const int64_t* __restrict__ soff, int64_t n, int k, unsigned long long* __restrict__ keys) {
  // one block per document (grid-stride); shingle p of document d starts at byte off[d] + p
  for (int64_t d = blockIdx.x; d < n; d += gridDim.x) {
    const uint8_t* s = res + off[d];
    const int64_t cnt = soff[d + 1] - soff[d];
    unsigned long long* out = keys + soff[d];
    for (int64_t p = threadIdx.x; p < cnt; p += blockDim.x) {
      unsigned long long key = 0;
      for (int b = 0; b < k; ++b) key = (key << 8) | s[p + b];
      out[p] = key;
    }
  }
}

__global__ void vocab_rank_kernel(const unsigned long long* __restrict__ keys, int64_t total,
                                  const unsigned long long* __restrict__ vocab, int64_t V, int32_t* __restrict__ ranks) {
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (int64_t)gridDim.x * blockDim.x) {
    const unsigned long long key = keys[g];
    int64_t lo = 0, hi = V - 1;
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if (vocab[mid] < key) lo = mid + 1;
      else hi = mid;
    }
    ranks[g] = (int32_t)(lo + 1);  // 1-based, as R's row index into the characteristic matrix
  }
}

}  // namespace
}  // namespace dyna

using namespace dyna;

extern "C" int dyna_minhash_vocab_ranks(const uint8_t* residues, const int64_t* offsets, int64_t n, int k,
                                        uint64_t* vocab_keys_out, int64_t vocab_capacity, int64_t* vocab_size_out,
                                        int32_t* ranks_out, int64_t* rank_offsets_out) {
  if (n <= 0) return fail(DYNA_ERR_INVALID, "Input sequences vector cannot be empty");
  DYNA_TRY(check_offsets(offsets, n, "dyna_minhash_vocab_ranks"));
  // shingle()'s argument check, raised for the first offending sequence as lapply would (R/minHash.R:15-16)
  for (int64_t d = 0; d < n; ++d) {
    const int64_t L = offsets[d + 1] - offsets[d];
    if (k < 1 || k > L) return fail(DYNA_ERR_INVALID, "'k' must be a positive integer between 1 and %lld", (long long)L);
  }
  if (k > 8) return fail(DYNA_ERR_UNSUPPORTED, "k > 8 is not supported by the device vocabulary (64-bit keys)");
  std::vector<int64_t> soff((size_t)n + 1, 0);
  for (int64_t d = 0; d < n; ++d) soff[(size_t)d + 1] = soff[(size_t)d] + (offsets[d + 1] - offsets[d] - k + 1);
  const int64_t total = soff[(size_t)n];
  if (total >= (1ll << 31)) return fail(DYNA_ERR_UNSUPPORTED, "more than 2^31 shingles in one call is not supported");
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0)
    return fail(DYNA_ERR_CUDA, "DynaAlign CUDA: no usable CUDA device; there is no CPU fallback");

  DevBuf<uint8_t> d_res;
  DevBuf<int64_t> d_off, d_soff;
  DevBuf<unsigned long long> d_keys, d_sorted, d_vocab;
  DevBuf<int32_t> d_ranks;
  DevBuf<long long> d_nsel;
  DYNA_TRY(d_res.alloc((size_t)offsets[n] + 8));
  DYNA_TRY(d_off.alloc((size_t)n + 1));
  DYNA_TRY(d_soff.alloc((size_t)n + 1));
  DYNA_TRY(d_keys.alloc((size_t)total));
  DYNA_TRY(d_sorted.alloc((size_t)total));
  DYNA_TRY(d_vocab.alloc((size_t)total));
  DYNA_TRY(d_ranks.alloc((size_t)total));
  DYNA_TRY(d_nsel.alloc(1));
  DYNA_CUDA(cudaMemcpy(d_res.p, residues, (size_t)offsets[n], cudaMemcpyHostToDevice));
  DYNA_CUDA(cudaMemcpy(d_off.p, offsets, sizeof(int64_t) * (size_t)(n + 1), cudaMemcpyHostToDevice));
  DYNA_CUDA(cudaMemcpy(d_soff.p, soff.data(), sizeof(int64_t) * (size_t)(n + 1), cudaMemcpyHostToDevice));

  vocab_pack_kernel<<<(int)std::min<int64_t>(n, (int64_t)kNumSMsB200 * 16), 128>>>(d_res.p, d_off.p, d_soff.p, n, k, d_keys.p);
  DYNA_CUDA(cudaGetLastError());

  size_t tb1 = 0, tb2 = 0;
  cub::DeviceRadixSort::SortKeys(nullptr, tb1, d_keys.p, d_sorted.p, (int)total, 0, 8 * k);
  cub::DeviceSelect::Unique(nullptr, tb2, d_sorted.p, d_vocab.p, d_nsel.p, (int)total);
  DevBuf<uint8_t> d_tmp;
  DYNA_TRY(d_tmp.alloc(std::max(tb1, tb2) + 16));
  DYNA_CUDA(cub::DeviceRadixSort::SortKeys(d_tmp.p, tb1, d_keys.p, d_sorted.p, (int)total, 0, 8 * k));
  DYNA_CUDA(cub::DeviceSelect::Unique(d_tmp.p, tb2, d_sorted.p, d_vocab.p, d_nsel.p, (int)total));
  long long V = 0;
  DYNA_CUDA(cudaMemcpy(&V, d_nsel.p, sizeof V, cudaMemcpyDeviceToHost));
  if (vocab_size_out) *vocab_size_out = V;
  vocab_rank_kernel<<<(int)std::min<int64_t>((total + 255) / 256, (int64_t)kNumSMsB200 * 16), 256>>>(d_keys.p, total, d_vocab.p, V,
                                                                                               d_ranks.p);
  DYNA_CUDA(cudaGetLastError());
  if (rank_offsets_out) memcpy(rank_offsets_out, soff.data(), sizeof(int64_t) * (size_t)(n + 1));
  if (ranks_out) DYNA_CUDA(cudaMemcpy(ranks_out, d_ranks.p, sizeof(int32_t) * (size_t)total, cudaMemcpyDeviceToHost));
  if (vocab_keys_out) {
    if (V > vocab_capacity)
      return fail(DYNA_ERR_INVALID, "vocabulary buffer too small: %lld entries, capacity %lld", V, (long long)vocab_capacity);
    DYNA_CUDA(cudaMemcpy(vocab_keys_out, d_vocab.p, sizeof(uint64_t) * (size_t)V, cudaMemcpyDeviceToHost));
  }
  DYNA_CUDA(cudaDeviceSynchronize());
  return DYNA_OK;
}
