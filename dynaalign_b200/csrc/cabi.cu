// extern "C" entry points of libdynaalign_b200 (see include/dynaalign_b200.h) and the host-side driver logic that
// replaces the reference's loops: input validation with the reference's error strings, HashFamily seed stream,
// residue encoding, work partitioning (row blocks of the upper triangle), device plans, result expansion.
// There is no CPU compute path in this file: every similarity value is produced by the kernels in
// mh_kernels.cu / nw_kernels.cu, and a missing or failing CUDA device is an error.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <map>
#include <memory>
#include <mutex>
#include <random>
#include <thread>
#include <vector>

#include "blosum_tables.h"
#include "common.cuh"
#include "gather.cuh"
#include "mh_kernels.cuh"
#include "mh_sparse.cuh"
#include "nw_kernels.cuh"
#include "nw_post.cuh"

using namespace dyna;

namespace {
struct PhaseTimer {  // DYNA_TIMING=1: per-phase wall clock of the host entry points on stderr
  bool on = getenv("DYNA_TIMING") != nullptr;
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  void lap(const char* what) {
    if (!on) return;
    auto t1 = std::chrono::steady_clock::now();
    fprintf(stderr, "[dyna timing] %-28s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
    t0 = t1;
  }
};


thread_local int g_device = 0;

int use_device(int device) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count <= 0)
    return fail(DYNA_ERR_CUDA, "DynaAlign CUDA: no usable CUDA device (%s); there is no CPU fallback",
                e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
  if (device < 0 || device >= count) return fail(DYNA_ERR_CUDA, "DynaAlign CUDA: device %d out of range (0..%d)", device, count - 1);
  DYNA_CUDA(cudaSetDevice(device));
  dev_pool_keep_cached(device);
  return DYNA_OK;
}

int resolve_gpus(int n_gpus, int* out) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count <= 0)
    return fail(DYNA_ERR_CUDA, "DynaAlign CUDA: no usable CUDA device (%s); there is no CPU fallback",
                e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
  *out = (n_gpus <= 0 || n_gpus > count) ? count : n_gpus;
  return DYNA_OK;
}

// One process driving several GPUs: every device must be able to read the result slabs of the others (the expansion
// kernels gather them with peer loads over NVLink).  Plain cudaDeviceEnablePeerAccess covers cudaMalloc memory; the
// stream-ordered pools the buffers come from need cudaMemPoolSetAccess as well.  Returns false if any pair of the
// first `gpus` devices cannot reach each other.
bool enable_peer_mesh(int gpus) {
  static std::mutex mu;
  static int done_for = 0;
  static bool ok_cached = true;
  std::lock_guard<std::mutex> lock(mu);
  if (gpus <= done_for) return ok_cached;
  bool ok = true;
  int prev = 0;
  cudaGetDevice(&prev);
  for (int a = 0; a < gpus && ok; ++a) {
    if (cudaSetDevice(a) != cudaSuccess) { ok = false; break; }
    dev_pool_keep_cached(a);
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, a) != cudaSuccess) { ok = false; break; }
    std::vector<cudaMemAccessDesc> descs;
    for (int b = 0; b < gpus; ++b) {
      if (b == a) continue;
      int can = 0;
      if (cudaDeviceCanAccessPeer(&can, a, b) != cudaSuccess || !can) { ok = false; break; }
      const cudaError_t e = cudaDeviceEnablePeerAccess(b, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { ok = false; break; }
      cudaGetLastError();
      cudaMemAccessDesc d{};
      d.location.type = cudaMemLocationTypeDevice;
      d.location.id = b;  // device b may read and write allocations of device a's pool
      d.flags = cudaMemAccessFlagsProtReadWrite;
      descs.push_back(d);
    }
    if (ok && !descs.empty() && cudaMemPoolSetAccess(pool, descs.data(), descs.size()) != cudaSuccess) ok = false;
  }
  cudaGetLastError();
  cudaSetDevice(prev);
  done_for = gpus;
  ok_cached = ok;
  return ok;
}

// columns [bounds[g], bounds[g+1]) of the n x n result are expanded and copied to the host by device g
void column_blocks(int64_t n, int gpus, std::vector<int64_t>& bounds) {
  bounds.resize((size_t)gpus + 1);
  for (int g = 0; g <= gpus; ++g) bounds[(size_t)g] = n * g / gpus;
}

// Runs fn(g) on one host thread per device (inline when there is one), collects the first error.
template <class Fn>
int for_each_gpu(int gpus, Fn fn) {
  std::vector<int> rcs((size_t)gpus, DYNA_OK);
  std::vector<std::string> errs((size_t)gpus);
  if (gpus == 1) {
    rcs[0] = fn(0);
    if (rcs[0] != DYNA_OK) errs[0] = err_slot();
  } else {
    std::vector<std::thread> th;
    for (int g = 0; g < gpus; ++g)
      th.emplace_back([&, g]() {
        rcs[(size_t)g] = fn(g);
        if (rcs[(size_t)g] != DYNA_OK) errs[(size_t)g] = err_slot();  // err_slot() is thread-local
      });
    for (auto& t : th) t.join();
  }
  for (int g = 0; g < gpus; ++g)
    if (rcs[(size_t)g] != DYNA_OK) {
      err_slot() = errs[(size_t)g];
      err_code_slot() = rcs[(size_t)g];
      return rcs[(size_t)g];
    }
  return DYNA_OK;
}

// ---- substitution tables: expand the packed lower triangles once
struct Tables {
  int8_t full[kNumTables][576];
  int8_t aa[256];
  Tables() {
    for (int t = 0; t < kNumTables; ++t)
      for (int i = 0; i < 24; ++i)
        for (int j = 0; j < 24; ++j) {
          const int hi = std::max(i, j), lo = std::min(i, j);
          full[t][i * 24 + j] = kTableTri[t][hi * (hi + 1) / 2 + lo];
        }
    for (int c = 0; c < 256; ++c) aa[c] = -1;
    for (int i = 0; i < 24; ++i) aa[(unsigned char)kAlphabet[i]] = (int8_t)i;
  }
};
const Tables& tables() {
  static const Tables t;
  return t;
}
int find_table(const char* name) {
  for (int t = 0; t < kNumTables; ++t)
    if (strcmp(name, kTableNames[t]) == 0) return t;
  return -1;
}

// first residue error in the order the reference discovers it (src/pairwiseSeqAlign.cpp:239-250 inside the
// i-major / j>=i pair loop :340-346): rows that are empty validate nothing; the first non-empty sequence i0 is
// checked as sequence1 at its first residue, then as sequence2 (pair (i0,i0)); every later sequence is checked as
// sequence2 against i0.  Returns DYNA_OK or the reference's message.
int validate_residues(const uint8_t* res, const int64_t* off, int64_t n) {
  const int8_t* aa = tables().aa;
  int64_t i0 = 0;
  while (i0 < n && off[i0 + 1] == off[i0]) ++i0;
  if (i0 >= n) return DYNA_OK;
  const uint8_t* s = res + off[i0];
  const int64_t L = off[i0 + 1] - off[i0];
  if (aa[s[0]] < 0) return fail(DYNA_ERR_INVALID, "Invalid amino acid in sequence1: %c", (char)s[0]);
  for (int64_t p = 0; p < L; ++p)
    if (aa[s[p]] < 0) return fail(DYNA_ERR_INVALID, "Invalid amino acid in sequence2: %c", (char)s[p]);
  for (int64_t j = i0 + 1; j < n; ++j)
    for (int64_t p = off[j]; p < off[j + 1]; ++p)
      if (aa[res[p]] < 0) return fail(DYNA_ERR_INVALID, "Invalid amino acid in sequence2: %c", (char)res[p]);
  return DYNA_OK;
}

// std::mt19937 stream exactly as HashFamily draws it (src/minHash.cpp:73-80)
void hashfamily_seeds(uint32_t seed, int n_hash, uint32_t* out) {
  std::mt19937 gen(seed);
  std::uniform_int_distribution<uint32_t> dis;
  for (int i = 0; i < n_hash; ++i) out[i] = dis(gen);
}

int check_mh_args(int64_t n, int k, int n_hash) {
  // reference messages and order: src/minHash.cpp:121-131
  if (n == 0) return fail(DYNA_ERR_INVALID, "Input sequences vector cannot be empty");
  if (k <= 0) return fail(DYNA_ERR_INVALID, "'k' must be a positive integer");
  if (n_hash <= 0) return fail(DYNA_ERR_INVALID, "Number of hash functions must be positive");
  if (n_hash > 65535) return fail(DYNA_ERR_UNSUPPORTED, "n_hash > 65535 is not supported (match counts are 16-bit)");
  if (k > 4096) return fail(DYNA_ERR_UNSUPPORTED, "k > 4096 is not supported");
  return DYNA_OK;
}

}  // namespace

// =====================================================================================================
// MinHash plan
// =====================================================================================================
struct dyna_mh_plan {
  int device = 0;
  int64_t n = 0;
  int n_hash = 0, hrows = 0;
  int64_t npitch = 0, row_begin = 0, row_end = 0, pairs = 0;
  int k = 0;
  int64_t max_len = 0;
  int launches = 0;
  DevBuf<uint8_t> res;
  DevBuf<int64_t> off;
  DevBuf<uint32_t> seeds, sig, sigT;
  DevBuf<uint16_t> counts;
  // 16-bit relabelled copy for the HSET2 match path (large inputs only)
  bool use16 = false;
  DevBuf<uint32_t> keys_out, vals_in, vals_out, sigP;
  DevBuf<int> seg_begin, seg_end, overflow;
  DevBuf<uint8_t> cub_temp;
  MhRelabelWork work;
  bool have_sequences = false, have_sig = false, have_sigT = false;
  // The dense u16 triangle is allocated when a dense producer first needs it (a plan that only ever uses the sparse
  // join below can describe inputs whose n(n-1)/2 counts would not fit any memory).
  bool counts_valid = false;
  // keys_out / vals_out hold EVERY hash row sorted by value (the relabelling ran over all code rows on this device)
  bool sorted_rows = false;
  // result of the sparse join (mh_sparse.cu): (pair key i*n+j, match count) for every pair of the row range with count >= 1
  bool sparse_valid = false;
  int64_t sp_runs = 0, sp_incidences = 0;
  DevBuf<unsigned long long> sp_keys;
  DevBuf<uint32_t> sp_counts;
  // Buffers are allocated and released through the stream-ordered allocator on the legacy default stream, while the
  // work runs on whatever stream the caller passes.  Every entry point records that stream here and nothing is
  // released (plan destruction, re-upload) before it has drained, so a released block can never still be in use --
  // also when the caller's stream is a non-blocking one that the legacy stream does not wait for.
  cudaStream_t last_stream = nullptr;
  ~dyna_mh_plan() { cudaStreamSynchronize(last_stream); }
};

namespace {
// threshold below which the relabelling overhead is not worth it
constexpr int64_t kMhPack16MinN = 2048;

int mh_plan_setup16(dyna_mh_plan* p) {
  const int64_t items = p->npitch * (int64_t)p->hrows;
  bool want = p->n >= kMhPack16MinN && items < (1ll << 31) && p->row_end > p->row_begin;
  if (const char* e = getenv("DYNA_MH_PACK16")) want = (atoi(e) != 0) && items < (1ll << 31) && p->row_end > p->row_begin;
  p->use16 = want;
  if (!want) return DYNA_OK;
  const int hrows2 = mh_hrows2(p->n_hash);
  DYNA_TRY(p->keys_out.alloc((size_t)items));
  DYNA_TRY(p->vals_in.alloc((size_t)items));
  DYNA_TRY(p->vals_out.alloc((size_t)items));
  DYNA_TRY(p->sigP.alloc((size_t)hrows2 * p->npitch));
  DYNA_TRY(p->seg_begin.alloc((size_t)p->hrows));
  DYNA_TRY(p->seg_end.alloc((size_t)p->hrows));
  DYNA_TRY(p->overflow.alloc(1));
  const size_t tb = mh_relabel_temp_bytes(p->npitch, p->hrows);
  DYNA_TRY(p->cub_temp.alloc(tb + 16));
  std::vector<int> b((size_t)p->hrows), e((size_t)p->hrows);
  for (int h = 0; h < p->hrows; ++h) {
    b[(size_t)h] = (int)(h * p->npitch);
    e[(size_t)h] = (int)(h * p->npitch + p->n);
  }
  DYNA_CUDA(cudaMemcpy(p->seg_begin.p, b.data(), sizeof(int) * b.size(), cudaMemcpyHostToDevice));
  DYNA_CUDA(cudaMemcpy(p->seg_end.p, e.data(), sizeof(int) * e.size(), cudaMemcpyHostToDevice));
  DYNA_TRY(launch_mh_iota(p->vals_in.p, p->npitch, p->hrows, nullptr));
  DYNA_CUDA(cudaDeviceSynchronize());
  p->work.temp = p->cub_temp.p;
  p->work.temp_bytes = tb;
  p->work.keys_out = p->keys_out.p;
  p->work.vals_in = p->vals_in.p;
  p->work.vals_out = p->vals_out.p;
  p->work.seg_begin = p->seg_begin.p;
  p->work.seg_end = p->seg_end.p;
  p->work.sigP = p->sigP.p;
  p->work.overflow = p->overflow.p;
  return DYNA_OK;
}

// transpose (+ relabel when the 16-bit path is on): everything the match kernel needs, from sig[n][n_hash]
int mh_plan_prepare_match_inputs(dyna_mh_plan* p, cudaStream_t st, int* launches, int code_row_begin = 0,
                                 int code_row_end = -1) {
  DYNA_TRY(launch_mh_transpose(p->sig.p, p->n, p->n_hash, p->sigT.p, p->npitch, p->hrows, st));
  int l = 1;
  if (p->use16) {
    int lr = 0;
    if (code_row_end < 0) code_row_end = mh_hrows2(p->n_hash);
    DYNA_TRY(launch_mh_relabel(p->sigT.p, p->n, p->n_hash, p->npitch, p->hrows, p->work, code_row_begin, code_row_end, st, &lr));
    l += lr;
    p->sorted_rows = code_row_begin == 0 && code_row_end == mh_hrows2(p->n_hash);
  }
  p->counts_valid = p->sparse_valid = false;  // new signatures: earlier match results are stale
  if (launches) *launches = l;
  return DYNA_OK;
}

int mh_plan_ensure_counts(dyna_mh_plan* p) {
  if (p->counts.p) return DYNA_OK;
  DYNA_TRY(check_device_fits(2.0 * (double)p->pairs, p->device, "the dense u16 match-count triangle"));
  DYNA_TRY(p->counts.alloc((size_t)p->pairs));
  DYNA_CUDA(cudaStreamSynchronize(0));  // stream-ordered on the legacy stream: visible to the caller's stream from here on
  return DYNA_OK;
}

// the dense triangle for a consumer: already there, or scattered from the sparse join's runs
int mh_plan_need_dense(dyna_mh_plan* p, cudaStream_t st, const char* who) {
  if (p->counts_valid) return DYNA_OK;
  if (!p->sparse_valid) return fail(DYNA_ERR_INVALID, "%s: no match counts on the device (run a match first)", who);
  DYNA_TRY(mh_plan_ensure_counts(p));
  DYNA_TRY(mh_sparse_densify(p->sp_keys.p, p->sp_counts.p, p->sp_runs, p->n, tri_strict_rows(p->n, p->row_begin), p->pairs,
                             p->counts.p, st));
  p->counts_valid = true;
  return DYNA_OK;
}
}  // namespace

extern "C" dyna_mh_plan* dyna_mh_plan_create(int64_t n, int n_hash, int64_t row_begin, int64_t row_end, int device) {
  if (n <= 0 || n_hash <= 0 || n_hash > 65535 || row_begin < 0 || row_end > n || row_begin > row_end) {
    fail(DYNA_ERR_INVALID, "dyna_mh_plan_create: bad arguments");
    return nullptr;
  }
  if (use_device(device) != DYNA_OK) return nullptr;
  std::unique_ptr<dyna_mh_plan> p(new dyna_mh_plan);
  p->device = device;
  p->n = n;
  p->n_hash = n_hash;
  p->hrows = mh_hrows(n_hash);
  p->npitch = mh_npitch(n);
  p->row_begin = row_begin;
  p->row_end = row_end;
  p->pairs = tri_strict_rows(n, row_end) - tri_strict_rows(n, row_begin);
  if (p->sig.alloc((size_t)n * n_hash) || p->sigT.alloc((size_t)p->hrows * p->npitch)) return nullptr;
  if (mh_plan_setup16(p.get()) != DYNA_OK) return nullptr;
  // allocations are stream-ordered on the legacy default stream: make them visible to any stream the caller uses
  if (cudaStreamSynchronize(0) != cudaSuccess) {
    fail(DYNA_ERR_CUDA, "DynaAlign CUDA: plan allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
    return nullptr;
  }
  return p.release();
}

extern "C" int dyna_mh_plan_upload_sequences(dyna_mh_plan* p, const uint8_t* residues, const int64_t* offsets, int k,
                                             const uint32_t* seeds, void* stream) {
  if (!p) return fail(DYNA_ERR_INVALID, "null plan");
  DYNA_TRY(use_device(p->device));
  if (k <= 0) return fail(DYNA_ERR_INVALID, "'k' must be a positive integer");
  if (k > 4096) return fail(DYNA_ERR_UNSUPPORTED, "k > 4096 is not supported");
  DYNA_TRY(check_offsets(offsets, p->n, "dyna_mh_plan_upload_sequences"));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaStream_t prev = p->last_stream;
  p->last_stream = st;
  WorkStreamGuard work_stream_guard(st);
  const int64_t total = offsets[p->n];
  p->max_len = 0;
  for (int64_t i = 0; i < p->n; ++i) p->max_len = std::max(p->max_len, offsets[i + 1] - offsets[i]);
  if (p->have_sequences) DYNA_CUDA(cudaStreamSynchronize(prev));  // the old sequence buffers may still be in use
  DYNA_TRY(p->res.alloc((size_t)total + 8));
  DYNA_TRY(p->off.alloc((size_t)p->n + 1));
  DYNA_TRY(p->seeds.alloc((size_t)p->n_hash));
  DYNA_CUDA(cudaMemcpyAsync(p->res.p, residues, (size_t)total, cudaMemcpyHostToDevice, st));
  DYNA_CUDA(cudaMemcpyAsync(p->off.p, offsets, sizeof(int64_t) * (size_t)(p->n + 1), cudaMemcpyHostToDevice, st));
  DYNA_CUDA(cudaMemcpyAsync(p->seeds.p, seeds, sizeof(uint32_t) * (size_t)p->n_hash, cudaMemcpyHostToDevice, st));
  p->k = k;
  p->have_sequences = true;
  p->have_sig = p->have_sigT = false;
  return DYNA_OK;
}

extern "C" int dyna_mh_plan_upload_signatures(dyna_mh_plan* p, const uint32_t* sig, void* stream) {
  if (!p) return fail(DYNA_ERR_INVALID, "null plan");
  DYNA_TRY(use_device(p->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  p->last_stream = st;
  WorkStreamGuard work_stream_guard(st);
  DYNA_CUDA(cudaMemcpyAsync(p->sig.p, sig, sizeof(uint32_t) * (size_t)p->n * p->n_hash, cudaMemcpyHostToDevice, st));
  int l = 0;
  DYNA_TRY(mh_plan_prepare_match_inputs(p, st, &l));
  p->have_sig = p->have_sigT = true;
  p->launches = l;
  return DYNA_OK;
}

extern "C" int dyna_mh_plan_run_signatures(dyna_mh_plan* p, void* stream) {
  if (!p) return fail(DYNA_ERR_INVALID, "null plan");
  if (!p->have_sequences) return fail(DYNA_ERR_INVALID, "dyna_mh_plan_run_signatures: no sequences uploaded");
  DYNA_TRY(use_device(p->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  p->last_stream = st;
  WorkStreamGuard work_stream_guard(st);
  DYNA_TRY(launch_mh_signature_murmur3(p->res.p, p->off.p, p->n, p->max_len, p->k, p->seeds.p, p->n_hash, p->sig.p, st));
  int l = 0;
  DYNA_TRY(mh_plan_prepare_match_inputs(p, st, &l));
  p->have_sig = p->have_sigT = true;
  p->launches = 1 + l;
  return DYNA_OK;
}

// Multi-rank form (all-gather variant): signatures and the layout transform for every hash row, the
// 16-bit relabelling only for this rank's share of the packed code rows.
extern "C" int dyna_mh_plan_run_signatures_shard(dyna_mh_plan* p, int code_row_begin, int code_row_end, void* stream) {
  if (!p) return fail(DYNA_ERR_INVALID, "null plan");
  if (!p->have_sequences) return fail(DYNA_ERR_INVALID, "dyna_mh_plan_run_signatures_shard: no sequences uploaded");
  if (!p->use16) return fail(DYNA_ERR_INVALID, "dyna_mh_plan_run_signatures_shard: the plan has no code table (dyna_mh_plan_code_rows() == 0)");
  if (code_row_begin < 0 || code_row_end < code_row_begin || code_row_end > mh_hrows2(p->n_hash))
    return fail(DYNA_ERR_INVALID, "dyna_mh_plan_run_signatures_shard: bad code row range");
  DYNA_TRY(use_device(p->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  p->last_stream = st;
  WorkStreamGuard work_stream_guard(st);
  DYNA_TRY(launch_mh_signature_murmur3(p->res.p, p->off.p, p->n, p->max_len, p->k, p->seeds.p, p->n_hash, p->sig.p, st));
  int l = 0;
  DYNA_TRY(mh_plan_prepare_match_inputs(p, st, &l, code_row_begin, code_row_end));
  p->have_sig = p->have_sigT = true;
  p->launches = 1 + l;
  return DYNA_OK;
}
extern "C" int dyna_mh_plan_code_rows(const dyna_mh_plan* p) { return (p && p->use16) ? mh_hrows2(p->n_hash) : 0; }
extern "C" int64_t dyna_mh_plan_code_row_bytes(const dyna_mh_plan* p) { return p ? (int64_t)sizeof(uint32_t) * p->npitch : 0; }
extern "C" void* dyna_mh_plan_codes_device_ptr(dyna_mh_plan* p) { return (p && p->use16) ? p->sigP.p : nullptr; }
extern "C" void* dyna_mh_plan_overflow_device_ptr(dyna_mh_plan* p) { return (p && p->use16) ? p->overflow.p : nullptr; }

extern "C" int dyna_mh_plan_run_match(dyna_mh_plan* p, void* stream) {
  if (!p) return fail(DYNA_ERR_INVALID, "null plan");
  if (!p->have_sigT) return fail(DYNA_ERR_INVALID, "dyna_mh_plan_run_match: no signatures on the device");
  DYNA_TRY(use_device(p->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  p->last_stream = st;
  WorkStreamGuard work_stream_guard(st);
  int l = 0;
  DYNA_TRY(mh_plan_ensure_counts(p));
  DYNA_TRY(launch_mh_match(p->sigT.p, p->npitch, p->hrows, p->n_hash, p->n, p->row_begin, p->row_end, p->counts.p,
                           p->use16 ? p->sigP.p : nullptr, p->use16 ? p->overflow.p : nullptr, st, &l));
  p->launches = l;
  p->counts_valid = true;
  p->sparse_valid = false;
  return DYNA_OK;
}

// The same match counts by joining the sorted hash rows on equal signature values (mh_sparse.cu): work and memory
// proportional to the number of (pair, hash function) matches instead of n^2 * n_hash.  *n_incidences_out = that number;
// *done_out = 1 if the join ran (n_incidences <= max_incidences; 0 = the default cap from free device memory), 0 if the
// caller should run the all-pairs kernel instead (too many matches, or the plan has no sorted rows: n < 2048 or a
// sharded relabelling).  Afterwards count_histogram / threshold_edges / checksum read the join's result; fetch_counts
// and counts_device_ptr scatter it into the dense triangle first.
extern "C" int dyna_mh_plan_run_match_sparse(dyna_mh_plan* p, int64_t max_incidences, int64_t* n_incidences_out, int* done_out,
                                             void* stream) {
  if (!p || !done_out) return fail(DYNA_ERR_INVALID, "dyna_mh_plan_run_match_sparse: null argument");
  *done_out = 0;
  if (n_incidences_out) *n_incidences_out = -1;
  if (!p->have_sigT) return fail(DYNA_ERR_INVALID, "dyna_mh_plan_run_match_sparse: no signatures on the device");
  if (!p->use16 || !p->sorted_rows) return DYNA_OK;
  DYNA_TRY(use_device(p->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  p->last_stream = st;
  WorkStreamGuard work_stream_guard(st);
  p->launches = 0;
  if (p->row_end <= p->row_begin || p->n < 2) {
    p->sp_runs = p->sp_incidences = 0;
    p->sparse_valid = true;
    p->counts_valid = false;
    *done_out = 1;
    if (n_incidences_out) *n_incidences_out = 0;
    return DYNA_OK;
  }
  DevBuf<unsigned long long> row_elems, row_pairs, row_eoff, row_poff, totals;
  DYNA_TRY(row_elems.alloc((size_t)p->n_hash));
  DYNA_TRY(row_pairs.alloc((size_t)p->n_hash));
  DYNA_TRY(row_eoff.alloc((size_t)p->n_hash));
  DYNA_TRY(row_poff.alloc((size_t)p->n_hash));
  DYNA_TRY(totals.alloc(2));
  DYNA_CUDA(cudaStreamSynchronize(0));
  DYNA_TRY(mh_sparse_count_incidences(p->keys_out.p, p->n, p->n_hash, p->npitch, row_elems.p, row_pairs.p, row_eoff.p, row_poff.p,
                                      totals.p, st));
  unsigned long long tot[2] = {0, 0};
  DYNA_CUDA(cudaMemcpyAsync(tot, totals.p, sizeof tot, cudaMemcpyDeviceToHost, st));
  DYNA_CUDA(cudaStreamSynchronize(st));
  p->launches = 3;
  const int64_t n_elems = (int64_t)tot[0], n_inc = (int64_t)tot[1];
  if (n_incidences_out) *n_incidences_out = n_inc;
  int64_t cap = max_incidences;
  if (cap <= 0) {  // ~40 bytes per incidence across the emit / sort / encode buffers
    size_t free_b = 0, total_b = 0;
    cap = (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) ? (int64_t)(free_b / 64) : (1ll << 27);
  }
  cap = std::min<int64_t>(cap, (1ll << 31) - 2);
  if (n_inc > cap) return DYNA_OK;  // dense data: the all-pairs kernel is the right tool
  if (n_inc == 0) {
    p->sp_runs = p->sp_incidences = 0;
    p->sparse_valid = true;
    p->counts_valid = false;
    *done_out = 1;
    return DYNA_OK;
  }
  DevBuf<uint32_t> el_pos, el_r;
  DevBuf<unsigned long long> el_off, pair_keys, sorted, d_runs;
  DevBuf<uint8_t> temp;
  DYNA_TRY(el_pos.alloc((size_t)n_elems));
  DYNA_TRY(el_r.alloc((size_t)n_elems));
  DYNA_TRY(el_off.alloc((size_t)n_elems));
  DYNA_TRY(pair_keys.alloc((size_t)n_inc));
  DYNA_TRY(sorted.alloc((size_t)n_inc));
  DYNA_TRY(d_runs.alloc(1));
  const size_t tb = mh_sparse_sort_temp_bytes(n_inc, p->n);
  DYNA_TRY(temp.alloc(tb));
  if (p->sparse_valid) DYNA_CUDA(cudaStreamSynchronize(st));
  p->sparse_valid = false;
  DYNA_TRY(p->sp_keys.alloc((size_t)n_inc));    // at most one run per incidence
  DYNA_TRY(p->sp_counts.alloc((size_t)n_inc));
  DYNA_CUDA(cudaStreamSynchronize(0));
  DYNA_TRY(mh_sparse_emit(p->keys_out.p, p->vals_out.p, p->n, p->n_hash, p->npitch, row_eoff.p, row_poff.p, el_pos.p, el_r.p,
                          el_off.p, n_elems, n_inc, p->row_begin, p->row_end, pair_keys.p, st));
  DYNA_TRY(mh_sparse_sort_encode(pair_keys.p, sorted.p, n_inc, p->n, temp.p, tb, p->sp_keys.p, p->sp_counts.p, d_runs.p, st));
  unsigned long long runs = 0, last_key = 0;
  DYNA_CUDA(cudaMemcpyAsync(&runs, d_runs.p, sizeof runs, cudaMemcpyDeviceToHost, st));
  DYNA_CUDA(cudaStreamSynchronize(st));
  if (runs > 0) {
    DYNA_CUDA(cudaMemcpyAsync(&last_key, p->sp_keys.p + (runs - 1), sizeof last_key, cudaMemcpyDeviceToHost, st));
    DYNA_CUDA(cudaStreamSynchronize(st));
    if (last_key == (unsigned long long)p->n * (unsigned long long)p->n) --runs;  // pairs of other ranks' rows
  }
  p->launches += 2 + 6;  // two of ours + the library's sort / encode passes
  p->sp_runs = (int64_t)runs;
  p->sp_incidences = n_inc;
  p->sparse_valid = true;
  p->counts_valid = false;
  *done_out = 1;
  return DYNA_OK;
}

extern "C" int dyna_mh_plan_fetch_signatures(dyna_mh_plan* p, uint32_t* sig_out, void* stream) {
  if (!p || !p->have_sig) return fail(DYNA_ERR_INVALID, "dyna_mh_plan_fetch_signatures: nothing to fetch");
  DYNA_TRY(use_device(p->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  p->last_stream = st;
  WorkStreamGuard work_stream_guard(st);
  DYNA_CUDA(cudaMemcpyAsync(sig_out, p->sig.p, sizeof(uint32_t) * (size_t)p->n * p->n_hash, cudaMemcpyDeviceToHost, st));
  DYNA_CUDA(cudaStreamSynchronize(st));
  return DYNA_OK;
}

extern "C" int dyna_mh_plan_fetch_counts(dyna_mh_plan* p, uint16_t* counts_out, void* stream) {
  if (!p) return fail(DYNA_ERR_INVALID, "null plan");
  DYNA_TRY(use_device(p->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  p->last_stream = st;
  WorkStreamGuard work_stream_guard(st);
  DYNA_TRY(mh_plan_need_dense(p, st, "dyna_mh_plan_fetch_counts"));
  if (p->pairs > 0)
    DYNA_CUDA(cudaMemcpyAsync(counts_out, p->counts.p, sizeof(uint16_t) * (size_t)p->pairs, cudaMemcpyDeviceToHost, st));
  DYNA_CUDA(cudaStreamSynchronize(st));
  return DYNA_OK;
}

// ---- sub-cluster plans: clusterbreak calls sim_fn again on every oversized cluster (R/clusterbreak.R:250-254); the
// signatures of a subset are a subset of the signatures, so the child plan gathers rows instead of re-hashing
extern "C" dyna_mh_plan* dyna_mh_plan_create_subset(dyna_mh_plan* parent, const int64_t* indices, int64_t m, int64_t row_begin,
                                                    int64_t row_end) {
  if (!parent || !parent->have_sig || !indices || m <= 0) {
    fail(DYNA_ERR_INVALID, "dyna_mh_plan_create_subset: parent has no signatures or bad arguments");
    return nullptr;
  }
  for (int64_t r = 0; r < m; ++r)
    if (indices[r] < 0 || indices[r] >= parent->n) {
      fail(DYNA_ERR_INVALID, "dyna_mh_plan_create_subset: index %lld out of range", (long long)indices[r]);
      return nullptr;
    }
  dyna_mh_plan* c = dyna_mh_plan_create(m, parent->n_hash, row_begin, row_end, parent->device);
  if (!c) return nullptr;
  DevBuf<int64_t> d_idx;
  int rc = d_idx.alloc((size_t)m);
  if (rc == DYNA_OK && cudaMemcpy(d_idx.p, indices, sizeof(int64_t) * (size_t)m, cudaMemcpyHostToDevice) != cudaSuccess)
    rc = fail(DYNA_ERR_CUDA, "DynaAlign CUDA: host-to-device copy failed");
  if (rc == DYNA_OK) rc = launch_mh_gather_rows(parent->sig.p, d_idx.p, m, parent->n_hash, c->sig.p, nullptr);
  int l = 0;
  if (rc == DYNA_OK) rc = mh_plan_prepare_match_inputs(c, nullptr, &l);
  if (rc == DYNA_OK && cudaDeviceSynchronize() != cudaSuccess) rc = fail(DYNA_ERR_CUDA, "DynaAlign CUDA: subset gather failed");
  if (rc != DYNA_OK) {
    dyna_mh_plan_destroy(c);
    return nullptr;
  }
  c->have_sig = c->have_sigT = true;
  c->launches = 1 + l;
  return c;
}

// ---- threshold + sparsify: the step right after the hot path in clusterbreak (R/clusterbreak.R:219-221)
extern "C" int dyna_mh_plan_count_histogram(dyna_mh_plan* p, uint64_t* hist_out, void* stream) {
  if (!p) return fail(DYNA_ERR_INVALID, "null plan");
  DYNA_TRY(use_device(p->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  p->last_stream = st;
  WorkStreamGuard work_stream_guard(st);
  DevBuf<unsigned long long> d_hist;
  DYNA_TRY(d_hist.alloc((size_t)p->n_hash + 1));
  const bool sparse = p->sparse_valid && !p->counts_valid;
  if (sparse) {
    DYNA_TRY(mh_sparse_histogram(p->sp_counts.p, p->sp_runs, p->n_hash, d_hist.p, st));
  } else {
    if (!p->counts_valid) return fail(DYNA_ERR_INVALID, "dyna_mh_plan_count_histogram: no match counts on the device (run a match first)");
    DYNA_TRY(launch_mh_count_hist(p->counts.p, p->pairs, p->n_hash, d_hist.p, st));
  }
  DYNA_CUDA(cudaMemcpyAsync(hist_out, d_hist.p, sizeof(uint64_t) * (size_t)(p->n_hash + 1), cudaMemcpyDeviceToHost, st));
  DYNA_CUDA(cudaStreamSynchronize(st));
  if (sparse) hist_out[0] = (uint64_t)(p->pairs - p->sp_runs);  // every pair without a run has no matching hash function
  return DYNA_OK;
}

extern "C" int dyna_quantile_type7_counts(const uint64_t* hist, int n_hash, double prob, double* threshold_out, int* min_count_out) {
  // R's quantile(x, prob, type = 7) (the default used at R/clusterbreak.R:219) for x = count / n_hash with the given
  // multiplicities:  index = 1 + (N-1)*prob; lo = floor(index); hi = ceiling(index); qs = x[lo];
  //                  if (index > lo && x[hi] != qs) qs = (1-h)*qs + h*x[hi], h = index - lo
  if (n_hash <= 0 || !(prob >= 0.0 && prob <= 1.0)) return fail(DYNA_ERR_INVALID, "'probs' outside [0,1]");
  long double total = 0;
  for (int c = 0; c <= n_hash; ++c) total += (long double)hist[c];
  if (total < 1) return fail(DYNA_ERR_INVALID, "quantile of an empty set of pairs");
  const double N = (double)total;
  const double index = 1.0 + std::max(N - 1.0, 0.0) * prob;
  const double lo = std::floor(index), hi = std::ceil(index);
  auto value_at_rank = [&](double r) {  // r-th smallest, 1-based
    uint64_t cum = 0;
    for (int c = 0; c <= n_hash; ++c) {
      cum += hist[c];
      if ((double)cum >= r) return static_cast<double>(c) / n_hash;
    }
    return 1.0;
  };
  double qs = value_at_rank(lo);
  const double xhi = value_at_rank(hi);
  if (index > lo && xhi != qs) {
    const double h = index - lo;
    qs = (1.0 - h) * qs + h * xhi;
  }
  if (threshold_out) *threshold_out = qs;
  if (min_count_out) {  // `sim[sim < threshold] <- 0` keeps count/n_hash >= threshold
    int c = 0;
    while (c <= n_hash && static_cast<double>(c) / n_hash < qs) ++c;
    *min_count_out = c;
  }
  return DYNA_OK;
}

extern "C" int dyna_mh_plan_threshold_edges(dyna_mh_plan* p, int min_count, int64_t max_edges, int32_t* i_out, int32_t* j_out,
                                            uint16_t* count_out, int64_t* n_edges_out, void* stream) {
  if (!p) return fail(DYNA_ERR_INVALID, "null plan");
  DYNA_TRY(use_device(p->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  p->last_stream = st;
  WorkStreamGuard work_stream_guard(st);
  const int64_t rows = p->row_end - p->row_begin;
  if (n_edges_out) *n_edges_out = 0;
  if (rows <= 0 || p->pairs <= 0) return DYNA_OK;
  const uint32_t mc = (uint32_t)std::max(min_count, 1);  // a zero count is never an edge (weight 0 in the adjacency matrix)
  if (p->sparse_valid && !p->counts_valid) {  // the join's runs are already in row-major pair order: select and unpack
    if (p->sp_runs == 0) return DYNA_OK;
    DevBuf<uint32_t> sel;
    DevBuf<unsigned long long> nsel;
    DevBuf<uint8_t> temp;
    const size_t tb = mh_sparse_select_temp_bytes(p->sp_runs);
    DYNA_TRY(sel.alloc((size_t)p->sp_runs));
    DYNA_TRY(nsel.alloc(1));
    DYNA_TRY(temp.alloc(tb));
    DYNA_CUDA(cudaStreamSynchronize(0));
    DYNA_TRY(mh_sparse_select(p->sp_counts.p, p->sp_runs, mc, temp.p, tb, sel.p, nsel.p, st));
    unsigned long long total = 0;
    DYNA_CUDA(cudaMemcpyAsync(&total, nsel.p, sizeof total, cudaMemcpyDeviceToHost, st));
    DYNA_CUDA(cudaStreamSynchronize(st));
    if (n_edges_out) *n_edges_out = (int64_t)total;
    if ((int64_t)total > max_edges)
      return fail(DYNA_ERR_INVALID, "edge buffer too small: %lld edges, capacity %lld", (long long)total, (long long)max_edges);
    if (total == 0) return DYNA_OK;
    DevBuf<int32_t> di, dj;
    DevBuf<uint16_t> dc;
    DYNA_TRY(di.alloc((size_t)total));
    DYNA_TRY(dj.alloc((size_t)total));
    DYNA_TRY(dc.alloc((size_t)total));
    DYNA_CUDA(cudaStreamSynchronize(0));
    DYNA_TRY(mh_sparse_gather_edges(p->sp_keys.p, p->sp_counts.p, sel.p, (int64_t)total, p->n, di.p, dj.p, dc.p, st));
    DYNA_CUDA(cudaMemcpyAsync(i_out, di.p, sizeof(int32_t) * (size_t)total, cudaMemcpyDeviceToHost, st));
    DYNA_CUDA(cudaMemcpyAsync(j_out, dj.p, sizeof(int32_t) * (size_t)total, cudaMemcpyDeviceToHost, st));
    DYNA_CUDA(cudaMemcpyAsync(count_out, dc.p, sizeof(uint16_t) * (size_t)total, cudaMemcpyDeviceToHost, st));
    DYNA_CUDA(cudaStreamSynchronize(st));
    return DYNA_OK;
  }
  if (!p->counts_valid) return fail(DYNA_ERR_INVALID, "dyna_mh_plan_threshold_edges: no match counts on the device (run a match first)");
  DevBuf<unsigned long long> rc, ro, tot;
  DYNA_TRY(rc.alloc((size_t)rows));
  DYNA_TRY(ro.alloc((size_t)rows));
  DYNA_TRY(tot.alloc(1));
  DYNA_TRY(launch_mh_edges_count(p->counts.p, p->n, p->row_begin, p->row_end, mc, rc.p, ro.p, tot.p, st));
  unsigned long long total = 0;
  DYNA_CUDA(cudaMemcpyAsync(&total, tot.p, sizeof total, cudaMemcpyDeviceToHost, st));
  DYNA_CUDA(cudaStreamSynchronize(st));
  if (n_edges_out) *n_edges_out = (int64_t)total;
  if ((int64_t)total > max_edges)
    return fail(DYNA_ERR_INVALID, "edge buffer too small: %lld edges, capacity %lld", (long long)total, (long long)max_edges);
  if (total == 0) return DYNA_OK;
  DevBuf<int32_t> di, dj;
  DevBuf<uint16_t> dc;
  DYNA_TRY(di.alloc((size_t)total));
  DYNA_TRY(dj.alloc((size_t)total));
  DYNA_TRY(dc.alloc((size_t)total));
  DYNA_TRY(launch_mh_edges_fill(p->counts.p, p->n, p->row_begin, p->row_end, mc, ro.p, di.p, dj.p, dc.p, st));
  DYNA_CUDA(cudaMemcpyAsync(i_out, di.p, sizeof(int32_t) * (size_t)total, cudaMemcpyDeviceToHost, st));
  DYNA_CUDA(cudaMemcpyAsync(j_out, dj.p, sizeof(int32_t) * (size_t)total, cudaMemcpyDeviceToHost, st));
  DYNA_CUDA(cudaMemcpyAsync(count_out, dc.p, sizeof(uint16_t) * (size_t)total, cudaMemcpyDeviceToHost, st));
  DYNA_CUDA(cudaStreamSynchronize(st));
  return DYNA_OK;
}

// match + fetch with the device-to-host copy of finished row chunks overlapped with the matching of the next ones:
// the counts slab of a row range is contiguous, so the plan's range is cut into pair-balanced chunks, each matched on
// `stream` and copied on a second stream as soon as its kernel has finished.  counts_out should be pinned memory.
extern "C" int dyna_mh_plan_run_match_fetch(dyna_mh_plan* p, uint16_t* counts_out, void* stream) {
  if (!p) return fail(DYNA_ERR_INVALID, "null plan");
  if (!p->have_sigT) return fail(DYNA_ERR_INVALID, "dyna_mh_plan_run_match_fetch: no signatures on the device");
  DYNA_TRY(use_device(p->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  p->last_stream = st;
  WorkStreamGuard work_stream_guard(st);
  const int64_t rows = p->row_end - p->row_begin;
  if (rows <= 0 || p->pairs <= 0) return DYNA_OK;
  DYNA_TRY(mh_plan_ensure_counts(p));
  p->counts_valid = true;  // every chunk below writes its part of the triangle
  p->sparse_valid = false;
  const int nchunks = (int)std::max<int64_t>(1, std::min<int64_t>(16, p->pairs / (64ll << 20)));
  // pair-balanced chunk boundaries inside [row_begin, row_end)
  std::vector<int64_t> b((size_t)nchunks + 1);
  b[0] = p->row_begin;
  const int64_t base = tri_strict_rows(p->n, p->row_begin);
  int64_t r = p->row_begin;
  for (int c = 1; c < nchunks; ++c) {
    const int64_t target = base + p->pairs * c / nchunks;
    while (r < p->row_end && tri_strict_rows(p->n, r) < target) ++r;
    b[(size_t)c] = r;
  }
  b[(size_t)nchunks] = p->row_end;
  cudaStream_t copy_st;
  DYNA_CUDA(cudaStreamCreateWithFlags(&copy_st, cudaStreamNonBlocking));
  std::vector<cudaEvent_t> ev((size_t)nchunks);
  int rc = DYNA_OK, launches = 0;
  for (int c = 0; c < nchunks && rc == DYNA_OK; ++c) {
    cudaEventCreateWithFlags(&ev[(size_t)c], cudaEventDisableTiming);
    const int64_t r0 = b[(size_t)c], r1 = b[(size_t)c + 1];
    if (r1 <= r0) continue;
    const int64_t off = tri_strict_rows(p->n, r0) - base, cnt = tri_strict_rows(p->n, r1) - tri_strict_rows(p->n, r0);
    int l = 0;
    rc = launch_mh_match(p->sigT.p, p->npitch, p->hrows, p->n_hash, p->n, r0, r1, p->counts.p + off,
                         p->use16 ? p->sigP.p : nullptr, p->use16 ? p->overflow.p : nullptr, st, &l);
    launches += l;
    if (rc != DYNA_OK) break;
    cudaEventRecord(ev[(size_t)c], st);
    cudaStreamWaitEvent(copy_st, ev[(size_t)c], 0);
    if (cnt > 0 && cudaMemcpyAsync(counts_out + off, p->counts.p + off, sizeof(uint16_t) * (size_t)cnt, cudaMemcpyDeviceToHost,
                                   copy_st) != cudaSuccess)
      rc = fail(DYNA_ERR_CUDA, "DynaAlign CUDA: device-to-host copy failed: %s", cudaGetErrorString(cudaGetLastError()));
  }
  if (cudaStreamSynchronize(copy_st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess)
    if (rc == DYNA_OK) rc = fail(DYNA_ERR_CUDA, "DynaAlign CUDA: %s", cudaGetErrorString(cudaGetLastError()));
  for (auto& e : ev)
    if (e) cudaEventDestroy(e);
  cudaStreamDestroy(copy_st);
  p->launches = launches;
  return rc;
}

// Position-weighted checksum of the plan's counts slab (see gather.cu): additive over the slabs of a partition.
extern "C" int dyna_mh_plan_checksum(dyna_mh_plan* p, uint64_t* sum_out, void* stream) {
  if (!p || !sum_out) return fail(DYNA_ERR_INVALID, "dyna_mh_plan_checksum: null argument");
  DYNA_TRY(use_device(p->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  p->last_stream = st;
  WorkStreamGuard work_stream_guard(st);
  DevBuf<unsigned long long> d_sum;
  DYNA_TRY(d_sum.alloc(1));
  if (p->sparse_valid && !p->counts_valid) {
    DYNA_TRY(mh_sparse_checksum(p->sp_keys.p, p->sp_counts.p, p->sp_runs, p->n, d_sum.p, st));
  } else {
    if (!p->counts_valid) return fail(DYNA_ERR_INVALID, "dyna_mh_plan_checksum: no match counts on the device (run a match first)");
    DYNA_TRY(launch_checksum_u16(p->counts.p, p->pairs, tri_strict_rows(p->n, p->row_begin), d_sum.p, st));
  }
  unsigned long long h = 0;
  DYNA_CUDA(cudaMemcpyAsync(&h, d_sum.p, sizeof h, cudaMemcpyDeviceToHost, st));
  DYNA_CUDA(cudaStreamSynchronize(st));
  *sum_out = h;
  return DYNA_OK;
}

// run_match_fetch in the narrow host form: one byte per pair (counts saturated at 255) plus the exact (pair index,
// count) of every pair that reached 255 -- lossless, and half the device-to-host traffic of the u16 triangle, which is
// what bounds the dense end-to-end rate (10 GB at BASELINE config 4).  Escapes arrive in no particular order.  If more
// than esc_capacity pairs escape, *n_esc_out reports how many there are and the call fails with DYNA_ERR_INVALID.
extern "C" int dyna_mh_plan_run_match_fetch8(dyna_mh_plan* p, uint8_t* counts8_out, int64_t esc_capacity, int64_t* esc_index_out,
                                             uint16_t* esc_count_out, int64_t* n_esc_out, void* stream) {
  if (!p) return fail(DYNA_ERR_INVALID, "null plan");
  if (!p->have_sigT) return fail(DYNA_ERR_INVALID, "dyna_mh_plan_run_match_fetch8: no signatures on the device");
  if (esc_capacity < 0 || (esc_capacity > 0 && (!esc_index_out || !esc_count_out)))
    return fail(DYNA_ERR_INVALID, "dyna_mh_plan_run_match_fetch8: bad escape buffers");
  DYNA_TRY(use_device(p->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  p->last_stream = st;
  WorkStreamGuard work_stream_guard(st);
  if (n_esc_out) *n_esc_out = 0;
  const int64_t rows = p->row_end - p->row_begin;
  if (rows <= 0 || p->pairs <= 0) return DYNA_OK;
  DYNA_TRY(mh_plan_ensure_counts(p));
  p->counts_valid = true;  // every chunk below writes its part of the triangle
  p->sparse_valid = false;
  const int nchunks = (int)std::max<int64_t>(1, std::min<int64_t>(16, p->pairs / (64ll << 20)));
  std::vector<int64_t> b((size_t)nchunks + 1);
  b[0] = p->row_begin;
  const int64_t base = tri_strict_rows(p->n, p->row_begin);
  int64_t r = p->row_begin;
  for (int c = 1; c < nchunks; ++c) {
    const int64_t target = base + p->pairs * c / nchunks;
    while (r < p->row_end && tri_strict_rows(p->n, r) < target) ++r;
    b[(size_t)c] = r;
  }
  b[(size_t)nchunks] = p->row_end;
  DevBuf<uint8_t> d8;
  DevBuf<long long> d_ei;
  DevBuf<uint16_t> d_ec;
  DevBuf<unsigned long long> d_en;
  DYNA_TRY(d8.alloc((size_t)p->pairs));
  DYNA_TRY(d_ei.alloc((size_t)std::max<int64_t>(esc_capacity, 1)));
  DYNA_TRY(d_ec.alloc((size_t)std::max<int64_t>(esc_capacity, 1)));
  DYNA_TRY(d_en.alloc(1));
  DYNA_CUDA(cudaMemsetAsync(d_en.p, 0, sizeof(unsigned long long), st));
  cudaStream_t copy_st;
  DYNA_CUDA(cudaStreamCreateWithFlags(&copy_st, cudaStreamNonBlocking));
  std::vector<cudaEvent_t> ev((size_t)nchunks, nullptr);
  int rc = DYNA_OK, launches = 0;
  for (int c = 0; c < nchunks && rc == DYNA_OK; ++c) {
    cudaEventCreateWithFlags(&ev[(size_t)c], cudaEventDisableTiming);
    const int64_t r0 = b[(size_t)c], r1 = b[(size_t)c + 1];
    if (r1 <= r0) continue;
    const int64_t off = tri_strict_rows(p->n, r0) - base, cnt = tri_strict_rows(p->n, r1) - tri_strict_rows(p->n, r0);
    int l = 0;
    rc = launch_mh_match(p->sigT.p, p->npitch, p->hrows, p->n_hash, p->n, r0, r1, p->counts.p + off,
                         p->use16 ? p->sigP.p : nullptr, p->use16 ? p->overflow.p : nullptr, st, &l);
    launches += l;
    if (rc == DYNA_OK)
      rc = launch_mh_narrow8(p->counts.p + off, cnt, base + off, d8.p + off, esc_capacity, d_ei.p, d_ec.p, d_en.p, st);
    ++launches;
    if (rc != DYNA_OK) break;
    cudaEventRecord(ev[(size_t)c], st);
    cudaStreamWaitEvent(copy_st, ev[(size_t)c], 0);
    if (cnt > 0 && cudaMemcpyAsync(counts8_out + off, d8.p + off, (size_t)cnt, cudaMemcpyDeviceToHost, copy_st) != cudaSuccess)
      rc = fail(DYNA_ERR_CUDA, "DynaAlign CUDA: device-to-host copy failed: %s", cudaGetErrorString(cudaGetLastError()));
  }
  unsigned long long n_esc = 0;
  if (rc == DYNA_OK && cudaMemcpyAsync(&n_esc, d_en.p, sizeof n_esc, cudaMemcpyDeviceToHost, st) != cudaSuccess)
    rc = fail(DYNA_ERR_CUDA, "DynaAlign CUDA: %s", cudaGetErrorString(cudaGetLastError()));
  if (cudaStreamSynchronize(st) != cudaSuccess || cudaStreamSynchronize(copy_st) != cudaSuccess)
    if (rc == DYNA_OK) rc = fail(DYNA_ERR_CUDA, "DynaAlign CUDA: %s", cudaGetErrorString(cudaGetLastError()));
  if (rc == DYNA_OK) {
    if (n_esc_out) *n_esc_out = (int64_t)n_esc;
    if ((int64_t)n_esc > esc_capacity)
      rc = fail(DYNA_ERR_INVALID, "escape buffer too small: %lld pairs have a count >= 255, capacity %lld", (long long)n_esc,
                (long long)esc_capacity);
    else if (n_esc > 0) {
      static_assert(sizeof(long long) == sizeof(int64_t), "escape index type");
      if (cudaMemcpy(esc_index_out, d_ei.p, sizeof(int64_t) * n_esc, cudaMemcpyDeviceToHost) != cudaSuccess ||
          cudaMemcpy(esc_count_out, d_ec.p, sizeof(uint16_t) * n_esc, cudaMemcpyDeviceToHost) != cudaSuccess)
        rc = fail(DYNA_ERR_CUDA, "DynaAlign CUDA: %s", cudaGetErrorString(cudaGetLastError()));
    }
  }
  for (auto& e : ev)
    if (e) cudaEventDestroy(e);
  cudaStreamDestroy(copy_st);
  p->launches = launches;
  return rc;
}

extern "C" int64_t dyna_mh_plan_pairs(const dyna_mh_plan* p) { return p ? p->pairs : 0; }
extern "C" int dyna_mh_plan_launches(const dyna_mh_plan* p) { return p ? p->launches : 0; }
extern "C" void* dyna_mh_plan_counts_device_ptr(dyna_mh_plan* p) {
  if (!p || use_device(p->device) != DYNA_OK || mh_plan_need_dense(p, p->last_stream, "dyna_mh_plan_counts_device_ptr") != DYNA_OK)
    return nullptr;
  return p->counts.p;
}
extern "C" void dyna_mh_plan_destroy(dyna_mh_plan* p) {
  if (!p) return;
  cudaSetDevice(p->device);
  delete p;
}

// =====================================================================================================
// NW plan
// =====================================================================================================
struct NwClass {
  int kind;  // 0 empty rows, 1 thread kernel, 2 warp kernel, 3 warp multipass, 4 two-pairs-per-warp 16-bit, 5 two-pairs-per-thread 16-bit,
             // 6 two-pairs-per-warp 16-bit, several passes, 7 two-pairs-per-warp-pair 16-bit (cooperating warps),
             // 8 two ROWS per warp 16-bit (units[].row and row+1 against the same column sequences), 9 its cooperative form,
             // 10 two rows per thread (short probes), 12 two rows per warp, several passes (rows 769..3072)
  int R;
  int64_t work = 0;  // DP cells of the class (launch order: largest first)
  std::vector<NwUnit> units;
  DevBuf<NwUnit> d_units;
};

constexpr int kNwSideStreams = 3;

constexpr int kNwLayoutOnly = -12345;  // `device` of a plan that only holds the work units (dyna_nw_plan_layout)

struct dyna_nw_plan {
  int device = 0;
  int64_t n = 0, row_begin = 0, row_end = 0, pairs = 0, cells = 0;
  int gap_open = 0, gap_ext = 0;
  bool slant = false;
  uint32_t bias16 = 0;  // offset of nw_rows2_kernel's unsigned 16-bit domain; 0 = signed lanes
  int launches = 0;
  int max_cols = 0;
  std::vector<std::unique_ptr<NwClass>> classes;
  DevBuf<uint8_t> codes;
  DevBuf<int32_t> off;
  DevBuf<int8_t> sub;
  DevBuf<uint32_t> matches, length;
  DevBuf<int32_t> scratch;
  DevBuf<uint4> scratch2;  // packed multi-pass kernel: boundary rows
  // The kernel classes of one plan are independent (disjoint pairs): the largest runs on the caller's stream, the
  // others on side streams forked from and joined back into it, so that small classes fill the tail of the large one
  // instead of each paying its own tail (BASELINE config 2: one 45 ms launch and four launches of ~1 ms).
  cudaStream_t side[kNwSideStreams] = {nullptr, nullptr, nullptr};
  cudaEvent_t ev_fork = nullptr, ev_join[kNwSideStreams] = {nullptr, nullptr, nullptr};
  cudaStream_t last_stream = nullptr;  // synchronised before any buffer is released (see DevBuf)
  bool owner_synced = false;           // the owner has already waited for the last work that touches this plan's buffers
  ~dyna_nw_plan() {
    if (device == kNwLayoutOnly) return;  // never touched a device
    if (!owner_synced) cudaStreamSynchronize(last_stream);
    for (int i = 0; i < kNwSideStreams; ++i) {
      if (side[i]) cudaStreamSynchronize(side[i]), cudaStreamDestroy(side[i]);
      if (ev_join[i]) cudaEventDestroy(ev_join[i]);
    }
    if (ev_fork) cudaEventDestroy(ev_fork);
  }
};

extern "C" dyna_nw_plan* dyna_nw_plan_create(const uint8_t* residues, const int64_t* offsets, int64_t n,
                                             const char* matrix_name, int gap_open, int gap_ext, int64_t row_begin,
                                             int64_t row_end, int device) {
  // argument checks in the reference's order: substitution matrix name first (src/pairwiseSeqAlign.cpp:338), then
  // residues as the pair loop meets them
  PhaseTimer timer;
  const int t = matrix_name ? find_table(matrix_name) : -1;
  if (t < 0) {
    fail(DYNA_ERR_INVALID, "Invalid substitution matrix name: %s", matrix_name ? matrix_name : "(null)");
    return nullptr;
  }
  if (n < 0 || row_begin < 0 || row_end > n || row_begin > row_end) {
    fail(DYNA_ERR_INVALID, "dyna_nw_plan_create: bad row range");
    return nullptr;
  }
  if (check_offsets(offsets, n, "dyna_nw_plan_create") != DYNA_OK) return nullptr;
  if (validate_residues(residues, offsets, n) != DYNA_OK) return nullptr;
  const int64_t total = n ? offsets[n] : 0;
  if (total >= (1ll << 31)) {
    fail(DYNA_ERR_UNSUPPORTED, "more than 2^31 residues in one call is not supported");
    return nullptr;
  }
  int64_t max_len = 0;
  for (int64_t i = 0; i < n; ++i) max_len = std::max(max_len, offsets[i + 1] - offsets[i]);
  if (max_len > 65535) {
    fail(DYNA_ERR_UNSUPPORTED, "sequences longer than 65535 residues are not supported");
    return nullptr;
  }
  // int32 head-room: every DP value (plus the slant) must stay far from INT_MIN/2 and INT_MAX
  const double mag = std::fabs((double)gap_open) + 2.0 * (double)max_len * (std::fabs((double)gap_ext) * 2.0 + 20.0);
  if (mag > 2.5e8) {
    fail(DYNA_ERR_UNSUPPORTED, "gap penalties / sequence lengths exceed the exact int32 range of the DP");
    return nullptr;
  }
  timer.lap("nw plan: validation");
  const bool layout_only = (device == kNwLayoutOnly);  // dyna_nw_plan_layout: the work units only, no device touched
  if (!layout_only && use_device(device) != DYNA_OK) return nullptr;
  timer.lap("nw plan: device");

  std::unique_ptr<dyna_nw_plan> p(new dyna_nw_plan);
  p->device = device;
  p->n = n;
  p->row_begin = row_begin;
  p->row_end = row_end;
  p->gap_open = gap_open;
  p->gap_ext = gap_ext;
  p->pairs = tri_diag_rows(n, row_end) - tri_diag_rows(n, row_begin);
  p->max_cols = (int)std::max<int64_t>(max_len, 1);
  // slanted recurrence needs score + 2*ge to fit the profile's int8 score byte
  int smin = 127, smax = -128;
  for (int i = 0; i < 576; ++i) {
    smin = std::min<int>(smin, tables().full[t][i]);
    smax = std::max<int>(smax, tables().full[t][i]);
  }
  p->slant = (smin + 2 * (int64_t)gap_ext >= -128) && (smax + 2 * (int64_t)gap_ext <= 127);
  if (const char* e = getenv("DYNA_NW_SLANT")) p->slant = p->slant && (atoi(e) != 0);
  // 16-bit two-pairs-per-warp kernel: every slanted DP value of a (row, column) pair lies in
  //   [smin - 3*go + 2*ge,  max(smax,0)*min(m,n) + (m+n)*ge]   (DESIGN.md section 4), which must fit int16 with room
  // for the flat -30000 sentinel below and the "- go" of a gap opening above.
  bool pack16 = p->slant && gap_ext >= 0 && gap_open >= 0 && (smin - 3ll * gap_open + 2ll * gap_ext) >= -24000;
  if (const char* e = getenv("DYNA_NW_PACK16")) pack16 = pack16 && (atoi(e) != 0);
  // The two-rows kernels' unsigned domain: value + bias16 with bias16 = -(lower bound) + go + margin, so that the smallest
  // value still exceeds go (H - go never wraps) and 0 is "minus infinity"; the largest is 32000 + bias16 < 65536 because
  // the lower bound is above -24000 (and go below 8000 with it).  Negative table scores are fine: the score words are
  // integer sums of the two halves (score2_word in nw_kernels.cu).  DYNA_NW_U16=0 keeps the signed lanes (A/B measurements).
  if (pack16) {
    // (+ the most negative slanted score once more: the diagonal candidate diag + s is formed before the max)
    p->bias16 = (uint32_t)(-(smin - 3ll * gap_open + 2ll * gap_ext) + gap_open + 64 + std::max<int64_t>(0, -(smin + 2ll * gap_ext)));
    if (const char* e = getenv("DYNA_NW_U16")) if (atoi(e) == 0) p->bias16 = 0;
  }
  // rows above this length take the multi-pass form of the packed kernel (strips of <= 12 rows keep the fast
  // ping-pong / increment-table configuration); measured cross-over against the single-pass tall-strip form
  int mp_min_rows = 32 * kNwWarp2MaxR + 1;  // 385..640 rows: single pass with tall strips measured faster (2.60 vs 2.25 TCUPS)
  if (const char* e = getenv("DYNA_NW_MP_MINROWS")) mp_min_rows = std::max(32 * 12 + 1, atoi(e));
  bool need_scratch2 = false;
  // rows 385..768: two cooperating warps per pair-set (nw_warp2co_kernel); DYNA_NW_CO=0 restores the tall single-pass
  // strips (385..640) and the multi-pass kernel (641..768) for A/B measurements
  bool use_co = true;
  if (const char* e = getenv("DYNA_NW_CO")) use_co = atoi(e) != 0;
  // columns per unit of the packed single-pass kernel: wide units amortise the per-unit table build (+0.9 % at
  // BASELINE config 5), but a small input needs the finer grain to fill 296 CTA slots (config 2: 2975 vs 2880 GCUPS)
  // rows of 33..384 residues are taken two at a time (rows i, i+1 against the same columns: nw_rows2_kernel);
  // DYNA_NW_ROWS2=0 restores one row against two column sequences (nw_warp2_kernel) for A/B measurements
  bool use_rows2 = p->pairs >= 10000;
  if (const char* e = getenv("DYNA_NW_ROWS2")) use_rows2 = atoi(e) != 0;
  // its unit is the row pair against up to 256 column sequences on a whole SM; smaller inputs get narrower units so that
  // there are ~9 or more per SM (measured on 330-residue proteins, GCUPS at 32 / 64 / 128 / 256 columns: n = 300
  // 3608 / 3184 / 2411 / 2092, n = 600 3898 / 4002 / 3854 / 3197, n = 1000 3994 / 4084 / 4097 / 3938, n = 2000
  // 4034 / 4157 / 4210 / 4228; the two-columns kernel it replaces there: 2619, 3066, 3179, 3248)
  int rows2_cols = 32;
  while (rows2_cols < kNwRows2UnitCols && p->pairs / (2 * 2 * rows2_cols) >= 1300) rows2_cols *= 2;
  if (const char* e = getenv("DYNA_NW_ROWS2_COLS")) rows2_cols = std::min(kNwRows2UnitCols, std::max(1, atoi(e)));
  // short probes (rows <= 32 residues): two rows per thread; DYNA_NW_TROWS2=0 restores one row against two columns
  bool use_trows2 = pack16;
  if (const char* e = getenv("DYNA_NW_TROWS2")) use_trows2 = use_trows2 && atoi(e) != 0;
  // the cooperative form of the same (row pairs of 385..576 residues).  Units of 128 columns also on small inputs
  // (BASELINE config 2: 3337 GCUPS against 3291 at 64 and 3207 at 32 columns)
  bool use_rows2co = use_co;
  if (const char* e = getenv("DYNA_NW_ROWS2CO")) use_rows2co = use_rows2co && atoi(e) != 0;
  // row pairs beyond the cooperative form: the two-rows multi-pass kernel (DYNA_NW_ROWS2MP=0: one row against two columns)
  bool use_rows2mp = pack16;
  if (const char* e = getenv("DYNA_NW_ROWS2MP")) use_rows2mp = use_rows2mp && atoi(e) != 0;
  int rows2mp_cols = 64;  // four rounds of 16 column sequences: its units are long (several passes), keep them many
  if (const char* e = getenv("DYNA_NW_ROWS2MP_COLS")) rows2mp_cols = std::min(kNwRows2UnitCols, std::max(16, atoi(e)));
  int rows2co_cols = kNwCoUnitCols;
  if (const char* e = getenv("DYNA_NW_ROWS2CO_COLS")) rows2co_cols = std::min(kNwCoUnitCols, std::max(1, atoi(e)));
  int warp2_cols = p->pairs >= (int64_t)kNwWarp2UnitColsMax * kNwMultiPassGrid * 32 ? kNwWarp2UnitColsMax : kNwWarp2UnitCols;
  if (const char* e = getenv("DYNA_NW_UNITCOLS")) warp2_cols = std::min(kNwWarp2UnitColsMax, std::max(2, atoi(e) & ~1));

  // encode residues, 32-bit offsets
  std::vector<uint8_t> codes((size_t)total + 4);
  std::vector<int32_t> off32((size_t)n + 1);
  const int8_t* aa = tables().aa;
  for (int64_t q = 0; q < total; ++q) codes[(size_t)q] = (uint8_t)aa[residues[q]];
  for (int64_t i = 0; i <= n; ++i) off32[(size_t)i] = (int32_t)(n ? offsets[i] : 0);

  // suffix sums of lengths -> exact cell count of the row range
  {
    std::vector<int64_t> suffix((size_t)n + 1, 0);
    for (int64_t i = n - 1; i >= 0; --i) suffix[(size_t)i] = suffix[(size_t)i + 1] + (offsets[i + 1] - offsets[i]);
    int64_t cells = 0;
    for (int64_t i = row_begin; i < row_end; ++i) cells += (offsets[i + 1] - offsets[i]) * suffix[(size_t)i];
    p->cells = cells;
  }

  // work units per kernel class
  std::map<std::pair<int, int>, NwClass*> by_key;
  auto get_class = [&](int kind, int R) -> NwClass* {
    auto key = std::make_pair(kind, R);
    auto it = by_key.find(key);
    if (it != by_key.end()) return it->second;
    p->classes.emplace_back(new NwClass);
    NwClass* c = p->classes.back().get();
    c->kind = kind;
    c->R = R;
    by_key[key] = c;
    return c;
  };
  bool need_scratch = false;
  // range-maximum table over the sequence lengths: the packed (16-bit) kernels are chosen per UNIT from the longest
  // column sequence the unit contains, so a few very long sequences do not push every pair onto the 32-bit path
  std::vector<std::vector<int32_t>> rmq;
  {
    rmq.emplace_back((size_t)n);
    for (int64_t i = 0; i < n; ++i) rmq[0][(size_t)i] = (int32_t)(offsets[i + 1] - offsets[i]);
    for (int64_t w = 1; (2 * w) <= n; w *= 2) {
      const std::vector<int32_t>& prev = rmq.back();
      std::vector<int32_t> cur((size_t)(n - 2 * w + 1));
      for (int64_t i = 0; i + 2 * w <= n; ++i) cur[(size_t)i] = std::max(prev[(size_t)i], prev[(size_t)(i + w)]);
      rmq.push_back(std::move(cur));
    }
  }
  auto range_max = [&](int64_t lo, int64_t hi) -> int64_t {  // max length over [lo, hi), hi > lo
    int lvl = 0;
    while ((2ll << lvl) <= hi - lo) ++lvl;
    return std::max(rmq[(size_t)lvl][(size_t)lo], rmq[(size_t)lvl][(size_t)(hi - (1ll << lvl))]);
  };
  auto fits16u = [&](int m, int64_t nmax) {  // fits16() with the unit's own longest column
    const int64_t hi = (int64_t)std::max(smax, 0) * std::min<int64_t>(m, nmax) + ((int64_t)m + nmax) * gap_ext + gap_open;
    return pack16 && hi <= 32000 && (int64_t)m + nmax <= 65535;
  };
  const bool force_warp2 = getenv("DYNA_NW_FORCE_WARP2") != nullptr;
  std::vector<int64_t> len_prefix((size_t)n + 1, 0);
  for (int64_t i = 0; i < n; ++i) len_prefix[(size_t)i + 1] = len_prefix[(size_t)i] + (offsets[i + 1] - offsets[i]);
  // one row against the column sequences [j_lo, j_hi): the kernel is chosen per unit from the unit's longest column
  auto emit_single = [&](int64_t i, int64_t j_lo, int64_t j_hi) {
    const int m = (int)(offsets[i + 1] - offsets[i]);
    int64_t j = j_lo;
    while (j < j_hi) {
      int kind, R, step;
      const bool co_rows = use_co && !force_warp2 && m >= kNwCoMinRows && m <= kNwCoMaxRows;
      if (m == 0) { kind = 0; R = 0; step = 4096; }
      else {
        // try the packed kernels on the widest unit they use; fall back to the 32-bit kernels for this stretch
        const bool warp2_rows = m <= 32 * kNwWarp2MaxR && (m < mp_min_rows || force_warp2);
        const int pstep = nw_use_thread_kernel(m) && !force_warp2 ? 2 * kNwThreadUnitPairs
                          : co_rows ? kNwCoUnitCols : warp2_rows ? warp2_cols : 2 * kNwWarpUnitPairs;
        const int64_t nmax = range_max(j, std::min<int64_t>(j + pstep, j_hi));
        const bool p16 = fits16u(m, nmax);
        if (force_warp2 && p16 && nmax <= kNwWarp2MaxCols && m <= 32 * kNwWarp2MaxR) { kind = 4; R = std::max(2, nw_warp_R(m)); step = pstep; }
        else if (nw_use_thread_kernel(m) && p16) { kind = 5; R = nw_thread_R(m); step = pstep; }
        else if (nw_use_thread_kernel(m)) { kind = 1; R = nw_thread_R(m); step = kNwThreadUnitPairs; }
        else if (co_rows && p16 && nmax <= kNwWarp2MaxCols) { kind = 7; R = nw_co_R(m); step = pstep; }
        else if (p16 && nmax <= kNwWarp2MaxCols && m < mp_min_rows && m <= 32 * kNwWarp2MaxR) { kind = 4; R = nw_warp_R(m); step = pstep; }
        else if (p16 && nmax <= kNwWarp2MpMaxCols && m > 32 * 6 && m <= kNwWarp2MpMaxRows) {
          // long rows, or columns too long for the single-pass kernel's staging buffer: the multi-pass form
          const int npass = (m + 32 * 12 - 1) / (32 * 12);
          kind = 6; R = std::max(7, (m + 32 * npass - 1) / (32 * npass)); step = pstep; need_scratch2 = true;
        }
        else if (m <= 32 * kNwWarpMaxR) { kind = 2; R = nw_warp_R(m); step = kNwWarpUnitPairs; }
        else { kind = 3; R = kNwWarpMaxR; step = 8; need_scratch = true; }
      }
      // the multi-pass kernel keeps one scratch line per pair-set: at most 64 columns per unit
      if (kind == 6) step = std::min(step, 2 * kNwWarpUnitPairs);
      const int64_t cnt = std::min<int64_t>(step, j_hi - j);
      NwClass* cls = get_class(kind, R);
      cls->units.push_back(NwUnit{(int32_t)i, (int32_t)j, (int32_t)cnt});
      cls->work += (int64_t)std::max(m, 1) * (len_prefix[(size_t)(j + cnt)] - len_prefix[(size_t)j] + cnt);
      j += cnt;
    }
  };
  // Two-rows kernels: row i takes a partner row i2 > i of the same kernel family and similar length -- the next such row
  // within a window, not necessarily i + 1, so that inputs of mixed lengths pair up too.  Both rows then share the column
  // sequences j >= i2; the columns i <= j < i2 exist for row i only (that half is computed and dropped: at most the
  // window's width of them).  The decision is per UNIT: a unit whose longest column sequence leaves the 16-bit range or
  // the staging buffer is handed to the single-row path for both rows, the other units of the pair keep the fast kernel.
  // family: 1 short probes (nw_thread_rows2_kernel), 2 rows 33..384 (nw_rows2_kernel), 3 rows 385..768 (nw_rows2co_kernel),
  // 4 rows 769..3072 (nw_rows2mp_kernel)
  auto family_of = [&](int m) -> int {
    if (force_warp2) return 0;
    if (use_trows2 && m >= 1 && m <= kNwThreadMaxRows) return 1;
    if (use_rows2 && m > kNwThreadMaxRows && m <= 32 * 12) return 2;
    if (use_rows2co && m >= kNwCoMinRows && m <= kNwRows2CoMaxRows && nw_co_R(m) >= 7) return 3;
    if (use_rows2mp && p->bias16 != 0u && m > kNwRows2CoMaxRows && m <= kNwRows2MpMaxRows) return 4;
    return 0;
  };
  auto compatible = [&](int fam, int ma, int mb) -> bool {
    const int mx = std::max(ma, mb), mn = std::min(ma, mb);
    if (fam == 1) return true;
    if (fam == 2) return 4 * mn >= 3 * mx;
    if (fam == 4)  // multi-pass form: both rows end in the last pass of 32*R rows
      return 4 * mn >= 3 * mx && (mn - 1) / (32 * nw_rows2mp_R(mx)) == (mx - 1) / (32 * nw_rows2mp_R(mx));
    // cooperative form: both rows reach into the second warp's block of 32*R rows
    return nw_co_R(mx) >= 7 && (mn - 1) / (32 * nw_co_R(mx)) == 1;
  };
  // window: wide enough to find a row of nearly the same length, small against the number of columns a row meets (the
  // columns between the two rows are computed for the first row only: n = 300 loses 3 % with a 6-row window); measured
  // on the length mix, n = 3000: 32 / 64 / 128 rows 2965 / 3006 / 3023 GCUPS, config-5 sample 4314 / 4299 / 4295
  int pair_window = (int)std::min<int64_t>(32, std::max<int64_t>(1, n / 96));
  if (const char* e = getenv("DYNA_NW_PAIR_WINDOW")) pair_window = std::min(4096, std::max(1, atoi(e)));
  std::vector<uint8_t> covered((size_t)(row_end - row_begin), 0);
  auto plan_units = [&]() {
  for (int64_t i = row_begin; i < row_end; ++i) {
    if (covered[(size_t)(i - row_begin)]) continue;
    const int m = (int)(offsets[i + 1] - offsets[i]);
    const int fam = family_of(m);
    int64_t i2 = -1;
    if (fam != 0) {  // the compatible row of the closest length inside the window (ties: the nearest)
      int64_t best = INT64_MAX;
      for (int64_t q = i + 1; q < std::min<int64_t>(i + 1 + pair_window, row_end); ++q) {
        if (covered[(size_t)(q - row_begin)]) continue;
        const int mq = (int)(offsets[q + 1] - offsets[q]);
        if (family_of(mq) != fam || !compatible(fam, m, mq)) continue;
        // strips follow the longer row: what counts is the padding of the shorter one, in strip rows per lane
        const int64_t cost = (int64_t)std::abs(m - mq) * 4096 + (q - i);
        if (cost < best) { best = cost; i2 = q; }
      }
    }
    if (i2 < 0) {
      emit_single(i, i, n);
      continue;
    }
    covered[(size_t)(i2 - row_begin)] = 1;
    const int m2 = (int)(offsets[i2 + 1] - offsets[i2]);
    const int mx = std::max(m, m2);
    const int kind = fam == 1 ? 10 : fam == 2 ? 8 : fam == 3 ? 9 : 12;
    const int R = fam == 1 ? nw_thread_R(mx) : fam == 2 ? nw_warp_R(mx) : fam == 3 ? nw_co_R(mx) : nw_rows2mp_R(mx);
    const int cols = fam == 1 ? 2 * kNwThreadUnitPairs : fam == 3 ? rows2co_cols : fam == 4 ? rows2mp_cols : rows2_cols;
    NwClass* cls = get_class(kind, R);
    for (int64_t j = i; j < n; j += cols) {
      const int64_t cnt = std::min<int64_t>(cols, n - j);
      const int64_t nmax = range_max(j, j + cnt);
      if (fits16u(mx, nmax) && (fam == 1 || nmax <= (fam == 3 ? nw_rows2co_max_cols(R) : kNwRows2MaxCols))) {
        if (fam == 4) need_scratch2 = true;
        cls->units.push_back(NwUnit{(int32_t)i, (int32_t)j, nw_pack_count(cnt, i2 - i)});
        cls->work += (int64_t)(m + m2) * (len_prefix[(size_t)(j + cnt)] - len_prefix[(size_t)j] + cnt);
      } else {
        emit_single(i, j, j + cnt);
        if (j + cnt > i2) emit_single(i2, std::max(j, i2), j + cnt);
      }
    }
  }
  };
  plan_units();
  // The two-rows multi-pass kernel runs as one persistent CTA per SM with long units: as a small class next to the
  // others it costs more in its tail than it gains (length mix, n = 3000, 5 % of the cells: 2880 vs 2974 GCUPS without;
  // long proteins, 99 % of the cells: 3378 vs 2629).  Below a tenth of the plan's cells its rows go back to the one-row
  // multi-pass kernel.
  if (use_rows2mp && !getenv("DYNA_NW_ROWS2MP")) {
    int64_t total = 0, mp = 0;
    for (const auto& cl : p->classes) {
      total += cl->work;
      if (cl->kind == 12) mp += cl->work;
    }
    if (mp > 0 && mp * 10 < total) {
      use_rows2mp = false;
      p->classes.clear();
      by_key.clear();
      need_scratch = need_scratch2 = false;
      std::fill(covered.begin(), covered.end(), (uint8_t)0);
      plan_units();
    }
  }

  std::stable_sort(p->classes.begin(), p->classes.end(),
                   [](const std::unique_ptr<NwClass>& a, const std::unique_ptr<NwClass>& b) { return a->work > b->work; });
  timer.lap("nw plan: work units");
  if (layout_only) return p.release();
  if (p->codes.alloc(codes.size()) || p->off.alloc(off32.size()) || p->sub.alloc(576) ||
      p->matches.alloc((size_t)p->pairs) || p->length.alloc((size_t)p->pairs))
    return nullptr;
  if (need_scratch && p->scratch.alloc((size_t)kNwMultiPassGrid * 8 * 3 * (size_t)p->max_cols)) return nullptr;
  if (need_scratch2 && p->scratch2.alloc((size_t)kNwMultiPassGrid * 32 * (size_t)kNwWarp2MpMaxCols)) return nullptr;
  timer.lap("nw plan: device buffers");
  auto cp = [&](void* dst, const void* src, size_t bytes) {
    return cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice) == cudaSuccess;
  };
  bool ok = cp(p->codes.p, codes.data(), codes.size()) && cp(p->off.p, off32.data(), off32.size() * sizeof(int32_t)) &&
            cp(p->sub.p, tables().full[t], 576);
  for (auto& c : p->classes) {
    if (c->d_units.alloc(c->units.size())) return nullptr;
    ok = ok && cp(c->d_units.p, c->units.data(), c->units.size() * sizeof(NwUnit));
  }
  if (!ok) {
    fail(DYNA_ERR_CUDA, "DynaAlign CUDA: host-to-device copy failed: %s", cudaGetErrorString(cudaGetLastError()));
    return nullptr;
  }
  if (p->classes.size() > 1 && !getenv("DYNA_NW_SERIAL")) {
    bool sok = cudaEventCreateWithFlags(&p->ev_fork, cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; i < kNwSideStreams && sok; ++i)
      sok = cudaStreamCreateWithFlags(&p->side[i], cudaStreamNonBlocking) == cudaSuccess &&
            cudaEventCreateWithFlags(&p->ev_join[i], cudaEventDisableTiming) == cudaSuccess;
    if (!sok) {
      fail(DYNA_ERR_CUDA, "DynaAlign CUDA: stream creation failed: %s", cudaGetErrorString(cudaGetLastError()));
      return nullptr;
    }
  }
  timer.lap("nw plan: uploads");
  return p.release();
}

extern "C" int dyna_nw_plan_run(dyna_nw_plan* p, void* stream) {
  if (!p) return fail(DYNA_ERR_INVALID, "null plan");
  DYNA_TRY(use_device(p->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  NwDeviceData d;
  d.codes = p->codes.p;
  d.off = p->off.p;
  d.sub = p->sub.p;
  d.n = p->n;
  d.slab_base = tri_diag_rows(p->n, p->row_begin);
  d.matches = p->matches.p;
  d.length = p->length.p;
  d.gap_open = p->gap_open;
  d.gap_ext = p->gap_ext;
  d.one = 1u;
  d.zero = 0u;
  d.bias16 = p->bias16;
  p->launches = 0;
  p->last_stream = st;
  WorkStreamGuard work_stream_guard(st);
  const bool fork = p->side[0] != nullptr;
  bool used[kNwSideStreams] = {false, false, false};
  if (fork) DYNA_CUDA(cudaEventRecord(p->ev_fork, st));
  int next_side = 0;
  bool first = true;
  cudaStream_t kind_stream[8] = {};  // indexed by the scratch owner: 3, or 6 (kinds 6 and 12 share scratch2)
  bool kind_stream_set[8] = {};
  for (auto& c : p->classes) {
    const int nu = (int)c->units.size();
    if (nu == 0) continue;
    cudaStream_t cs = st;
    // the two multi-pass kernels keep per-plan scratch lines: all classes of one such kind share ONE stream (which is
    // the caller's stream if the largest class happens to be of that kind)
    const bool scratch_kind = (c->kind == 3 || c->kind == 6 || c->kind == 12);
    const int sk = c->kind == 12 ? 6 : c->kind;
    if (scratch_kind && kind_stream_set[sk]) {
      cs = kind_stream[sk];
    } else if (fork && !first) {
      const int si = next_side++ % kNwSideStreams;
      if (!used[si]) DYNA_CUDA(cudaStreamWaitEvent(p->side[si], p->ev_fork, 0));
      used[si] = true;
      cs = p->side[si];
    }
    if (scratch_kind) {
      kind_stream_set[sk] = true;
      kind_stream[sk] = cs;
    }
    first = false;
    switch (c->kind) {
      case 0: DYNA_TRY(launch_nw_empty_rows(d, c->d_units.p, nu, cs)); break;
      case 1: DYNA_TRY(launch_nw_thread(c->R, p->slant, d, c->d_units.p, nu, cs)); break;
      case 2: DYNA_TRY(launch_nw_warp(c->R, p->slant, false, d, c->d_units.p, nu, nullptr, 0, cs)); break;
      case 4: DYNA_TRY(launch_nw_warp2(c->R, d, c->d_units.p, nu, cs)); break;
      case 5: DYNA_TRY(launch_nw_thread2(c->R, d, c->d_units.p, nu, cs)); break;
      case 6: DYNA_TRY(launch_nw_warp2mp(c->R, d, c->d_units.p, nu, p->scratch2.p, cs)); break;
      case 7: DYNA_TRY(launch_nw_warp2co(c->R, d, c->d_units.p, nu, cs)); break;
      case 8: DYNA_TRY(launch_nw_rows2(c->R, d, c->d_units.p, nu, cs)); break;
      case 9: DYNA_TRY(launch_nw_rows2co(c->R, d, c->d_units.p, nu, cs)); break;
      case 10: DYNA_TRY(launch_nw_thread_rows2(c->R, d, c->d_units.p, nu, cs)); break;
      case 12: DYNA_TRY(launch_nw_rows2mp(c->R, d, c->d_units.p, nu, p->scratch2.p, cs)); break;
      default: DYNA_TRY(launch_nw_warp(c->R, p->slant, true, d, c->d_units.p, nu, p->scratch.p, p->max_cols, cs)); break;
    }
    ++p->launches;
  }
  for (int i = 0; i < kNwSideStreams; ++i)
    if (used[i]) {
      DYNA_CUDA(cudaEventRecord(p->ev_join[i], p->side[i]));
      DYNA_CUDA(cudaStreamWaitEvent(st, p->ev_join[i], 0));
    }
  return DYNA_OK;
}

extern "C" int dyna_nw_plan_fetch(dyna_nw_plan* p, uint32_t* matches_out, uint32_t* length_out, void* stream) {
  if (!p) return fail(DYNA_ERR_INVALID, "null plan");
  DYNA_TRY(use_device(p->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  p->last_stream = st;
  WorkStreamGuard work_stream_guard(st);
  if (p->pairs > 0) {
    DYNA_CUDA(cudaMemcpyAsync(matches_out, p->matches.p, sizeof(uint32_t) * (size_t)p->pairs, cudaMemcpyDeviceToHost, st));
    DYNA_CUDA(cudaMemcpyAsync(length_out, p->length.p, sizeof(uint32_t) * (size_t)p->pairs, cudaMemcpyDeviceToHost, st));
  }
  DYNA_CUDA(cudaStreamSynchronize(st));
  return DYNA_OK;
}

// Position-weighted checksums of the plan's (matches, length) slab: sum_out[0] over matches, sum_out[1] over length.
extern "C" int dyna_nw_plan_checksum(dyna_nw_plan* p, uint64_t* sum_out, void* stream) {
  if (!p || !sum_out) return fail(DYNA_ERR_INVALID, "dyna_nw_plan_checksum: null argument");
  DYNA_TRY(use_device(p->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  p->last_stream = st;
  WorkStreamGuard work_stream_guard(st);
  DevBuf<unsigned long long> d_sum;
  DYNA_TRY(d_sum.alloc(2));
  const int64_t first = tri_diag_rows(p->n, p->row_begin);
  DYNA_TRY(launch_checksum_u32(p->matches.p, p->pairs, first, d_sum.p, st));
  DYNA_TRY(launch_checksum_u32(p->length.p, p->pairs, first, d_sum.p + 1, st));
  unsigned long long h[2] = {0, 0};
  DYNA_CUDA(cudaMemcpyAsync(h, d_sum.p, sizeof h, cudaMemcpyDeviceToHost, st));
  DYNA_CUDA(cudaStreamSynchronize(st));
  sum_out[0] = h[0];
  sum_out[1] = h[1];
  return DYNA_OK;
}

// fetch in the narrow host form: one byte for matches and one for the alignment length per pair (2 B/pair instead of
// 8).  Only for plans whose every alignment length fits a byte: max_len_i + max_len_j <= 255 (short peptides -- the
// 100,000 x 16-mer target is 10 GB this way instead of 40 GB).
extern "C" int dyna_nw_plan_fetch_packed8(dyna_nw_plan* p, uint8_t* matches8_out, uint8_t* length8_out, void* stream) {
  if (!p) return fail(DYNA_ERR_INVALID, "null plan");
  if (2 * (int64_t)p->max_cols > 255)
    return fail(DYNA_ERR_UNSUPPORTED, "dyna_nw_plan_fetch_packed8: alignment lengths up to %lld do not fit one byte",
                2 * (long long)p->max_cols);
  DYNA_TRY(use_device(p->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  p->last_stream = st;
  WorkStreamGuard work_stream_guard(st);
  if (p->pairs <= 0) return DYNA_OK;
  DevBuf<uint8_t> m8, l8;
  DYNA_TRY(m8.alloc((size_t)p->pairs));
  DYNA_TRY(l8.alloc((size_t)p->pairs));
  DYNA_TRY(launch_nw_pack8(p->matches.p, p->length.p, p->pairs, m8.p, l8.p, st));
  DYNA_CUDA(cudaMemcpyAsync(matches8_out, m8.p, (size_t)p->pairs, cudaMemcpyDeviceToHost, st));
  DYNA_CUDA(cudaMemcpyAsync(length8_out, l8.p, (size_t)p->pairs, cudaMemcpyDeviceToHost, st));
  DYNA_CUDA(cudaStreamSynchronize(st));
  return DYNA_OK;
}

// ---- threshold + sparsify for sim_fn = similarityNW (R/clusterbreak.R:217-221); kernels in nw_post.cu
extern "C" int dyna_nw_plan_max_len(const dyna_nw_plan* p) { return p ? p->max_cols : 0; }

namespace {
constexpr int64_t kNwHistMaxBins = 1ll << 27;  // 1 GB of u64 bins: sequences up to ~8,000 residues

// node description of a plan; uploads the member list (strictly increasing plan indices) when there is one
int nw_make_node(dyna_nw_plan* p, const int32_t* members, int64_t n_members, DevBuf<int32_t>& d_members, cudaStream_t st,
                 NwNode* nd, const char* who) {
  nd->matches = p->matches.p;
  nd->length = p->length.p;
  nd->n = p->n;
  nd->row_begin = p->row_begin;
  nd->row_end = p->row_end;
  nd->slab_base = tri_diag_rows(p->n, p->row_begin);
  nd->members = nullptr;
  nd->n_node = p->n;
  if (!members) return DYNA_OK;
  if (n_members < 0) return fail(DYNA_ERR_INVALID, "%s: negative member count", who);
  for (int64_t t = 0; t < n_members; ++t)
    if (members[t] < 0 || members[t] >= p->n || (t > 0 && members[t] <= members[t - 1]))
      return fail(DYNA_ERR_INVALID, "%s: members must be strictly increasing indices in [0, %lld) (members[%lld] = %d)", who,
                  (long long)p->n, (long long)t, (int)members[t]);
  DYNA_TRY(d_members.alloc((size_t)n_members));
  if (n_members > 0)
    DYNA_CUDA(cudaMemcpyAsync(d_members.p, members, sizeof(int32_t) * (size_t)n_members, cudaMemcpyHostToDevice, st));
  nd->members = d_members.p;
  nd->n_node = n_members;
  return DYNA_OK;
}
}  // namespace

// hist_out: (max_len + 1) x (2 max_len + 1) counters, row-major [matches][length], over the strict upper triangle of the
// node (pairs a < b of `members`, or of all sequences when members == NULL) restricted to the plan's row range
extern "C" int dyna_nw_plan_stat_histogram(dyna_nw_plan* p, const int32_t* members, int64_t n_members, uint64_t* hist_out,
                                           void* stream) {
  if (!p || !hist_out) return fail(DYNA_ERR_INVALID, "dyna_nw_plan_stat_histogram: null argument");
  DYNA_TRY(use_device(p->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  p->last_stream = st;
  WorkStreamGuard work_stream_guard(st);
  const int64_t mdim = (int64_t)p->max_cols + 1, ldim = 2 * (int64_t)p->max_cols + 1;
  if (mdim * ldim > kNwHistMaxBins)
    return fail(DYNA_ERR_UNSUPPORTED, "dyna_nw_plan_stat_histogram: %lld x %lld (matches, length) bins exceed the supported %lld",
                (long long)mdim, (long long)ldim, (long long)kNwHistMaxBins);
  DevBuf<int32_t> d_members;
  NwNode nd;
  DYNA_TRY(nw_make_node(p, members, n_members, d_members, st, &nd, "dyna_nw_plan_stat_histogram"));
  DevBuf<unsigned long long> d_hist;
  DYNA_TRY(d_hist.alloc((size_t)(mdim * ldim)));
  DYNA_TRY(launch_nw_stat_hist(nd, mdim, ldim, d_hist.p, st));
  DYNA_CUDA(cudaMemcpyAsync(hist_out, d_hist.p, sizeof(uint64_t) * (size_t)(mdim * ldim), cudaMemcpyDeviceToHost, st));
  DYNA_CUDA(cudaStreamSynchronize(st));
  return DYNA_OK;
}

// (matches, length) of the node's self-alignments -- the diagonal of the similarity matrix, which
// graph_from_adjacency_matrix(mode = "upper") turns into self-loops (R/clusterbreak.R:122-124); zeros for rows outside
// the plan's row range, so ranks can add their outputs
extern "C" int dyna_nw_plan_fetch_diagonal(dyna_nw_plan* p, const int32_t* members, int64_t n_members, uint32_t* matches_out,
                                           uint32_t* length_out, void* stream) {
  if (!p || !matches_out || !length_out) return fail(DYNA_ERR_INVALID, "dyna_nw_plan_fetch_diagonal: null argument");
  DYNA_TRY(use_device(p->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  p->last_stream = st;
  WorkStreamGuard work_stream_guard(st);
  DevBuf<int32_t> d_members;
  NwNode nd;
  DYNA_TRY(nw_make_node(p, members, n_members, d_members, st, &nd, "dyna_nw_plan_fetch_diagonal"));
  if (nd.n_node > 0) {
    DevBuf<uint32_t> dm, dl;
    DYNA_TRY(dm.alloc((size_t)nd.n_node));
    DYNA_TRY(dl.alloc((size_t)nd.n_node));
    DYNA_TRY(launch_nw_diag(nd, dm.p, dl.p, st));
    DYNA_CUDA(cudaMemcpyAsync(matches_out, dm.p, sizeof(uint32_t) * (size_t)nd.n_node, cudaMemcpyDeviceToHost, st));
    DYNA_CUDA(cudaMemcpyAsync(length_out, dl.p, sizeof(uint32_t) * (size_t)nd.n_node, cudaMemcpyDeviceToHost, st));
    DYNA_CUDA(cudaStreamSynchronize(st));
  }
  return DYNA_OK;
}

// R's quantile(x, prob, type = 7) (R/clusterbreak.R:219) for x = (double)matches / (double)length with the multiplicities of
// a (matches, length) histogram; several ranks' histograms are simply added before the call
extern "C" int dyna_quantile_type7_identities(const uint64_t* hist, int64_t mdim, int64_t ldim, double prob, double* threshold_out) {
  if (!hist || mdim <= 0 || ldim <= 0) return fail(DYNA_ERR_INVALID, "dyna_quantile_type7_identities: bad histogram");
  if (!(prob >= 0.0 && prob <= 1.0)) return fail(DYNA_ERR_INVALID, "'probs' outside [0,1]");
  if (hist[0] != 0)  // two empty sequences: 0/0 = NaN in the reference's matrix (src/pairwiseSeqAlign.cpp:311)
    return fail(DYNA_ERR_INVALID, "missing values and NaN's not allowed if 'na.rm' is FALSE");
  std::vector<std::pair<double, uint64_t>> v;
  for (int64_t m = 0; m < mdim; ++m)
    for (int64_t l = 1; l < ldim; ++l)
      if (const uint64_t c = hist[m * ldim + l]) v.emplace_back((double)m / (double)l, c);
  if (v.empty()) return fail(DYNA_ERR_INVALID, "quantile of an empty set of pairs");
  std::sort(v.begin(), v.end());
  long double total = 0;
  for (const auto& e : v) total += (long double)e.second;
  const double N = (double)total;
  const double index = 1.0 + std::max(N - 1.0, 0.0) * prob;
  const double lo = std::floor(index), hi = std::ceil(index);
  auto value_at_rank = [&](double r) {  // r-th smallest, 1-based
    long double cum = 0;
    for (const auto& e : v) {
      cum += (long double)e.second;
      if ((double)cum >= r) return e.first;
    }
    return v.back().first;
  };
  double qs = value_at_rank(lo);
  const double xhi = value_at_rank(hi);
  if (index > lo && xhi != qs) {
    const double h = index - lo;
    qs = (1.0 - h) * qs + h * xhi;
  }
  if (threshold_out) *threshold_out = qs;
  return DYNA_OK;
}

// the pairs a < b of the node that survive `sim[sim < threshold] <- 0` with a non-zero similarity, row-major, as
// node-local 0-based indices plus (matches, length); weight = (double)matches / length
extern "C" int dyna_nw_plan_threshold_edges(dyna_nw_plan* p, const int32_t* members, int64_t n_members, double threshold,
                                            int64_t max_edges, int32_t* i_out, int32_t* j_out, uint32_t* matches_out,
                                            uint32_t* length_out, int64_t* n_edges_out, void* stream) {
  if (!p) return fail(DYNA_ERR_INVALID, "null plan");
  if (threshold != threshold) return fail(DYNA_ERR_INVALID, "dyna_nw_plan_threshold_edges: threshold is NaN");
  DYNA_TRY(use_device(p->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  p->last_stream = st;
  WorkStreamGuard work_stream_guard(st);
  if (n_edges_out) *n_edges_out = 0;
  DevBuf<int32_t> d_members;
  NwNode nd;
  DYNA_TRY(nw_make_node(p, members, n_members, d_members, st, &nd, "dyna_nw_plan_threshold_edges"));
  if (nd.n_node < 2 || p->pairs <= 0) {
    DYNA_CUDA(cudaStreamSynchronize(st));
    return DYNA_OK;
  }
  DevBuf<unsigned long long> rc, ro, tot;
  DYNA_TRY(rc.alloc((size_t)nd.n_node));
  DYNA_TRY(ro.alloc((size_t)nd.n_node));
  DYNA_TRY(tot.alloc(1));
  DYNA_TRY(launch_nw_edges_count(nd, threshold, rc.p, ro.p, tot.p, st));
  unsigned long long total = 0;
  DYNA_CUDA(cudaMemcpyAsync(&total, tot.p, sizeof total, cudaMemcpyDeviceToHost, st));
  DYNA_CUDA(cudaStreamSynchronize(st));
  if (n_edges_out) *n_edges_out = (int64_t)total;
  if ((int64_t)total > max_edges)
    return fail(DYNA_ERR_INVALID, "edge buffer too small: %lld edges, capacity %lld", (long long)total, (long long)max_edges);
  if (total == 0) return DYNA_OK;
  if (!i_out || !j_out || !matches_out || !length_out) return fail(DYNA_ERR_INVALID, "dyna_nw_plan_threshold_edges: null output");
  DevBuf<int32_t> di, dj;
  DevBuf<uint32_t> dm, dl;
  DYNA_TRY(di.alloc((size_t)total));
  DYNA_TRY(dj.alloc((size_t)total));
  DYNA_TRY(dm.alloc((size_t)total));
  DYNA_TRY(dl.alloc((size_t)total));
  DYNA_TRY(launch_nw_edges_fill(nd, threshold, ro.p, di.p, dj.p, dm.p, dl.p, st));
  DYNA_CUDA(cudaMemcpyAsync(i_out, di.p, sizeof(int32_t) * (size_t)total, cudaMemcpyDeviceToHost, st));
  DYNA_CUDA(cudaMemcpyAsync(j_out, dj.p, sizeof(int32_t) * (size_t)total, cudaMemcpyDeviceToHost, st));
  DYNA_CUDA(cudaMemcpyAsync(matches_out, dm.p, sizeof(uint32_t) * (size_t)total, cudaMemcpyDeviceToHost, st));
  DYNA_CUDA(cudaMemcpyAsync(length_out, dl.p, sizeof(uint32_t) * (size_t)total, cudaMemcpyDeviceToHost, st));
  DYNA_CUDA(cudaStreamSynchronize(st));
  return DYNA_OK;
}

// The planner's decisions without a device (host logic only; tests/test_nw_planner_host.py checks that the units of every
// class together cover each pair of the row range exactly once).
extern "C" dyna_nw_plan* dyna_nw_plan_layout(const uint8_t* residues, const int64_t* offsets, int64_t n, const char* matrix_name,
                                             int gap_open, int gap_ext, int64_t row_begin, int64_t row_end) {
  return dyna_nw_plan_create(residues, offsets, n, matrix_name, gap_open, gap_ext, row_begin, row_end, kNwLayoutOnly);
}
extern "C" int64_t dyna_nw_plan_unit_count(const dyna_nw_plan* p) {
  int64_t c = 0;
  if (p) for (const auto& cl : p->classes) c += (int64_t)cl->units.size();
  return c;
}
// out: 6 values per unit -- kernel kind (NwClass::kind), strip height R, row, second row (two-rows kinds; else -1),
// first column, column count
extern "C" int dyna_nw_plan_export_units(const dyna_nw_plan* p, int32_t* out) {
  if (!p || !out) return fail(DYNA_ERR_INVALID, "dyna_nw_plan_export_units: null argument");
  for (const auto& cl : p->classes) {
    const bool two_rows = (cl->kind >= 8 && cl->kind <= 10) || cl->kind == 12;
    for (const NwUnit& un : cl->units) {
      *out++ = cl->kind;
      *out++ = cl->R;
      *out++ = un.row;
      *out++ = two_rows ? un.row + (un.j_count >> 16) : -1;
      *out++ = un.j_begin;
      *out++ = two_rows ? (un.j_count & 0xFFFF) : un.j_count;
    }
  }
  return DYNA_OK;
}

extern "C" int64_t dyna_nw_plan_pairs(const dyna_nw_plan* p) { return p ? p->pairs : 0; }
extern "C" int64_t dyna_nw_plan_cells(const dyna_nw_plan* p) { return p ? p->cells : 0; }
extern "C" int dyna_nw_plan_launches(const dyna_nw_plan* p) { return p ? p->launches : 0; }
extern "C" void dyna_nw_plan_destroy(dyna_nw_plan* p) {
  if (!p) return;
  if (p->device != kNwLayoutOnly) cudaSetDevice(p->device);
  delete p;
}

// =====================================================================================================
// misc
// =====================================================================================================
extern "C" const char* dyna_last_error(void) { return err_slot().c_str(); }
extern "C" int dyna_version(void) { return 100; }
extern "C" int dyna_device_count(void) {
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return count;
}
// Device memory of finished calls stays cached in the stream-ordered pool (see common.cuh); this hands it back to
// the driver, e.g. before another process or another phase needs the device's memory.
extern "C" int dyna_release_cached_memory(int device) {
  DYNA_TRY(use_device(device));
  DYNA_CUDA(cudaDeviceSynchronize());
  cudaMemPool_t pool;
  DYNA_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
  DYNA_CUDA(cudaMemPoolTrimTo(pool, 0));
  return DYNA_OK;
}
extern "C" int dyna_set_device(int device) {
  DYNA_TRY(use_device(device));
  g_device = device;
  return DYNA_OK;
}

extern "C" int dyna_partition_rows(int64_t n, const int64_t* weights, int include_diagonal, int nshards,
                                   int64_t* bounds_out) {
  if (n < 0 || nshards <= 0 || !bounds_out) return fail(DYNA_ERR_INVALID, "dyna_partition_rows: bad arguments");
  // work of row i = w_i * sum_{j >= i (or > i)} w_j ; cut the prefix sums at equal shares
  std::vector<long double> rowwork((size_t)n);
  long double suffix = 0, total = 0;
  for (int64_t i = n - 1; i >= 0; --i) {
    const long double wi = weights ? (long double)weights[i] : 1.0L;
    const long double others = include_diagonal ? suffix + wi : suffix;
    rowwork[(size_t)i] = wi * others + 1e-9L;  // epsilon keeps zero-work rows ordered
    suffix += wi;
  }
  for (int64_t i = 0; i < n; ++i) total += rowwork[(size_t)i];
  bounds_out[0] = 0;
  long double acc = 0;
  int64_t i = 0;
  for (int s = 1; s < nshards; ++s) {
    const long double target = total * (long double)s / (long double)nshards;
    while (i < n && acc + rowwork[(size_t)i] * 0.5L < target) acc += rowwork[(size_t)i++];
    bounds_out[s] = i;
  }
  bounds_out[nshards] = n;
  return DYNA_OK;
}

extern "C" int dyna_hashfamily_seeds(uint32_t seed, int n_hash, uint32_t* seeds_out) {
  if (n_hash <= 0) return fail(DYNA_ERR_INVALID, "Number of hash functions must be positive");
  hashfamily_seeds(seed, n_hash, seeds_out);
  return DYNA_OK;
}
extern "C" uint32_t dyna_random_seed(void) { return std::random_device{}(); }

extern "C" int dyna_substitution_matrix(const char* name, int8_t* out576) {
  const int t = name ? find_table(name) : -1;
  if (t < 0) return fail(DYNA_ERR_INVALID, "Invalid substitution matrix name: %s", name ? name : "(null)");
  memcpy(out576, tables().full[t], 576);
  return DYNA_OK;
}
extern "C" void dyna_aa_index_table(int8_t* out256) { memcpy(out256, tables().aa, 256); }

// =====================================================================================================
// MinHash host entry points
// =====================================================================================================
namespace {
}  // namespace

static int plan_error_code_mh() { return err_code_slot() ? err_code_slot() : DYNA_ERR_CUDA; }

extern "C" int dyna_mh_signatures_murmur3(const uint8_t* residues, const int64_t* offsets, int64_t n, int k,
                                          const uint32_t* seeds, int n_hash, uint32_t* sig_out) {
  DYNA_TRY(check_mh_args(n, k, n_hash));
  dyna_mh_plan* p = dyna_mh_plan_create(n, n_hash, 0, 0, g_device);
  if (!p) return plan_error_code_mh();
  int rc = dyna_mh_plan_upload_sequences(p, residues, offsets, k, seeds, nullptr);
  if (rc == DYNA_OK) rc = dyna_mh_plan_run_signatures(p, nullptr);
  if (rc == DYNA_OK) rc = dyna_mh_plan_fetch_signatures(p, sig_out, nullptr);
  dyna_mh_plan_destroy(p);
  return rc;
}

extern "C" int dyna_mh_signatures_linear(const int32_t* ranks, const int64_t* rank_offsets, int64_t n, const int64_t* a,
                                         const int64_t* b, int64_t m, int n_hash, uint32_t* sig_out) {
  if (n <= 0) return fail(DYNA_ERR_INVALID, "Input sequences vector cannot be empty");
  if (n_hash < 1) return fail(DYNA_ERR_INVALID, "Number of hash functions must be positive");  // R/minHash.R:82
  if (m < 2) return fail(DYNA_ERR_INVALID, "Maximum value must be at least 2");                 // R/minHash.R:83
  if (m >= (1ll << 31)) return fail(DYNA_ERR_UNSUPPORTED, "vocabulary larger than 2^31-1 is not supported");
  for (int h = 0; h < n_hash; ++h)
    if (a[h] < 0 || b[h] < 0 || a[h] >= (1ll << 31) || b[h] >= (1ll << 31))
      return fail(DYNA_ERR_UNSUPPORTED, "hash parameters must be in [0, 2^31)");
  DYNA_TRY(check_offsets(rank_offsets, n, "dyna_mh_signatures_linear"));
  DYNA_TRY(use_device(g_device));
  const int64_t total = rank_offsets[n];
  for (int64_t q = 0; q < total; ++q)
    if (ranks[q] < 1) return fail(DYNA_ERR_INVALID, "vocabulary ranks are 1-based");
  DevBuf<int32_t> d_ranks;
  DevBuf<int64_t> d_roff, d_a, d_b;
  DevBuf<uint32_t> d_sig;
  DYNA_TRY(d_ranks.alloc((size_t)total));
  DYNA_TRY(d_roff.alloc((size_t)n + 1));
  DYNA_TRY(d_a.alloc((size_t)n_hash));
  DYNA_TRY(d_b.alloc((size_t)n_hash));
  DYNA_TRY(d_sig.alloc((size_t)n * n_hash));
  DYNA_CUDA(cudaMemcpy(d_ranks.p, ranks, sizeof(int32_t) * (size_t)total, cudaMemcpyHostToDevice));
  DYNA_CUDA(cudaMemcpy(d_roff.p, rank_offsets, sizeof(int64_t) * (size_t)(n + 1), cudaMemcpyHostToDevice));
  DYNA_CUDA(cudaMemcpy(d_a.p, a, sizeof(int64_t) * (size_t)n_hash, cudaMemcpyHostToDevice));
  DYNA_CUDA(cudaMemcpy(d_b.p, b, sizeof(int64_t) * (size_t)n_hash, cudaMemcpyHostToDevice));
  DYNA_TRY(launch_mh_signature_linear(d_ranks.p, d_roff.p, n, d_a.p, d_b.p, m, n_hash, d_sig.p, nullptr));
  DYNA_CUDA(cudaMemcpy(sig_out, d_sig.p, sizeof(uint32_t) * (size_t)n * n_hash, cudaMemcpyDeviceToHost));
  return DYNA_OK;
}

extern "C" int dyna_mh_match_counts(const uint32_t* sig, int64_t n, int n_hash, int64_t row_begin, int64_t row_end,
                                    uint16_t* counts_tri_out) {
  if (n <= 0) return fail(DYNA_ERR_INVALID, "Input sequences vector cannot be empty");
  if (n_hash <= 0) return fail(DYNA_ERR_INVALID, "Number of hash functions must be positive");
  if (n_hash > 65535) return fail(DYNA_ERR_UNSUPPORTED, "n_hash > 65535 is not supported (match counts are 16-bit)");
  dyna_mh_plan* p = dyna_mh_plan_create(n, n_hash, row_begin, row_end, g_device);
  if (!p) return err_code_slot() ? err_code_slot() : DYNA_ERR_CUDA;
  int rc = dyna_mh_plan_upload_signatures(p, sig, nullptr);
  if (rc == DYNA_OK) rc = dyna_mh_plan_run_match(p, nullptr);
  if (rc == DYNA_OK) rc = dyna_mh_plan_fetch_counts(p, counts_tri_out, nullptr);
  dyna_mh_plan_destroy(p);
  return rc;
}

namespace {

// value table for the expansion: reference arithmetic per kind
void mh_value_table(int n_hash, int kind, std::vector<double>& table, double* diag) {
  table.resize((size_t)n_hash + 1);
  for (int c = 0; c <= n_hash; ++c) {
    if (kind == DYNA_MH_DISTANCE) {
      // R: 1 - mean(logical): long double accumulate and divide, then double (R/minHash.R:174-175)
      const double sim = (double)((long double)c / (long double)n_hash);
      table[(size_t)c] = 1.0 - sim;
    } else {
      table[(size_t)c] = static_cast<double>(c) / n_hash;  // src/minHash.cpp:174
    }
  }
  *diag = (kind == DYNA_MH_DISTANCE) ? 0.0 : 1.0;
}

// The n x n double matrix from `gpus` devices of this process (gpus == 1 included).  Device g matches its pair-
// balanced row block; after all devices have finished, device g expands COLUMNS [n*g/G, n*(g+1)/G) -- reading every
// slab those columns need out of the owning device's memory (peer loads over NVLink) -- and copies its block, a
// contiguous range of the caller's matrix, to the host.  prepare(plan) must leave signatures resident on the plan's
// device.  No slab is staged through the host and the host does no arithmetic.
template <class Prepare>
int mh_matrix_gpus(int64_t n, int n_hash, int kind, int gpus, double* out, Prepare prepare) {
  if (gpus > kMaxSlabs) gpus = kMaxSlabs;
  if (gpus > 1 && !enable_peer_mesh(gpus)) gpus = 1;  // no peer access between the devices: one device does it all
  const int dev0 = g_device;
  std::vector<int64_t> bounds((size_t)gpus + 1), cols;
  DYNA_TRY(dyna_partition_rows(n, nullptr, 0, gpus, bounds.data()));
  column_blocks(n, gpus, cols);
  std::vector<double> table;
  double diag;
  mh_value_table(n_hash, kind, table, &diag);
  std::vector<dyna_mh_plan*> plans((size_t)gpus, nullptr);
  auto device_of = [&](int g) { return gpus == 1 ? dev0 : g; };
  int rc = for_each_gpu(gpus, [&](int g) {
    DYNA_TRY(use_device(device_of(g)));
    const int64_t cb = cols[(size_t)g + 1] - cols[(size_t)g];
    DYNA_TRY(check_device_fits(8.0 * (double)n * (double)cb + 2.0 * (double)(tri_strict_rows(n, bounds[(size_t)g + 1]) -
                                                                          tri_strict_rows(n, bounds[(size_t)g])),
                               device_of(g), "the n x n double matrix"));
    plans[(size_t)g] = dyna_mh_plan_create(n, n_hash, bounds[(size_t)g], bounds[(size_t)g + 1], device_of(g));
    if (!plans[(size_t)g]) return plan_error_code_mh();
    DYNA_TRY(prepare(plans[(size_t)g]));
    DYNA_TRY(dyna_mh_plan_run_match(plans[(size_t)g], nullptr));
    DYNA_CUDA(cudaStreamSynchronize(nullptr));  // this slab is complete before any device starts to read it
    return (int)DYNA_OK;
  });
  if (rc == DYNA_OK) {
    TriSlabs slabs{};
    slabs.nslabs = gpus;
    for (int g = 0; g < gpus; ++g) {
      slabs.row_begin[g] = bounds[(size_t)g];
      slabs.a[g] = plans[(size_t)g]->counts.p;
      slabs.b[g] = nullptr;
    }
    slabs.row_begin[gpus] = n;
    rc = for_each_gpu(gpus, [&](int g) {
      DYNA_TRY(use_device(device_of(g)));
      const int64_t c0 = cols[(size_t)g], c1 = cols[(size_t)g + 1];
      if (c1 <= c0) return (int)DYNA_OK;
      DevBuf<double> d_table, d_block;
      DYNA_TRY(d_table.alloc(table.size()));
      DYNA_TRY(d_block.alloc((size_t)n * (size_t)(c1 - c0)));
      DYNA_CUDA(cudaMemcpyAsync(d_table.p, table.data(), sizeof(double) * table.size(), cudaMemcpyHostToDevice, nullptr));
      DYNA_TRY(launch_mh_expand_block(slabs, n, c0, c1, d_table.p, diag, d_block.p, nullptr));
      DYNA_CUDA(cudaMemcpyAsync(out + c0 * n, d_block.p, sizeof(double) * (size_t)n * (size_t)(c1 - c0), cudaMemcpyDeviceToHost, nullptr));
      DYNA_CUDA(cudaStreamSynchronize(nullptr));
      return (int)DYNA_OK;
    });
  }
  for (int g = 0; g < gpus; ++g)
    if (plans[(size_t)g]) dyna_mh_plan_destroy(plans[(size_t)g]);
  if (gpus > 1) cudaSetDevice(dev0);
  return rc;
}

}  // namespace

extern "C" int dyna_mh_match_matrix(const uint32_t* sig, int64_t n, int n_hash, int kind, double* out, int n_gpus) {
  if (n <= 0) return fail(DYNA_ERR_INVALID, "Input sequences vector cannot be empty");
  if (n_hash <= 0) return fail(DYNA_ERR_INVALID, "Number of hash functions must be positive");
  if (n_hash > 65535) return fail(DYNA_ERR_UNSUPPORTED, "n_hash > 65535 is not supported (match counts are 16-bit)");
  if (kind != DYNA_MH_SIMILARITY && kind != DYNA_MH_DISTANCE) return fail(DYNA_ERR_INVALID, "unknown matrix kind %d", kind);
  int gpus = 1;
  DYNA_TRY(resolve_gpus(n_gpus == 0 ? 1 : n_gpus, &gpus));
  if (n < 2 * 128 * gpus) gpus = 1;
  return mh_matrix_gpus(n, n_hash, kind, gpus, out,
                        [&](dyna_mh_plan* p) { return dyna_mh_plan_upload_signatures(p, sig, nullptr); });
}

extern "C" int dyna_similarityMH(const uint8_t* residues, const int64_t* offsets, int64_t n, int k, int n_hash,
                                 const uint32_t* seeds, double* out, int n_gpus) {
  DYNA_TRY(check_mh_args(n, k, n_hash));
  DYNA_TRY(check_offsets(offsets, n, "dyna_similarityMH"));
  std::vector<uint32_t> own;
  if (!seeds) {  // reference behaviour: HashFamily(n_hash) seeded from std::random_device (src/minHash.cpp:73,137)
    own.resize((size_t)n_hash);
    hashfamily_seeds(dyna_random_seed(), n_hash, own.data());
    seeds = own.data();
  }
  int gpus = 1;
  DYNA_TRY(resolve_gpus(n_gpus == 0 ? 1 : n_gpus, &gpus));
  if (n < 2 * 128 * gpus) gpus = 1;
  auto prepare = [&](dyna_mh_plan* p) {
    int rc = dyna_mh_plan_upload_sequences(p, residues, offsets, k, seeds, nullptr);
    if (rc == DYNA_OK) rc = dyna_mh_plan_run_signatures(p, nullptr);  // every GPU rebuilds all signatures (< 1 ms)
    return rc;
  };
  return mh_matrix_gpus(n, n_hash, DYNA_MH_SIMILARITY, gpus, out, prepare);
}

// =====================================================================================================
// NW host entry points
// =====================================================================================================
extern "C" int dyna_nw_pair_stats(const uint8_t* residues, const int64_t* offsets, int64_t n, const char* matrix_name,
                                  int gap_open, int gap_ext, int64_t row_begin, int64_t row_end, uint32_t* matches_out,
                                  uint32_t* length_out) {
  PhaseTimer tm;
  dyna_nw_plan* p = dyna_nw_plan_create(residues, offsets, n, matrix_name, gap_open, gap_ext, row_begin, row_end, g_device);
  tm.lap("nw plan create");
  if (!p) return err_code_slot() ? err_code_slot() : DYNA_ERR_CUDA;
  int rc = dyna_nw_plan_run(p, nullptr);
  if (tm.on) cudaDeviceSynchronize();
  tm.lap("nw kernels");
  if (rc == DYNA_OK) rc = dyna_nw_plan_fetch(p, matches_out, length_out, nullptr);
  tm.lap("nw fetch (D2H)");
  dyna_nw_plan_destroy(p);
  tm.lap("nw plan destroy");
  return rc;
}

// dyna_nw_pair_stats with the narrow result form of dyna_nw_plan_fetch_packed8 (2 bytes per pair)
namespace {
int plan_error_code() { return err_code_slot() ? err_code_slot() : DYNA_ERR_CUDA; }  // code of the failed plan creation
}  // namespace

extern "C" int dyna_nw_pair_stats8(const uint8_t* residues, const int64_t* offsets, int64_t n, const char* matrix_name,
                                   int gap_open, int gap_ext, int64_t row_begin, int64_t row_end, uint8_t* matches8_out,
                                   uint8_t* length8_out) {
  // Large row ranges go in sixteen row blocks of equal DP cells, pipelined: while block k is aligned on one stream, block k-1 is
  // narrowed to bytes and copied to the host on another (the 100,000-peptide target: 0.29 s of kernels and 10 GB = 0.2 s
  // of PCIe no longer add up), and the host builds the plan of block k+1.  At most three blocks hold device memory.
  const int64_t total_pairs = (n >= 0 && row_begin >= 0 && row_end <= n && row_begin <= row_end)
                                  ? tri_diag_rows(n, row_end) - tri_diag_rows(n, row_begin) : 0;
  int blocks = total_pairs >= (64ll << 20) ? 16 : 1;  // 100,000 peptides: 1 / 4 / 8 / 16 blocks 0.53 / 0.40 / 0.35 / 0.325 s
  if (const char* e = getenv("DYNA_NW_STATS8_BLOCKS")) blocks = std::max(1, std::min(64, atoi(e)));
  if (blocks == 1 || check_offsets(offsets, n, "dyna_nw_pair_stats8") != DYNA_OK) {
    dyna_nw_plan* p = dyna_nw_plan_create(residues, offsets, n, matrix_name, gap_open, gap_ext, row_begin, row_end, g_device);
    if (!p) return err_code_slot() ? err_code_slot() : DYNA_ERR_CUDA;
    int rc = dyna_nw_plan_run(p, nullptr);
    if (rc == DYNA_OK) rc = dyna_nw_plan_fetch_packed8(p, matches8_out, length8_out, nullptr);
    dyna_nw_plan_destroy(p);
    return rc;
  }
  // row bounds: equal shares of sum_i len_i * (sum_{j >= i} len_j)  (+ one per pair, so that empty sequences count)
  std::vector<int64_t> bounds((size_t)blocks + 1, row_end);
  {
    std::vector<double> suffix((size_t)n + 1, 0.0);
    for (int64_t i = n - 1; i >= 0; --i) suffix[(size_t)i] = suffix[(size_t)i + 1] + (double)(offsets[i + 1] - offsets[i]) + 1.0;
    double total = 0.0;
    for (int64_t i = row_begin; i < row_end; ++i) total += ((double)(offsets[i + 1] - offsets[i]) + 1.0) * suffix[(size_t)i];
    double acc = 0.0;
    int b = 1;
    bounds[0] = row_begin;
    for (int64_t i = row_begin; i < row_end && b < blocks; ++i) {
      acc += ((double)(offsets[i + 1] - offsets[i]) + 1.0) * suffix[(size_t)i];
      while (b < blocks && acc >= total * b / blocks) bounds[(size_t)b++] = i + 1;
    }
  }
  DYNA_TRY(use_device(g_device));
  struct Block {
    dyna_nw_plan* plan = nullptr;
    DevBuf<uint8_t> m8, l8;
    cudaEvent_t done = nullptr;
  };
  std::vector<Block> blk((size_t)blocks);
  cudaStream_t sc = nullptr, sd = nullptr;
  cudaEvent_t ran = nullptr, allocated = nullptr;
  int rc = DYNA_OK;
  auto cu = [&](cudaError_t e) {
    if (e != cudaSuccess && rc == DYNA_OK) rc = fail(DYNA_ERR_CUDA, "DynaAlign CUDA: dyna_nw_pair_stats8: %s", cudaGetErrorString(e));
    return e == cudaSuccess;
  };
  auto retire = [&](Block& b) {  // wait for the block's copies, then hand its device memory back
    if (b.done) {  // recorded behind the block's last use of its buffers (kernels -> event -> pack -> copies)
      cudaEventSynchronize(b.done);
      cudaEventDestroy(b.done);
      b.done = nullptr;
      if (b.plan) b.plan->owner_synced = true;  // do not wait for the LATER blocks queued on the compute stream
    }
    b.m8.release();
    b.l8.release();
    if (b.plan) dyna_nw_plan_destroy(b.plan);
    b.plan = nullptr;
  };
  cu(cudaStreamCreateWithFlags(&sc, cudaStreamNonBlocking));
  cu(cudaStreamCreateWithFlags(&sd, cudaStreamNonBlocking));
  cu(cudaEventCreateWithFlags(&ran, cudaEventDisableTiming));
  cu(cudaEventCreateWithFlags(&allocated, cudaEventDisableTiming));
  // DevBuf allocates in the order of the legacy stream; the two streams here do not synchronise with it, so every use of
  // fresh memory on them is put behind an event recorded on the legacy stream after the allocation (a pool that has to
  // map new memory does so in stream order: without this the copy engine read unmapped pages once the blocks were
  // smaller than what the pool held cached -- an illegal access three blocks later)
  auto after_alloc = [&](cudaStream_t s) {
    cu(cudaEventRecord(allocated, cudaStreamLegacy));
    cu(cudaStreamWaitEvent(s, allocated, 0));
  };
  const int64_t first = tri_diag_rows(n, row_begin);
  for (int k = 0; k < blocks && rc == DYNA_OK; ++k) {
    const int64_t rb = bounds[(size_t)k], re = bounds[(size_t)k + 1];
    if (re <= rb) continue;
    if (k >= 3 && !getenv("DYNA_NW_STATS8_KEEP")) retire(blk[(size_t)k - 3]);
    Block& b = blk[(size_t)k];
    b.plan = dyna_nw_plan_create(residues, offsets, n, matrix_name, gap_open, gap_ext, rb, re, g_device);
    if (!b.plan) { rc = plan_error_code(); break; }
    if (2 * (int64_t)b.plan->max_cols > 255) {
      rc = fail(DYNA_ERR_UNSUPPORTED, "dyna_nw_pair_stats8: alignment lengths up to %lld do not fit one byte", 2 * (long long)b.plan->max_cols);
      break;
    }
    after_alloc(sc);
    if ((rc = dyna_nw_plan_run(b.plan, sc)) != DYNA_OK) {
      if (getenv("DYNA_TIMING")) fprintf(stderr, "dyna_nw_pair_stats8: block %d of %d (rows %lld..%lld) failed\n", k, blocks, (long long)rb, (long long)re);
      break;
    }
    const int64_t bp = b.plan->pairs, at = tri_diag_rows(n, rb) - first;
    if (bp <= 0) continue;
    if (b.m8.alloc((size_t)bp) || b.l8.alloc((size_t)bp)) { rc = DYNA_ERR_CUDA; break; }
    after_alloc(sd);
    cu(cudaEventRecord(ran, sc));
    cu(cudaStreamWaitEvent(sd, ran, 0));
    if ((rc = launch_nw_pack8(b.plan->matches.p, b.plan->length.p, bp, b.m8.p, b.l8.p, sd)) != DYNA_OK) break;
    cu(cudaMemcpyAsync(matches8_out + at, b.m8.p, (size_t)bp, cudaMemcpyDeviceToHost, sd));
    cu(cudaMemcpyAsync(length8_out + at, b.l8.p, (size_t)bp, cudaMemcpyDeviceToHost, sd));
    cu(cudaEventCreateWithFlags(&b.done, cudaEventDisableTiming));
    cu(cudaEventRecord(b.done, sd));
    b.plan->last_stream = sc;
  }
  if (sc) cudaStreamSynchronize(sc);
  if (sd) cudaStreamSynchronize(sd);
  for (auto& b : blk) retire(b);
  if (ran) cudaEventDestroy(ran);
  if (allocated) cudaEventDestroy(allocated);
  if (sc) cudaStreamDestroy(sc);
  if (sd) cudaStreamDestroy(sd);
  return rc;
}

extern "C" int dyna_similarityNW(const uint8_t* residues, const int64_t* offsets, int64_t n, const char* matrix_name,
                                 int gap_open, int gap_ext, double* out, int n_gpus) {
  if (n == 0) {  // the reference returns a 0 x 0 matrix, but still rejects an unknown matrix name first
    if (!matrix_name || find_table(matrix_name) < 0)
      return fail(DYNA_ERR_INVALID, "Invalid substitution matrix name: %s", matrix_name ? matrix_name : "(null)");
    return DYNA_OK;
  }
  int gpus = 1;
  if (!matrix_name || find_table(matrix_name) < 0)
    return fail(DYNA_ERR_INVALID, "Invalid substitution matrix name: %s", matrix_name ? matrix_name : "(null)");
  DYNA_TRY(check_offsets(offsets, n, "dyna_similarityNW"));
  DYNA_TRY(validate_residues(residues, offsets, n));
  DYNA_TRY(resolve_gpus(n_gpus == 0 ? 1 : n_gpus, &gpus));
  if (n < 64 * gpus) gpus = 1;
  // Row blocks balanced by DP cells (len_i * len_j), one host thread per device.  When every device has finished its
  // block, device g expands columns [n*g/G, n*(g+1)/G) of the result -- gathering the (matches, length) entries those
  // columns need from the other devices' slabs by peer loads over NVLink -- and copies its block, a contiguous range
  // of the caller's column-major matrix, to the host (src/pairwiseSeqAlign.cpp:349-350 writes both triangles).
  if (gpus > kMaxSlabs) gpus = kMaxSlabs;
  if (gpus > 1 && !enable_peer_mesh(gpus)) gpus = 1;
  const int dev0 = g_device;
  auto device_of = [&](int g) { return gpus == 1 ? dev0 : g; };
  std::vector<int64_t> lens((size_t)n), bounds((size_t)gpus + 1), cols;
  for (int64_t i = 0; i < n; ++i) lens[(size_t)i] = offsets[i + 1] - offsets[i];
  DYNA_TRY(dyna_partition_rows(n, lens.data(), 1, gpus, bounds.data()));
  column_blocks(n, gpus, cols);
  std::vector<dyna_nw_plan*> plans((size_t)gpus, nullptr);
  int rc = for_each_gpu(gpus, [&](int g) {
    DYNA_TRY(use_device(device_of(g)));
    const int64_t cb = cols[(size_t)g + 1] - cols[(size_t)g];
    DYNA_TRY(check_device_fits(8.0 * (double)n * (double)cb + 8.0 * (double)(tri_diag_rows(n, bounds[(size_t)g + 1]) -
                                                                          tri_diag_rows(n, bounds[(size_t)g])),
                               device_of(g), "similarityNW: the n x n double matrix"));
    plans[(size_t)g] = dyna_nw_plan_create(residues, offsets, n, matrix_name, gap_open, gap_ext, bounds[(size_t)g],
                                           bounds[(size_t)g + 1], device_of(g));
    if (!plans[(size_t)g]) return plan_error_code();
    DYNA_TRY(dyna_nw_plan_run(plans[(size_t)g], nullptr));
    DYNA_CUDA(cudaStreamSynchronize(nullptr));  // this slab is complete before any device starts to read it
    return (int)DYNA_OK;
  });
  if (rc == DYNA_OK) {
    TriSlabs slabs{};
    slabs.nslabs = gpus;
    for (int g = 0; g < gpus; ++g) {
      slabs.row_begin[g] = bounds[(size_t)g];
      slabs.a[g] = plans[(size_t)g]->matches.p;
      slabs.b[g] = plans[(size_t)g]->length.p;
    }
    slabs.row_begin[gpus] = n;
    rc = for_each_gpu(gpus, [&](int g) {
      DYNA_TRY(use_device(device_of(g)));
      const int64_t c0 = cols[(size_t)g], c1 = cols[(size_t)g + 1];
      if (c1 <= c0) return (int)DYNA_OK;
      DevBuf<double> d_block;
      DYNA_TRY(d_block.alloc((size_t)n * (size_t)(c1 - c0)));
      DYNA_TRY(launch_nw_expand_block(slabs, n, c0, c1, d_block.p, nullptr));
      DYNA_CUDA(cudaMemcpyAsync(out + c0 * n, d_block.p, sizeof(double) * (size_t)n * (size_t)(c1 - c0), cudaMemcpyDeviceToHost, nullptr));
      DYNA_CUDA(cudaStreamSynchronize(nullptr));
      return (int)DYNA_OK;
    });
  }
  for (int g = 0; g < gpus; ++g)
    if (plans[(size_t)g]) dyna_nw_plan_destroy(plans[(size_t)g]);
  if (gpus > 1) cudaSetDevice(dev0);
  return rc;
}
