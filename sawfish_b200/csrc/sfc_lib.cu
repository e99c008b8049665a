#include <chrono>
// sfc_lib.cu — host side of libsawfish_cuda.so: the C ABI declared in include/sawfish_cuda.h.
//
// One sfc_ctx owns one CUDA stream and every device buffer of one GPU.  A batch goes
//   upload (H2D of the ASCII arena + pair table, K3 pack)  ->  run (K1 score-only / K2 backtrace kernels)
//   ->  download (scores, status, run-length ops).
// There is no CPU fallback anywhere in this file: if CUDA is unavailable every entry point fails.
#include "../../include/sawfish_cuda.h"

#include <algorithm>
#include <climits>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "sfc_device.cuh"
#include "sfc_k_align.cuh"
#include "sfc_k_align2.cuh"
#include "sfc_k_bio.cuh"
#include "sfc_k_bio_score.cuh"
#include "sfc_k_pack.cuh"
#include "sfc_k_score.cuh"

using namespace sfc;

namespace {

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
};

struct AdjPen {
  int metric, x, o1, e1, o2, e2, match;
};

}  // namespace

struct sfc_ctx {
  int device = 0;
  int sm_count = 0;
  size_t smem_optin = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[8] = {};
  cudaStream_t job_stream[6] = {};  // concurrent K2 jobs (one per size class / cluster size)
  cudaEvent_t job_ev[7] = {};
  std::string err;
  size_t ws_limit = 0;
  // device buffers
  DevBuf d_arena, d_nib, d_seqs, d_seqwoff, d_pairs, d_order, d_scores, d_status, d_misc, d_pool, d_outoff, d_outlen, d_ws;
  DevBuf d_pages, d_pgstack, d_pairpages;  // K2 v2: the page pool of the wavefront history, its free stack, pages expected per pair
  std::vector<uint32_t> h_pairpages;
  std::vector<int> h_pgstack;
  DevBuf d_b_arena, d_b_pairs, d_b_out, d_b_pool;  // K4 (bio aligner) staging; the slabs reuse d_ws
  // batch state
  size_t n_pairs = 0;
  bool uploaded = false, ran = false;
  // Free device memory as of the last cudaMemGetInfo, kept up to date with this context's own allocations.  The query itself
  // is not repeated per batch: with ~150 GB allocated it takes 1-3 ms and now and then 20-90 ms (measured on B200: one
  // config2 step in ten took 100-170 ms instead of 76 with the kernels waiting for the host), so it is redone only after an
  // allocation failed (another context / process took the memory).
  size_t free_cache = 0;
  bool free_valid = false, free_fresh = false;
  int want_cigar = 0;
  std::vector<DevPair> h_pairs;
  std::vector<SeqDesc> h_seqs;
  std::vector<uint32_t> h_seqwoff;
  std::vector<uint32_t> h_order;
  std::vector<uint8_t> h_cls;  // class code of every pair of the batch (classify_pairs)
  size_t cls_begin[16] = {};   // first position of every class in h_order
  bool cls_valid = false, cls_k1_ok = false;
  std::vector<int32_t> h_status;
  std::vector<unsigned long long> h_outoff;
  std::vector<uint32_t> h_outlen;
  std::vector<uint32_t> h_pool;
  unsigned long long pool_used = 0;
  sfc_batch_stats stats = {};
};

struct sfc_aligner {
  sfc_ctx* ctx;
  sfc_penalties pen;
  std::vector<uint8_t> text;
  // the last alignment, for the debug view (sfc_aligner_matching): run-length ops + the pattern they refer to
  std::vector<uint32_t> last_ops;
  std::vector<uint8_t> last_pattern;
};

namespace {

int fail(sfc_ctx* c, int code, const char* fmt, ...) {
  if (c) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    c->err = buf;
  }
  return code;
}

#define CU(call)                                                                                        \
  do {                                                                                                  \
    cudaError_t e_ = (call);                                                                            \
    if (e_ != cudaSuccess) return fail(c, SFC_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

int ensure(sfc_ctx* c, DevBuf& b, size_t bytes) {
  if (bytes <= b.cap) return 0;
  if (b.p) {
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaFree(b.p));
    c->free_cache += b.cap;
    b.p = nullptr;
    b.cap = 0;
  }
  size_t want = bytes + bytes / 4 + 256;
  cudaError_t e = cudaMalloc(&b.p, want);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    want = bytes;
    e = cudaMalloc(&b.p, want);
  }
  if (e != cudaSuccess) {
    b.p = nullptr;
    c->free_valid = false;  // somebody else holds memory we counted on: ask the driver again next time
    return fail(c, SFC_ERR_OOM, "cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
  }
  b.cap = want;
  c->free_cache = c->free_cache > want ? c->free_cache - want : 0;
  return 0;
}

// free device memory (see sfc_ctx::free_cache)
size_t device_free_bytes(sfc_ctx* c) {
  c->free_fresh = false;
  if (!c->free_valid) {
    size_t f = 0, t = 0;
    if (cudaMemGetInfo(&f, &t) != cudaSuccess) { (void)cudaGetLastError(); f = 0; }
    c->free_cache = f;
    c->free_valid = true;
    c->free_fresh = true;
  }
  return c->free_cache;
}

void release(DevBuf& b) {
  if (b.p) cudaFree(b.p);
  b.p = nullptr;
  b.cap = 0;
}

int adjust_penalties(const sfc_penalties* pen, AdjPen* a) {
  if (!pen || (pen->metric != 0 && pen->metric != 1) || pen->match > 0) return -1;
  const int m = pen->match;
  a->metric = pen->metric;
  a->match = m;
  if (m == 0) {
    a->x = pen->mismatch; a->o1 = pen->gap_open1; a->e1 = pen->gap_ext1; a->o2 = pen->gap_open2; a->e2 = pen->gap_ext2;
  } else {  // Eizenga's adjustment, as WFA2-lib applies it for match < 0
    a->x = 2 * pen->mismatch - 2 * m;
    a->o1 = 2 * pen->gap_open1; a->e1 = 2 * pen->gap_ext1 - m;
    a->o2 = 2 * pen->gap_open2; a->e2 = 2 * pen->gap_ext2 - m;
  }
  if (a->metric == 0) { a->o1 = 0; a->o2 = 0; a->e2 = 0; }
  if (a->x <= 0 || a->e1 <= 0) return -1;
  if (a->metric == 1 && (a->e2 <= 0 || a->o1 < 0 || a->o2 < 0)) return -1;
  if (a->x > 4096 || a->o1 + a->e1 > 4096 || a->o2 + a->e2 > 4096) return -1;
  return 0;
}

// [off, off + len) inside [0, arena_len) without wrapping in uint64
inline bool in_arena(uint64_t off, uint64_t len, uint64_t arena_len) { return off <= arena_len && len <= arena_len - off; }

long long gap_cost(const AdjPen& a, long long L) {
  if (L <= 0) return 0;
  if (a.metric == 0) return (long long)a.e1 * L;
  return std::min((long long)a.o1 + (long long)a.e1 * L, (long long)a.o2 + (long long)a.e2 * L);
}

double estimate_work(const AdjPen& a, double plen, double tlen, double pbf, double tbf);

void init_lut_host(uint8_t* lut) {
  memset(lut, 0xFF, 256);
  const char* up = "ACGTN";
  const char* lo = "acgtn";
  for (int i = 0; i < 5; ++i) { lut[(uint8_t)up[i]] = (uint8_t)i; lut[(uint8_t)lo[i]] = (uint8_t)(5 + i); }
}

// ---- K1 size classes (shared-memory ring) ----
struct K1Class { int wcap; int wpp; };
// finer classes = less shared memory per pair = more pairs resident per SM; classes run concurrently
const K1Class kK1Classes[] = {{256, 1}, {384, 2}, {512, 2}, {640, 2}, {768, 4}, {896, 4}, {1024, 4}, {1152, 4},
                              {1536, 8}, {2304, 8}, {3328, 8}, {4352, 8}};
constexpr int kNumK1Classes = 12;

template <int WPP>
int launch_k1(sfc_ctx* c, const K1Params& P, int groups_per_block, int grid, size_t smem, cudaStream_t st) {
  // always the device maximum: the attribute is per device (shared by every ctx / thread on it), so a smaller value set
  // by one launch must never undercut another launch in flight
  CU(cudaFuncSetAttribute(k1_score<WPP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->smem_optin));
  // ask for the full shared-memory carveout: occupancy of this kernel is set by shared memory, not by L1
  CU(cudaFuncSetAttribute(k1_score<WPP>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  k1_score<WPP><<<grid, groups_per_block * WPP * 32, smem, st>>>(P);
  CU(cudaGetLastError());
  return 0;
}

__global__ void int_peak_kernel(unsigned* out, int iters, unsigned seed) {
  unsigned a0 = threadIdx.x + seed, a1 = a0 * 3u + 1u, a2 = a0 * 5u + 2u, a3 = a0 * 7u + 3u;
  unsigned a4 = a0 ^ 0x9E3779B9u, a5 = a1 ^ 0x7F4A7C15u, a6 = a2 ^ 0x85EBCA6Bu, a7 = a3 ^ 0xC2B2AE35u;
  const unsigned b = blockIdx.x | 1u, d = seed | 3u;
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {  // 8 chains x 8 x (IADD + LOP3) per iteration
      a0 = (a0 + b) ^ d; a1 = (a1 + b) ^ d; a2 = (a2 + b) ^ d; a3 = (a3 + b) ^ d;
      a4 = (a4 + b) ^ d; a5 = (a5 + b) ^ d; a6 = (a6 + b) ^ d; a7 = (a7 + b) ^ d;
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
}

}  // namespace

// ============================================================================ context
extern "C" int sfc_abi_version(void) { return SFC_ABI_VERSION; }

extern "C" int sfc_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { (void)cudaGetLastError(); return 0; }
  return n;
}

extern "C" int sfc_create(int device_ordinal, sfc_ctx** out) {
  if (!out) return SFC_ERR_INVALID;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) { (void)cudaGetLastError(); return SFC_ERR_NO_DEVICE; }
  if (device_ordinal < 0 || device_ordinal >= n) return SFC_ERR_INVALID;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device_ordinal) != cudaSuccess) return SFC_ERR_NO_DEVICE;
  if (prop.major != 10) return SFC_ERR_NO_DEVICE;  // built for sm_100a only
  sfc_ctx* c = new (std::nothrow) sfc_ctx();
  if (!c) return SFC_ERR_OOM;
  c->device = device_ordinal;
  c->sm_count = prop.multiProcessorCount;
  c->smem_optin = prop.sharedMemPerBlockOptin;
  if (cudaSetDevice(device_ordinal) != cudaSuccess || cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete c;
    return SFC_ERR_CUDA;
  }
  for (auto& e : c->ev) {
    if (cudaEventCreate(&e) != cudaSuccess) { sfc_destroy(c); return SFC_ERR_CUDA; }
  }
  {
    // job streams 0-3 have the highest priority, 4-5 the lowest: the K2 cluster jobs go on the former, the persistent
    // single-CTA job on the latter.  Without it the block scheduler now and then starts the single-CTA job first, its
    // CTAs take every slot for the whole step, and the clusters (which need 2-8 free slots in one GPC at once) only start
    // when it ends: sporadic steps of 150 ms instead of 76 (same per-CTA cycle counts, serialised jobs).
    int least = 0, greatest = 0;
    if (cudaDeviceGetStreamPriorityRange(&least, &greatest) != cudaSuccess) { least = greatest = 0; }
    for (int q = 0; q < 6; ++q) {
      if (cudaStreamCreateWithPriority(&c->job_stream[q], cudaStreamNonBlocking, q < 4 ? greatest : least) != cudaSuccess) { sfc_destroy(c); return SFC_ERR_CUDA; }
    }
  }
  for (auto& e : c->job_ev) {
    if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) { sfc_destroy(c); return SFC_ERR_CUDA; }
  }
  uint8_t lut[256];
  init_lut_host(lut);
  if (cudaMemcpyToSymbol(c_nib_lut, lut, 256) != cudaSuccess) { sfc_destroy(c); return SFC_ERR_CUDA; }
  *out = c;
  return SFC_OK;
}

extern "C" void sfc_destroy(sfc_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  for (DevBuf* b : {&c->d_arena, &c->d_nib, &c->d_seqs, &c->d_seqwoff, &c->d_pairs, &c->d_order, &c->d_scores, &c->d_status,
                    &c->d_misc, &c->d_pool, &c->d_outoff, &c->d_outlen, &c->d_ws, &c->d_pages, &c->d_pgstack, &c->d_pairpages, &c->d_b_arena, &c->d_b_pairs, &c->d_b_out,
                    &c->d_b_pool})
    release(*b);
  for (auto& e : c->ev) if (e) cudaEventDestroy(e);
  for (auto& e : c->job_ev) if (e) cudaEventDestroy(e);
  for (auto& st : c->job_stream) if (st) cudaStreamDestroy(st);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

extern "C" const char* sfc_last_error(const sfc_ctx* c) { return c ? c->err.c_str() : "null context"; }

extern "C" int sfc_set_workspace_limit(sfc_ctx* c, size_t bytes) {
  if (!c) return SFC_ERR_INVALID;
  c->ws_limit = bytes;
  return SFC_OK;
}

// ============================================================================ upload
// Work order of a batch: counting sort by class (K1 size classes 0..11, then small / large K2 pairs), two linear passes
// over the pairs — at 10^5-10^6 pairs per batch this is on the end-to-end path, so sfc_batch_upload runs it while the
// arena copy is in flight (for the K1-eligible case; sfc_batch_run redoes it when the penalties rule K1 out).
void classify_pairs(sfc_ctx* c, bool k1_ok) {
  constexpr int kSmall = kNumK1Classes, kLarge = kNumK1Classes + 1, kCodes = kNumK1Classes + 2;
  const size_t n = c->n_pairs;
  uint8_t cls_of[40];  // K1 class of a ring-row width W, by ceil(W / 128) (every class capacity is a multiple of 128)
  for (int q = 0; q < 40; ++q) {
    cls_of[q] = 0xFF;
    for (int j = 0; j < kNumK1Classes; ++j) if (q * 128 <= kK1Classes[j].wcap) { cls_of[q] = (uint8_t)j; break; }
  }
  c->h_cls.resize(n);
  size_t count[kCodes] = {};
  for (size_t i = 0; i < n; ++i) {
    const DevPair& d = c->h_pairs[i];
    const long long W = (long long)d.plen + d.tlen + 12;  // ring row: diagonals -plen-1..tlen+1 plus vector-group slack
    int code = W <= 4096 + 7 ? kSmall : kLarge;
    if (k1_ok && W <= 39 * 128 && d.plen <= K1_MAX_LEN && d.tlen <= K1_MAX_LEN && d.pbf <= K1_MAX_FREE && d.tbf <= K1_MAX_FREE) {
      const uint8_t cl = cls_of[(W + 127) >> 7];
      if (cl != 0xFF) code = cl;
    }
    c->h_cls[i] = (uint8_t)code;
    ++count[code];
  }
  c->cls_begin[0] = 0;
  for (int q = 0; q < kCodes; ++q) c->cls_begin[q + 1] = c->cls_begin[q] + count[q];
  c->h_order.resize(n);
  {
    size_t pos[kCodes];
    for (int q = 0; q < kCodes; ++q) pos[q] = c->cls_begin[q];
    for (size_t i = 0; i < n; ++i) c->h_order[pos[c->h_cls[i]]++] = (uint32_t)i;  // stable: no sort inside a K1 class (the groups pull pairs dynamically)
  }
  auto sort_heavy_first = [&](size_t b, size_t e) {
    std::stable_sort(c->h_order.begin() + b, c->h_order.begin() + e, [&](uint32_t x, uint32_t y) {
      const long long wx = (long long)c->h_pairs[x].plen + c->h_pairs[x].tlen, wy = (long long)c->h_pairs[y].plen + c->h_pairs[y].tlen;
      return (wx >> 6) > (wy >> 6);
    });
  };
  sort_heavy_first(c->cls_begin[kSmall], c->cls_begin[kLarge]);
  sort_heavy_first(c->cls_begin[kLarge], n);
  c->cls_valid = true;
  c->cls_k1_ok = k1_ok;
}

extern "C" int sfc_batch_upload(sfc_ctx* c, const uint8_t* arena, size_t arena_len, const sfc_pair* pairs, size_t n) {
  if (!c) return SFC_ERR_INVALID;
  if (!arena || !pairs || n == 0 || n > 0x7FFFFFF0u) return fail(c, SFC_ERR_INVALID, "null/empty batch");
  CU(cudaSetDevice(c->device));
  c->uploaded = c->ran = false;
  c->cls_valid = false;
  c->n_pairs = n;
  c->h_pairs.resize(n);
  c->h_seqs.resize(2 * n);
  c->h_seqwoff.resize(2 * n + 1);
  int rc;
  // the ASCII arena goes first: from pinned host memory its copy overlaps the host loop below (a few ms per 10^5 pairs)
  if ((rc = ensure(c, c->d_arena, arena_len))) return rc;
  CU(cudaEventRecord(c->ev[0], c->stream));
  CU(cudaMemcpyAsync(c->d_arena.p, arena, arena_len, cudaMemcpyHostToDevice, c->stream));
  CU(cudaEventRecord(c->ev[7], c->stream));  // waited for before this call returns: the caller may then reuse `arena`
  // an argument error below must not return while that copy still reads the caller's buffer
  auto bad = [&](const char* fmt, size_t i) { cudaStreamSynchronize(c->stream); return fail(c, SFC_ERR_INVALID, fmt, i); };
  uint64_t woff = 0;
  for (size_t i = 0; i < n; ++i) {
    const sfc_pair& p = pairs[i];
    if (p.pattern_len == 0 || p.text_len == 0) return bad("pair %zu: empty sequence", i);
    if (p.pattern_len > 0x3FFFFFFFu || p.text_len > 0x3FFFFFFFu) return bad("pair %zu: sequence too long", i);
    if (!in_arena(p.pattern_off, p.pattern_len, arena_len) || !in_arena(p.text_off, p.text_len, arena_len))
      return bad("pair %zu: sequence outside arena", i);
    if (p.pattern_begin_free < 0 || p.pattern_end_free < 0 || p.text_begin_free < 0 || p.text_end_free < 0 ||
        (uint32_t)p.pattern_begin_free > p.pattern_len || (uint32_t)p.pattern_end_free > p.pattern_len ||
        (uint32_t)p.text_begin_free > p.text_len || (uint32_t)p.text_end_free > p.text_len)
      return bad("pair %zu: ends-free allowance outside [0, len]", i);
    const uint32_t rev = (p.flags & SFC_PAIR_REVERSE) ? 1u : 0u;
    DevPair& d = c->h_pairs[i];
    d.plen = (int32_t)p.pattern_len; d.tlen = (int32_t)p.text_len;
    d.pbf = p.pattern_begin_free; d.pef = p.pattern_end_free; d.tbf = p.text_begin_free; d.tef = p.text_end_free;
    c->h_seqs[2 * i] = SeqDesc{p.pattern_off, p.pattern_len, rev};
    c->h_seqwoff[2 * i] = (uint32_t)woff; d.p_woff = (uint32_t)woff;
    woff += (p.pattern_len + 7) / 8 + 2;
    c->h_seqs[2 * i + 1] = SeqDesc{p.text_off, p.text_len, rev | 2u};
    c->h_seqwoff[2 * i + 1] = (uint32_t)woff; d.t_woff = (uint32_t)woff;
    woff += (p.text_len + 7) / 8 + 2;
    if (woff > 0xFFFFFFF0ull) return bad("batch too large (packed arena exceeds 2^32 words; pair %zu)", i);
  }
  c->h_seqwoff[2 * n] = (uint32_t)woff;
  if ((rc = ensure(c, c->d_nib, (size_t)woff * 4))) { cudaStreamSynchronize(c->stream); return rc; }
  if ((rc = ensure(c, c->d_seqs, 2 * n * sizeof(SeqDesc)))) { cudaStreamSynchronize(c->stream); return rc; }
  if ((rc = ensure(c, c->d_seqwoff, (2 * n + 1) * 4))) { cudaStreamSynchronize(c->stream); return rc; }
  if ((rc = ensure(c, c->d_pairs, n * sizeof(DevPair)))) { cudaStreamSynchronize(c->stream); return rc; }
  if ((rc = ensure(c, c->d_misc, 256))) { cudaStreamSynchronize(c->stream); return rc; }
  CU(cudaMemcpyAsync(c->d_seqs.p, c->h_seqs.data(), 2 * n * sizeof(SeqDesc), cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemcpyAsync(c->d_seqwoff.p, c->h_seqwoff.data(), (2 * n + 1) * 4, cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemcpyAsync(c->d_pairs.p, c->h_pairs.data(), n * sizeof(DevPair), cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemsetAsync(c->d_misc.p, 0, 256, c->stream));
  CU(cudaEventRecord(c->ev[1], c->stream));
  {
    const uint32_t total_words = (uint32_t)woff;
    const int grid = (int)std::min<uint64_t>((total_words + 255) / 256, (uint64_t)c->sm_count * 16);
    k3_pack<<<grid, 256, 0, c->stream>>>((const uint8_t*)c->d_arena.p, (const SeqDesc*)c->d_seqs.p, (const uint32_t*)c->d_seqwoff.p,
                                         (uint32_t)(2 * n), total_words, (uint32_t*)c->d_nib.p, (uint32_t*)c->d_misc.p + 32);
    CU(cudaGetLastError());
  }
  CU(cudaEventRecord(c->ev[2], c->stream));
  c->stats = sfc_batch_stats{};
  c->stats.h2d_bytes = arena_len + 2 * n * sizeof(SeqDesc) + (2 * n + 1) * 4 + n * sizeof(DevPair);
  c->stats.launches = 1;
  classify_pairs(c, true);  // hidden behind the arena copy
  // the caller's arena has been read completely (include/sawfish_cuda.h: buffer lifetime); the pack kernel keeps running
  CU(cudaEventSynchronize(c->ev[7]));
  c->uploaded = true;
  return SFC_OK;
}

// ============================================================================ run
// d_misc layout (uint32 words): [2..3] cells (u64), [4..5] pool_used (u64), [16..27] K1 class work counters,
// [8] K2 work counter, [32] bad-symbol flag.
static int batch_run_impl(sfc_ctx* c, const sfc_penalties* pen, int want_cigar);

extern "C" int sfc_batch_run(sfc_ctx* c, const sfc_penalties* pen, int want_cigar) {
  if (!c) return SFC_ERR_INVALID;
  const int rc = batch_run_impl(c, pen, want_cigar);
  // out of memory while planning with a remembered free-memory figure (sfc_ctx::free_cache): the failed allocation has
  // invalidated it, so plan once more against what the driver reports now
  if (rc == SFC_ERR_OOM && !c->free_fresh && c->uploaded) return batch_run_impl(c, pen, want_cigar);
  return rc;
}

static int batch_run_impl(sfc_ctx* c, const sfc_penalties* pen, int want_cigar) {
  if (!c) return SFC_ERR_INVALID;
  if (!c->uploaded) return fail(c, SFC_ERR_INVALID, "sfc_batch_run before sfc_batch_upload");
  AdjPen a;
  if (adjust_penalties(pen, &a)) return fail(c, SFC_ERR_INVALID, "invalid penalties");
  const auto t_run_entry = std::chrono::steady_clock::now();
  CU(cudaSetDevice(c->device));
  const size_t n = c->n_pairs;
  int rc;
  if ((rc = ensure(c, c->d_scores, n * 4))) return rc;
  if ((rc = ensure(c, c->d_status, n * 4))) return rc;
  if ((rc = ensure(c, c->d_order, n * 4))) return rc;
  if ((rc = ensure(c, c->d_outoff, n * 8))) return rc;
  if ((rc = ensure(c, c->d_outlen, n * 4))) return rc;
  c->want_cigar = want_cigar ? 1 : 0;

  // ---- classify pairs ----
  const int lookL = std::max(a.x, a.e1);
  const bool k1_ok = (a.metric == 0) && !want_cigar && lookL <= 32;
  // (sfc_batch_upload has classified the batch for the K1-eligible case while the arena copy was in flight)
  if (!c->cls_valid || c->cls_k1_ok != k1_ok) classify_pairs(c, k1_ok);
  constexpr int kSmall = kNumK1Classes, kLarge = kNumK1Classes + 1;
  size_t k1_begin[kNumK1Classes + 1];
  for (int j = 0; j <= kNumK1Classes; ++j) k1_begin[j] = c->cls_begin[j];
  const size_t k2s_begin = c->cls_begin[kSmall], k2l_begin = c->cls_begin[kLarge];
  c->stats.n_score_kernel = (uint32_t)k2s_begin;
  c->stats.n_align_kernel = (uint32_t)(n - k2s_begin);

  uint32_t* misc = (uint32_t*)c->d_misc.p;
  CU(cudaEventRecord(c->ev[3], c->stream));
  CU(cudaMemcpyAsync(c->d_order.p, c->h_order.data(), n * 4, cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemsetAsync(misc, 0, 64, c->stream));  // counters, cells, pool_used (not the bad-symbol flag)
  CU(cudaMemsetAsync(misc + 40, 0, 32, c->stream));  // optional phase timers
  c->stats.h2d_bytes += n * 4;

  // ---- K1 launches: one per size class, concurrently on the job streams (fork / join with events) ----
  CU(cudaMemsetAsync(misc + 16, 0, 64, c->stream));  // per-class work counters
  CU(cudaEventRecord(c->job_ev[6], c->stream));
  int k1_launched = 0;
  for (int j = 0; j < kNumK1Classes; ++j) {
    const size_t cnt = k1_begin[j + 1] - k1_begin[j];
    if (!cnt) continue;
    const int wcap = kK1Classes[j].wcap;
    int wpp = kK1Classes[j].wpp;
    if (const char* e = getenv("SFC_K1_WPP")) {  // tuning knob: warps cooperating on one pair
      const int v = atoi(e);
      if (v == 1 || v == 2 || v == 4 || v == 8) wpp = v;
    }
    K1Params P;
    P.pairs = (const DevPair*)c->d_pairs.p;
    P.order = (const uint32_t*)c->d_order.p + k1_begin[j];
    P.n = (uint32_t)cnt;
    P.nib = (const uint32_t*)c->d_nib.p;
    P.scores = (int32_t*)c->d_scores.p;
    P.status = (int32_t*)c->d_status.p;
    P.x = a.x; P.ip = a.e1; P.match = a.match;
    P.rd = lookL + 1;
    P.wcap = wcap;
    P.win_cap = (wcap >> 3) + 4;
    P.max_score = INT_MAX / 4;
    P.cells = (unsigned long long*)(misc + 2);
    size_t gb = (size_t)P.rd * wcap * 4 + (size_t)wcap * 4 /* per-base windows of both sequences */ + (size_t)(2 * K1P_N + 2 * P.rd + 2) * 4;  // plan records, slot ranges, work index / terminating diagonal
    gb = (gb + 15) & ~(size_t)15;
    P.group_bytes = (int)gb;
    const size_t smem_max = c->smem_optin;
    int max_groups = (int)(smem_max / gb);
    if (max_groups < 1) return fail(c, SFC_ERR_UNSUPPORTED, "K1 class needs %zu B shared memory", gb);
    int gpb = (wpp >= 2) ? 1 : std::min(max_groups, 256 / (wpp * 32));  // one group per block unless warp-per-pair
    gpb = std::min(gpb, 15);
    // prefer two resident blocks per SM when a block would otherwise monopolise shared memory with few warps
    const size_t smem = gb * gpb;
    const int blocks_per_sm = std::max(1, (int)((smem_max + 1024) / (smem + 1024)));
    const size_t need_groups = cnt;
    int grid = (int)std::min<size_t>((need_groups + gpb - 1) / gpb, (size_t)c->sm_count * blocks_per_sm);
    P.counter = misc + 16 + j;
    cudaStream_t st = c->job_stream[k1_launched % 6];
    if (k1_launched < 6) CU(cudaStreamWaitEvent(st, c->job_ev[6], 0));
    switch (wpp) {
      case 1: rc = launch_k1<1>(c, P, gpb, grid, smem, st); break;
      case 2: rc = launch_k1<2>(c, P, gpb, grid, smem, st); break;
      case 4: rc = launch_k1<4>(c, P, gpb, grid, smem, st); break;
      default: rc = launch_k1<8>(c, P, gpb, grid, smem, st); break;
    }
    if (rc) return rc;
    ++k1_launched;
    c->stats.launches++;
  }
  for (int q = 0; q < std::min(k1_launched, 6); ++q) {
    CU(cudaEventRecord(c->job_ev[q], c->job_stream[q]));
    CU(cudaStreamWaitEvent(c->stream, c->job_ev[q], 0));
  }

  // ---- K2 launches: jobs by (size class, cluster size), heaviest first, with workspace escalation ----
  c->pool_used = 0;
  if (k2s_begin < n) {
    size_t free_b = 0;
    std::string laps;
    auto lap = [&](const char* what) {
      if (!getenv("SFC_K2_PROF")) return;
      char buf[64];
      snprintf(buf, sizeof buf, " %s %.2f", what, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_run_entry).count());
      laps += buf;
    };
    lap("k2-begin");
    free_b = device_free_bytes(c);
    lap("meminfo");
    size_t avail = (size_t)((double)(free_b + c->d_ws.cap + c->d_pool.cap + c->d_pages.cap) * 0.85);
    if (c->ws_limit && avail > c->ws_limit) avail = c->ws_limit;
    size_t pool_cap = want_cigar ? std::max<size_t>((n - k2s_begin) * 64, (size_t)1 << 20) : 16;
    if ((rc = ensure(c, c->d_pool, pool_cap * 4))) return rc;
    pool_cap = c->d_pool.cap / 4;
    const int nM = (a.metric == 1 ? std::max(a.x, std::max(a.o1 + a.e1, a.o2 + a.e2)) : lookL) + 1;
    const int nG1 = a.e1 + 1, nG2 = a.e2 + 1;
    const int nslots = a.metric == 1 ? nM + 2 * nG1 + 2 * nG2 : nM;

    struct Job { std::vector<uint32_t> list; int threads; int C; double max_est = 0; bool small_job = false; };
    auto kernel_of = [&](bool wide) -> const void* {
      if (a.metric == 1) return wide ? (const void*)k2_align<true, int32_t> : (const void*)k2_align<true, int16_t>;
      return wide ? (const void*)k2_align<false, int32_t> : (const void*)k2_align<false, int16_t>;
    };
    // K2 v2 launch shapes (sfc_k_align2.cuh): 128 threads for small pairs; for the rest SFC_K2_SHAPE = "threads[,CTAs per SM]"
    // picks one of the compiled (threads, CTAs / SM) variants (default 384 x 1)
    auto v2_kernel = [&](int threads, int minb) -> const void* {
      if (threads == 128 && minb < 2) return (const void*)k2v2_align<128, 1>;
      if (threads == 160) return (const void*)k2v2_align<160, 3>;
      if (threads == 128 && minb >= 2) return (const void*)k2v2_align<128, 4>;
      if (threads == 256) return (const void*)k2v2_align<256, 2>;
      if (threads == 320) return (const void*)k2v2_align<320, 2>;
      if (threads == 512) return (const void*)k2v2_align<512, 1>;
      if (threads == 384 && minb >= 2) return (const void*)k2v2_align<384, 2>;
      return (const void*)k2v2_align<384, 1>;
    };
    int big_threads = K2_LB_T, big_minb = 1;
    if (a.metric == 1 && !getenv("SFC_K2_V1")) {
      // default: two 256-thread CTAs per SM — one pair's serial trim / plan section and its barrier waits overlap the other
      // pair's cell phase (config2 77 vs 83 ms, config3 70 vs 74 ms against one 384-thread CTA per SM)
      big_threads = 256; big_minb = 2;
      if (const char* e = getenv("SFC_K2_SHAPE")) {
        int t = 0, b = 0;
        const int got = sscanf(e, "%d,%d", &t, &b);
        if (got >= 1 && (t == 128 || t == 160 || t == 256 || t == 320 || t == 384 || t == 512)) {
          big_threads = t;
          big_minb = t == 128 ? 4 : t == 160 ? 3 : (t == 256 || t == 320) ? 2 : (t == 384 && got >= 2 && b >= 2) ? 2 : 1;
        }
      }
    }
    std::vector<Job> jobs;
    // small pairs: one 128-thread CTA each
    if (k2l_begin > k2s_begin) jobs.push_back(Job{std::vector<uint32_t>(c->h_order.begin() + k2s_begin, c->h_order.begin() + k2l_begin), 128, 1, 0, true});
    // large pairs: 384-thread CTAs; a pair whose estimated work exceeds the balanced per-SM share is spread over a
    // thread-block cluster of 2/4/8 CTAs so that the heaviest alignment does not set the batch time
    if (n > k2l_begin) {
      const size_t nl = n - k2l_begin;
      std::vector<double> est(nl);
      double total = 0;
      for (size_t q = 0; q < nl; ++q) {
        const DevPair& d = c->h_pairs[c->h_order[k2l_begin + q]];
        est[q] = estimate_work(a, d.plen, d.tlen, d.pbf, d.tbf);
        total += est[q];
      }
      int force_c = 0;
      if (const char* e = getenv("SFC_K2_CLUSTER")) force_c = atoi(e);
      // a pair heavier than the balanced share of one resident CTA is spread over a cluster (K2 v2 may keep several CTAs
      // per SM: each then has a fraction of the SM's throughput)
      const int occ_guess = (a.metric == 1 && !getenv("SFC_K2_V1")) ? big_minb : 1;
      // (measured on config2 / config3, profiles/r02_k2v2_shape_sweep.md: cutting finer than the balanced share costs more
      // in cluster barriers than it saves in tail)
      double share_div = 1.0;
      if (const char* e = getenv("SFC_K2_SHARE_DIV")) share_div = std::max(0.25, atof(e));  // experiment knob
      const double share = std::max(total / ((double)c->sm_count * occ_guess * share_div), 2e6);
      std::vector<uint32_t> byc[4];
      // [0..3]: cluster size 1/2/4/8; [4..7]: the same for pairs whose text is beyond K2 v2's 16-bit offsets — they get jobs
      // of their own (round-1 kernel), so that one such pair does not take the rest of its class off K2 v2 with it
      std::vector<std::pair<double, uint32_t>> tmp[8];
      const bool split_v2 = a.metric == 1 && !getenv("SFC_K2_V1");
      for (size_t q = 0; q < nl; ++q) {
        int ci = 0;
        double r = est[q] / share;
        while (ci < 3 && r > 1.0) { r *= 0.5; ++ci; }
        if (force_c == 1) ci = 0; else if (force_c == 2) ci = 1; else if (force_c == 4) ci = 2; else if (force_c == 8) ci = 3;
        const uint32_t pi = c->h_order[k2l_begin + q];
        const bool over = split_v2 && c->h_pairs[pi].tlen > V2_MAX_TLEN;
        // (such pairs are long: clusters of 4 or 8 only, which also keeps the job count within the eight work counters)
        tmp[over ? 4 + std::max(ci, force_c ? ci : 2) : ci].push_back({est[q], pi});
      }
      for (int cj = 7; cj >= 0; --cj) {
        const int ci = cj & 3;
        if (tmp[cj].empty()) continue;
        std::stable_sort(tmp[cj].begin(), tmp[cj].end(), [](const std::pair<double, uint32_t>& x, const std::pair<double, uint32_t>& y) { return x.first > y.first; });
        Job j; j.threads = big_threads; j.C = 1 << ci;
        if (const char* e = getenv("SFC_K2_THREADS")) { if (a.metric != 1 || getenv("SFC_K2_V1")) j.threads = std::max(64, std::min(K2_LB_T, atoi(e))); }  // experiment knob (v1 kernel)
        for (auto& t : tmp[cj]) j.list.push_back(t.second);
        j.max_est = tmp[cj].front().first;  // sorted heaviest first
        jobs.push_back(std::move(j));
      }
    }

    // size one job's launch: returns grid (in clusters), slab bytes, shared bytes, win_cap
    struct Plan { size_t n_cl, slab, smem, max_fixed, rows_want; int win_cap, stages, minb; bool wide; bool v2; };
    constexpr int kPageShift = 23;          // 8 MB pages
    constexpr size_t kMaxPages = 15360;     // 120 GB: row origins are 32-bit counts of 64-byte units
    size_t v2_pages = 0;                    // pages of the pool of this run (0: no v2 job)
    // max_est: estimated wavefront cells of the heaviest pair of the list (0 = unknown).  One choice byte is kept per cell,
    // so the slab's choice-byte budget follows it: a pair that runs out is re-run from scratch in an escalation pass, at
    // the tail of the batch, which costs far more than the memory does (180 GB of HBM).
    auto plan_job = [&](const std::vector<uint32_t>& list, int threads, int C, size_t budget, int divisor, Plan* pl, double max_est = 0, bool small_job = false) -> int {
      size_t max_fixed = 0, max_win = 0, max_W = 0;
      bool wide = false;
      for (uint32_t pi : list) wide = wide || c->h_pairs[pi].plen > K2_MAX_LEN16 || c->h_pairs[pi].tlen > K2_MAX_LEN16;
      // int32 rings are slightly faster (no 16-bit pack/unpack); int16 halves the workspace and is used on request
      if (!getenv("SFC_K2_NARROW")) wide = true;
      pl->wide = wide;
      const size_t esz = wide ? 4 : 2;
      const void* kfn = kernel_of(wide);
      for (uint32_t pi : list) {
        const DevPair& d = c->h_pairs[pi];
        const size_t Wp = ((size_t)d.plen + d.tlen + 16 + 7) & ~(size_t)7;
        long long scap = gap_cost(a, d.plen) + gap_cost(a, d.tlen) + 1;
        scap = std::min<long long>(scap, INT_MAX / 4);
        const size_t fixed = (size_t)nslots * Wp * esz + ((size_t)scap + 1) * 12 + 2 * (((size_t)d.plen + d.tlen + 8) * 4) + 256;
        max_fixed = std::max(max_fixed, fixed);
        max_win = std::max(max_win, (size_t)(d.plen >> 3) + (d.tlen >> 3) + 2);
        max_W = std::max(max_W, Wp);
      }
      // ---- K2 v2 (sfc_k_align2.cuh): affine 2-piece, texts up to 64 000 bases, wavefront history kept (10 B per cell) ----
      pl->v2 = false;
      if (a.metric == 1 && !getenv("SFC_K2_V1")) {
        bool fits = true;
        size_t v2_fixed = 0;
        for (uint32_t pi : list) {
          const DevPair& d = c->h_pairs[pi];
          fits = fits && d.tlen <= V2_MAX_TLEN;
          long long scap = gap_cost(a, d.plen) + gap_cost(a, d.tlen) + 1;
          scap = std::min<long long>(scap, INT_MAX / 4);
          v2_fixed = std::max(v2_fixed, ((size_t)scap + 1) * V2_HIST * 4 + 2 * (((size_t)d.plen + d.tlen + 8) * 4) + 256);
        }
        const size_t nslots2 = ((size_t)nslots + 1) & ~(size_t)1;
        const size_t fixed_sm2 = (max_win * 8 + nslots2 * 20 + V2_MISC_BYTES + 15) & ~(size_t)15;
        const size_t per_stage2 = (size_t)(threads / 32) * 7 * V2_ROWB;
        int st2 = 2;
        if (const char* e = getenv("SFC_K2_STAGES")) st2 = std::max(1, std::min(3, atoi(e)));
        // shared memory per CTA such that the variant's CTAs per SM all fit (228 KB per SM, 1 KB reserved per CTA)
        const int minb = small_job ? 1 : big_minb;
        const size_t sm_cap = std::min<size_t>(c->smem_optin, ((size_t)228 * 1024) / (size_t)minb - 1024);
        while (st2 > 1 && fixed_sm2 + st2 * per_stage2 > sm_cap) --st2;
        if (fits && fixed_sm2 + st2 * per_stage2 <= c->smem_optin) {
          pl->v2 = true;
          pl->minb = minb;
          pl->wide = false;
          pl->win_cap = (int)max_win;
          pl->stages = st2;
          pl->smem = fixed_sm2 + st2 * per_stage2;
          pl->max_fixed = v2_fixed;
          const void* kv2 = v2_kernel(threads, minb);
          CU(cudaFuncSetAttribute(kv2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->smem_optin));
          CU(cudaFuncSetAttribute(kv2, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
          int occ = 1;
          CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kv2, threads, pl->smem));
          occ = std::max(1, occ);
          size_t slots = (size_t)c->sm_count * occ / C;
          if (C > 1) slots = std::max<size_t>(1, slots * 7 / 8);
          const size_t n_cl = std::max<size_t>(1, std::min<size_t>(list.size(), slots) / divisor);
          // rows: 10 B per wavefront cell, in pages of the shared pool (sized below from every job's wish: 1.5 x the work
          // estimate of its heaviest pair for every resident pair); the slab only holds the per-slot tables
          size_t rows_want = std::max<size_t>((size_t)2048 * max_W * 10, (size_t)16 << 20);
          if (max_est > 0) rows_want = std::max<size_t>(rows_want, (size_t)(15.0 * max_est) + ((size_t)4 << 20));
          pl->rows_want = rows_want;
          size_t slab = (v2_fixed + kMaxPages * 4 + 64 + 255) & ~(size_t)255;
          size_t ncl = n_cl;
          while (ncl > 1 && slab * ncl > budget) ncl = (ncl + 1) / 2;
          if (slab * ncl > budget) return fail(c, SFC_ERR_OOM, "device workspace too small: need %zu B for one pair", slab);
          pl->slab = slab;
          pl->n_cl = ncl;
          return 0;
        }
      }
      const size_t base = std::min(max_win * 8 + (size_t)nslots * 8 + K2_MISC_BYTES, c->smem_optin);
      pl->win_cap = (int)((base - (size_t)nslots * 8 - K2_MISC_BYTES) / 8);
      // cp.async stages for the interior chunks (int32 rings): two per warp when shared memory allows, else one, else none
      const size_t fixed_sm = ((size_t)pl->win_cap * 8 + (size_t)nslots * 8 + K2_MISC_BYTES + 15) & ~(size_t)15;
      const size_t per_stage = (size_t)(threads / 32) * (a.metric == 1 ? K2_SR_N : 2) * K2_SE * 4;
      int stages = wide ? 2 : 0;
      if (const char* e = getenv("SFC_K2_STAGES")) stages = wide ? std::max(0, std::min(3, atoi(e))) : 0;  // experiment knob
      while (stages > 0 && fixed_sm + stages * per_stage > c->smem_optin) --stages;
      pl->stages = stages;
      const size_t smem = stages ? fixed_sm + stages * per_stage : base;
      pl->smem = smem;
      pl->max_fixed = max_fixed;
      // always the device maximum (see launch_k1): per-device attribute shared by every ctx / thread / job
      CU(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->smem_optin));
      int occ = 1;
      CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kfn, threads, smem));
      occ = std::max(1, occ);
      size_t slots = (size_t)c->sm_count * occ / C;
      if (C > 1) slots = std::max<size_t>(1, slots * 7 / 8);  // clusters cannot straddle GPCs
      if (const char* e = getenv("SFC_K2_SLOT_DIV")) slots = std::max<size_t>(1, slots / std::max(1, atoi(e)));  // experiment knob
      size_t n_cl = std::max<size_t>(1, std::min<size_t>(list.size(), slots) / divisor);
      size_t bt_want = std::min<size_t>(std::max<size_t>((size_t)2048 * max_W * divisor, (size_t)4 << 20), (size_t)8 << 30);
      if (max_est > 0 && !getenv("SFC_K2_NO_EST_SLAB"))
        bt_want = std::min<size_t>(std::max<size_t>(bt_want, (size_t)(1.5 * max_est) + ((size_t)1 << 20)), (size_t)8 << 30);
      size_t slab = max_fixed + bt_want;
      while (n_cl > 1 && slab * n_cl > budget) n_cl = (n_cl + 1) / 2;
      if (slab * n_cl > budget) slab = budget;
      if (slab < max_fixed + 4096) return fail(c, SFC_ERR_OOM, "device workspace too small: need %zu B for one pair", max_fixed);
      pl->slab = slab & ~(size_t)255;
      pl->n_cl = n_cl;
      return 0;
    };
    auto launch_job = [&](const uint32_t* order_dev, size_t count, int threads, int C, const Plan& pl, unsigned char* ws, uint32_t* counter,
                          cudaStream_t st) -> int {
      K2Params P;
      P.pairs = (const DevPair*)c->d_pairs.p;
      P.order = order_dev;
      P.n = (uint32_t)count;
      P.counter = counter;
      P.nib = (const uint32_t*)c->d_nib.p;
      P.scores = (int32_t*)c->d_scores.p;
      P.status = (int32_t*)c->d_status.p;
      P.metric = a.metric; P.x = a.x; P.o1 = a.o1; P.e1 = a.e1; P.o2 = a.o2; P.e2 = a.e2; P.match = a.match;
      P.nM = nM; P.nG1 = nG1; P.nG2 = nG2;
      P.slabs = ws;
      P.slab_bytes = pl.slab;
      P.want_cigar = c->want_cigar;
      P.pool = (uint32_t*)c->d_pool.p;
      P.pool_cap = pool_cap;
      P.pool_used = (unsigned long long*)(misc + 4);
      P.out_off = (unsigned long long*)c->d_outoff.p;
      P.out_len = (uint32_t*)c->d_outlen.p;
      P.cells = (unsigned long long*)(misc + 2);
      P.win_cap = pl.win_cap;
      P.stages = pl.stages;
      P.max_score = INT_MAX / 4;
      P.prof = getenv("SFC_K2_PROF") ? (unsigned long long*)(misc + 40) : nullptr;
      P.pages = (unsigned char*)c->d_pages.p; P.page_shift = kPageShift; P.n_pages = (int)v2_pages; P.pgstack = (int*)c->d_pgstack.p; P.pair_pages = (const uint32_t*)c->d_pairpages.p;
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3((unsigned)(pl.n_cl * C));
      cfg.blockDim = dim3((unsigned)threads);
      cfg.dynamicSmemBytes = pl.smem;
      cfg.stream = st;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = (unsigned)C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      if (pl.v2) {
        if (threads == 128 && pl.minb < 2) CU(cudaLaunchKernelEx(&cfg, k2v2_align<128, 1>, P));
        else if (threads == 128) CU(cudaLaunchKernelEx(&cfg, k2v2_align<128, 4>, P));
        else if (threads == 160) CU(cudaLaunchKernelEx(&cfg, k2v2_align<160, 3>, P));
        else if (threads == 256) CU(cudaLaunchKernelEx(&cfg, k2v2_align<256, 2>, P));
        else if (threads == 320) CU(cudaLaunchKernelEx(&cfg, k2v2_align<320, 2>, P));
        else if (threads == 512) CU(cudaLaunchKernelEx(&cfg, k2v2_align<512, 1>, P));
        else if (pl.minb >= 2) CU(cudaLaunchKernelEx(&cfg, k2v2_align<384, 2>, P));
        else CU(cudaLaunchKernelEx(&cfg, k2v2_align<384, 1>, P));
      } else if (a.metric == 1) {
        if (pl.wide) CU(cudaLaunchKernelEx(&cfg, k2_align<true, int32_t>, P)); else CU(cudaLaunchKernelEx(&cfg, k2_align<true, int16_t>, P));
      } else {
        if (pl.wide) CU(cudaLaunchKernelEx(&cfg, k2_align<false, int32_t>, P)); else CU(cudaLaunchKernelEx(&cfg, k2_align<false, int16_t>, P));
      }
      c->stats.launches++;
      return 0;
    };

    // The page pool of the K2 v2 jobs: what they wish for (never less than 2 GB when the device has it, nor than three pages
    // per resident pair), at most what the device has left next to the slabs.  Grow-only across batches.
    auto size_pool = [&](size_t wish, size_t resident, size_t slab_bytes) -> int {
      const size_t left = avail > slab_bytes ? avail - slab_bytes : 0;
      size_t bytes = std::max(std::max(wish, (size_t)2 << 30), (resident * 3 + 8) << kPageShift);
      bytes = std::max(std::min(bytes, left), std::min<size_t>(c->d_pages.cap, left));
      size_t pages = std::min(bytes >> kPageShift, kMaxPages);
      if (pages < 4) return fail(c, SFC_ERR_OOM, "device workspace too small for the wavefront page pool");
      int rc2 = ensure(c, c->d_pages, pages << kPageShift);
      if (rc2) return rc2;
      v2_pages = std::min(std::max(pages, c->d_pages.cap >> kPageShift), kMaxPages);
      return ensure(c, c->d_pgstack, (kMaxPages + 4) * 4);
    };
    // ---- first pass: all jobs back to back on the stream (heaviest clusters first) ----
    // Each job owns a slice of the workspace, of the order array and its own work counter.
    lap("jobs-built");
    std::vector<Plan> plans(jobs.size());
    {
      // split the budget in proportion to demand, then allocate once
      size_t demand_total = 0;
      std::vector<size_t> demand(jobs.size());
      for (size_t j = 0; j < jobs.size(); ++j) {
        if ((rc = plan_job(jobs[j].list, jobs[j].threads, jobs[j].C, avail, 1, &plans[j], jobs[j].max_est, jobs[j].small_job))) return rc;
        demand[j] = plans[j].slab * plans[j].n_cl;
        demand_total += demand[j];
      }
      if (demand_total > avail) {
        for (size_t j = 0; j < jobs.size(); ++j) {
          const size_t budget = std::max<size_t>((size_t)((double)avail * ((double)demand[j] / (double)demand_total)), plans[j].max_fixed + (1 << 20));
          if ((rc = plan_job(jobs[j].list, jobs[j].threads, jobs[j].C, budget, 1, &plans[j], jobs[j].max_est, jobs[j].small_job))) return rc;
        }
      }
      size_t need = 0;
      for (size_t j = 0; j < jobs.size(); ++j) need += plans[j].slab * plans[j].n_cl;
      if ((rc = ensure(c, c->d_ws, need))) return rc;
      // the page pool of the v2 jobs: what they wish for, at least two pages per resident pair, at most what is left
      size_t wish = 0, resident = 0;
      for (size_t j = 0; j < jobs.size(); ++j)
        if (plans[j].v2) { wish += plans[j].rows_want * plans[j].n_cl; resident += plans[j].n_cl; }
      if (resident) {
        if ((rc = size_pool(wish, resident, need))) return rc;
        // more wished for than the pool holds: fewer pairs resident (a pair that finds the pool empty is re-run from
        // scratch in an escalation pass, which costs far more than a shorter queue).  The wish already carries a 1.5 x
        // margin and the resident pairs are never all at their largest at once (measured on config2: a wish of 1.7 x the
        // pool runs without a single retry), so the queue only shrinks beyond 2 x.
        const double have = (double)(v2_pages << kPageShift);
        if ((double)wish > 2.0 * have) {
          const double scale = 3.0 * have / (double)wish;
          for (size_t j = 0; j < jobs.size(); ++j)
            if (plans[j].v2 && scale < 1.0) plans[j].n_cl = std::max<size_t>(1, (size_t)((double)plans[j].n_cl * scale));
        }
      }
    }
    auto reset_pool = [&]() -> int {  // every page free (stream-ordered before the launches of a pass)
      if (!v2_pages) return 0;
      c->h_pgstack.resize(v2_pages + 4);
      c->h_pgstack[0] = 0; c->h_pgstack[1] = (int)v2_pages; c->h_pgstack[2] = 0; c->h_pgstack[3] = 0;
      for (size_t i = 0; i < v2_pages; ++i) c->h_pgstack[4 + i] = (int)(v2_pages - 1 - i);
      CU(cudaMemcpyAsync(c->d_pgstack.p, c->h_pgstack.data(), (v2_pages + 4) * 4, cudaMemcpyHostToDevice, c->stream));
      // pages each pair is expected to need: 10 B per wavefront cell at 1.4 x the work estimate (the estimate was at most
      // 1.37 x short on the synthetic workloads), plus the two pages every pair holds from the start
      c->h_pairpages.assign(n, 2);
      for (size_t q = k2s_begin; q < n; ++q) {
        const uint32_t pi = c->h_order[q];
        const DevPair& d = c->h_pairs[pi];
        const double cells = 1.4 * estimate_work(a, d.plen, d.tlen, d.pbf, d.tbf);
        c->h_pairpages[pi] = (uint32_t)std::min<double>(cells * 10.0 / (double)((size_t)1 << kPageShift) + 3.0, (double)kMaxPages);
      }
      int rc3 = ensure(c, c->d_pairpages, n * 4);
      if (rc3) return rc3;
      CU(cudaMemcpyAsync(c->d_pairpages.p, c->h_pairpages.data(), n * 4, cudaMemcpyHostToDevice, c->stream));
      return 0;
    };
    lap("planned+pool");
    if ((rc = reset_pool())) return rc;
    lap("pool-reset");
    CU(cudaMemsetAsync(misc + 8, 0, 32, c->stream));
    {
      // rewrite the K2 part of the order array job by job (a later run on the same upload classifies afresh)
      c->cls_valid = false;
      size_t pos = k2s_begin;
      for (auto& j : jobs) { std::copy(j.list.begin(), j.list.end(), c->h_order.begin() + pos); pos += j.list.size(); }
      CU(cudaMemcpyAsync((uint32_t*)c->d_order.p + k2s_begin, c->h_order.data() + k2s_begin, (n - k2s_begin) * 4, cudaMemcpyHostToDevice, c->stream));
      size_t ws_off = 0;
      pos = k2s_begin;
      // jobs run concurrently on their own streams (fork from / join to the ctx stream with events): cluster jobs
      // take whole SMs first, single-CTA jobs fill whatever is free and take over as clusters drain
      CU(cudaEventRecord(c->job_ev[6], c->stream));
      int n_hi = 0, n_lo = 0;
      const bool prof_jobs = getenv("SFC_K2_PROF") != nullptr;
      std::vector<cudaEvent_t> prof_ev;
      std::vector<std::string> prof_desc;
      if (prof_jobs) {  // the fork point on the GPU's time line, and how long the host took to get here
        cudaEvent_t e;
        CU(cudaEventCreate(&e));
        CU(cudaEventRecord(e, c->stream));
        prof_ev.push_back(e);
        char buf[96];
        snprintf(buf, sizeof buf, "fork; host %.2f ms after run() entry;", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_run_entry).count());
        prof_desc.push_back(std::string(buf) + laps);
      }
      for (size_t j = 0; j < jobs.size(); ++j) {
        const int sq = jobs[j].C > 1 ? (n_hi++ % 4) : 4 + (n_lo++ % 2);  // clusters first (stream priority, see sfc_create)
        cudaStream_t st = c->job_stream[sq];
        CU(cudaStreamWaitEvent(st, c->job_ev[6], 0));
        if ((rc = launch_job((const uint32_t*)c->d_order.p + pos, jobs[j].list.size(), jobs[j].threads, jobs[j].C, plans[j],
                             (unsigned char*)c->d_ws.p + ws_off, misc + 8 + j, st))) return rc;
        CU(cudaEventRecord(c->job_ev[sq], st));
        CU(cudaStreamWaitEvent(c->stream, c->job_ev[sq], 0));
        if (prof_jobs) {  // when did each job END, counted from the start of the run's kernel section
          cudaEvent_t e;
          CU(cudaEventCreate(&e));
          CU(cudaEventRecord(e, st));
          prof_ev.push_back(e);
          prof_desc.push_back(std::string("C=") + std::to_string(jobs[j].C) + " thr=" + std::to_string(jobs[j].threads) + " pairs=" + std::to_string(jobs[j].list.size()) +
                              " grid=" + std::to_string(plans[j].n_cl * jobs[j].C));
        }
        ws_off += plans[j].slab * plans[j].n_cl;
        pos += jobs[j].list.size();
      }
      if (prof_jobs) {
        std::string line = "[k2 jobs] end of each job after the run's start (ms):";
        for (size_t q = 0; q < prof_ev.size(); ++q) {
          CU(cudaEventSynchronize(prof_ev[q]));
          float ms = 0;
          CU(cudaEventElapsedTime(&ms, c->ev[3], prof_ev[q]));
          char buf[64];
          snprintf(buf, sizeof buf, " %.1f", ms);
          line += "  [" + prof_desc[q] + "]" + buf;
          cudaEventDestroy(prof_ev[q]);
        }
        fprintf(stderr, "%s\n", line.c_str());
      }
    }
    // ---- escalation passes for pairs that ran out of slab (-6) or of pool (-4) ----
    std::vector<uint32_t> retry, cur(c->h_order.begin() + k2s_begin, c->h_order.end());
    int divisor = 1;
    const int force_c_env = getenv("SFC_K2_CLUSTER") ? atoi(getenv("SFC_K2_CLUSTER")) : 0;
    // (a pass that only grows the ops pool does not count against the 12 workspace escalations: it always makes progress)
    for (int pass = 0, pool_passes = 0; pass < 12 && pool_passes < 64; ++pass) {
      c->h_status.resize(n);
      CU(cudaMemcpyAsync(c->h_status.data(), c->d_status.p, n * 4, cudaMemcpyDeviceToHost, c->stream));
      CU(cudaStreamSynchronize(c->stream));
      retry.clear();
      bool pool_short = false;
      for (uint32_t pi : cur) {
        const int stt = c->h_status[pi];
        if (stt == SFC_STATUS_OOM) retry.push_back(pi);
        else if (stt == SFC_STATUS_POOL) { retry.push_back(pi); pool_short = true; }
      }
      if (retry.empty()) break;
      if (divisor >= (1 << 20)) break;  // nothing left to escalate: statuses stay -6
      c->stats.retries++;
      if (pool_short) {
        unsigned long long used = 0;
        CU(cudaMemcpy(&used, misc + 4, 8, cudaMemcpyDeviceToHost));
        DevBuf bigger;
        const size_t new_cap = std::max<size_t>(pool_cap * 4, (size_t)used + ((size_t)1 << 22));
        if ((rc = ensure(c, bigger, new_cap * 4))) return rc;
        CU(cudaMemcpy(bigger.p, c->d_pool.p, std::min<size_t>((size_t)used, pool_cap) * 4, cudaMemcpyDeviceToDevice));
        // failed reservations leave holes above the old capacity but never overlap valid runs
        c->free_cache += c->d_pool.cap;
        release(c->d_pool);
        c->d_pool = bigger;
        pool_cap = c->d_pool.cap / 4;
      }
      bool oom_any = false;
      for (uint32_t pi : retry) oom_any = oom_any || c->h_status[pi] == SFC_STATUS_OOM;
      if (oom_any) divisor *= 4;
      else { --pass; ++pool_passes; }  // pool growth only: same slabs, bigger pool
      Plan pl;
      // the pairs that ran out of choice bytes are the heaviest ones: when few are left, spread each over a cluster
      int Cr = 1;
      for (int cc = 8; cc > 1; cc >>= 1)
        if (retry.size() * (size_t)cc <= (size_t)c->sm_count * 7 / 8) { Cr = cc; break; }
      if (force_c_env > 0) Cr = force_c_env;
      if ((rc = plan_job(retry, big_threads, Cr, avail, divisor, &pl))) return rc;
      if ((rc = ensure(c, c->d_ws, pl.slab * pl.n_cl))) return rc;
      if (pl.v2) {
        // a pair ran out of pages: fewer pairs resident (divisor) AND a pool grown towards what the device has
        const size_t prev = v2_pages;
        if ((rc = size_pool(std::max(pl.rows_want * pl.n_cl, (prev << kPageShift) * 4), pl.n_cl, pl.slab * pl.n_cl))) return rc;
        if ((rc = reset_pool())) return rc;
      }
      cur = retry;
      CU(cudaMemcpyAsync((uint32_t*)c->d_order.p + k2s_begin, retry.data(), retry.size() * 4, cudaMemcpyHostToDevice, c->stream));
      CU(cudaMemsetAsync(misc + 8, 0, 4, c->stream));
      if ((rc = launch_job((const uint32_t*)c->d_order.p + k2s_begin, retry.size(), big_threads, Cr, pl, (unsigned char*)c->d_ws.p, misc + 8, c->stream))) return rc;
      if (pl.n_cl == 1 && (pl.v2 || pl.slab >= avail - 4096) && !pool_short) divisor = 1 << 20;  // last resort was tried
    }
  }
  CU(cudaEventRecord(c->ev[4], c->stream));
  c->ran = true;
  return SFC_OK;
}

// ============================================================================ download
extern "C" int sfc_batch_download(sfc_ctx* c, int32_t* scores, int32_t* status, uint32_t* ops_rle, size_t ops_cap, uint64_t* ops_off) {
  if (!c) return SFC_ERR_INVALID;
  if (!c->ran) return fail(c, SFC_ERR_INVALID, "sfc_batch_download before sfc_batch_run");
  if (!scores || !status) return fail(c, SFC_ERR_INVALID, "null output");
  if (c->want_cigar && (!ops_off || (!ops_rle && ops_cap))) return fail(c, SFC_ERR_INVALID, "null ops output");
  CU(cudaSetDevice(c->device));
  const size_t n = c->n_pairs;
  uint32_t* misc = (uint32_t*)c->d_misc.p;
  uint32_t h_misc[64];
  CU(cudaEventRecord(c->ev[5], c->stream));
  CU(cudaMemcpyAsync(scores, c->d_scores.p, n * 4, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaMemcpyAsync(status, c->d_status.p, n * 4, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaMemcpyAsync(h_misc, misc, 256, cudaMemcpyDeviceToHost, c->stream));
  c->stats.d2h_bytes = n * 8 + 256;
  const bool k2_any = c->stats.n_align_kernel > 0;
  if (c->want_cigar && k2_any) {
    c->h_outoff.resize(n);
    c->h_outlen.resize(n);
    CU(cudaMemcpyAsync(c->h_outoff.data(), c->d_outoff.p, n * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(c->h_outlen.data(), c->d_outlen.p, n * 4, cudaMemcpyDeviceToHost, c->stream));
    c->stats.d2h_bytes += n * 12;
  }
  CU(cudaStreamSynchronize(c->stream));
  unsigned long long cells, pool_used;
  memcpy(&cells, h_misc + 2, 8);
  memcpy(&pool_used, h_misc + 4, 8);
  c->stats.cells = cells;
  if (getenv("SFC_K2_PROF")) {
    unsigned long long pr[8];
    memcpy(pr, h_misc + 40, 64);
    fprintf(stderr, "[k2 prof] cycles of thread 0 summed over %llu CTAs: cells %.3e  barrier waits %.3e  backtrace+fetch %.3e  setup %.3e  trim %.3e  plan %.3e | busiest CTA %.3e\n", pr[7], (double)pr[0], (double)pr[1], (double)pr[2], (double)pr[3], (double)pr[4], (double)pr[5], (double)pr[6]);
  }
  if (h_misc[32]) return fail(c, SFC_ERR_UNSUPPORTED, "sequence symbol outside {A,C,G,T,N,a,c,g,t,n}");
  int ret = SFC_OK;
  if (c->want_cigar) {
    if (!k2_any) return fail(c, SFC_ERR_INVALID, "internal: cigar requested but no backtrace kernel ran");
    const size_t pool_valid = std::min<size_t>((size_t)pool_used, c->d_pool.cap / 4);
    c->h_pool.resize(pool_valid);
    if (pool_valid) {
      CU(cudaMemcpyAsync(c->h_pool.data(), c->d_pool.p, pool_valid * 4, cudaMemcpyDeviceToHost, c->stream));
      CU(cudaStreamSynchronize(c->stream));
      c->stats.d2h_bytes += pool_valid * 4;
    }
    uint64_t acc = 0;
    for (size_t i = 0; i < n; ++i) { ops_off[i] = acc; acc += (status[i] == 0) ? c->h_outlen[i] : 0; }
    ops_off[n] = acc;
    if (acc > ops_cap) ret = fail(c, SFC_ERR_CAPACITY, "ops buffer too small: need %llu entries", (unsigned long long)acc);
    else {
      for (size_t i = 0; i < n; ++i) {
        if (status[i] != 0 || !c->h_outlen[i]) continue;
        memcpy(ops_rle + ops_off[i], c->h_pool.data() + c->h_outoff[i], (size_t)c->h_outlen[i] * 4);
      }
    }
  }
  CU(cudaEventRecord(c->ev[6], c->stream));
  CU(cudaEventSynchronize(c->ev[6]));
  float ms = 0;
  CU(cudaEventElapsedTime(&ms, c->ev[0], c->ev[1])); c->stats.h2d_ms = ms;
  CU(cudaEventElapsedTime(&ms, c->ev[1], c->ev[2])); c->stats.pack_ms = ms;
  CU(cudaEventElapsedTime(&ms, c->ev[3], c->ev[4])); c->stats.kernel_ms = ms;
  CU(cudaEventElapsedTime(&ms, c->ev[5], c->ev[6])); c->stats.d2h_ms = ms;
  return ret;
}

extern "C" int sfc_batch_get_stats(const sfc_ctx* c, sfc_batch_stats* out) {
  if (!c || !out) return SFC_ERR_INVALID;
  *out = c->stats;
  return SFC_OK;
}

extern "C" int sfc_align_batch(sfc_ctx* c, const sfc_penalties* pen, const uint8_t* arena, size_t arena_len, const sfc_pair* pairs,
                               size_t n, int want_cigar, int32_t* scores, int32_t* status, uint32_t* ops_rle, size_t ops_cap,
                               uint64_t* ops_off) {
  int rc = sfc_batch_upload(c, arena, arena_len, pairs, n);
  if (rc) return rc;
  rc = sfc_batch_run(c, pen, want_cigar);
  if (rc) return rc;
  return sfc_batch_download(c, scores, status, ops_rle, ops_cap, ops_off);
}

// ============================================================================ CIGAR helpers (host)
extern "C" int64_t sfc_expand_ops(const uint32_t* ops_rle, size_t n_runs, uint8_t* wfa_ops, size_t cap) {
  static const char kOp[4] = {'M', 'X', 'I', 'D'};
  size_t pos = 0;
  for (size_t i = 0; i < n_runs; ++i) {
    const size_t len = ops_rle[i] >> 2;
    if (pos + len > cap) return SFC_ERR_CAPACITY;
    memset(wfa_ops + pos, kOp[ops_rle[i] & 3u], len);
    pos += len;
  }
  return (int64_t)pos;
}

namespace {
// State machine of process_wfa2_cigar_string (reference src/wfa2_utils.rs:13-95), fed run by run in REVERSE op order.
struct Translator {
  int first_match = 0;
  int64_t ref_offset = 0;
  uint32_t* cigar;
  size_t cap, n = 0;
  uint8_t last = 0;
  uint64_t count = 0;
  int err = 0;
  void flush(bool final_state) {
    if (!count) return;
    int op = -1;
    switch (last) {
      case 'M': op = 7; break;  // Equal
      case 'X': op = 8; break;  // Diff
      case 'D': op = (!first_match || final_state) ? 4 : 1; break;  // SoftClip : Ins
      case 'I':
        if (!first_match) ref_offset += (int64_t)count;
        else if (!final_state) op = 2;  // Del
        break;
      default: err = SFC_ERR_INVALID; return;
    }
    if (op >= 0) {
      if (n >= cap) { err = SFC_ERR_CAPACITY; return; }
      cigar[n++] = (uint32_t)(count << 4) | (uint32_t)op;
    }
  }
  void feed(uint8_t state, uint64_t len) {
    if (!len) return;
    if (state != last) { flush(false); count = 0; }
    // first_al_match is raised by the first M/X *before* its own run is flushed, exactly as in the reference
    if (!first_match && (state == 'M' || state == 'X')) first_match = 1;
    count += len;
    last = state;
  }
  void finish() { flush(true); count = 0; }
};
}  // namespace

extern "C" int sfc_translate_cigar(const uint8_t* wfa_ops, size_t n_ops, int64_t* ref_offset, uint32_t* cigar, size_t cigar_cap,
                                   size_t* n_cigar, int* has_match) {
  if ((!wfa_ops && n_ops) || (!cigar && cigar_cap)) return SFC_ERR_INVALID;
  Translator t;
  t.cigar = cigar;
  t.cap = cigar_cap;
  for (size_t i = n_ops; i-- > 0;) {
    const uint8_t s = wfa_ops[i];
    if (s != 'M' && s != 'X' && s != 'I' && s != 'D') return SFC_ERR_INVALID;
    t.feed(s, 1);
    if (t.err) return t.err;
  }
  t.finish();
  if (t.err) return t.err;
  if (ref_offset) *ref_offset = t.ref_offset;
  if (n_cigar) *n_cigar = t.n;
  if (has_match) *has_match = t.first_match;
  return SFC_OK;
}

extern "C" int sfc_translate_rle(const uint32_t* ops_rle, size_t n_runs, int64_t* ref_offset, uint32_t* cigar, size_t cigar_cap,
                                 size_t* n_cigar, int* has_match) {
  static const uint8_t kOp[4] = {'M', 'X', 'I', 'D'};
  if ((!ops_rle && n_runs) || (!cigar && cigar_cap)) return SFC_ERR_INVALID;
  Translator t;
  t.cigar = cigar;
  t.cap = cigar_cap;
  for (size_t i = n_runs; i-- > 0;) {
    t.feed(kOp[ops_rle[i] & 3u], ops_rle[i] >> 2);
    if (t.err) return t.err;
  }
  t.finish();
  if (t.err) return t.err;
  if (ref_offset) *ref_offset = t.ref_offset;
  if (n_cigar) *n_cigar = t.n;
  if (has_match) *has_match = t.first_match;
  return SFC_OK;
}

// ============================================================================ PairwiseAligner mirror
static int aligner_new(sfc_ctx* c, const sfc_penalties& pen, sfc_aligner** out) {
  if (!c || !out) return SFC_ERR_INVALID;
  sfc_aligner* a = new (std::nothrow) sfc_aligner();
  if (!a) return SFC_ERR_OOM;
  a->ctx = c;
  a->pen = pen;
  *out = a;
  return SFC_OK;
}

extern "C" int sfc_aligner_new_consensus(sfc_ctx* c, sfc_aligner** out) {
  const sfc_penalties pen = {0, -1, 3, 0, 2, 0, 0};  // src/wfa2_utils.rs:183-189
  return aligner_new(c, pen, out);
}

extern "C" int sfc_aligner_new_large_indel(sfc_ctx* c, int is_long_contig, sfc_aligner** out) {
  // is_long_contig selects WFA2's MemoryMed vs MemoryHigh in the reference (src/wfa2_utils.rs:153-158); both are
  // designed to return the same alignment and this implementation has a single (choice-byte) backtrace, so it is unused.
  (void)is_long_contig;
  const sfc_penalties pen = {1, -1, 3, 1, 2, 60, 0};  // src/wfa2_utils.rs:159-168
  return aligner_new(c, pen, out);
}

extern "C" void sfc_aligner_free(sfc_aligner* a) { delete a; }

extern "C" int sfc_aligner_set_text(sfc_aligner* a, const uint8_t* text, size_t len) {
  if (!a) return SFC_ERR_INVALID;
  if (!text || len == 0) return fail(a->ctx, SFC_ERR_INVALID, "set_text: empty text");  // assert!(!text.is_empty())
  a->text.assign(text, text + len);
  return SFC_OK;
}

extern "C" int sfc_aligner_align(sfc_aligner* a, const uint8_t* pattern, size_t len, int mode, double text_clip_frac,
                                 double pattern_clip_frac, int32_t* score, int* has_alignment, int64_t* ref_offset, uint32_t* cigar,
                                 size_t cigar_cap, size_t* n_cigar) {
  if (!a || !score) return SFC_ERR_INVALID;
  sfc_ctx* c = a->ctx;
  if (a->text.empty()) return fail(c, SFC_ERR_INVALID, "align: text not set");
  if (!pattern || len == 0) return fail(c, SFC_ERR_INVALID, "align: empty pattern");
  std::vector<uint8_t> arena(a->text);
  arena.insert(arena.end(), pattern, pattern + len);
  sfc_pair pr;
  memset(&pr, 0, sizeof(pr));
  pr.text_off = 0; pr.text_len = (uint32_t)a->text.size();
  pr.pattern_off = a->text.size(); pr.pattern_len = (uint32_t)len;
  pr.flags = SFC_PAIR_REVERSE;
  if (mode == 0) {  // src/wfa2_utils.rs:112-116
    const int p_clip = (int)((double)len * 0.4), t_clip = (int)((double)a->text.size() * 0.4);
    const int clip = std::min(p_clip, t_clip);
    pr.pattern_begin_free = pr.pattern_end_free = pr.text_begin_free = pr.text_end_free = clip;
  } else if (mode == 1) {  // src/wfa2_utils.rs:127-130
    const int p_clip = (int)((double)len * pattern_clip_frac), t_clip = (int)((double)a->text.size() * text_clip_frac);
    pr.pattern_begin_free = pr.pattern_end_free = p_clip;
    pr.text_begin_free = pr.text_end_free = t_clip;
  } else if (mode != 2) {
    return fail(c, SFC_ERR_INVALID, "align: unknown mode %d", mode);
  }
  int32_t status = 0;
  std::vector<uint32_t> ops(len + a->text.size() + 8);
  uint64_t off[2] = {0, 0};
  int rc = sfc_align_batch(c, &a->pen, arena.data(), arena.size(), &pr, 1, 1, score, &status, ops.data(), ops.size(), off);
  if (rc) return rc;
  if (status != 0)  // per-pair status -> sfc_error: only -5 / -6 share their meaning with the error enum
    return fail(c, (status == SFC_STATUS_UNSUPPORTED || status == SFC_STATUS_OOM) ? status : SFC_ERR_INTERNAL, "align: pair status %d", status);
  a->last_ops.assign(ops.begin(), ops.begin() + (size_t)off[1]);
  a->last_pattern.assign(pattern, pattern + len);
  int hm = 0;
  size_t nc = 0;
  int64_t roff = 0;
  rc = sfc_translate_rle(ops.data(), (size_t)off[1], &roff, cigar, cigar_cap, &nc, &hm);
  if (rc) return fail(c, rc, "align: cigar translation failed (%d)", rc);
  if (has_alignment) *has_alignment = hm;
  if (ref_offset) *ref_offset = roff;
  if (n_cigar) *n_cigar = nc;
  return SFC_OK;
}

// Debug view of the last alignment (reference src/wfa2_utils.rs:253-255: PairwiseAligner::matching -> WFAligner::matching):
// the gapped pattern, a line of '|' under matches, the gapped text, in the order the sequences were aligned (reversed).
extern "C" int sfc_aligner_matching(sfc_aligner* a, char* pattern_line, char* mid_line, char* text_line, size_t cap, size_t* out_len) {
  if (!a || !out_len) return SFC_ERR_INVALID;
  size_t n = 0;
  for (uint32_t w : a->last_ops) n += w >> 2;
  *out_len = n;
  if (n == 0) return a->last_ops.empty() && a->last_pattern.empty() ? fail(a->ctx, SFC_ERR_INVALID, "matching: no alignment yet") : SFC_OK;
  if (!pattern_line || !mid_line || !text_line || cap < n) return fail(a->ctx, SFC_ERR_CAPACITY, "matching: need %zu bytes per line", n);
  const size_t pl = a->last_pattern.size(), tl = a->text.size();
  size_t v = 0, h = 0, pos = 0;  // positions in the REVERSED pattern / text
  for (uint32_t w : a->last_ops) {
    const uint32_t code = w & 3u;
    for (uint32_t i = 0; i < (w >> 2); ++i, ++pos) {
      const char pc = (code != SFC_OP_I && v < pl) ? (char)a->last_pattern[pl - 1 - v] : '-';
      const char tc = (code != SFC_OP_D && h < tl) ? (char)a->text[tl - 1 - h] : '-';
      pattern_line[pos] = pc; text_line[pos] = tc; mid_line[pos] = code == SFC_OP_M ? '|' : ' ';
      if (code != SFC_OP_I) ++v;
      if (code != SFC_OP_D) ++h;
    }
  }
  return SFC_OK;
}

// ============================================================================ sharding helpers
namespace {
double estimate_work(const AdjPen& a, double plen, double tlen, double pbf, double tbf) {
  double dl = std::abs(plen - tlen);
  static const int mode = getenv("SFC_EST") ? atoi(getenv("SFC_EST")) : 0;
  if (mode == 1) dl = std::max(0.0, tlen > plen ? (tlen - plen) - 2.0 * tbf : (plen - tlen) - 2.0 * pbf);
  const double es = (double)gap_cost(a, (long long)dl) + 0.012 * 0.5 * (plen + tlen) + 8.0;  // expected final score
  if (mode == 1) return es * std::min(pbf + tbf + 1.0 + es, plen + tlen + 1.0);
  return es * (pbf + tbf + 1.0 + 0.4 * es);                                                   // x mean wavefront width
}
}  // namespace

extern "C" int sfc_estimate_pair_work(const sfc_penalties* pen, const sfc_pair* pairs, size_t n, uint64_t* work) {
  AdjPen a;
  if (adjust_penalties(pen, &a) || (!pairs && n) || (!work && n)) return SFC_ERR_INVALID;
  for (size_t i = 0; i < n; ++i)
    work[i] = (uint64_t)estimate_work(a, pairs[i].pattern_len, pairs[i].text_len, pairs[i].pattern_begin_free, pairs[i].text_begin_free);
  return SFC_OK;
}

extern "C" int sfc_shard_plan(const uint64_t* group_work, size_t n_groups, int n_shards, uint32_t* shard_of_group) {
  if (n_shards <= 0 || (!group_work && n_groups) || (!shard_of_group && n_groups)) return SFC_ERR_INVALID;
  std::vector<uint32_t> idx(n_groups);
  for (size_t i = 0; i < n_groups; ++i) idx[i] = (uint32_t)i;
  std::stable_sort(idx.begin(), idx.end(), [&](uint32_t x, uint32_t y) { return group_work[x] > group_work[y]; });
  std::vector<unsigned long long> load((size_t)n_shards, 0);
  for (uint32_t g : idx) {
    int best = 0;
    for (int s = 1; s < n_shards; ++s) if (load[s] < load[best]) best = s;
    shard_of_group[g] = (uint32_t)best;
    load[best] += group_work[g] + 1;
  }
  return SFC_OK;
}

// ============================================================================ one process, several GPUs
struct sfc_multi {
  // `split` contexts (and submitter threads) per device: a score-only batch is cut into n_dev x split shards so that one
  // sub-shard's host staging (compaction, upload) overlaps the other's kernels on the same GPU.  ctx[d * split + j].
  int n_dev = 0, split = 1;
  std::vector<sfc_ctx*> ctx;
  std::vector<double> kernel_ms;  // per context
  std::string err;
  // per-shard pinned staging (grow-only): compacted arena, pair table, results
  struct Stage {
    uint8_t* arena = nullptr; size_t arena_cap = 0;
    sfc_pair* pairs = nullptr; size_t pairs_cap = 0;
    int32_t* scores = nullptr; size_t scores_cap = 0;
    int32_t* status = nullptr; size_t status_cap = 0;
  };
  std::vector<Stage> stage;
};

namespace {
template <typename T>
int pinned_reserve(T*& p, size_t& cap, size_t n) {
  if (n <= cap) return 0;
  if (p) cudaFreeHost(p);
  p = nullptr;
  cap = 0;
  const size_t want = n + n / 4 + 64;
  if (cudaMallocHost((void**)&p, want * sizeof(T)) != cudaSuccess) { (void)cudaGetLastError(); p = nullptr; return SFC_ERR_OOM; }
  cap = want;
  return 0;
}
}  // namespace

extern "C" int sfc_multi_create(const int* devs, int n, sfc_multi** out) {
  if (!devs || n <= 0 || !out) return SFC_ERR_INVALID;
  *out = nullptr;
  sfc_multi* m = new (std::nothrow) sfc_multi();
  if (!m) return SFC_ERR_OOM;
  m->n_dev = n;
  m->split = 2;
  if (const char* e = getenv("SFC_MULTI_SPLIT")) m->split = std::max(1, std::min(4, atoi(e)));
  for (int i = 0; i < n; ++i) {
    for (int j = 0; j < m->split; ++j) {
      sfc_ctx* c = nullptr;
      const int rc = sfc_create(devs[i], &c);
      if (rc) { sfc_multi_destroy(m); return rc; }
      m->ctx.push_back(c);
    }
  }
  m->kernel_ms.assign(m->ctx.size(), 0.0);
  m->stage.resize(m->ctx.size());
  *out = m;
  return SFC_OK;
}

extern "C" void sfc_multi_destroy(sfc_multi* m) {
  if (!m) return;
  for (auto& st : m->stage) {
    if (st.arena) cudaFreeHost(st.arena);
    if (st.pairs) cudaFreeHost(st.pairs);
    if (st.scores) cudaFreeHost(st.scores);
    if (st.status) cudaFreeHost(st.status);
  }
  for (sfc_ctx* c : m->ctx) sfc_destroy(c);
  delete m;
}

extern "C" int sfc_multi_device_count(const sfc_multi* m) { return m ? m->n_dev : 0; }
extern "C" const char* sfc_multi_last_error(const sfc_multi* m) { return m ? m->err.c_str() : "null"; }

extern "C" int sfc_multi_last_kernel_ms(const sfc_multi* m, double* out) {
  if (!m || !out) return SFC_ERR_INVALID;
  // per device: the sum over its contexts (their kernels share the GPU, so this is an upper bound of the time it was busy)
  for (int d = 0; d < m->n_dev; ++d) {
    out[d] = 0.0;
    for (int j = 0; j < m->split; ++j) out[d] += m->kernel_ms[(size_t)d * m->split + j];
  }
  return SFC_OK;
}

extern "C" int sfc_multi_align_batch(sfc_multi* m, const sfc_penalties* pen, const uint8_t* arena, size_t arena_len,
                                     const sfc_pair* pairs, size_t n, const uint32_t* group_of_pair, int want_cigar,
                                     int32_t* scores, int32_t* status, uint32_t* ops_rle, size_t ops_cap, uint64_t* ops_off) {
  if (!m || !pen || !arena || !pairs || !scores || !status || n == 0) return SFC_ERR_INVALID;
  if (want_cigar && !ops_off) return SFC_ERR_INVALID;
  // shards: `split` per device for a score-only gap-linear batch (K1: no workspace, kernels of two contexts share a GPU
  // well), one per device otherwise (two K2 runs would each size their workspace for the whole device)
  std::fill(m->kernel_ms.begin(), m->kernel_ms.end(), 0.0);
  const int split = (pen->metric == 0 && !want_cigar) ? m->split : 1;
  const int nd = m->n_dev * split;
  auto ctx_of = [&](int d) -> size_t { return (size_t)(d / split) * m->split + (size_t)(d % split); };
  // ---- plan: LPT over groups by estimated work.  Without caller groups, runs of consecutive pairs form the groups
  // (consecutive pairs share reads / loci in the callers), which keeps the plan O(n) instead of a sort over every pair.
  // (the per-pair estimate and the bounds check are split over a few host threads: at 10^5-10^6 pairs per batch the
  // serial part of this call is what limits strong scaling over 8 GPUs)
  std::vector<uint64_t> work(n);
  int rc = 0;
  {
    AdjPen a_chk;
    if (adjust_penalties(pen, &a_chk)) { m->err = "invalid penalties"; return SFC_ERR_INVALID; }
    const size_t T = std::max<size_t>(1, std::min<size_t>(16, n / 32768));
    std::vector<int> trc(T, 0);
    std::vector<std::thread> pt;
    for (size_t t = 0; t < T; ++t) {
      pt.emplace_back([&, t]() {
        const size_t b = n * t / T, e = n * (t + 1) / T;
        trc[t] = sfc_estimate_pair_work(pen, pairs + b, e - b, work.data() + b);
        for (size_t i = b; i < e && !trc[t]; ++i)
          if (!in_arena(pairs[i].pattern_off, pairs[i].pattern_len, arena_len) || !in_arena(pairs[i].text_off, pairs[i].text_len, arena_len)) trc[t] = -100;
      });
    }
    for (auto& t : pt) t.join();
    for (int v : trc) {
      if (v == -100) { m->err = "sequence outside arena"; return SFC_ERR_INVALID; }
      if (v) { m->err = "invalid penalties"; return v; }
    }
  }
  const size_t run = group_of_pair ? 1 : std::max<size_t>(1, std::min<size_t>(256, n / ((size_t)nd * 64) + 1));
  auto group_of = [&](size_t i) -> uint32_t { return group_of_pair ? group_of_pair[i] : (uint32_t)(i / run); };
  uint32_t n_groups = 0;
  for (size_t i = 0; i < n; ++i) n_groups = std::max(n_groups, group_of(i) + 1);
  std::vector<uint64_t> gwork(n_groups, 0);
  for (size_t i = 0; i < n; ++i) gwork[group_of(i)] += work[i];
  std::vector<uint32_t> shard_of_group(n_groups);
  if ((rc = sfc_shard_plan(gwork.data(), n_groups, nd, shard_of_group.data()))) return rc;
  // ---- per shard: pair list, then a compacted copy of just its sequences in pinned staging ----
  struct Shard {
    std::vector<uint32_t> idx;
    size_t arena_bytes = 0;
    std::vector<uint32_t> ops;
    std::vector<uint64_t> ops_off;
    int rc = 0;
  };
  std::vector<Shard> sh((size_t)nd);
  // ---- one submitter thread per GPU: sizing (pass 1), compaction (pass 2), upload, kernels, download ----
  std::vector<std::thread> th;
  for (int d = 0; d < nd; ++d) {
    th.emplace_back([&, d]() {
      Shard& s = sh[(size_t)d];
      const size_t ci = ctx_of(d);
      sfc_multi::Stage& st = m->stage[ci];
      m->kernel_ms[ci] = 0.0;
      // this shard's pair list: every submitter scans the plan itself (a serial fill of all lists would sit in front of
      // every GPU's work)
      s.idx.reserve(n / (size_t)nd + n / 8 + 16);
      for (size_t i = 0; i < n; ++i) if (shard_of_group[group_of(i)] == (uint32_t)d) s.idx.push_back((uint32_t)i);
      if (s.idx.empty()) return;
      {
        uint64_t lt_off = ~0ull;
        uint32_t lt_len = 0;
        size_t bytes = 0;
        for (uint32_t i : s.idx) {  // pass 1: size (consecutive pairs usually share their text: one read flank, several alleles)
          const sfc_pair& q = pairs[i];
          if (!(q.text_off == lt_off && q.text_len == lt_len)) { bytes += q.text_len; lt_off = q.text_off; lt_len = q.text_len; }
          bytes += q.pattern_len;
        }
        cudaSetDevice(m->ctx[ci]->device);
        if (pinned_reserve(st.arena, st.arena_cap, bytes) || pinned_reserve(st.pairs, st.pairs_cap, s.idx.size()) ||
            pinned_reserve(st.scores, st.scores_cap, s.idx.size()) || pinned_reserve(st.status, st.status_cap, s.idx.size())) {
          s.rc = SFC_ERR_OOM;
          return;
        }
        s.arena_bytes = bytes;
      }
      // pass 2: new offsets (serial, a few ns per pair), then the copies themselves on a few threads: ~0.1 us per pair in
      // two small memcpys is 50 ms for a 5 x 10^5-pair shard on one thread, more than the shard's share of the kernels
      uint64_t last_t_off = ~0ull, last_t_new = 0, pos = 0;
      uint32_t last_t_len = 0;
      std::vector<uint8_t> own_text(s.idx.size());
      for (size_t j = 0; j < s.idx.size(); ++j) {
        sfc_pair q = pairs[s.idx[j]];
        if (q.text_off == last_t_off && q.text_len == last_t_len) { q.text_off = last_t_new; own_text[j] = 0; }
        else {
          last_t_off = q.text_off; last_t_len = q.text_len; last_t_new = pos;
          q.text_off = pos; pos += q.text_len;
          own_text[j] = 1;
        }
        q.pattern_off = pos; pos += q.pattern_len;
        st.pairs[j] = q;
      }
      {
        const size_t np_ = s.idx.size();
        const size_t T = std::max<size_t>(1, std::min<size_t>(4, np_ / 16384));
        auto copy_range = [&](size_t b, size_t e) {
          for (size_t j = b; j < e; ++j) {
            const sfc_pair& o = pairs[s.idx[j]];
            const sfc_pair& q = st.pairs[j];
            if (own_text[j]) memcpy(st.arena + q.text_off, arena + o.text_off, o.text_len);
            memcpy(st.arena + q.pattern_off, arena + o.pattern_off, o.pattern_len);
          }
        };
        std::vector<std::thread> ct;
        for (size_t t = 1; t < T; ++t) ct.emplace_back(copy_range, np_ * t / T, np_ * (t + 1) / T);
        copy_range(0, np_ / T);
        for (auto& t : ct) t.join();
      }
      if (want_cigar) { s.ops_off.resize(s.idx.size() + 1); s.ops.resize(std::max<size_t>(s.idx.size() * 64, 1 << 16)); }
      sfc_ctx* c = m->ctx[ci];
      s.rc = sfc_batch_upload(c, st.arena, s.arena_bytes, st.pairs, s.idx.size());
      if (!s.rc) s.rc = sfc_batch_run(c, pen, want_cigar);
      if (!s.rc) {
        s.rc = sfc_batch_download(c, st.scores, st.status, want_cigar ? s.ops.data() : nullptr, s.ops.size(),
                                  want_cigar ? s.ops_off.data() : nullptr);
        if (s.rc == SFC_ERR_CAPACITY) {
          s.ops.resize((size_t)s.ops_off[s.idx.size()] + 16);
          s.rc = sfc_batch_download(c, st.scores, st.status, s.ops.data(), s.ops.size(), s.ops_off.data());
        }
      }
      if (!s.rc) {  // host-side gather of this shard's integer results into the caller's pair order
        for (size_t j = 0; j < s.idx.size(); ++j) { scores[s.idx[j]] = st.scores[j]; status[s.idx[j]] = st.status[j]; }
      }
      sfc_batch_stats bs;
      if (!sfc_batch_get_stats(c, &bs)) m->kernel_ms[ci] = bs.kernel_ms;
    });
  }
  for (auto& t : th) t.join();
  for (int d = 0; d < nd; ++d) {
    if (sh[(size_t)d].rc == SFC_ERR_OOM && sh[(size_t)d].arena_bytes == 0) { m->err = "pinned staging allocation failed"; return SFC_ERR_OOM; }
    if (sh[(size_t)d].rc) { m->err = std::string("device ") + std::to_string(d / split) + ": " + sfc_last_error(m->ctx[ctx_of(d)]); return sh[(size_t)d].rc; }
  }
  if (want_cigar) {
    std::vector<uint32_t> where(n);
    std::vector<uint8_t> shard_of_pair(n);
    for (int d = 0; d < nd; ++d)
      for (size_t j = 0; j < sh[(size_t)d].idx.size(); ++j) { where[sh[(size_t)d].idx[j]] = (uint32_t)j; shard_of_pair[sh[(size_t)d].idx[j]] = (uint8_t)d; }
    uint64_t acc = 0;
    for (size_t i = 0; i < n; ++i) {
      const Shard& s = sh[shard_of_pair[i]];
      ops_off[i] = acc;
      acc += s.ops_off[where[i] + 1] - s.ops_off[where[i]];
    }
    ops_off[n] = acc;
    if (acc > ops_cap || (!ops_rle && acc)) { m->err = "ops buffer too small"; return SFC_ERR_CAPACITY; }
    for (size_t i = 0; i < n; ++i) {
      const Shard& s = sh[shard_of_pair[i]];
      const uint64_t b0 = s.ops_off[where[i]], e0 = s.ops_off[where[i] + 1];
      if (e0 > b0) memcpy(ops_rle + ops_off[i], s.ops.data() + b0, (size_t)(e0 - b0) * 4);
    }
  }
  return SFC_OK;
}

// ============================================================================ integer-ALU peak
extern "C" int sfc_measure_int_peak(sfc_ctx* c, double* gops, double* mhz_est) {
  if (!c || !gops) return SFC_ERR_INVALID;
  CU(cudaSetDevice(c->device));
  const int blocks = c->sm_count * 8, threads = 256, iters = 4096;
  DevBuf out;
  int rc = ensure(c, out, (size_t)blocks * threads * 4);
  if (rc) return rc;
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0));
  CU(cudaEventCreate(&e1));
  double best = 0;
  for (int rep = 0; rep < 5; ++rep) {
    CU(cudaEventRecord(e0, c->stream));
    int_peak_kernel<<<blocks, threads, 0, c->stream>>>((unsigned*)out.p, iters, 12345u + rep);
    CU(cudaEventRecord(e1, c->stream));
    CU(cudaEventSynchronize(e1));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, e0, e1));
    const double ops = (double)blocks * threads * (double)iters * 8.0 * 8.0 * 2.0;
    best = std::max(best, ops / (ms * 1e-3) / 1e9);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  c->free_cache += out.cap;
  release(out);
  *gops = best;
  if (mhz_est) *mhz_est = best * 1e3 / ((double)c->sm_count * 64.0);  // assuming 64 INT32 lanes/SM/clk
  return SFC_OK;
}

// ============================================================================ K4: rust-bio style aligner
extern "C" int sfc_bio_align_batch(sfc_ctx* c, const sfc_bio_scoring* sco, const uint8_t* arena, size_t arena_len,
                                   const sfc_pair* pairs, size_t n, int32_t* scores, int32_t* status, int32_t* bounds,
                                   uint32_t* ops, size_t ops_cap, uint64_t* ops_off) {
  if (!c) return SFC_ERR_INVALID;
  if (!sco || !arena || !pairs || n == 0 || n > 0x7FFFFFF0u || !scores || !status || !bounds)
    return fail(c, SFC_ERR_INVALID, "null/empty batch");
  if (sco->gap_open > 0 || sco->gap_extend > 0) return fail(c, SFC_ERR_INVALID, "gap penalties must be <= 0");
  if ((ops == nullptr) != (ops_off == nullptr)) return fail(c, SFC_ERR_INVALID, "ops and ops_off go together");
  CU(cudaSetDevice(c->device));
  std::vector<K4Pair> hp(n);
  size_t max_need = 0, tot_ops = 0;
  for (size_t i = 0; i < n; ++i) {
    const sfc_pair& p = pairs[i];
    if (p.pattern_len == 0 || p.text_len == 0) return fail(c, SFC_ERR_INVALID, "pair %zu: empty sequence", i);  // assert!(!text.is_empty())
    if (!in_arena(p.pattern_off, p.pattern_len, arena_len) || !in_arena(p.text_off, p.text_len, arena_len))
      return fail(c, SFC_ERR_INVALID, "pair %zu outside the arena", i);
    if (p.pattern_len > 1000000u || p.text_len > 1000000u) return fail(c, SFC_ERR_UNSUPPORTED, "pair %zu: sequence too long", i);
    hp[i] = K4Pair{p.pattern_off, p.text_off, (int)p.pattern_len, (int)p.text_len};
    max_need = std::max(max_need, k4_slab_need(hp[i].m, hp[i].n));
    tot_ops += (size_t)p.pattern_len + p.text_len + 4;
  }
  const size_t free_b = device_free_bytes(c);
  size_t avail = (size_t)((double)(free_b + c->d_ws.cap) * 0.8);
  if (c->ws_limit && avail > c->ws_limit) avail = c->ws_limit;
  if (max_need > avail) return fail(c, SFC_ERR_OOM, "device workspace too small: need %zu B for one pair", max_need);
  const size_t slab = (max_need + 255) & ~(size_t)255;
  size_t ctas = std::min<size_t>(std::min<size_t>(n, (size_t)c->sm_count * 2), std::max<size_t>(1, avail / slab));
  int rc;
  if ((rc = ensure(c, c->d_ws, slab * ctas))) return rc;
  if ((rc = ensure(c, c->d_b_arena, arena_len))) return rc;
  if ((rc = ensure(c, c->d_b_pairs, n * sizeof(K4Pair)))) return rc;
  // out block: scores[n] status[n] bounds[4n] out_len[n] (int32) | out_off[n] (u64) | counter, pool_used
  const size_t out_i32 = 7 * n, out_bytes = ((out_i32 * 4 + 15) & ~(size_t)15) + 8 * n + 32;
  if ((rc = ensure(c, c->d_b_out, out_bytes))) return rc;
  const size_t pool_cap = std::min<size_t>(tot_ops, (size_t)1 << 31);
  if ((rc = ensure(c, c->d_b_pool, pool_cap * 4))) return rc;
  unsigned char* ob = (unsigned char*)c->d_b_out.p;
  int32_t* d_scores = (int32_t*)ob;
  int32_t* d_status = d_scores + n;
  int32_t* d_bounds = d_status + n;
  uint32_t* d_outlen = (uint32_t*)(d_bounds + 4 * n);
  unsigned long long* d_outoff = (unsigned long long*)(ob + ((out_i32 * 4 + 15) & ~(size_t)15));
  unsigned long long* d_tail = d_outoff + n;  // [0] pool_used, [1] work counter (low 32 bits)
  CU(cudaMemcpyAsync(c->d_b_arena.p, arena, arena_len, cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemcpyAsync(c->d_b_pairs.p, hp.data(), n * sizeof(K4Pair), cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemsetAsync(d_tail, 0, 16, c->stream));
  K4Params P;
  P.arena = (const uint8_t*)c->d_b_arena.p;
  P.pairs = (const K4Pair*)c->d_b_pairs.p;
  P.n = (uint32_t)n;
  P.counter = (uint32_t*)(d_tail + 1);
  P.sc = K4Scoring{sco->gap_open, sco->gap_extend, sco->match, sco->mismatch, sco->xclip_prefix, sco->xclip_suffix,
                   sco->yclip_prefix, sco->yclip_suffix};
  P.slabs = (unsigned char*)c->d_ws.p;
  P.slab_bytes = slab;
  P.scores = d_scores; P.status = d_status; P.bounds = d_bounds;
  P.pool = (uint32_t*)c->d_b_pool.p; P.pool_cap = pool_cap; P.pool_used = d_tail;
  P.out_off = d_outoff; P.out_len = d_outlen;
  k4_bio_align<<<(unsigned)ctas, 256, 0, c->stream>>>(P);
  CU(cudaGetLastError());
  c->stats.launches++;
  CU(cudaMemcpyAsync(scores, d_scores, n * 4, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaMemcpyAsync(status, d_status, n * 4, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaMemcpyAsync(bounds, d_bounds, n * 16, cudaMemcpyDeviceToHost, c->stream));
  if (ops_off) {
    std::vector<uint32_t> h_len(n);
    std::vector<unsigned long long> h_off(n);
    CU(cudaMemcpyAsync(h_len.data(), d_outlen, n * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(h_off.data(), d_outoff, n * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    size_t total = 0;
    for (size_t i = 0; i < n; ++i) total += (status[i] == 0) ? h_len[i] : 0;
    if (total > ops_cap) return fail(c, SFC_ERR_CAPACITY, "ops_cap %zu < %zu words", ops_cap, total);
    // one copy of the used part of the pool, scattered into pair order on the host
    unsigned long long used = 0;
    CU(cudaMemcpy(&used, d_tail, 8, cudaMemcpyDeviceToHost));
    const size_t valid = std::min<size_t>((size_t)used, pool_cap);
    c->h_pool.resize(valid);
    if (valid) CU(cudaMemcpy(c->h_pool.data(), c->d_b_pool.p, valid * 4, cudaMemcpyDeviceToHost));
    size_t pos = 0;
    for (size_t i = 0; i < n; ++i) {
      ops_off[i] = pos;
      if (status[i] == 0 && h_len[i]) {
        if (h_off[i] + h_len[i] > valid) return fail(c, SFC_ERR_INTERNAL, "pair %zu: ops outside the pool", i);
        memcpy(ops + pos, c->h_pool.data() + h_off[i], (size_t)h_len[i] * 4);
        pos += h_len[i];
      }
    }
    ops_off[n] = pos;
  }
  CU(cudaStreamSynchronize(c->stream));
  return SFC_OK;
}

// ============================================================================ K5: rust-bio style aligner, score only
extern "C" int sfc_bio_score_batch(sfc_ctx* c, const sfc_bio_scoring* sco, const uint8_t* arena, size_t arena_len,
                                   const sfc_pair* pairs, size_t n, int32_t* scores, int32_t* status) {
  if (!c) return SFC_ERR_INVALID;
  if (!sco || !arena || !pairs || n == 0 || n > 0x7FFFFFF0u || !scores || !status) return fail(c, SFC_ERR_INVALID, "null/empty batch");
  if (sco->gap_open > 0 || sco->gap_extend > 0) return fail(c, SFC_ERR_INVALID, "gap penalties must be <= 0");
  const bool xs_on = sco->xclip_suffix > K4_MIN_SCORE / 2;
  if (xs_on || sco->xclip_prefix > 0 || sco->yclip_prefix > 0 || sco->yclip_suffix > 0) {
    // x-suffix clipping (new_pattern_suffix_aligner) needs the per-column maximum K4 keeps: same scores, slower kernel
    std::vector<int32_t> bounds(4 * n);
    return sfc_bio_align_batch(c, sco, arena, arena_len, pairs, n, scores, status, bounds.data(), nullptr, 0, nullptr);
  }
  CU(cudaSetDevice(c->device));
  std::vector<K4Pair> hp(n);
  std::vector<uint32_t> order(n);
  size_t max_n = 0;
  for (size_t i = 0; i < n; ++i) {
    const sfc_pair& p = pairs[i];
    if (p.pattern_len == 0 || p.text_len == 0) return fail(c, SFC_ERR_INVALID, "pair %zu: empty sequence", i);  // assert!(!text.is_empty())
    if (!in_arena(p.pattern_off, p.pattern_len, arena_len) || !in_arena(p.text_off, p.text_len, arena_len))
      return fail(c, SFC_ERR_INVALID, "pair %zu outside the arena", i);
    if (p.pattern_len > 1000000u || p.text_len > 1000000u) return fail(c, SFC_ERR_UNSUPPORTED, "pair %zu: sequence too long", i);
    hp[i] = K4Pair{p.pattern_off, p.text_off, (int)p.pattern_len, (int)p.text_len};
    max_n = std::max<size_t>(max_n, p.text_len);
    order[i] = (uint32_t)i;
  }
  std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
    return (uint64_t)hp[a].m * (uint64_t)hp[a].n > (uint64_t)hp[b].m * (uint64_t)hp[b].n;
  });
  const size_t stride = ((max_n + 1 + 31) & ~(size_t)31) + 32;
  const size_t per_cta = (size_t)K5_NBUF * stride * sizeof(int2);
  const size_t free_b = device_free_bytes(c);
  size_t avail = (size_t)((double)(free_b + c->d_ws.cap) * 0.8);
  if (c->ws_limit && avail > c->ws_limit) avail = c->ws_limit;
  if (per_cta > avail) return fail(c, SFC_ERR_OOM, "device workspace too small: need %zu B for one pair", per_cta);
  const size_t ctas = std::min<size_t>(std::min<size_t>(n, (size_t)c->sm_count * 4), avail / per_cta);
  int rc;
  if ((rc = ensure(c, c->d_ws, per_cta * ctas))) return rc;
  if ((rc = ensure(c, c->d_b_arena, arena_len))) return rc;
  if ((rc = ensure(c, c->d_b_pairs, n * sizeof(K4Pair)))) return rc;
  // out block: scores[n] status[n] order[n] | cells (u64), counter
  const size_t out_bytes = ((3 * n * 4 + 15) & ~(size_t)15) + 32;
  if ((rc = ensure(c, c->d_b_out, out_bytes))) return rc;
  unsigned char* ob = (unsigned char*)c->d_b_out.p;
  int32_t* d_scores = (int32_t*)ob;
  int32_t* d_status = d_scores + n;
  uint32_t* d_order = (uint32_t*)(d_status + n);
  unsigned long long* d_tail = (unsigned long long*)(ob + ((3 * n * 4 + 15) & ~(size_t)15));  // [0] cells, [1] work counter
  c->stats = sfc_batch_stats{};
  CU(cudaEventRecord(c->ev[0], c->stream));
  CU(cudaMemcpyAsync(c->d_b_arena.p, arena, arena_len, cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemcpyAsync(c->d_b_pairs.p, hp.data(), n * sizeof(K4Pair), cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemcpyAsync(d_order, order.data(), n * 4, cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemsetAsync(d_tail, 0, 16, c->stream));
  CU(cudaEventRecord(c->ev[1], c->stream));
  K5Params P;
  P.arena = (const uint8_t*)c->d_b_arena.p;
  P.pairs = (const K4Pair*)c->d_b_pairs.p;
  P.order = d_order;
  P.n = (uint32_t)n;
  P.counter = (uint32_t*)(d_tail + 1);
  P.sc = K4Scoring{sco->gap_open, sco->gap_extend, sco->match, sco->mismatch, sco->xclip_prefix, sco->xclip_suffix,
                   sco->yclip_prefix, sco->yclip_suffix};
  P.bnd = (int2*)c->d_ws.p;
  P.bnd_stride = stride;
  P.scores = d_scores; P.status = d_status; P.cells = d_tail;
  const unsigned grid = (unsigned)ctas;
  if (sco->xclip_prefix > K4_MIN_SCORE / 2) k5_bio_score<true><<<grid, K5_WARPS * 32, 0, c->stream>>>(P);
  else k5_bio_score<false><<<grid, K5_WARPS * 32, 0, c->stream>>>(P);
  CU(cudaGetLastError());
  CU(cudaEventRecord(c->ev[2], c->stream));
  unsigned long long h_cells = 0;
  CU(cudaMemcpyAsync(scores, d_scores, n * 4, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaMemcpyAsync(status, d_status, n * 4, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaMemcpyAsync(&h_cells, d_tail, 8, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaEventRecord(c->ev[3], c->stream));
  CU(cudaStreamSynchronize(c->stream));
  float ms = 0;
  CU(cudaEventElapsedTime(&ms, c->ev[0], c->ev[1])); c->stats.h2d_ms = ms;
  CU(cudaEventElapsedTime(&ms, c->ev[1], c->ev[2])); c->stats.kernel_ms = ms;
  CU(cudaEventElapsedTime(&ms, c->ev[2], c->ev[3])); c->stats.d2h_ms = ms;
  c->stats.h2d_bytes = arena_len + n * (sizeof(K4Pair) + 4);
  c->stats.d2h_bytes = n * 8 + 8;
  c->stats.cells = h_cells;  // full-matrix cells |x| * |y|
  c->stats.launches = 1;
  return SFC_OK;
}
