// sfc_k_align.cuh — K2: ends-free wavefront alignment WITH backtrace (gap-affine 2-piece or gap-linear).
//
// Replaces WFA2-lib's WFAlignerGapAffine2Pieces / WFAlignerGapLinear in AlignmentScope::Alignment as
// sawfish's refine_sv drives them: PairwiseAligner::new_large_indel_aligner + set_text + align
// (reference src/wfa2_utils.rs:149-175,198-219), called per contig at
// src/refine_sv/align/mod.rs:439-451 and :2027-2049 where the full CIGAR is consumed.
//
// Mapping: one CTA per pair (persistent CTAs pull pairs from a global work counter).
//   * sequences: 4-bit, overlapped 8-byte windows in shared memory (one LDS.64 + one funnel shift per
//     8-base compare; warp-cooperative 256-base rounds for long match runs);
//   * wavefront rings (M depth max(x',o1'+e1',o2'+e2')+1, I1/D1 depth e1'+1, I2/D2 depth e2'+1) live in
//     the CTA's slab of device memory (HBM/L2) — a 20 kb x 20 kb pair needs ~22 MB of ring; a warp's chunk of
//     every source row is staged global -> shared with cp.async one chunk ahead of the arithmetic;
//   * reads of source wavefronts are range-guarded from per-wavefront (lo,hi) metadata kept in shared
//     memory, so rings are never initialised or cleared; trimming of every component to its first/last
//     in-matrix diagonal is metadata only (warp 0 probes the row ends between the two barriers of a score);
//   * everything a score step needs besides the cells (ring slots, ranges, chunk classification) is planned once
//     per CTA by warp 0, one lane per ring row, and published in shared memory;
//   * backtrace: instead of keeping every wavefront (WFA2 MemoryHigh) the kernel records ONE choice byte per
//     (score, diagonal) at compute time using WFA2's tie-break priority
//     (X > D2 > D1 > I2 > I1, extend >= open), walks the bytes back to score 0 (also yields k0 for the
//     score formula), then replays forward with greedy match extension to emit run-length ops.
//     This is the same path WFA2's offset-comparing backtrace takes (validated against the oracle's
//     LITERAL mode).
// Bound by integer ALU issue + L2/HBM traffic of the rings (~50 B per diagonal per score); no tensor cores.
#pragma once
#include <cooperative_groups.h>

#include "sfc_device.cuh"

namespace sfc {
namespace cg = cooperative_groups;

enum { K2_SEL_X = 0, K2_SEL_I1 = 1, K2_SEL_I2 = 2, K2_SEL_D1 = 3, K2_SEL_D2 = 4 };
constexpr uint32_t K2_BIT_I1E = 0x08, K2_BIT_I2E = 0x10, K2_BIT_D1E = 0x20, K2_BIT_D2E = 0x40;
enum { K2_E_X = 0, K2_E_IO = 1, K2_E_IE = 2, K2_E_DO = 3, K2_E_DE = 4 };  // backward edit codes
enum { K2_C_M = 0, K2_C_I1 = 1, K2_C_D1 = 2, K2_C_I2 = 3, K2_C_D2 = 4 };
// per-score plan in shared memory (ints from misc + 4), see k2_align
enum { K2_D_S = 0, K2_D_STATUS, K2_D_LO, K2_D_HI, K2_D_HAS2, K2_D_G1A, K2_D_G1B, K2_D_G2A, K2_D_G2B, K2_D_M2A, K2_D_M2B,
       K2_D_Q = 12 /* 12 row offsets */, K2_D_RLO = 24 /* 7 */, K2_D_RCNT = 31 /* 7 */, K2_D_OSLOT = 38 /* 5 */,
       K2_D_CLO = 43 /* 5: raw (untrimmed) range of each output component, */, K2_D_CHI = 48 /* 5 */,
       K2_D_WIDE = 220 /* 10: width of the out-of-matrix run each trim scan crossed at the previous score */,
       K2_D_CAND = 230 /* 10: first hit of the cooperative wide scans */, K2_D_ANYW = 240 /* any K2_D_WIDE > 3 */,
       K2_D_BT = 58 /* index from misc: two 64-bit counters */, K2_D_TAB = 64 /* index from misc: plan table [32 lanes][5] */ };
constexpr int K2_MISC_BYTES = 1024;  // misc area behind the range metadata: pair index, terminating diagonal, plan

struct K2Params {
  const DevPair* pairs;
  const uint32_t* order;
  uint32_t n;
  uint32_t* counter;
  const uint32_t* nib;
  int32_t* scores;
  int32_t* status;
  int metric, x, o1, e1, o2, e2, match;  // adjusted penalties (linear: e1 = indel), original match
  int nM, nG1, nG2;                      // ring depths
  unsigned char* slabs;
  unsigned long long slab_bytes;
  int want_cigar;
  uint32_t* pool;  // run-length ops pool
  unsigned long long pool_cap;
  unsigned long long* pool_used;
  unsigned long long* out_off;  // [n] offset of pair's runs in pool
  uint32_t* out_len;            // [n]
  unsigned long long* cells;
  int win_cap;  // uint2 windows available in shared memory for both sequences together
  int stages;   // cp.async stages per warp for the interior chunks (0 = load straight from global; int32 rings only)
  int max_score;
  unsigned long long* prof;  // optional phase timers (clock64 of thread 0): [0] cell loop [1] barriers+trim [2] backtrace [3] setup
  // K2 v2 (sfc_k_align2.cuh): the wavefront history of every resident pair lives in pages of one shared pool
  unsigned char* pages;  // n_pages pages of (1 << page_shift) bytes
  int page_shift, n_pages;
  int* pgstack;         // [0] lock, [1] free pages, [2] pages reserved by the resident pairs, [4 ..] free page ids
  const uint32_t* pair_pages;  // [pair index] pages the pair is expected to need (admission: see k2v2_align)
};

struct K2Range {
  int lo, hi;
  __device__ __forceinline__ bool any() const { return lo <= hi; }
  __device__ __forceinline__ int cnt() const { return max(hi - lo + 1, 0); }
};

__device__ __forceinline__ int32_t k2_guard(const int32_t* __restrict__ row, int k, int lo, int cnt) {
  return ((unsigned)(k - lo) < (unsigned)cnt) ? row[k] : WNULL;
}
__device__ __forceinline__ bool k2_in_matrix(int32_t off, int k, int plen, int tlen) {
  return (unsigned)off <= (unsigned)tlen && (unsigned)(off - k) <= (unsigned)plen;
}
__device__ __forceinline__ long long k2_gap_cost(const K2Params& P, int L) {
  if (L <= 0) return 0;
  if (P.metric == 0) return (long long)P.e1 * L;
  const long long a = (long long)P.o1 + (long long)P.e1 * L, b = (long long)P.o2 + (long long)P.e2 * L;
  return a < b ? a : b;
}

// Ring element type: int16 when both sequences are <= K2_MAX_LEN16 bases (halves the L2/HBM traffic of the rings,
// which is what bounds the kernel), int32 otherwise.  Stored values are clamped: negative -> null, and
// out-of-matrix insertion offsets (which only ever grow) saturate at a value that is still out of matrix.
constexpr int K2_MAX_LEN16 = 32000;
template <typename T> struct K2Off;
template <> struct K2Off<int32_t> {
  static __device__ __forceinline__ int32_t enc(int32_t v) { return v; }
};
template <> struct K2Off<int16_t> {
  static __device__ __forceinline__ int16_t enc(int32_t v) { return (int16_t)(v < 0 ? -16384 : min(v, 32767)); }
};

template <typename T> struct K2Vec;
template <> struct K2Vec<int32_t> {
  static __device__ __forceinline__ void load(const int32_t* p, int32_t (&o)[4]) {
    const int4 v = *reinterpret_cast<const int4*>(p);
    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
  }
  static __device__ __forceinline__ void store(int32_t* p, const int32_t (&o)[4]) {
    *reinterpret_cast<int4*>(p) = make_int4(o[0], o[1], o[2], o[3]);
  }
};
template <> struct K2Vec<int16_t> {
  static __device__ __forceinline__ void load(const int16_t* p, int32_t (&o)[4]) {
    const short4 v = *reinterpret_cast<const short4*>(p);
    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
  }
  static __device__ __forceinline__ void store(int16_t* p, const int32_t (&o)[4]) {
    short4 v;
    v.x = K2Off<int16_t>::enc(o[0]); v.y = K2Off<int16_t>::enc(o[1]); v.z = K2Off<int16_t>::enc(o[2]); v.w = K2Off<int16_t>::enc(o[3]);
    *reinterpret_cast<short4*>(p) = v;
  }
};


// Continuation of a match run that already covered its first 8-base window (rare): one more private window,
// then warp-cooperative 256-base rounds.  Warp-uniform call; `go` selects the lanes that continue.
__device__ __forceinline__ int k2_extend_more(const uint2* __restrict__ wp, const uint2* __restrict__ wt, int plen, int tlen,
                                              int k, int off, bool go, int lane) {
  bool pending = false;
  if (go) {
    const uint32_t x2 = window8(wp, off - k) ^ window8(wt, off);
    if (x2) off += lead_eq(x2);
    else { off += 8; pending = true; }
  }
  unsigned pm = __ballot_sync(FULL, pending);
  while (pm) {
    const int src = __ffs((int)pm) - 1;
    pm &= pm - 1;
    const int ko = __shfl_sync(FULL, k, src), ho = __shfl_sync(FULL, off, src);
    const int adv = coop_extend(wp, wt, plen, tlen, ko, ho, lane);
    if (lane == src) off += adv;
  }
  return off;
}

template <typename T>
struct K2Rows {  // 32-bit element offsets (from R) of the source / destination ring rows of one score step
  T* R;
  unsigned pMX, pMO1, pMO2, pI1E, pD1E, pI2E, pD2E;
  unsigned oM, oI1, oD1, oI2, oD2;
};

// One interior chunk (128 diagonals per warp, 4 per thread): every lane is active and every read of every source
// lies inside that source's live range, so there are no guards and no predicates — straight-line code.
__device__ __forceinline__ void k2_prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// HAS2: piece-2 sources exist at this score; MO2: this chunk lies inside the live range of M[s-o2'-e2'] (outside it the
// gap-open-2 candidates are null and that row is neither read nor prefetched).
// ---- shared-memory staging of the interior chunks (int32 rings) ----
// A warp's chunk needs, from each of up to 7 source rows, its 128 elements plus one neighbour on each side.  The
// 16-byte units [e0w-4, e0w+132) of every row are copied global -> shared with cp.async (LDGSTS, L2 -> smem, no
// registers held across the latency) one or two chunks ahead of the compute, so the ~1 us ring-load latency that
// used to stall every chunk (45% of all warp samples: long scoreboard) overlaps the previous chunk's arithmetic.
constexpr int K2_SE = 136;                                     // staged elements per row: 4 + 128 + 4 (544 bytes, see k2_stage_issue)
enum { K2_SR_MX = 0, K2_SR_MO1, K2_SR_I1E, K2_SR_D1E, K2_SR_I2E, K2_SR_D2E, K2_SR_MO2, K2_SR_N };
__device__ __forceinline__ void k2_cp16(const int32_t* smem_dst, const int32_t* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void k2_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void k2_cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void k2_cp_wait_n(int n) {  // wait until at most n of this thread's groups are pending
  if (n <= 0) k2_cp_wait<0>(); else if (n == 1) k2_cp_wait<1>(); else k2_cp_wait<2>();
}
// lane L copies its own group (unit L+1) of every needed row; lanes 0 / 31 also copy the unit left / right of the chunk.
// All copies of a chunk are issued from ONE asm block each (own rows / edge units) so that every source address
// sits in its own register pair: ptxas otherwise recycles one pair, and each LDGSTS then waits (long scoreboard) for
// its predecessor to release it.
template <bool AFF>
__device__ __forceinline__ void k2_stage_issue(int32_t* stg, const K2Rows<int32_t>& q, unsigned e0w, bool has2, bool mo2, int lane) {
  const unsigned eo = e0w + ((unsigned)lane << 2);
  const unsigned d = (unsigned)__cvta_generic_to_shared(stg + 4 + (lane << 2));
  const int32_t* R = q.R;
  if (AFF) {
    asm volatile(
        "{\n .reg .pred p2, pm;\n setp.ne.b32 p2, %8, 0;\n setp.ne.b32 pm, %9, 0;\n"
        " cp.async.cg.shared.global [%0], [%1], 16;\n"
        " cp.async.cg.shared.global [%0+544], [%2], 16;\n"
        " cp.async.cg.shared.global [%0+1088], [%3], 16;\n"
        " cp.async.cg.shared.global [%0+1632], [%4], 16;\n"
        " @p2 cp.async.cg.shared.global [%0+2176], [%5], 16;\n"
        " @p2 cp.async.cg.shared.global [%0+2720], [%6], 16;\n"
        " @pm cp.async.cg.shared.global [%0+3264], [%7], 16;\n}"
        ::"r"(d), "l"(R + (q.pMX + eo)), "l"(R + (q.pMO1 + eo)), "l"(R + (q.pI1E + eo)), "l"(R + (q.pD1E + eo)),
          "l"(R + (q.pI2E + eo)), "l"(R + (q.pD2E + eo)), "l"(R + (q.pMO2 + eo)), "r"((int)has2), "r"((int)(has2 && mo2))
        : "memory");
    const bool left = lane == 0;
    if ((left && e0w >= 4u) || lane == 31) {  // the leftmost boundary chunk can start at element 0 of a row
      const unsigned xo = left ? (unsigned)-4 : 4u;
      const unsigned r1 = left ? K2_SR_I1E : K2_SR_D1E, r2 = left ? K2_SR_I2E : K2_SR_D2E;
      const unsigned p1 = left ? q.pI1E : q.pD1E, p2 = left ? q.pI2E : q.pD2E;
      const unsigned dx = d + xo * 4u;
      asm volatile(
          "{\n .reg .pred p2, pm;\n setp.ne.b32 p2, %8, 0;\n setp.ne.b32 pm, %9, 0;\n"
          " cp.async.cg.shared.global [%0], [%1], 16;\n"
          " cp.async.cg.shared.global [%2], [%3], 16;\n"
          " @p2 cp.async.cg.shared.global [%4], [%5], 16;\n"
          " @pm cp.async.cg.shared.global [%6], [%7], 16;\n}"
          ::"r"(dx + K2_SR_MO1 * K2_SE * 4u), "l"(R + (q.pMO1 + eo + xo)), "r"(dx + r1 * K2_SE * 4u), "l"(R + (p1 + eo + xo)),
            "r"(dx + r2 * K2_SE * 4u), "l"(R + (p2 + eo + xo)), "r"(dx + K2_SR_MO2 * K2_SE * 4u), "l"(R + (q.pMO2 + eo + xo)),
            "r"((int)has2), "r"((int)(has2 && mo2))
          : "memory");
    }
  } else {
    asm volatile(
        "{\n cp.async.cg.shared.global [%0], [%1], 16;\n cp.async.cg.shared.global [%0+544], [%2], 16;\n}"
        ::"r"(d), "l"(R + (q.pMX + eo)), "l"(R + (q.pMO1 + eo)) : "memory");
    const bool left = lane == 0;
    if ((left && e0w >= 4u) || lane == 31) {
      const unsigned xo = left ? (unsigned)-4 : 4u;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + xo * 4u + K2_SR_MO1 * K2_SE * 4u), "l"(R + (q.pMO1 + eo + xo)) : "memory");
    }
  }
}
__device__ __forceinline__ void k2_lds4(const int32_t* p, int32_t (&o)[4]) {
  const int4 v = *reinterpret_cast<const int4*>(p);
  o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
}

// STG: the sources were staged in shared memory at `stg` (this warp's stage, see k2_stage_issue) instead of being
// loaded from global here.
template <bool AFF, bool HAS2, bool MO2, bool STG, typename T>
__device__ __forceinline__ void k2_fast_chunk(unsigned char* __restrict__ btq, const K2Rows<T>& q, const int32_t* __restrict__ stg,
                                              unsigned e0, unsigned pf, int k0, int lane,
                                              const uint2* __restrict__ wp, const uint2* __restrict__ wt,
                                              int plen, int tlen, int pef, int tef, int& term_k) {
  int32_t mx[4], m1[4], best[4];
  uint32_t ch[4];
  const int32_t* sb = STG ? stg + 4 + (lane << 2) : nullptr;
  if (!STG && pf) {  // pull this warp's NEXT chunk of every source row towards L2 while the current chunk is computed
    k2_prefetch_l2(q.R + (q.pMX + e0 + pf)); k2_prefetch_l2(q.R + (q.pMO1 + e0 + pf));
    if (AFF) {
      k2_prefetch_l2(q.R + (q.pI1E + e0 + pf)); k2_prefetch_l2(q.R + (q.pD1E + e0 + pf));
      if (HAS2) { if (MO2) k2_prefetch_l2(q.R + (q.pMO2 + e0 + pf)); k2_prefetch_l2(q.R + (q.pI2E + e0 + pf)); k2_prefetch_l2(q.R + (q.pD2E + e0 + pf)); }
    }
  }
  int32_t m1l, m1r;
  if (STG) {
    k2_lds4(sb + K2_SR_MX * K2_SE, mx);
    k2_lds4(sb + K2_SR_MO1 * K2_SE, m1);
    m1l = sb[K2_SR_MO1 * K2_SE - 1]; m1r = sb[K2_SR_MO1 * K2_SE + 4];
  } else {
    K2Vec<T>::load(q.R + (q.pMX + e0), mx);
    K2Vec<T>::load(q.R + (q.pMO1 + e0), m1);
    m1l = __shfl_up_sync(FULL, m1[3], 1); m1r = __shfl_down_sync(FULL, m1[0], 1);
    if (lane == 0) m1l = q.R[q.pMO1 + e0 - 1u];
    if (lane == 31) m1r = q.R[q.pMO1 + e0 + 4u];
  }
  int32_t ins1[4], del1[4], ins2[4], del2[4];
  if (AFF) {
    int32_t a1[4], d1[4], a1l, d1r;
    int32_t m2[4], a2[4], d2[4], m2l = WNULL, m2r = WNULL, a2l = WNULL, d2r = WNULL;
    if (STG) {
      k2_lds4(sb + K2_SR_I1E * K2_SE, a1); a1l = sb[K2_SR_I1E * K2_SE - 1];
      k2_lds4(sb + K2_SR_D1E * K2_SE, d1); d1r = sb[K2_SR_D1E * K2_SE + 4];
      if (HAS2) {
        k2_lds4(sb + K2_SR_I2E * K2_SE, a2); a2l = sb[K2_SR_I2E * K2_SE - 1];
        k2_lds4(sb + K2_SR_D2E * K2_SE, d2); d2r = sb[K2_SR_D2E * K2_SE + 4];
        if (MO2) {
          k2_lds4(sb + K2_SR_MO2 * K2_SE, m2);
          m2l = sb[K2_SR_MO2 * K2_SE - 1]; m2r = sb[K2_SR_MO2 * K2_SE + 4];
        } else {
          m2[0] = m2[1] = m2[2] = m2[3] = WNULL;
        }
      }
    } else {
      K2Vec<T>::load(q.R + (q.pI1E + e0), a1);
      K2Vec<T>::load(q.R + (q.pD1E + e0), d1);
      a1l = __shfl_up_sync(FULL, a1[3], 1); d1r = __shfl_down_sync(FULL, d1[0], 1);
      if (lane == 0) a1l = q.R[q.pI1E + e0 - 1u];
      if (lane == 31) d1r = q.R[q.pD1E + e0 + 4u];
      if (HAS2) {
        K2Vec<T>::load(q.R + (q.pI2E + e0), a2);
        K2Vec<T>::load(q.R + (q.pD2E + e0), d2);
        a2l = __shfl_up_sync(FULL, a2[3], 1); d2r = __shfl_down_sync(FULL, d2[0], 1);
        if (lane == 0) a2l = q.R[q.pI2E + e0 - 1u];
        if (lane == 31) d2r = q.R[q.pD2E + e0 + 4u];
        if (MO2) {
          K2Vec<T>::load(q.R + (q.pMO2 + e0), m2);
          m2l = __shfl_up_sync(FULL, m2[3], 1); m2r = __shfl_down_sync(FULL, m2[0], 1);
          if (lane == 0) m2l = q.R[q.pMO2 + e0 - 1u];
          if (lane == 31) m2r = q.R[q.pMO2 + e0 + 4u];
        } else {
          m2[0] = m2[1] = m2[2] = m2[3] = WNULL;
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int32_t i1o = j == 0 ? m1l : m1[j > 0 ? j - 1 : 0], d1o = j == 3 ? m1r : m1[j < 3 ? j + 1 : 3];
      const int32_t i1e = j == 0 ? a1l : a1[j > 0 ? j - 1 : 0], d1e = j == 3 ? d1r : d1[j < 3 ? j + 1 : 3];
      uint32_t c = 0;
      const int32_t vi1 = max(i1o, i1e) + 1; if (i1e >= i1o) c |= K2_BIT_I1E;
      const int32_t vd1 = max(d1o, d1e);     if (d1e >= d1o) c |= K2_BIT_D1E;
      int32_t bb = vi1; uint32_t sel = K2_SEL_I1;
      if (HAS2) {
        const int32_t i2o = j == 0 ? m2l : m2[j > 0 ? j - 1 : 0], d2o = j == 3 ? m2r : m2[j < 3 ? j + 1 : 3];
        const int32_t i2e = j == 0 ? a2l : a2[j > 0 ? j - 1 : 0], d2e = j == 3 ? d2r : d2[j < 3 ? j + 1 : 3];
        const int32_t vi2 = max(i2o, i2e) + 1; if (i2e >= i2o) c |= K2_BIT_I2E;
        const int32_t vd2 = max(d2o, d2e);     if (d2e >= d2o) c |= K2_BIT_D2E;
        if (vi2 >= bb) { bb = vi2; sel = K2_SEL_I2; }
        if (vd1 >= bb) { bb = vd1; sel = K2_SEL_D1; }
        if (vd2 >= bb) { bb = vd2; sel = K2_SEL_D2; }
        ins2[j] = vi2; del2[j] = vd2;
      } else {
        // piece-2 sources do not exist: ins2 = del2 = null.  The reference comparison chain still visits them
        // (null >= null is true for the ext bits; null never wins the max), so only the ext bits are affected.
        c |= K2_BIT_I2E | K2_BIT_D2E;
        if (vd1 >= bb) { bb = vd1; sel = K2_SEL_D1; }
      }
      const int32_t mis = mx[j] + 1;
      if (mis >= bb) { bb = mis; sel = K2_SEL_X; }
      best[j] = bb; ch[j] = c | sel;
      ins1[j] = vi1; del1[j] = vd1;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int32_t vi = (j == 0 ? m1l : m1[j > 0 ? j - 1 : 0]) + 1, vd = j == 3 ? m1r : m1[j < 3 ? j + 1 : 3];
      const int32_t mis = mx[j] + 1;
      int32_t bb = vi; uint32_t sel = K2_SEL_I1;
      if (vd >= bb) { bb = vd; sel = K2_SEL_D1; }
      if (mis >= bb) { bb = mis; sel = K2_SEL_X; }
      best[j] = bb; ch[j] = sel;
    }
  }
  // choice bytes and I/D rows do not depend on the extension: store them first
  *reinterpret_cast<uint32_t*>(btq + e0) = ch[0] | (ch[1] << 8) | (ch[2] << 16) | (ch[3] << 24);
  if (AFF) {
    K2Vec<T>::store(q.R + (q.oI1 + e0), ins1);
    K2Vec<T>::store(q.R + (q.oD1 + e0), del1);
    if (HAS2) { K2Vec<T>::store(q.R + (q.oI2 + e0), ins2); K2Vec<T>::store(q.R + (q.oD2 + e0), del2); }
  }
  // bounds + first 8-base window of the four cells (independent chains)
  uint32_t xw[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int k = k0 + j;
    if (!k2_in_matrix(best[j], k, plen, tlen)) best[j] = WNULL;
    const bool valid = best[j] >= 0;
    xw[j] = window8(wp, valid ? best[j] - k : 0) ^ window8(wt, valid ? best[j] : 0);
  }
  bool more = false, edge = false;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const bool valid = best[j] >= 0;
    if (valid) best[j] += xw[j] ? lead_eq(xw[j]) : 8;
    more = more || (valid && xw[j] == 0u);
  }
  if (__any_sync(FULL, more)) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      best[j] = k2_extend_more(wp, wt, plen, tlen, k0 + j, best[j], best[j] >= 0 && xw[j] == 0u, lane);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) edge = edge || best[j] >= tlen || best[j] - (k0 + j) >= plen;
  if (edge) {  // a cell touches the end of a sequence: full ends-free termination test (rare)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + j, off = best[j];
      if (off >= 0) {
        const int v = off - k;
        if ((off >= tlen && plen - v <= pef) || (v >= plen && tlen - off <= tef)) term_k = min(term_k, k);
      }
    }
  }
  K2Vec<T>::store(q.R + (q.oM + e0), best);
}

template <bool AFF, typename T>
#ifndef K2_LB_T
#define K2_LB_T 384
#define K2_LB_B 1
#endif
__global__ void __launch_bounds__(K2_LB_T, K2_LB_B) k2_align(const K2Params P) {
  extern __shared__ __align__(16) unsigned char k2_smem[];
  // A pair is aligned by one CTA, or by a thread-block cluster of C CTAs (C SMs) for long haplotypes: the
  // CTAs of a cluster stride the diagonals of each wavefront together, share the rings / choice bytes in
  // global memory (cluster-scope release/acquire at each barrier) and exchange the few words of per-score
  // metadata through distributed shared memory.
  cg::cluster_group cluster = cg::this_cluster();
  const int C = (int)cluster.num_blocks(), crank = (int)cluster.block_rank();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, NT = blockDim.x, NW = NT >> 5;
  auto csync = [&]() { if (C > 1) cluster.sync(); else __syncthreads(); };
  const int nslots = AFF ? P.nM + 2 * P.nG1 + 2 * P.nG2 : P.nM;
  uint2* win = reinterpret_cast<uint2*>(k2_smem);
  int* meta = reinterpret_cast<int*>(win + P.win_cap);  // [nslots][2]
  int* misc = meta + 2 * nslots;                        // [0] pair idx, [1] terminating diagonal
  // per-warp cp.async stages behind the fixed part (host: K2 plan_job sizes both)
  constexpr bool CAN_STAGE = sizeof(T) == 4;
  const int NR = AFF ? (int)K2_SR_N : 2;  // staged rows per chunk
  const int S = CAN_STAGE ? P.stages : 0;
  int32_t* wstg = reinterpret_cast<int32_t*>(k2_smem + (((size_t)P.win_cap * 8 + (size_t)nslots * 8 + K2_MISC_BYTES + 15) & ~(size_t)15)) +
                  (size_t)(threadIdx.x >> 5) * S * NR * K2_SE;
  unsigned char* slab = P.slabs + (size_t)(blockIdx.x / C) * P.slab_bytes;
  int* misc0 = C > 1 ? cluster.map_shared_rank(misc, 0) : misc;  // rank 0's copy: work index + terminating diagonal
  unsigned long long cells_acc = 0;
  long long t_cells = 0, t_sync = 0, t_bt = 0, t_setup = 0, t_trim = 0, t_plan = 0, t_mark = clock64();
#define K2_TICK(acc) do { const long long now_ = clock64(); acc += now_ - t_mark; t_mark = now_; } while (0)

  for (;;) {
    csync();
    K2_TICK(t_bt);
    if (crank == 0 && tid == 0) { misc[0] = (int)atomicAdd(P.counter, 1u); misc[1] = misc[2] = INT_MAX; }
    csync();
    const uint32_t wi = (uint32_t)misc0[0];
    if (wi >= P.n) { csync(); break; }  // no CTA may exit while a peer can still read its shared memory
    const uint32_t pi = P.order[wi];
    const DevPair pr = P.pairs[pi];
    const int plen = pr.plen, tlen = pr.tlen;
    const int nwp = (plen >> 3) + 1, nwt = (tlen >> 3) + 1;
    uint2* wp = win;
    uint2* wt = win + nwp;
    // ---- carve the slab ----
    const int Wp = (plen + tlen + 16 + 7) & ~7, kbase = plen + 2;  // rows are vector-aligned and have slack for the edge groups
    long long scap_ll = k2_gap_cost(P, plen) + k2_gap_cost(P, tlen) + 1;
    if (scap_ll > P.max_score) scap_ll = P.max_score;
    const int S_cap = (int)scap_ll;
    const size_t ring_bytes = ((size_t)nslots * Wp * sizeof(T) + 15) & ~(size_t)15;
    const size_t rowoff_bytes = (((size_t)S_cap + 1) * 8 + 15) & ~(size_t)15;
    const size_t rowlo_bytes = (((size_t)S_cap + 1) * 4 + 15) & ~(size_t)15;
    const size_t E_cap = (size_t)plen + tlen + 8;
    const size_t tmp_bytes = (E_cap * 4 + 15) & ~(size_t)15;
    const size_t fixed = ring_bytes + rowoff_bytes + rowlo_bytes + 2 * tmp_bytes;
    int res_status = 0, res_score = 0;
    if (nwp + nwt > P.win_cap) res_status = -5;
    else if (fixed + 1024 > P.slab_bytes) res_status = -6;
    if (res_status != 0) {
      if (crank == 0 && tid == 0) { P.scores[pi] = 0; P.status[pi] = res_status; P.out_len[pi] = 0; P.out_off[pi] = 0; }
      continue;
    }
    T* rings = reinterpret_cast<T*>(slab);
    unsigned long long* rowoff = reinterpret_cast<unsigned long long*>(slab + ring_bytes);
    int* rowlo = reinterpret_cast<int*>(slab + ring_bytes + rowoff_bytes);
    uint32_t* tmp_b = reinterpret_cast<uint32_t*>(slab + ring_bytes + rowoff_bytes + rowlo_bytes);
    uint32_t* tmp_f = tmp_b + (tmp_bytes >> 2);
    unsigned char* bt = slab + fixed;
    const unsigned long long bt_cap = P.slab_bytes - fixed;
    {
      const uint32_t* gp = P.nib + pr.p_woff;
      const uint32_t* gt = P.nib + pr.t_woff;
      for (int i = tid; i < nwp; i += NT) wp[i] = make_uint2(gp[i], gp[i + 1]);
      for (int i = tid; i < nwt; i += NT) wt[i] = make_uint2(gt[i], gt[i + 1]);
    }
    __syncthreads();

    // ---- per-score plan ----
    // Ring slots, source ranges, the wavefront's own range, the chunk classification and the choice-row offset of a
    // score are computed ONCE per CTA by warp 0 (lane r handles ring row r) between the two barriers of the previous
    // score and published in shared memory (D); the other warps only read them.
    const int sbM = 0, sbI1 = P.nM, sbD1 = P.nM + P.nG1, sbI2 = P.nM + 2 * P.nG1, sbD2 = P.nM + 2 * P.nG1 + P.nG2;
    int* D = misc + 4;
    unsigned long long* D64 = reinterpret_cast<unsigned long long*>(misc + K2_D_BT);  // [0] choice bytes used, [1] cells
    // cluster-wide min of terminating diagonals (rank 0's shared memory, reset at pair fetch); two slots used alternately
    // because only barrier 1 is cluster-wide: a CTA may start the next score while a peer still reads this one's slot
    const int ncomp = AFF ? 5 : 1;
    auto plan = [&](int s_from) {  // warp 0: find the next existing score >= s_from and publish its plan
      // per-lane row constants from the table built at pair start: ring depth, first slot, score distance, range shifts
      int* tab = misc + K2_D_TAB + lane * 5;
      const int back = tab[0], sbase = tab[1], depth = tab[2], dpk = tab[3];
      int rel = tab[4];  // (sN - back) mod depth, kept incrementally (an integer modulo costs ~40 dependent instructions)
      const int dl = (int)(signed char)(dpk & 0xFF), dh = (int)(signed char)((dpk >> 8) & 0xFF);
      const int il = (int)(signed char)((dpk >> 16) & 0xFF), ih = (int)(signed char)((dpk >> 24) & 0xFF);
      const bool is_src = lane < (AFF ? 7 : 2), is_out = lane >= 7 && lane < 7 + ncomp;
      unsigned long long bt_used = D64[0];
      for (int sN = s_from;; ++sN) {
        if (sN > S_cap) { if (lane == 0) D[K2_D_STATUS] = -7; break; }
        const int sc = sN - back;
        const int slot = sbase + rel;
        rel = rel + 1 == depth ? 0 : rel + 1;
        int rlo = 1, rhi = -1;
        if (is_src && sc >= 0 && sN > 0) { rlo = meta[2 * slot]; rhi = meta[2 * slot + 1]; }
        const bool any = rlo <= rhi;
        const unsigned anym = __ballot_sync(FULL, any);
        int lo = warp_min(any ? rlo + dl : INT_MAX), hi = warp_max(any ? rhi + dh : INT_MIN);
        if (sN == 0) { lo = -pr.pbf; hi = pr.tbf; }
        if (lo > hi) {  // wavefront sN does not exist: its rows are empty
          if (is_out) { meta[2 * slot] = 1; meta[2 * slot + 1] = -1; }
          __syncwarp();
          continue;
        }
        const bool has2 = AFF && (anym & 0x70u) != 0u;
        // interior 1: every read of every existing-piece source in range; interior 2: everything but M[s-o2'-e2']
        const unsigned set1 = has2 ? 0x7Fu : (AFF ? 0x0Fu : 0x03u), set2 = 0x3Fu;
        const bool in1 = (set1 >> lane) & 1u, in2 = (set2 >> lane) & 1u;
        int ilo = warp_max(in1 ? rlo + il : INT_MIN), ihi = warp_min(in1 ? rhi + ih : INT_MAX);
        int ilo_b = warp_max(in2 ? rlo + il : INT_MIN), ihi_b = warp_min(in2 ? rhi + ih : INT_MAX);
        if ((anym & set1) != set1 || sN == 0) { ilo = 1; ihi = 0; }
        if (!(AFF && has2) || (anym & set2) != set2 || sN == 0) { ilo_b = 1; ihi_b = 0; }
        const int flo = max(ilo, lo), fhi = min(ihi, hi), flo_b = max(ilo_b, lo), fhi_b = min(ihi_b, hi);
        const int g_lo = (lo + kbase) >> 2, g_hi = (hi + kbase) >> 2;
        const unsigned long long bt_row_bytes = 4ull * (unsigned long long)(g_hi - g_lo + 1);
        if (bt_used + bt_row_bytes > bt_cap) { if (lane == 0) D[K2_D_STATUS] = -6; break; }
        const int m2lo = __shfl_sync(FULL, rlo, K2_SR_MO2), m2hi = __shfl_sync(FULL, rhi, K2_SR_MO2);
        if (lane < 12) D[K2_D_Q + lane] = (int)((unsigned)slot * (unsigned)Wp);
        if (is_out) D[K2_D_OSLOT + lane - 7] = slot;
        if (lane < 7) { D[K2_D_RLO + lane] = rlo; D[K2_D_RCNT + lane] = max(rhi - rlo + 1, 0); }
        {  // raw (untrimmed) diagonal range of each output component: where its two sources can be non-null
          const int mo1lo = __shfl_sync(FULL, rlo, K2_SR_MO1), mo1hi = __shfl_sync(FULL, rhi, K2_SR_MO1);
          // lane c+1 (c = 1..4 -> rows I1E, D1E, I2E, D2E hold the extension source of component c)
          const bool g2 = lane == K2_SR_I2E || lane == K2_SR_D2E;
          const int olo = g2 ? m2lo : mo1lo, ohi = g2 ? m2hi : mo1hi;  // the gap-open source M row
          const bool oany = olo <= ohi;
          int clo = any ? (oany ? min(rlo, olo) : rlo) : (oany ? olo : 1), chi = any ? (oany ? max(rhi, ohi) : rhi) : (oany ? ohi : -1);
          if (any || oany) { clo += dl; chi += dl; }  // insertions read k-1 (range shifts by +1), deletions k+1 (-1); dl == dh for these rows
          if (lane >= K2_SR_I1E && lane <= K2_SR_D2E) { D[K2_D_CLO + lane - 1] = max(clo, lo); D[K2_D_CHI + lane - 1] = min(chi, hi); }
          if (lane == 0) { D[K2_D_CLO] = lo; D[K2_D_CHI] = hi; }
        }
        if (lane == 0) {
          D[K2_D_S] = sN; D[K2_D_LO] = lo; D[K2_D_HI] = hi; D[K2_D_HAS2] = has2;
          // chunk kinds as ranges of the chunk's first group gb (chunk diagonals 4*gb - kbase .. + 127)
          const bool e1 = flo + 127 > fhi, e2 = flo_b + 127 > fhi_b;
          D[K2_D_G1A] = e1 ? INT_MAX : (flo + kbase + 3) >> 2;  D[K2_D_G1B] = e1 ? INT_MIN : (fhi - 127 + kbase) >> 2;
          D[K2_D_G2A] = e2 ? INT_MAX : (flo_b + kbase + 3) >> 2; D[K2_D_G2B] = e2 ? INT_MIN : (fhi_b - 127 + kbase) >> 2;
          const bool m2any = m2lo <= m2hi;  // kind 2 must be clear of M[s-o2'-e2'] (gap-open-2 candidates all null)
          D[K2_D_M2A] = m2any ? (m2lo + kbase - 129) >> 2 : INT_MAX;  // gb <= M2A: chunk entirely left of it
          D[K2_D_M2B] = m2any ? ((m2hi + kbase + 1) >> 2) + 1 : INT_MIN;  // gb >= M2B: entirely right of it
        }
        break;
      }
      tab[4] = rel;
      __syncwarp();
    };
    if (warp == 0) {
      if (lane == 0) { D[K2_D_STATUS] = 0; D64[0] = 0; D64[1] = 0; D[K2_D_ANYW] = 0; }
      if (lane < 10) { D[K2_D_WIDE + lane] = 0; D[K2_D_CAND + lane] = (lane & 1) ? INT_MIN : INT_MAX; }
      {  // plan table: rows 0..6 sources in staging order (MX, MO1, I1E, D1E, I2E, D2E, MO2); rows 7..11 outputs (M, I1, D1, I2, D2)
        int comp = K2_C_M, back = 0, dl = 0, dh = 0, il = 0, ih = 0;
        switch (lane) {
          case 0: back = P.x; break;
          case 1: back = AFF ? P.o1 + P.e1 : P.e1; dl = -1; dh = 1; il = 1; ih = -1; break;
          case 2: comp = K2_C_I1; back = P.e1; dl = dh = 1; il = ih = 1; break;
          case 3: comp = K2_C_D1; back = P.e1; dl = dh = -1; il = ih = -1; break;
          case 4: comp = K2_C_I2; back = P.e2; dl = dh = 1; il = ih = 1; break;
          case 5: comp = K2_C_D2; back = P.e2; dl = dh = -1; il = ih = -1; break;
          case 6: back = P.o2 + P.e2; dl = -1; dh = 1; il = 1; ih = -1; break;
          case 8: comp = K2_C_I1; break;
          case 9: comp = K2_C_D1; break;
          case 10: comp = K2_C_I2; break;
          case 11: comp = K2_C_D2; break;
          default: break;
        }
        const int depth = max(comp == K2_C_M ? P.nM : (comp == K2_C_I1 || comp == K2_C_D1) ? P.nG1 : P.nG2, 1);
        const int sbase = comp == K2_C_M ? sbM : comp == K2_C_I1 ? sbI1 : comp == K2_C_D1 ? sbD1 : comp == K2_C_I2 ? sbI2 : sbD2;
        int* tab = misc + K2_D_TAB + lane * 5;
        tab[0] = back; tab[1] = sbase; tab[2] = depth;
        tab[3] = (dl & 0xFF) | ((dh & 0xFF) << 8) | ((il & 0xFF) << 16) | ((ih & 0xFF) << 24);
        tab[4] = (depth - back % depth) % depth;  // (0 - back) mod depth
      }
      __syncwarp();
      plan(0);
    }

    int s = 0, end_k = 0;
    bool done = false;
    K2_TICK(t_setup);
    for (int it = 0;; ++it) {
      int* term_slot = misc0 + 1 + (it & 1);
      __syncthreads();  // barrier 2: plan of the next score (and the trimmed ranges behind it) visible; CTA-local
      K2_TICK(t_sync);
      { const int st = D[K2_D_STATUS]; if (st != 0) { res_status = st; break; } }
      s = D[K2_D_S];
      const int lo = D[K2_D_LO], hi = D[K2_D_HI];
      const bool has2 = D[K2_D_HAS2] != 0;
      const bool anyw_s = D[K2_D_ANYW] != 0;  // read here: warp 0 rewrites it while other warps may still be behind barrier 1
      const int G1A = D[K2_D_G1A], G1B = D[K2_D_G1B], G2A = D[K2_D_G2A], G2B = D[K2_D_G2B], M2A = D[K2_D_M2A], M2B = D[K2_D_M2B];
      const unsigned long long bt_used = D64[0];
      const T* __restrict__ R = rings;
      const unsigned qMX = (unsigned)D[K2_D_Q + 0], qMO1 = (unsigned)D[K2_D_Q + 1], qI1E = (unsigned)D[K2_D_Q + 2], qD1E = (unsigned)D[K2_D_Q + 3];
      const unsigned qI2E = (unsigned)D[K2_D_Q + 4], qD2E = (unsigned)D[K2_D_Q + 5], qMO2 = (unsigned)D[K2_D_Q + 6];
      const unsigned wM = (unsigned)D[K2_D_Q + 7], wI1 = (unsigned)D[K2_D_Q + 8], wD1 = (unsigned)D[K2_D_Q + 9];
      const unsigned wI2 = (unsigned)D[K2_D_Q + 10], wD2 = (unsigned)D[K2_D_Q + 11];
      unsigned char* btrow = bt + bt_used;  // byte index = (k + kbase) - 4*g_lo
      int term_k = INT_MAX;

      // ---- cell loop: each thread owns a group of 4 consecutive diagonals (one 16-byte vector per ring row),
      // a warp a chunk of 128 ----
      const int g_lo = (lo + kbase) >> 2, g_hi = (hi + kbase) >> 2;
      auto in_rng = [&](int k, int row) -> bool { return (unsigned)(k - D[K2_D_RLO + row]) < (unsigned)D[K2_D_RCNT + row]; };
      // vector of a row at my group, plus the element just left / right of it (from the neighbouring lanes)
      auto load_row = [&](unsigned q, unsigned e0, bool ok, bool last_group, bool want_l, bool want_r, int32_t (&v)[4], int32_t& left, int32_t& right) {
        v[0] = v[1] = v[2] = v[3] = WNULL;
        if (ok) K2Vec<T>::load(R + q + e0, v);
        if (want_l) {
          left = __shfl_up_sync(FULL, v[3], 1);
          if (lane == 0) left = (ok && e0 >= 1u) ? (int32_t)R[q + e0 - 1u] : WNULL;
        }
        if (want_r) {
          right = __shfl_down_sync(FULL, v[0], 1);
          if (lane == 31 || last_group) right = ok ? (int32_t)R[q + e0 + 4u] : WNULL;  // lane+1 holds nothing useful
        }
      };
      K2Rows<T> rows;
      rows.R = rings;
      rows.pMX = qMX; rows.pMO1 = qMO1; rows.pMO2 = qMO2; rows.pI1E = qI1E; rows.pD1E = qD1E; rows.pI2E = qI2E; rows.pD2E = qD2E;
      rows.oM = wM; rows.oI1 = wI1; rows.oD1 = wD1; rows.oI2 = wI2; rows.oD2 = wD2;
      unsigned char* btq = btrow - ((size_t)g_lo << 2);  // indexed by the element offset e = k + kbase
      // 1: interior of every source; 2: interior of everything but M[s-o2'-e2'] and clear of it; 0: boundary (guarded)
      auto kind_of = [&](int gb) -> int {
        if (gb >= G1A && gb <= G1B) return 1;
        if (AFF && gb >= G2A && gb <= G2B && (gb <= M2A || gb >= M2B)) return 2;
        return 0;
      };
      const int gstep = C * NW * 32;
      int gi = g_lo + (crank * NW + warp) * 32, is = 0, cs = 0;  // next chunk to stage, its stage, stage of the chunk being computed
      auto stage_next = [&]() {
        if constexpr (CAN_STAGE) {
          // boundary chunks are staged too (their out-of-range elements are masked by the range checks below)
          if (gi <= g_hi && s > 0) k2_stage_issue<AFF>(wstg + is * NR * K2_SE, rows, (unsigned)gi << 2, has2, kind_of(gi) != 2, lane);
          k2_cp_commit();  // one group per chunk slot (possibly empty) keeps the wait distance constant
          gi += gstep;
          is = is + 1 == S ? 0 : is + 1;
        }
      };
      if (S > 0) for (int p = 0; p < S - 1; ++p) stage_next();
      for (int gb = g_lo + (crank * NW + warp) * 32; gb <= g_hi; gb += gstep) {
        const int g = gb + lane;
        const bool gact = g <= g_hi;
        const unsigned e0 = (unsigned)g << 2;
        const int k0 = (int)e0 - kbase;
        const int kind = kind_of(gb);
        const int32_t* stg = nullptr;
        if (S > 0) {
          __syncwarp();  // every lane is done reading the stage that is refilled now
          stage_next();
          k2_cp_wait_n(S - 1);
          __syncwarp();  // the other lanes' copies of this chunk have landed too
          stg = wstg + cs * NR * K2_SE;
          cs = cs + 1 == S ? 0 : cs + 1;
        }
        if (kind == 1) {  // interior chunk
          if (S > 0) {
            if (has2) k2_fast_chunk<AFF, true, true, true, T>(btq, rows, stg, e0, 0u, k0, lane, wp, wt, plen, tlen, pr.pef, pr.tef, term_k);
            else k2_fast_chunk<AFF, false, false, true, T>(btq, rows, stg, e0, 0u, k0, lane, wp, wt, plen, tlen, pr.pef, pr.tef, term_k);
          } else {
            const unsigned pf = (gb + gstep + 31 <= g_hi) ? (unsigned)(gstep << 2) : 0u;  // next chunk of this warp, in elements
            if (has2) k2_fast_chunk<AFF, true, true, false, T>(btq, rows, nullptr, e0, pf, k0, lane, wp, wt, plen, tlen, pr.pef, pr.tef, term_k);
            else k2_fast_chunk<AFF, false, false, false, T>(btq, rows, nullptr, e0, pf, k0, lane, wp, wt, plen, tlen, pr.pef, pr.tef, term_k);
          }
          continue;
        }
        if (AFF && kind == 2) {
          if (S > 0) {
            k2_fast_chunk<AFF, true, false, true, T>(btq, rows, stg, e0, 0u, k0, lane, wp, wt, plen, tlen, pr.pef, pr.tef, term_k);
          } else {
            const unsigned pf = (gb + gstep + 31 <= g_hi) ? (unsigned)(gstep << 2) : 0u;
            k2_fast_chunk<AFF, true, false, false, T>(btq, rows, nullptr, e0, pf, k0, lane, wp, wt, plen, tlen, pr.pef, pr.tef, term_k);
          }
          continue;
        }
        int32_t best[4], ins1[4], ins2[4], del1[4], del2[4];
        uint32_t ch[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { best[j] = ins1[j] = ins2[j] = del1[j] = del2[j] = WNULL; ch[j] = 0; }
        if (s == 0) {
#pragma unroll
          for (int j = 0; j < 4; ++j) { const int k = k0 + j; if (k >= lo && k <= hi) best[j] = max(k, 0); }
        } else {
          const bool lastg = (g == g_hi);
          int32_t mx[4], m1[4], m1l, m1r, a1[4], a1l, d1[4], d1r, dummy = 0;
          int32_t m2[4], m2l = WNULL, m2r = WNULL, a2[4], a2l = WNULL, d2[4], d2r = WNULL;
          if (S > 0) {  // staged: garbage outside the live ranges, masked below (boundary chunks are never `fast`)
            const int32_t* sb = stg + 4 + (lane << 2);
            k2_lds4(sb + K2_SR_MX * K2_SE, mx);
            k2_lds4(sb + K2_SR_MO1 * K2_SE, m1); m1l = sb[K2_SR_MO1 * K2_SE - 1]; m1r = sb[K2_SR_MO1 * K2_SE + 4];
            if (AFF) {
              k2_lds4(sb + K2_SR_I1E * K2_SE, a1); a1l = sb[K2_SR_I1E * K2_SE - 1];
              k2_lds4(sb + K2_SR_D1E * K2_SE, d1); d1r = sb[K2_SR_D1E * K2_SE + 4];
              if (has2) {
                k2_lds4(sb + K2_SR_MO2 * K2_SE, m2); m2l = sb[K2_SR_MO2 * K2_SE - 1]; m2r = sb[K2_SR_MO2 * K2_SE + 4];
                k2_lds4(sb + K2_SR_I2E * K2_SE, a2); a2l = sb[K2_SR_I2E * K2_SE - 1];
                k2_lds4(sb + K2_SR_D2E * K2_SE, d2); d2r = sb[K2_SR_D2E * K2_SE + 4];
              } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) m2[j] = a2[j] = d2[j] = WNULL;
              }
            }
          } else {
          load_row(qMX, e0, gact, lastg, false, false, mx, dummy, dummy);
          load_row(qMO1, e0, gact, lastg, true, true, m1, m1l, m1r);
          }
          if (AFF && S == 0) {
            load_row(qI1E, e0, gact, lastg, true, false, a1, a1l, dummy);
            load_row(qD1E, e0, gact, lastg, false, true, d1, dummy, d1r);
            if (has2) {
              load_row(qMO2, e0, gact, lastg, true, true, m2, m2l, m2r);
              load_row(qI2E, e0, gact, lastg, true, false, a2, a2l, dummy);
              load_row(qD2E, e0, gact, lastg, false, true, d2, dummy, d2r);
            } else {
#pragma unroll
              for (int j = 0; j < 4; ++j) m2[j] = a2[j] = d2[j] = WNULL;
            }
          } else if (!AFF) {
            a1l = d1r = WNULL;
#pragma unroll
            for (int j = 0; j < 4; ++j) a1[j] = d1[j] = m2[j] = a2[j] = d2[j] = WNULL;
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int k = k0 + j;
            int32_t vmx = mx[j];
            int32_t i1o = j == 0 ? m1l : m1[j > 0 ? j - 1 : 0], d1o = j == 3 ? m1r : m1[j < 3 ? j + 1 : 3];
            int32_t i1e = j == 0 ? a1l : a1[j > 0 ? j - 1 : 0], d1e = j == 3 ? d1r : d1[j < 3 ? j + 1 : 3];
            int32_t i2o = j == 0 ? m2l : m2[j > 0 ? j - 1 : 0], d2o = j == 3 ? m2r : m2[j < 3 ? j + 1 : 3];
            int32_t i2e = j == 0 ? a2l : a2[j > 0 ? j - 1 : 0], d2e = j == 3 ? d2r : d2[j < 3 ? j + 1 : 3];
            {  // boundary chunk: every read is checked against its source wavefront's live range
              const bool act = k >= lo && k <= hi;
              vmx = (act && in_rng(k, K2_SR_MX)) ? vmx : WNULL;
              i1o = (act && in_rng(k - 1, K2_SR_MO1)) ? i1o : WNULL;
              d1o = (act && in_rng(k + 1, K2_SR_MO1)) ? d1o : WNULL;
              if (AFF) {
                i1e = (act && in_rng(k - 1, K2_SR_I1E)) ? i1e : WNULL;
                d1e = (act && in_rng(k + 1, K2_SR_D1E)) ? d1e : WNULL;
                i2o = (act && in_rng(k - 1, K2_SR_MO2)) ? i2o : WNULL;
                d2o = (act && in_rng(k + 1, K2_SR_MO2)) ? d2o : WNULL;
                i2e = (act && in_rng(k - 1, K2_SR_I2E)) ? i2e : WNULL;
                d2e = (act && in_rng(k + 1, K2_SR_D2E)) ? d2e : WNULL;
              }
            }
            const int32_t mis = vmx + 1;
            if (AFF) {
              uint32_t c = 0;
              const int32_t vi1 = max(i1o, i1e) + 1; if (i1e >= i1o) c |= K2_BIT_I1E;
              const int32_t vi2 = max(i2o, i2e) + 1; if (i2e >= i2o) c |= K2_BIT_I2E;
              const int32_t vd1 = max(d1o, d1e);     if (d1e >= d1o) c |= K2_BIT_D1E;
              const int32_t vd2 = max(d2o, d2e);     if (d2e >= d2o) c |= K2_BIT_D2E;
              int32_t bb = vi1; uint32_t sel = K2_SEL_I1;
              if (vi2 >= bb) { bb = vi2; sel = K2_SEL_I2; }
              if (vd1 >= bb) { bb = vd1; sel = K2_SEL_D1; }
              if (vd2 >= bb) { bb = vd2; sel = K2_SEL_D2; }
              if (mis >= bb) { bb = mis; sel = K2_SEL_X; }
              best[j] = bb; ch[j] = c | sel;
              ins1[j] = vi1; ins2[j] = vi2; del1[j] = vd1; del2[j] = vd2;
            } else {
              const int32_t vi = i1o + 1, vd = d1o;
              int32_t bb = vi; uint32_t sel = K2_SEL_I1;
              if (vd >= bb) { bb = vd; sel = K2_SEL_D1; }
              if (mis >= bb) { bb = mis; sel = K2_SEL_X; }
              best[j] = bb; ch[j] = sel;
            }
          }
        }
        // ---- bounds, extension along matches, termination test ----
        // First 8-base window of all four cells as straight-line code (independent LDS/SHF/XOR chains overlap);
        // a cell that matches all 8 bases (1 in 65k on unrelated diagonals) takes the slow continuation below.
        uint32_t xw[4];
        bool more = false;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int k = k0 + j;
          if (!k2_in_matrix(best[j], k, plen, tlen)) best[j] = WNULL;
          const bool valid = best[j] >= 0;
          xw[j] = window8(wp, valid ? best[j] - k : 0) ^ window8(wt, valid ? best[j] : 0);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const bool valid = best[j] >= 0;
          const int n = xw[j] ? lead_eq(xw[j]) : 8;
          if (valid) best[j] += n;
          more = more || (valid && xw[j] == 0u);
        }
        if (__any_sync(FULL, more)) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int k = k0 + j;
            int off = best[j];
            bool pending = false;
            if (off >= 0 && xw[j] == 0u) {  // second private window, then warp-cooperative rounds
              const uint32_t x2 = window8(wp, off - k) ^ window8(wt, off);
              if (x2) off += lead_eq(x2);
              else { off += 8; pending = true; }
            }
            unsigned pm = __ballot_sync(FULL, pending);
            while (pm) {
              const int src = __ffs((int)pm) - 1;
              pm &= pm - 1;
              const int ko = __shfl_sync(FULL, k, src), ho = __shfl_sync(FULL, off, src);
              const int adv = coop_extend(wp, wt, plen, tlen, ko, ho, lane);
              if (lane == src) off += adv;
            }
            best[j] = off;
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int k = k0 + j, off = best[j];
          if (off >= 0) {
            const int v = off - k;
            if ((off >= tlen && plen - v <= pr.pef) || (v >= plen && tlen - off <= pr.tef)) term_k = min(term_k, k);
          }
        }
        if (gact) {
          K2Vec<T>::store(rings + wM + e0, best);
          if (s > 0) {
            *reinterpret_cast<uint32_t*>(btrow + (e0 - ((unsigned)g_lo << 2))) = ch[0] | (ch[1] << 8) | (ch[2] << 16) | (ch[3] << 24);
            if (AFF) {
              K2Vec<T>::store(rings + wI1 + e0, ins1);
              K2Vec<T>::store(rings + wD1 + e0, del1);
              if (has2) { K2Vec<T>::store(rings + wI2 + e0, ins2); K2Vec<T>::store(rings + wD2 + e0, del2); }
            }
          }
        }
      }
      K2_TICK(t_cells);
      // ---- barrier 1: rows of score s complete (cluster-wide); did any diagonal terminate? ----
      term_k = warp_min(term_k);
      if (term_k != INT_MAX && lane == 0) atomicMin(term_slot, term_k);
      csync();
      const int tk = *term_slot;
      if (tk != INT_MAX) {
        end_k = tk; done = true;
        if (crank == 0 && tid == 0) { rowoff[s] = bt_used; rowlo[s] = (g_lo << 2) - kbase; D64[1] += (unsigned long long)(hi - lo + 1); }
        break;
      }
      // ---- warp 0 (of every CTA of a cluster, redundantly): trim the rows of score s, then plan the next score ----
      K2_TICK(t_sync);
      // Behind a sequence end (pattern or text exhausted, common when the text window is much longer than the contig)
      // every gap entry is out of matrix on hundreds to thousands of diagonals, and each score's trim has to cross
      // that run again.  A scan that crossed a long run at the previous score is therefore done by ALL warps together
      // (a slice each, one load latency), sized by that run, before warp 0 finishes the trims.
      if (anyw_s) {
        for (int t = 0; t < 2 * ncomp; ++t) {
          const int w = D[K2_D_WIDE + t];
          if (w <= 3) continue;
          const int c = t >> 1, side = t & 1;
          if (!((c == K2_C_M || s > 0) && !((c == K2_C_I2 || c == K2_C_D2) && !has2))) continue;
          const int clo = D[K2_D_CLO + c], chi = D[K2_D_CHI + c];
          const unsigned q = (unsigned)D[K2_D_Q + 7 + c] + (unsigned)kbase;
          const int win = min(w + 64, chi - clo + 1);
          const int per = min((win + NW * 32 - 1) / (NW * 32), 8);  // 32-diagonal blocks per warp
          const int o = warp * per * 32;                              // my slice, in scan order from the edge
          int32_t pv8[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int off = o + 32 * u + lane;
            const int kk = side ? chi - off : clo + off;
            pv8[u] = (u < per && off < win) ? (int32_t)rings[q + (unsigned)kk] : WNULL;
          }
          bool got = false;
          int pos = 0;
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int off = o + 32 * u + lane;
            const int kk = side ? chi - off : clo + off;
            const unsigned b = __ballot_sync(FULL, u < per && off < win && k2_in_matrix(pv8[u], kk, plen, tlen));
            if (b && !got) { got = true; const int fo = o + 32 * u + __ffs((int)b) - 1; pos = side ? chi - fo : clo + fo; }
          }
          if (got && lane == 0) { if (side) atomicMax(D + K2_D_CAND + t, pos); else atomicMin(D + K2_D_CAND + t, pos); }
        }
        __syncthreads();
      }
      if (warp == 0) {
        // trim every component of score s to its first / last in-matrix diagonal (metadata only).  All ten scans
        // (5 components x 2 sides) probe their first three diagonals with ONE load: lane 3t+j reads diagonal j of scan t,
        // starting from the component's raw range.  A scan with no hit there (rare) continues 32 diagonals at a time.
        {
          const int t = lane / 3, j = lane - 3 * t, c = min(t >> 1, 4), side = t & 1;
          const bool en = t < 2 * ncomp && (c == K2_C_M || s > 0) && !((c == K2_C_I2 || c == K2_C_D2) && !has2);
          const int clo = en ? D[K2_D_CLO + c] : 1, chi = en ? D[K2_D_CHI + c] : -1;
          const unsigned q = (unsigned)D[K2_D_Q + 7 + c] + (unsigned)kbase;
          const int k = side ? chi - j : clo + j;
          const int wprev = (t < 10 && anyw_s) ? D[K2_D_WIDE + t] : 0;
          const bool wide = en && wprev > 3;  // scanned by all warps above: first hit (if any) in K2_D_CAND
          const bool ok = !wide && k >= clo && k <= chi;
          const int32_t v = ok ? (int32_t)rings[q + (unsigned)k] : WNULL;
          const unsigned hits = __ballot_sync(FULL, ok && k2_in_matrix(v, k, plen, tlen));
          const unsigned mine = (hits >> (3 * t)) & 7u;  // t <= 10: shift <= 30
          int found = side ? -1 : 1;                      // empty encoding: lo = 1, hi = -1
          bool got = mine != 0u;
          int skip = 3;                                   // diagonals already examined from the edge
          if (mine) found = side ? chi - (__ffs((int)mine) - 1) : clo + __ffs((int)mine) - 1;
          if (wide) {
            const int cand = D[K2_D_CAND + t];
            got = side ? cand != INT_MIN : cand != INT_MAX;
            if (got) found = cand;
            const int win = min(wprev + 64, chi - clo + 1);
            skip = min(win, min((win + NW * 32 - 1) / (NW * 32), 8) * NW * 32);
          }
          unsigned unres = __ballot_sync(FULL, j == 0 && en && !got && chi - clo + 1 > skip);
          if (j == 0 && t < 2 * ncomp) {
            meta[2 * D[K2_D_OSLOT + c] + side] = found;
            // remember how long an out-of-matrix run this scan crossed: the next score's scan is sized by it
            D[K2_D_WIDE + t] = !en ? 0 : got ? (side ? chi - found : found - clo) : chi - clo + 1;
            D[K2_D_CAND + t] = side ? INT_MIN : INT_MAX;
          }
          while (unres) {  // rare: warp-wide continuation of one scan
            const int src = __ffs((int)unres) - 1;
            unres &= unres - 1;
            const int c2 = __shfl_sync(FULL, c, src), side2 = __shfl_sync(FULL, side, src);
            const int lo2 = __shfl_sync(FULL, clo, src), hi2 = __shfl_sync(FULL, chi, src);
            const unsigned q2 = __shfl_sync(FULL, q, src);
            const int skip2 = __shfl_sync(FULL, skip, src), t2 = __shfl_sync(FULL, t, src);
            int f2 = side2 ? -1 : 1;
            // 256 diagonals per round (eight loads in flight): the ranges behind a sequence end, where every gap entry
            // is out of matrix, can be hundreds of diagonals wide and are rescanned at every score
            for (int base = side2 ? hi2 - skip2 : lo2 + skip2; side2 ? base >= lo2 : base <= hi2; base += side2 ? -256 : 256) {
              ++t_setup;
              int32_t pv8[8];
#pragma unroll
              for (int u = 0; u < 8; ++u) {
                const int kk = side2 ? base - 32 * u - lane : base + 32 * u + lane;
                pv8[u] = (kk >= lo2 && kk <= hi2) ? (int32_t)rings[q2 + (unsigned)kk] : WNULL;
              }
              bool got = false;
#pragma unroll
              for (int u = 0; u < 8; ++u) {
                const int kk = side2 ? base - 32 * u - lane : base + 32 * u + lane;
                const unsigned b = __ballot_sync(FULL, kk >= lo2 && kk <= hi2 && k2_in_matrix(pv8[u], kk, plen, tlen));
                if (b && !got) { got = true; f2 = side2 ? base - 32 * u - (__ffs((int)b) - 1) : base + 32 * u + __ffs((int)b) - 1; }
              }
              if (got) break;
            }
            if (lane == 0) {
              meta[2 * D[K2_D_OSLOT + c2] + side2] = f2;
              if (side2 ? f2 >= lo2 : f2 <= hi2) D[K2_D_WIDE + t2] = side2 ? hi2 - f2 : f2 - lo2;  // found (else the whole range was recorded)
            }
          }
          __syncwarp();
          const bool anyw = lane < 2 * ncomp && D[K2_D_WIDE + min(lane, 9)] > 3;
          const unsigned aw = __ballot_sync(FULL, anyw);
          if (lane == 0) D[K2_D_ANYW] = aw != 0u;
        }
        if (lane == 0) {
          if (crank == 0) { rowoff[s] = bt_used; rowlo[s] = (g_lo << 2) - kbase; }
          D64[1] += (unsigned long long)(hi - lo + 1);
          if (s > 0) D64[0] = bt_used + 4ull * (unsigned long long)(g_hi - g_lo + 1);
        }
        __syncwarp();
        K2_TICK(t_trim);
        plan(s + 1);
        K2_TICK(t_plan);
      }
    }

    K2_TICK(t_sync);
    if (crank == 0 && tid == 0) cells_acc += D64[1];
    // ---- backtrace (warp 0): choice walk to score 0, then forward replay ----
    if (done) {
      csync();  // rowoff/rowlo/bt/oM of the last step visible to warp 0 of rank 0
      if (crank == 0 && warp == 0) {
        const T* oM = rings + (unsigned)D[K2_D_Q + 7] + kbase;
        const int end_off = oM[end_k];
        int k0 = 0, nb = 0, bad = 0;
        if (lane == 0) {
          int sc = s, k = end_k, comp = K2_C_M;
          int last_code = -1;
          uint32_t run = 0;
          auto push = [&](int code) {
            if (code == last_code) { ++run; return; }
            if (run) { if ((size_t)nb < E_cap) tmp_b[nb++] = (run << 3) | (uint32_t)last_code; else bad = 1; }
            last_code = code; run = 1;
          };
          while (sc > 0 && !bad) {
            const uint32_t c = bt[rowoff[sc] + (unsigned long long)(k - rowlo[sc])];
            if (!AFF) {
              const uint32_t sel = c & 7u;
              if (sel == K2_SEL_X) { push(K2_E_X); sc -= P.x; }
              else if (sel == K2_SEL_I1) { push(K2_E_IO); sc -= P.e1; k -= 1; }
              else { push(K2_E_DO); sc -= P.e1; k += 1; }
              continue;
            }
            switch (comp) {
              case K2_C_M: {
                const uint32_t sel = c & 7u;
                if (sel == K2_SEL_X) { push(K2_E_X); sc -= P.x; }
                else if (sel == K2_SEL_I1) comp = K2_C_I1;
                else if (sel == K2_SEL_I2) comp = K2_C_I2;
                else if (sel == K2_SEL_D1) comp = K2_C_D1;
                else comp = K2_C_D2;
                break;
              }
              case K2_C_I1:
                if (c & K2_BIT_I1E) { push(K2_E_IE); sc -= P.e1; } else { push(K2_E_IO); sc -= P.o1 + P.e1; comp = K2_C_M; }
                k -= 1;
                break;
              case K2_C_I2:
                if (c & K2_BIT_I2E) { push(K2_E_IE); sc -= P.e2; } else { push(K2_E_IO); sc -= P.o2 + P.e2; comp = K2_C_M; }
                k -= 1;
                break;
              case K2_C_D1:
                if (c & K2_BIT_D1E) { push(K2_E_DE); sc -= P.e1; } else { push(K2_E_DO); sc -= P.o1 + P.e1; comp = K2_C_M; }
                k += 1;
                break;
              default:
                if (c & K2_BIT_D2E) { push(K2_E_DE); sc -= P.e2; } else { push(K2_E_DO); sc -= P.o2 + P.e2; comp = K2_C_M; }
                k += 1;
                break;
            }
          }
          if (run) { if ((size_t)nb < E_cap) tmp_b[nb++] = (run << 3) | (uint32_t)last_code; else bad = 1; }
          if (sc != 0 || comp != K2_C_M) bad = 1;
          k0 = k;
        }
        k0 = __shfl_sync(FULL, k0, 0);
        nb = __shfl_sync(FULL, nb, 0);
        bad = __shfl_sync(FULL, bad, 0);
        __syncwarp();
        const int h_end = end_off, v_end = end_off - end_k;
        const int al_t = h_end - max(k0, 0), al_p = v_end - max(-k0, 0);
        res_score = (P.match == 0) ? -s : ((-P.match) * (al_p + al_t) - s) / 2;
        unsigned long long o_off = 0;
        uint32_t o_len = 0;
        if (P.want_cigar && !bad) {
          // forward replay; lane 0 keeps the run-length output, all lanes extend matches together
          int h = max(k0, 0), v = max(-k0, 0), nf = 0, cur_code = -1;
          uint32_t cur_len = 0;
          auto emit = [&](int code, uint32_t len) {
            if (len == 0) return;
            if (code == cur_code) { cur_len += len; return; }
            if (cur_len && lane == 0) { if ((size_t)nf < E_cap) tmp_f[nf] = (cur_len << 2) | (uint32_t)cur_code; }
            if (cur_len) ++nf;
            cur_code = code; cur_len = len;
          };
          emit(2, (uint32_t)h);
          emit(3, (uint32_t)v);
          for (int i = nb - 1; i >= 0; --i) {
            const uint32_t e = tmp_b[i];
            const int code = (int)(e & 7u);
            const uint32_t cnt = e >> 3;
            if (code == K2_E_IE) { emit(2, cnt); h += (int)cnt; continue; }
            if (code == K2_E_DE) { emit(3, cnt); v += (int)cnt; continue; }
            for (uint32_t j = 0; j < cnt; ++j) {
              const int m = coop_extend(wp, wt, plen, tlen, h - v, h, lane);
              emit(0, (uint32_t)m); h += m; v += m;
              if (code == K2_E_X) { emit(1, 1); ++h; ++v; }
              else if (code == K2_E_IO) { emit(2, 1); ++h; }
              else { emit(3, 1); ++v; }
            }
          }
          {
            const int m = coop_extend(wp, wt, plen, tlen, h - v, h, lane);
            emit(0, (uint32_t)m); h += m; v += m;
          }
          if (h != h_end || v != v_end) bad = 1;
          emit(2, (uint32_t)(tlen - h));
          emit(3, (uint32_t)(plen - v));
          emit(-2, 1);  // flush
          if ((size_t)nf > E_cap) bad = 1;
          if (!bad) {
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(P.pool_used, (unsigned long long)nf);
            base = __shfl_sync(FULL, base, 0);
            __syncwarp();
            if (base + (unsigned long long)nf > P.pool_cap) { res_status = -10; /* SFC_STATUS_POOL */ }
            else {
              for (int i = lane; i < nf; i += 32) P.pool[base + i] = tmp_f[i];
              o_off = base; o_len = (uint32_t)nf;
            }
          }
        }
        if (bad) res_status = -9;
        if (lane == 0) { P.scores[pi] = res_score; P.status[pi] = res_status; P.out_off[pi] = o_off; P.out_len[pi] = o_len; }
      }
    } else if (crank == 0 && tid == 0) {
      P.scores[pi] = 0; P.status[pi] = res_status; P.out_off[pi] = 0; P.out_len[pi] = 0;
    }
  }
  if (crank == 0 && tid == 0 && cells_acc) atomicAdd(P.cells, cells_acc);
  if (tid == 0 && P.prof) {
    K2_TICK(t_bt);
    atomicAdd(P.prof + 0, (unsigned long long)t_cells); atomicAdd(P.prof + 1, (unsigned long long)t_sync);
    atomicAdd(P.prof + 2, (unsigned long long)t_bt); atomicAdd(P.prof + 3, (unsigned long long)t_setup);
    atomicAdd(P.prof + 4, (unsigned long long)t_trim); atomicAdd(P.prof + 5, (unsigned long long)t_plan);
    atomicMax(P.prof + 6, (unsigned long long)(t_cells + t_sync + t_bt + t_setup + t_trim + t_plan));  // busiest CTA
    atomicAdd(P.prof + 7, 1ull);
  }
#undef K2_TICK
}

}  // namespace sfc
