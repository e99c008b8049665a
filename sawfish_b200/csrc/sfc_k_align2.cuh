// sfc_k_align2.cuh — K2 v2: gap-affine 2-piece ends-free wavefront alignment WITH backtrace, wavefront HISTORY kept.
//
// Same contract as k2_align<true, *> (sfc_k_align.cuh; replaces WFA2-lib's WFAlignerGapAffine2Pieces in
// AlignmentScope::Alignment as driven by reference src/wfa2_utils.rs:149-175,198-219 from
// src/refine_sv/align/mod.rs:439-451,2027-2049), different memory design:
//
//   * offsets are 16 bit: stored value = h - 32000 (text position, biased), null = -32768.  Any text up to 64 000
//     bases fits (the pattern may be longer); longer texts go to the int32 kernel.  Eight diagonals per thread, two per
//     register: the recurrence runs on packed s16x2 DPX instructions (VIMNMX / VIMNMX3 / VIADDMNMX .S16x2, 7 per cell
//     instead of 27), only the match extension is scalar;
//   * every wavefront of every component is KEPT (append-only rows, 10 B per cell — what WFA2's MemoryHigh does, laid
//     out for 180 GB of HBM: one pool of 8 MB pages shared by every resident pair, a pair takes pages as its history
//     grows and returns them when it is done) instead of ring rows + one choice byte per cell.  Nothing
//     is recorded at compute time: the backtrace re-derives each decision on its path from the stored source values
//     with the compute-time formula (same inputs, same tie-break order X > D2 > D1 > I2 > I1, extend >= open), a warp
//     at a time (nine candidates of an M step in parallel, 32 speculative steps of a gap run per round trip);
//   * a warp's chunk is 256 diagonals x 7 source rows (+ one 16-byte unit either side), staged global -> shared with
//     cp.async one chunk ahead; interior chunks are straight-line code, boundary chunks a scalar range-checked loop;
//   * the per-score plan / trim machinery (warp 0, one lane per row) is the one of k2_align.
#pragma once
#include "sfc_k_align.cuh"

namespace sfc {

constexpr int V2_BIAS = 32000;      // stored int16 = h - V2_BIAS; multiple of 8 (sequence windows are addressed with it)
constexpr int V2_NULL = -32768;
constexpr int V2_VMIN = -32000;     // smallest valid stored value (h = 0); anything below is null
constexpr int V2_CAP = 32766;       // stored values never exceed this (out-of-matrix insertion offsets saturate here)
constexpr int V2_MAX_TLEN = 64000;  // V2_CAP + V2_BIAS must stay outside the matrix
constexpr int V2_ROWB = 544;        // staged bytes per source row: 34 units of 16 B (unit -1, 32 own units, unit +32)
constexpr int V2_ROWE = 272;        // ... in int16 elements
constexpr int V2_HIST = 16;         // ints per score in the history table
constexpr int V2_MISC_BYTES = 1280;
// plan area, indices from `misc`
enum { V2_S = 4, V2_STATUS, V2_LO, V2_HI, V2_HAS2, V2_G1A, V2_G1B, V2_G2A, V2_G2B, V2_M2A, V2_M2B, V2_RENORM, V2_G0, V2_NU, V2_ANYW,
       V2_Q = 20 /* 7 source + 5 output rows: virtual origin in 16-byte units */, V2_RLO = 32 /* 7 */, V2_RCNT = 39 /* 7 */,
       V2_SG0 = 46 /* 7: first stored unit of each source row */, V2_SNG = 53 /* 7: stored units */, V2_OSLOT = 60 /* 5 */,
       V2_CLO = 65 /* 5 */, V2_CHI = 70 /* 5 */, V2_D64 = 76 /* two u64: units used, cells */, V2_TAB = 80 /* [32][5] */,
       V2_WIDE = 240 /* 10 */, V2_CAND = 250 /* 10 */, V2_PGSEQ = 260 /* 4: ring of this pair's page ids (rank 0's copy is read) */,
       V2_PGI = 264 /* index of the current page in that sequence */, V2_PGN = 265 /* pages acquired (rank 0) */, V2_NEED = 266 /* 64-byte units of this score's rows */,
       V2_NUP = 267 /* 64-byte units per row (padded) */, V2_RSV = 268 /* pages reserved for the current pair (rank 0) */ };

// ---- packed s16x2 helpers: DPX on the device, plain C++ on the host (tests/cpp/test_v2_recur.cu) ----
#if defined(__CUDA_ARCH__)
#define V2_UNROLL _Pragma("unroll")
__device__ __forceinline__ uint32_t v2_max(uint32_t a, uint32_t b) { return __vmaxs2(a, b); }
__device__ __forceinline__ uint32_t v2_max3(uint32_t a, uint32_t b, uint32_t c) { return __vimax3_s16x2(a, b, c); }
__device__ __forceinline__ uint32_t v2_inc_cap(uint32_t a) { return __viaddmin_s16x2(a, 0x00010001u, 0x7ffe7ffeu); }  // min(a + 1, 32766)
__device__ __forceinline__ uint32_t v2_inc(uint32_t a) { return __vadd2(a, 0x00010001u); }
__device__ __forceinline__ uint32_t v2_shl(uint32_t lo, uint32_t hi) { return __funnelshift_l(lo, hi, 16); }
#else
#define V2_UNROLL
inline int v2h_lo(uint32_t w) { return (int)(int16_t)(w & 0xFFFFu); }
inline int v2h_hi(uint32_t w) { return (int)(int16_t)(w >> 16); }
inline uint32_t v2h_pack(int lo, int hi) { return ((uint32_t)lo & 0xFFFFu) | ((uint32_t)hi << 16); }
inline int v2h_mx(int a, int b) { return a > b ? a : b; }
inline uint32_t v2_max(uint32_t a, uint32_t b) { return v2h_pack(v2h_mx(v2h_lo(a), v2h_lo(b)), v2h_mx(v2h_hi(a), v2h_hi(b))); }
inline uint32_t v2_max3(uint32_t a, uint32_t b, uint32_t c) { return v2_max(v2_max(a, b), c); }
inline uint32_t v2_inc_cap(uint32_t a) {
  const int l = (int)(int16_t)(v2h_lo(a) + 1), h = (int)(int16_t)(v2h_hi(a) + 1);
  return v2h_pack(l < 32766 ? l : 32766, h < 32766 ? h : 32766);
}
inline uint32_t v2_inc(uint32_t a) { return v2h_pack(v2h_lo(a) + 1, v2h_hi(a) + 1); }
inline uint32_t v2_shl(uint32_t lo, uint32_t hi) { return (hi << 16) | (lo >> 16); }
#endif

// One thread's eight consecutive diagonals e0 .. e0+7 of the seven source rows, two per word (even element in the low
// half), plus the word left of them (elements e0-2, e0-1) for the rows read at k-1 and the word right of them
// (e0+8, e0+9) for the rows read at k+1.
struct V2Src {
  uint32_t mx[4], mo1[4], i1e[4], d1e[4], i2e[4], d2e[4], mo2[4];
  uint32_t mo1L, mo1R, i1eL, d1eR, i2eL, d2eR, mo2L, mo2R;
};

// The affine 2-piece recurrence on packed words.  ins(k) = max(M_open(k-1), I_ext(k-1)) + 1 is formed at the SOURCE
// diagonal (aligned words) and moved one element up by a funnel shift; del(k) = max(M_open(k+1), D_ext(k+1)) likewise
// one element down.  best = max(mismatch + 1, ins1, ins2, del1, del2) before bounds and extension.
template <bool HAS2, bool MO2>
__host__ __device__ __forceinline__ void v2_recur8(const V2Src& s, uint32_t (&best)[4], uint32_t (&ins1)[4], uint32_t (&del1)[4],
                                                   uint32_t (&ins2)[4], uint32_t (&del2)[4]) {
  const uint32_t NULL2 = 0x80008000u;
  {
    uint32_t t[5], u[5];
    t[0] = v2_max(s.mo1L, s.i1eL);
V2_UNROLL
    for (int i = 0; i < 4; ++i) { t[i + 1] = v2_max(s.mo1[i], s.i1e[i]); u[i] = v2_max(s.mo1[i], s.d1e[i]); }
    u[4] = v2_max(s.mo1R, s.d1eR);
V2_UNROLL
    for (int i = 0; i < 4; ++i) { ins1[i] = v2_inc_cap(v2_shl(t[i], t[i + 1])); del1[i] = v2_shl(u[i], u[i + 1]); }
  }
  if (HAS2) {
    uint32_t t[5], u[5];
    if (MO2) {
      t[0] = v2_max(s.mo2L, s.i2eL);
V2_UNROLL
      for (int i = 0; i < 4; ++i) { t[i + 1] = v2_max(s.mo2[i], s.i2e[i]); u[i] = v2_max(s.mo2[i], s.d2e[i]); }
      u[4] = v2_max(s.mo2R, s.d2eR);
    } else {  // this chunk lies clear of M[s-o2'-e2']: the gap-open-2 candidates are null
      t[0] = s.i2eL;
V2_UNROLL
      for (int i = 0; i < 4; ++i) { t[i + 1] = s.i2e[i]; u[i] = s.d2e[i]; }
      u[4] = s.d2eR;
    }
V2_UNROLL
    for (int i = 0; i < 4; ++i) { ins2[i] = v2_inc_cap(v2_shl(t[i], t[i + 1])); del2[i] = v2_shl(u[i], u[i + 1]); }
  } else {
V2_UNROLL
    for (int i = 0; i < 4; ++i) { ins2[i] = NULL2; del2[i] = NULL2; }
  }
V2_UNROLL
  for (int i = 0; i < 4; ++i) best[i] = v2_max3(v2_max3(ins1[i], ins2[i], del1[i]), del2[i], v2_inc(s.mx[i]));
}

#if defined(__CUDACC__)
__device__ __forceinline__ int v2_lo16(uint32_t w) { return (int)(short)(w & 0xFFFFu); }  // sign-extended low half (one PRMT)
__device__ __forceinline__ int v2_hi16(uint32_t w) { return (int)w >> 16; }
__device__ __forceinline__ uint32_t v2_pack(int lo, int hi) { return __byte_perm((uint32_t)lo, (uint32_t)hi, 0x5410); }
// values below V2_VMIN are nulls that crept up by the "+1" of the insertion rows: back to the exact null
__device__ __forceinline__ uint32_t v2_renull(uint32_t w) {
  const int l = v2_lo16(w), h = v2_hi16(w);
  return v2_pack(l < V2_VMIN ? V2_NULL : l, h < V2_VMIN ? V2_NULL : h);
}
#define V2_ROW(R, q, g) ((R) + (((long long)(q) << 2) + (long long)(g)))
__device__ __forceinline__ void v2_cp16(unsigned dst_smem, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(gsrc) : "memory");
}

// ---- page pool (one thread at a time behind a spin lock; a pair takes a page every few dozen scores) ----
__device__ __forceinline__ void v2_pool_lock(int* st) { while (atomicCAS(st, 0, 1) != 0) __nanosleep(64); __threadfence(); }
__device__ __forceinline__ void v2_pool_unlock(int* st) { __threadfence(); atomicExch(st, 0); }
__device__ __forceinline__ int v2_page_acquire(int* st) {  // -1: the pool is empty (the pair fails with SFC_STATUS_OOM and is re-run later)
  v2_pool_lock(st);
  volatile int* v = st;
  const int n = v[1];
  int id = -1;
  if (n > 0) { id = v[3 + n]; v[1] = n - 1; }
  v2_pool_unlock(st);
  return id;
}
__device__ __forceinline__ void v2_pages_release(int* st, const int* list, int cnt) {
  if (cnt <= 0) return;
  v2_pool_lock(st);
  volatile int* v = st;
  int n = v[1];
  for (int i = 0; i < cnt; ++i) { const int id = list[i]; if (id >= 0) { v[4 + n] = id; ++n; } }
  v[1] = n;
  v2_pool_unlock(st);
}

// Row addressing: a row's origin q is in 64-byte units from the pool base (32 bit: 128 GB), a diagonal group g in 16-byte
// units from the origin: address = pool + q * 64 + g * 16.
// Interior chunk: lane L copies unit gb+L of every needed row; lanes 0 / 31 also the unit left / right of the chunk for
// the rows read at k-1 / k+1.  One asm block per group of copies so that every source address has its own register
// pair (see k2_stage_issue).
__device__ __forceinline__ void v2_stage_fast(unsigned char* stg, const int4* __restrict__ R, const int (&q)[7], int gb, bool has2, bool mo2, int lane) {
  const int g = gb + lane;
  const unsigned d = (unsigned)__cvta_generic_to_shared(stg) + 16u + ((unsigned)lane << 4);
  asm volatile(
      "{\n .reg .pred p2, pm;\n setp.ne.b32 p2, %8, 0;\n setp.ne.b32 pm, %9, 0;\n"
      " cp.async.cg.shared.global [%0], [%1], 16;\n"
      " cp.async.cg.shared.global [%0+544], [%2], 16;\n"
      " cp.async.cg.shared.global [%0+1088], [%3], 16;\n"
      " cp.async.cg.shared.global [%0+1632], [%4], 16;\n"
      " @p2 cp.async.cg.shared.global [%0+2176], [%5], 16;\n"
      " @p2 cp.async.cg.shared.global [%0+2720], [%6], 16;\n"
      " @pm cp.async.cg.shared.global [%0+3264], [%7], 16;\n}"
      ::"r"(d), "l"(V2_ROW(R, q[K2_SR_MX], g)), "l"(V2_ROW(R, q[K2_SR_MO1], g)), "l"(V2_ROW(R, q[K2_SR_I1E], g)), "l"(V2_ROW(R, q[K2_SR_D1E], g)),
        "l"(V2_ROW(R, q[K2_SR_I2E], g)), "l"(V2_ROW(R, q[K2_SR_D2E], g)), "l"(V2_ROW(R, q[K2_SR_MO2], g)), "r"((int)has2), "r"((int)(has2 && mo2))
      : "memory");
  if (lane == 0 || lane == 31) {
    const bool left = lane == 0;
    const int xo = left ? -1 : 1;
    const unsigned r1 = left ? K2_SR_I1E : K2_SR_D1E, r2 = left ? K2_SR_I2E : K2_SR_D2E;
    const int q1 = left ? q[K2_SR_I1E] : q[K2_SR_D1E], q2 = left ? q[K2_SR_I2E] : q[K2_SR_D2E];
    const unsigned dx = left ? d - 16u : d + 16u;
    asm volatile(
        "{\n .reg .pred p2, pm;\n setp.ne.b32 p2, %8, 0;\n setp.ne.b32 pm, %9, 0;\n"
        " cp.async.cg.shared.global [%0], [%1], 16;\n"
        " cp.async.cg.shared.global [%2], [%3], 16;\n"
        " @p2 cp.async.cg.shared.global [%4], [%5], 16;\n"
        " @pm cp.async.cg.shared.global [%6], [%7], 16;\n}"
        ::"r"(dx + K2_SR_MO1 * V2_ROWB), "l"(V2_ROW(R, q[K2_SR_MO1], g + xo)), "r"(dx + r1 * V2_ROWB), "l"(V2_ROW(R, q1, g + xo)),
          "r"(dx + r2 * V2_ROWB), "l"(V2_ROW(R, q2, g + xo)), "r"(dx + K2_SR_MO2 * V2_ROWB), "l"(V2_ROW(R, q[K2_SR_MO2], g + xo)),
          "r"((int)has2), "r"((int)(has2 && mo2))
        : "memory");
  }
}

// Boundary chunk: a unit is copied only if the source row has it in storage (rows hold exactly the 16-byte units of
// their own wavefront); what is not copied is never read (the range checks of the scalar path mask it).
__device__ __forceinline__ void v2_stage_guarded(unsigned char* stg, const int4* __restrict__ R, const int (&q)[7], const int* __restrict__ misc,
                                                 int gb, int lane) {
  const int g = gb + lane;
  const unsigned d = (unsigned)__cvta_generic_to_shared(stg) + 16u + ((unsigned)lane << 4);
#pragma unroll
  for (int r = 0; r < 7; ++r) {
    const int g0 = misc[V2_SG0 + r], ng = misc[V2_SNG + r];
    if ((unsigned)(g - g0) < (unsigned)ng) v2_cp16(d + r * V2_ROWB, V2_ROW(R, q[r], g));
    if (r != K2_SR_MX) {
      if (lane == 0 && (unsigned)(g - 1 - g0) < (unsigned)ng) v2_cp16(d - 16u + r * V2_ROWB, V2_ROW(R, q[r], g - 1));
      if (lane == 31 && (unsigned)(g + 1 - g0) < (unsigned)ng) v2_cp16(d + 16u + r * V2_ROWB, V2_ROW(R, q[r], g + 1));
    }
  }
}

// Launch shapes: (threads per CTA, CTAs per SM).  One 384- or 512-thread CTA per SM, or two 256-thread CTAs (two pairs
// per SM: one pair's serial trim / plan section overlaps the other's cell phase).
template <int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) k2v2_align(const K2Params P) {
  extern __shared__ __align__(16) unsigned char k2_smem[];
  cg::cluster_group cluster = cg::this_cluster();
  const int C = (int)cluster.num_blocks(), crank = (int)cluster.block_rank();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, NT = blockDim.x, NW = NT >> 5;
  auto csync = [&]() { if (C > 1) cluster.sync(); else __syncthreads(); };
  const int nslots = (P.nM + 2 * P.nG1 + 2 * P.nG2 + 1) & ~1;  // even: keeps `misc` 8-byte aligned
  uint2* win = reinterpret_cast<uint2*>(k2_smem);
  int* meta = reinterpret_cast<int*>(win + P.win_cap);  // [nslots][5]: lo, hi (trimmed), row origin, first stored unit, stored units
  int* misc = meta + 5 * nslots;
  const int S = P.stages;  // >= 1 (host)
  unsigned char* wstg = k2_smem + (((size_t)P.win_cap * 8 + (size_t)nslots * 20 + V2_MISC_BYTES + 15) & ~(size_t)15) +
                        (size_t)warp * S * 7 * V2_ROWB;
  unsigned char* slab = P.slabs + (size_t)(blockIdx.x / C) * P.slab_bytes;
  int* misc0 = C > 1 ? cluster.map_shared_rank(misc, 0) : misc;
  unsigned long long cells_acc = 0;
  long long t_cells = 0, t_sync = 0, t_bt = 0, t_setup = 0, t_trim = 0, t_plan = 0, t_mark = clock64();
#define V2_TICK(acc) do { const long long now_ = clock64(); acc += now_ - t_mark; t_mark = now_; } while (0)

  if (tid == 0) { misc[V2_PGN] = 0; misc[V2_RSV] = 0; }
  for (;;) {
    csync();
    V2_TICK(t_bt);
    if (crank == 0 && tid == 0) {
      int* pglist = reinterpret_cast<int*>(slab);
      v2_pages_release(P.pgstack, pglist, misc[V2_PGN]);  // the previous pair's history
      if (misc[V2_RSV] > 0) atomicSub(P.pgstack + 2, misc[V2_RSV]);
      misc[V2_PGN] = 0; misc[V2_RSV] = 0;
      misc[0] = (int)atomicAdd(P.counter, 1u); misc[1] = misc[2] = INT_MAX;
      if ((uint32_t)misc[0] < P.n) {
        // admission: the pair reserves the pages its work estimate says it will need and waits while the pool is
        // promised to the pairs already resident (a pair that found the pool empty half-way would be re-run from scratch);
        // a waiting pair holds nothing, and a pair alone on the device always starts
        const int want = min((int)P.pair_pages[P.order[misc[0]]], P.n_pages);
        for (;;) {
          const int before = atomicAdd(P.pgstack + 2, want);
          if (before == 0 || before + want <= P.n_pages) break;
          atomicSub(P.pgstack + 2, want);
          __nanosleep(2000);
        }
        misc[V2_RSV] = want;
        // the first two pages of the pair (the second one is the reserve, see plan)
        const int p0 = v2_page_acquire(P.pgstack), p1 = v2_page_acquire(P.pgstack);
        pglist[0] = p0; pglist[1] = p1; misc[V2_PGN] = 2;
        misc[V2_PGSEQ] = p0; misc[V2_PGSEQ + 1] = p1; misc[V2_PGSEQ + 2] = misc[V2_PGSEQ + 3] = -1;
      }
    }
    csync();
    const uint32_t wi = (uint32_t)misc0[0];
    if (wi >= P.n) { csync(); break; }
    const uint32_t pi = P.order[wi];
    const DevPair pr = P.pairs[pi];
    const int plen = pr.plen, tlen = pr.tlen;
    const int nwp = (plen >> 3) + 1, nwt = (tlen >> 3) + 1;
    uint2* wp = win;
    uint2* wt = win + nwp;
    const uint2* wtb = wt + (V2_BIAS >> 3);  // text windows addressed with the biased offset
    const int kbase = plen + 8;
    long long scap_ll = k2_gap_cost(P, plen) + k2_gap_cost(P, tlen) + 1;
    if (scap_ll > P.max_score) scap_ll = P.max_score;
    const int S_cap = (int)scap_ll;
    const size_t hist_bytes = (((size_t)S_cap + 1) * V2_HIST * 4 + 15) & ~(size_t)15;
    const size_t E_cap = (size_t)plen + tlen + 8;
    const size_t tmp_bytes = (E_cap * 4 + 15) & ~(size_t)15;
    const size_t pgl_bytes = ((size_t)P.n_pages * 4 + 15) & ~(size_t)15;  // this slot's list of pages taken (for the release)
    const size_t fixed = pgl_bytes + hist_bytes + 2 * tmp_bytes;
    int res_status = 0, res_score = 0;
    if (nwp + nwt > P.win_cap || tlen > V2_MAX_TLEN) res_status = -5;
    else if (fixed > P.slab_bytes) res_status = -6;
    if (res_status != 0) {
      if (crank == 0 && tid == 0) { P.scores[pi] = 0; P.status[pi] = res_status; P.out_len[pi] = 0; P.out_off[pi] = 0; }
      continue;
    }
    int* pglist = reinterpret_cast<int*>(slab);
    int* hist = reinterpret_cast<int*>(slab + pgl_bytes);  // [S_cap+1][16]: exists, M-row origin, 64-byte units per row, first unit, lo/hi x 5
    uint32_t* tmp_b = reinterpret_cast<uint32_t*>(slab + pgl_bytes + hist_bytes);
    uint32_t* tmp_f = tmp_b + (tmp_bytes >> 2);
    int16_t* rows16 = reinterpret_cast<int16_t*>(P.pages);
    const int4* rowsU = reinterpret_cast<const int4*>(P.pages);
    const int pu64 = 1 << (P.page_shift - 6);  // 64-byte units per page
    {
      const uint32_t* gp = P.nib + pr.p_woff;
      const uint32_t* gt = P.nib + pr.t_woff;
      for (int i = tid; i < nwp; i += NT) wp[i] = make_uint2(gp[i], gp[i + 1]);
      for (int i = tid; i < nwt; i += NT) wt[i] = make_uint2(gt[i], gt[i + 1]);
    }
    __syncthreads();

    const int sbM = 0, sbI1 = P.nM, sbD1 = P.nM + P.nG1, sbI2 = P.nM + 2 * P.nG1, sbD2 = P.nM + 2 * P.nG1 + P.nG2;
    unsigned long long* D64 = reinterpret_cast<unsigned long long*>(misc + V2_D64);  // [0] row units used, [1] cells
    // element of row-origin q (16-byte units) at diagonal k
    auto rd = [&](int q, int k) -> int { return (int)__ldcg(rows16 + (long long)q * 32 + (k + kbase)); };
    auto in_mat = [&](int vb, int k) -> bool { return k2_in_matrix(vb + V2_BIAS, k, plen, tlen); };

    // ---- per-score plan (warp 0, lane r = row r: 0..6 sources MX, MO1, I1E, D1E, I2E, D2E, MO2; 7..11 outputs) ----
    auto plan = [&](int s_from) {
      int* tab = misc + V2_TAB + lane * 5;
      const int back = tab[0], sbase = tab[1], depth = tab[2], dpk = tab[3];
      int rel = tab[4];
      const int dl = (int)(signed char)(dpk & 0xFF), dh = (int)(signed char)((dpk >> 8) & 0xFF);
      const int il = (int)(signed char)((dpk >> 16) & 0xFF), ih = (int)(signed char)((dpk >> 24) & 0xFF);
      const bool is_src = lane < 7, is_out = lane >= 7 && lane < 12;
      int used = (int)D64[0], pgi = misc[V2_PGI];  // 64-byte units taken in the current page, its index in the pair's page sequence
      for (int sN = s_from;; ++sN) {
        if (sN > S_cap) { if (lane == 0) misc[V2_STATUS] = -7; break; }
        const int sc = sN - back;
        const int slot = sbase + rel;
        rel = rel + 1 == depth ? 0 : rel + 1;
        int rlo = 1, rhi = -1, rq = 0, rg0 = 0, rng = 0;
        if (is_src && sc >= 0 && sN > 0) {
          rlo = meta[5 * slot]; rhi = meta[5 * slot + 1];
          if (rlo <= rhi) { rq = meta[5 * slot + 2]; rg0 = meta[5 * slot + 3]; rng = meta[5 * slot + 4]; }
        }
        const bool any = rlo <= rhi;
        const unsigned anym = __ballot_sync(FULL, any);
        int lo = warp_min(any ? rlo + dl : INT_MAX), hi = warp_max(any ? rhi + dh : INT_MIN);
        if (sN == 0) { lo = -pr.pbf; hi = pr.tbf; }
        if (lo > hi) {  // wavefront sN does not exist
          if (is_out) { meta[5 * slot] = 1; meta[5 * slot + 1] = -1; }
          if (lane == 0 && crank == 0) hist[(size_t)sN * V2_HIST] = 0;
          __syncwarp();
          continue;
        }
        const bool has2 = (anym & 0x70u) != 0u;
        const unsigned set1 = has2 ? 0x7Fu : 0x0Fu, set2 = 0x3Fu;
        const bool in1 = (set1 >> lane) & 1u, in2 = (set2 >> lane) & 1u;
        int ilo = warp_max(in1 ? rlo + il : INT_MIN), ihi = warp_min(in1 ? rhi + ih : INT_MAX);
        int ilo_b = warp_max(in2 ? rlo + il : INT_MIN), ihi_b = warp_min(in2 ? rhi + ih : INT_MAX);
        if ((anym & set1) != set1 || sN == 0) { ilo = 1; ihi = 0; }
        if (!has2 || (anym & set2) != set2 || sN == 0) { ilo_b = 1; ihi_b = 0; }
        const int flo = max(ilo, lo), fhi = min(ihi, hi), flo_b = max(ilo_b, lo), fhi_b = min(ihi_b, hi);
        // rows start on 64-byte boundaries: first unit a multiple of 4, row pitch a multiple of 4 units
        const int g_lo = ((lo + kbase) >> 3) & ~3, g_hi = (hi + kbase) >> 3, nu = g_hi - g_lo + 1, nup = (nu + 3) >> 2;
        const int need = 5 * nup;
        const bool sw = used + need > pu64;  // the rows of a score never straddle pages: move to the reserve page
        if (sw) { ++pgi; used = 0; }
        const int page = misc0[V2_PGSEQ + (pgi & 3)];  // published by rank 0 at least one cluster barrier ago
        if (need > pu64 || page < 0) { if (lane == 0) misc[V2_STATUS] = -6; break; }
        if (sw && lane == 0 && crank == 0) {  // take the next reserve page now: it is first needed at a later score
          const int np = v2_page_acquire(P.pgstack);
          if (np >= 0 && misc[V2_PGN] < P.n_pages) { pglist[misc[V2_PGN]] = np; misc[V2_PGN] += 1; }
          misc[V2_PGSEQ + ((pgi + 1) & 3)] = np;
        }
        const int base64 = page * pu64 + used;
        const int m2lo = __shfl_sync(FULL, rlo, K2_SR_MO2), m2hi = __shfl_sync(FULL, rhi, K2_SR_MO2);
        if (lane < 7) {
          misc[V2_Q + lane] = rq; misc[V2_RLO + lane] = rlo; misc[V2_RCNT + lane] = max(rhi - rlo + 1, 0);
          misc[V2_SG0 + lane] = rg0; misc[V2_SNG + lane] = rng;
        }
        if (is_out) {
          const int c = lane - 7;
          const int qo = base64 + c * nup - (g_lo >> 2);
          misc[V2_Q + lane] = qo; misc[V2_OSLOT + c] = slot;
          meta[5 * slot + 2] = qo; meta[5 * slot + 3] = g_lo; meta[5 * slot + 4] = nu;  // lo / hi follow at trim time
        }
        {  // raw (untrimmed) diagonal range of each output component: where its two sources can be non-null
          const int mo1lo = __shfl_sync(FULL, rlo, K2_SR_MO1), mo1hi = __shfl_sync(FULL, rhi, K2_SR_MO1);
          const bool g2 = lane == K2_SR_I2E || lane == K2_SR_D2E;
          const int olo = g2 ? m2lo : mo1lo, ohi = g2 ? m2hi : mo1hi;
          const bool oany = olo <= ohi;
          int clo = any ? (oany ? min(rlo, olo) : rlo) : (oany ? olo : 1), chi = any ? (oany ? max(rhi, ohi) : rhi) : (oany ? ohi : -1);
          if (any || oany) { clo += dl; chi += dl; }
          if (lane >= K2_SR_I1E && lane <= K2_SR_D2E) { misc[V2_CLO + lane - 1] = max(clo, lo); misc[V2_CHI + lane - 1] = min(chi, hi); }
          if (lane == 0) { misc[V2_CLO] = lo; misc[V2_CHI] = hi; }
        }
        if (lane == 0) {
          misc[V2_S] = sN; misc[V2_LO] = lo; misc[V2_HI] = hi; misc[V2_HAS2] = has2; misc[V2_G0] = g_lo; misc[V2_NU] = nu;
          misc[V2_NEED] = need; misc[V2_NUP] = nup; misc[V2_PGI] = pgi; D64[0] = (unsigned long long)used;
          misc[V2_RENORM] = (sN & 255) == 0;
          // chunk kinds as ranges of the chunk's first unit gb (chunk diagonals 8*gb - kbase .. + 255)
          const bool e1 = flo + 255 > fhi, e2 = flo_b + 255 > fhi_b;
          misc[V2_G1A] = e1 ? INT_MAX : (flo + kbase + 7) >> 3;   misc[V2_G1B] = e1 ? INT_MIN : (fhi - 255 + kbase) >> 3;
          misc[V2_G2A] = e2 ? INT_MAX : (flo_b + kbase + 7) >> 3; misc[V2_G2B] = e2 ? INT_MIN : (fhi_b - 255 + kbase) >> 3;
          const bool m2any = m2lo <= m2hi;
          misc[V2_M2A] = m2any ? (m2lo + kbase - 257) >> 3 : INT_MAX;        // gb <= M2A: chunk entirely left of M[s-o2'-e2'] (reads k-1 .. k+1)
          misc[V2_M2B] = m2any ? ((m2hi + kbase + 1) >> 3) + 1 : INT_MIN;    // gb >= M2B: entirely right of it
        }
        break;
      }
      tab[4] = rel;
      __syncwarp();
    };
    if (warp == 0) {
      if (lane == 0) { misc[V2_STATUS] = 0; D64[0] = 0; D64[1] = 0; misc[V2_ANYW] = 0; misc[V2_PGI] = 0; }
      if (lane < 10) { misc[V2_WIDE + lane] = 0; misc[V2_CAND + lane] = (lane & 1) ? INT_MIN : INT_MAX; }
      {
        int comp = K2_C_M, back = 0, dl = 0, dh = 0, il = 0, ih = 0;
        switch (lane) {
          case 0: back = P.x; break;
          case 1: back = P.o1 + P.e1; dl = -1; dh = 1; il = 1; ih = -1; break;
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
        int* tab = misc + V2_TAB + lane * 5;
        tab[0] = back; tab[1] = sbase; tab[2] = depth;
        tab[3] = (dl & 0xFF) | ((dh & 0xFF) << 8) | ((il & 0xFF) << 16) | ((ih & 0xFF) << 24);
        tab[4] = (depth - back % depth) % depth;
      }
      __syncwarp();
      plan(0);
    }

    int s = 0, end_k = 0;
    bool done = false;
    V2_TICK(t_setup);
    for (int it = 0;; ++it) {
      int* term_slot = misc0 + 1 + (it & 1);
      __syncthreads();  // barrier 2: plan of the next score visible; CTA-local
      V2_TICK(t_sync);
      { const int st = misc[V2_STATUS]; if (st != 0) { res_status = st; break; } }
      s = misc[V2_S];
      const int lo = misc[V2_LO], hi = misc[V2_HI];
      const bool has2 = misc[V2_HAS2] != 0, renorm = misc[V2_RENORM] != 0;
      const bool anyw_s = misc[V2_ANYW] != 0;
      const int G1A = misc[V2_G1A], G1B = misc[V2_G1B], G2A = misc[V2_G2A], G2B = misc[V2_G2B], M2A = misc[V2_M2A], M2B = misc[V2_M2B];
      const int g_lo = misc[V2_G0], nu = misc[V2_NU], g_hi = g_lo + nu - 1;
      int qs[7];
#pragma unroll
      for (int r = 0; r < 7; ++r) qs[r] = misc[V2_Q + r];
      const int qM = misc[V2_Q + 7], qI1 = misc[V2_Q + 8], qD1 = misc[V2_Q + 9], qI2 = misc[V2_Q + 10], qD2 = misc[V2_Q + 11];
      int4* outU = reinterpret_cast<int4*>(P.pages);
      const int nup = misc[V2_NUP];
      int term_k = INT_MAX;

      auto kind_of = [&](int gb) -> int {
        if (gb >= G1A && gb <= G1B) return 1;
        if (gb >= G2A && gb <= G2B && (gb <= M2A || gb >= M2B)) return 2;
        return 0;
      };
      const int gstep = C * NW * 32;
      int gi = g_lo + (crank * NW + warp) * 32, is = 0, cs = 0;
      auto stage_next = [&]() {
        if (gi <= g_hi && s > 0) {
          const int kd = kind_of(gi);
          if (kd != 0) v2_stage_fast(wstg + is * 7 * V2_ROWB, rowsU, qs, gi, has2, kd != 2, lane);
          else v2_stage_guarded(wstg + is * 7 * V2_ROWB, rowsU, qs, misc, gi, lane);
        }
        k2_cp_commit();
        gi += gstep;
        is = is + 1 == S ? 0 : is + 1;
      };
      for (int p = 0; p < S - 1; ++p) stage_next();
      for (int gb = g_lo + (crank * NW + warp) * 32; gb <= g_hi; gb += gstep) {
        const int g = gb + lane;
        const bool gact = g <= g_hi;
        const int k0 = (g << 3) - kbase;
        const int kind = kind_of(gb);
        __syncwarp();  // every lane is done reading the stage that is refilled now
        stage_next();
        k2_cp_wait_n(S - 1);
        __syncwarp();  // the other lanes' copies of this chunk have landed too
        const unsigned char* stg = wstg + cs * 7 * V2_ROWB;
        cs = cs + 1 == S ? 0 : cs + 1;

        uint32_t bw[4], ins1[4], del1[4], ins2[4], del2[4];
        int hb[8];
        if (kind != 0) {  // interior chunk: no guards, packed arithmetic
          V2Src sv;
          const unsigned char* sb = stg + 16 + (lane << 4);
          auto ld4 = [&](int r, uint32_t (&o)[4]) { const uint4 v = *reinterpret_cast<const uint4*>(sb + r * V2_ROWB); o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; };
          auto ldw = [&](int r, int off) -> uint32_t { return *reinterpret_cast<const uint32_t*>(sb + r * V2_ROWB + off); };
          ld4(K2_SR_MX, sv.mx);
          ld4(K2_SR_MO1, sv.mo1); sv.mo1L = ldw(K2_SR_MO1, -4); sv.mo1R = ldw(K2_SR_MO1, 16);
          ld4(K2_SR_I1E, sv.i1e); sv.i1eL = ldw(K2_SR_I1E, -4);
          ld4(K2_SR_D1E, sv.d1e); sv.d1eR = ldw(K2_SR_D1E, 16);
          if (has2) {
            ld4(K2_SR_I2E, sv.i2e); sv.i2eL = ldw(K2_SR_I2E, -4);
            ld4(K2_SR_D2E, sv.d2e); sv.d2eR = ldw(K2_SR_D2E, 16);
            if (kind == 1) {
              ld4(K2_SR_MO2, sv.mo2); sv.mo2L = ldw(K2_SR_MO2, -4); sv.mo2R = ldw(K2_SR_MO2, 16);
              v2_recur8<true, true>(sv, bw, ins1, del1, ins2, del2);
            } else {
              v2_recur8<true, false>(sv, bw, ins1, del1, ins2, del2);
            }
          } else {
            v2_recur8<false, false>(sv, bw, ins1, del1, ins2, del2);
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) { hb[2 * i] = v2_lo16(bw[i]); hb[2 * i + 1] = v2_hi16(bw[i]); }
        } else if (s == 0) {
#pragma unroll
          for (int j = 0; j < 8; ++j) { const int k = k0 + j; hb[j] = (k >= lo && k <= hi) ? max(k, 0) - V2_BIAS : V2_NULL; }
#pragma unroll
          for (int i = 0; i < 4; ++i) ins1[i] = del1[i] = ins2[i] = del2[i] = 0x80008000u;
        } else {
          // boundary chunk: elements outside their source wavefront's live range are overwritten with nulls in the staged
          // copy (lane L owns unit L of every row, lanes 0 / 31 also the unit left / right of the chunk), then the chunk
          // runs the packed recurrence like an interior one.  A diagonal outside this wavefront's own [lo, hi] has every
          // source out of range by construction (lo / hi are the union of the shifted source ranges).
          {
            unsigned char* sw = const_cast<unsigned char*>(stg);
            const int c0 = kbase - (gb << 3) + 8;  // staged element position of diagonal k: k + c0 (0 .. 271)
#pragma unroll
            for (int r = 0; r < 7; ++r) {
              const int pa = misc[V2_RLO + r] + c0, pb = pa + misc[V2_RCNT + r] - 1;  // live positions pa .. pb
#pragma unroll
              for (int part = 0; part < 2; ++part) {
                if (part == 1 && !(lane == 0 || lane == 31)) continue;
                const int p0 = part == 0 ? 8 + (lane << 3) : (lane == 0 ? 0 : 264);
                if (p0 >= pa && p0 + 7 <= pb) continue;  // fully live
                uint4* u = reinterpret_cast<uint4*>(sw + r * V2_ROWB + p0 * 2);
                if (p0 + 7 < pa || p0 > pb) { *u = make_uint4(0x80008000u, 0x80008000u, 0x80008000u, 0x80008000u); continue; }
                uint4 v = *u;
                uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const int e = p0 + 2 * i;
                  const bool okl = e >= pa && e <= pb, okh = e + 1 >= pa && e + 1 <= pb;
                  w[i] = (okl ? (w[i] & 0xFFFFu) : 0x8000u) | (okh ? (w[i] & 0xFFFF0000u) : 0x80000000u);
                }
                *u = make_uint4(w[0], w[1], w[2], w[3]);
              }
            }
            __syncwarp();
          }
          V2Src sv;
          const unsigned char* sb = stg + 16 + (lane << 4);
          auto ld4 = [&](int r, uint32_t (&o)[4]) { const uint4 v = *reinterpret_cast<const uint4*>(sb + r * V2_ROWB); o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; };
          auto ldw = [&](int r, int off) -> uint32_t { return *reinterpret_cast<const uint32_t*>(sb + r * V2_ROWB + off); };
          ld4(K2_SR_MX, sv.mx);
          ld4(K2_SR_MO1, sv.mo1); sv.mo1L = ldw(K2_SR_MO1, -4); sv.mo1R = ldw(K2_SR_MO1, 16);
          ld4(K2_SR_I1E, sv.i1e); sv.i1eL = ldw(K2_SR_I1E, -4);
          ld4(K2_SR_D1E, sv.d1e); sv.d1eR = ldw(K2_SR_D1E, 16);
          ld4(K2_SR_I2E, sv.i2e); sv.i2eL = ldw(K2_SR_I2E, -4);
          ld4(K2_SR_D2E, sv.d2e); sv.d2eR = ldw(K2_SR_D2E, 16);
          ld4(K2_SR_MO2, sv.mo2); sv.mo2L = ldw(K2_SR_MO2, -4); sv.mo2R = ldw(K2_SR_MO2, 16);
          v2_recur8<true, true>(sv, bw, ins1, del1, ins2, del2);
#pragma unroll
          for (int i = 0; i < 4; ++i) { hb[2 * i] = v2_lo16(bw[i]); hb[2 * i + 1] = v2_hi16(bw[i]); }
        }
        // the gap rows do not depend on the extension: store them first
        if (gact && s > 0) {
          if (renorm) {
#pragma unroll
            for (int i = 0; i < 4; ++i) { ins1[i] = v2_renull(ins1[i]); ins2[i] = v2_renull(ins2[i]); }
          }
          *V2_ROW(outU, qI1, g) = make_int4((int)ins1[0], (int)ins1[1], (int)ins1[2], (int)ins1[3]);
          *V2_ROW(outU, qD1, g) = make_int4((int)del1[0], (int)del1[1], (int)del1[2], (int)del1[3]);
          if (has2) {
            *V2_ROW(outU, qI2, g) = make_int4((int)ins2[0], (int)ins2[1], (int)ins2[2], (int)ins2[3]);
            *V2_ROW(outU, qD2, g) = make_int4((int)del2[0], (int)del2[1], (int)del2[2], (int)del2[3]);
          }
        }
        // ---- bounds, extension along matches, termination test (biased offsets: h = hb + BIAS, v = hb - (k - BIAS)) ----
        // Straight-line code, no branches (selects only): the eight cells' chains interleave.
        uint32_t xw[8];
        int okm[8];
        const int kb0 = k0 - V2_BIAS;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int v = hb[j] - (kb0 + j);
          const int ok = (int)((unsigned)(hb[j] + V2_BIAS) <= (unsigned)tlen) & (int)((unsigned)v <= (unsigned)plen);
          okm[j] = ok;
          hb[j] = ok ? hb[j] : V2_NULL;
          xw[j] = window8(wp, ok ? v : 0) ^ window8(wtb, ok ? hb[j] : -V2_BIAS);
        }
        int advmax = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int adv = (int)min(((unsigned)(__ffs((int)xw[j]) - 1)) >> 2, 8u);  // equal bases of the window; 8 when xw == 0
          const int a = okm[j] ? adv : 0;
          hb[j] += a;
          advmax = max(advmax, a);
        }
        if (__any_sync(FULL, advmax == 8)) {  // rare: a run longer than 8 bases (the true diagonal)
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const bool go = okm[j] != 0 && xw[j] == 0u;
            const int off = k2_extend_more(wp, wt, plen, tlen, k0 + j, go ? hb[j] + V2_BIAS : 0, go, lane);
            if (go) hb[j] = off - V2_BIAS;
          }
        }
        // a cell touches the end of a sequence <=> the largest text / pattern position of the eight does (a null cell has
        // h = -768 and v = -768 - k <= plen - 760: never at an end)
        int hmax = hb[0], vmax = hb[0] - kb0;
#pragma unroll
        for (int j = 1; j < 8; ++j) { hmax = max(hmax, hb[j]); vmax = max(vmax, hb[j] - (kb0 + j)); }
        if (hmax + V2_BIAS >= tlen || vmax >= plen) {  // full ends-free termination test (rare)
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (hb[j] >= V2_VMIN) {
              const int k = k0 + j, off = hb[j] + V2_BIAS, v = off - k;
              if ((off >= tlen && plen - v <= pr.pef) || (v >= plen && tlen - off <= pr.tef)) term_k = min(term_k, k);
            }
          }
        }
        if (gact) *V2_ROW(outU, qM, g) = make_int4((int)v2_pack(hb[0], hb[1]), (int)v2_pack(hb[2], hb[3]), (int)v2_pack(hb[4], hb[5]), (int)v2_pack(hb[6], hb[7]));
      }
      k2_cp_wait<0>();
      V2_TICK(t_cells);
      // ---- barrier 1: rows of score s complete (cluster-wide); did any diagonal terminate? ----
      term_k = warp_min(term_k);
      if (term_k != INT_MAX && lane == 0) atomicMin(term_slot, term_k);
      csync();
      const int tk = *term_slot;
      const int ncomp = 5;
      if (tk != INT_MAX) {
        end_k = tk; done = true;
        if (crank == 0 && tid == 0) {
          int* h = hist + (size_t)s * V2_HIST;
          h[0] = 1; h[1] = qM; h[2] = nup; h[3] = g_lo;  // the final wavefront is only read at (s, end_k) itself
          D64[1] += (unsigned long long)(hi - lo + 1);
        }
        break;
      }
      V2_TICK(t_sync);
      // ---- trims (see k2_align): scans that crossed a long out-of-matrix run at the previous score are done by all warps ----
      // warp 0's narrow probes (the first three diagonals of every scan that was short at the previous score) are issued
      // first so that their round trip overlaps the cooperative wide scans below
      const int tr_t = lane / 3, tr_j = lane - 3 * tr_t, tr_c = min(tr_t >> 1, 4), tr_side = tr_t & 1;
      int tr_v = V2_NULL, tr_clo = 1, tr_chi = -1, tr_q = 0, tr_k = 0, tr_wprev = 0;
      bool tr_en = false, tr_wide = false, tr_ok = false;
      if (warp == 0) {
        tr_en = tr_t < 2 * ncomp && (tr_c == K2_C_M || s > 0) && !((tr_c == K2_C_I2 || tr_c == K2_C_D2) && !has2);
        tr_clo = tr_en ? misc[V2_CLO + tr_c] : 1; tr_chi = tr_en ? misc[V2_CHI + tr_c] : -1;
        tr_q = misc[V2_Q + 7 + tr_c];
        tr_k = tr_side ? tr_chi - tr_j : tr_clo + tr_j;
        tr_wprev = (tr_t < 10 && anyw_s) ? misc[V2_WIDE + tr_t] : 0;
        tr_wide = tr_en && tr_wprev > 3;
        tr_ok = !tr_wide && tr_k >= tr_clo && tr_k <= tr_chi;
        tr_v = tr_ok ? rd(tr_q, tr_k) : V2_NULL;
      }
      if (anyw_s) {
        // first block (32 diagonals per warp) of every wide scan: all loads in flight together
        int pv0[10];
#pragma unroll
        for (int t = 0; t < 10; ++t) {
          pv0[t] = V2_NULL;
          const int w = misc[V2_WIDE + t];
          const int c = t >> 1, side = t & 1;
          if (w > 3 && (c == K2_C_M || s > 0) && !((c == K2_C_I2 || c == K2_C_D2) && !has2)) {
            const int clo = misc[V2_CLO + c], chi = misc[V2_CHI + c];
            const int win_n = min(w + 64, chi - clo + 1);
            const int per = min((win_n + NW * 32 - 1) / (NW * 32), 8);
            const int off = warp * per * 32 + lane;
            if (off < win_n) pv0[t] = rd(misc[V2_Q + 7 + c], side ? chi - off : clo + off);
          }
        }
#pragma unroll
        for (int t = 0; t < 10; ++t) {
          const int w = misc[V2_WIDE + t];
          if (w <= 3) continue;
          const int c = t >> 1, side = t & 1;
          if (!((c == K2_C_M || s > 0) && !((c == K2_C_I2 || c == K2_C_D2) && !has2))) continue;
          const int clo = misc[V2_CLO + c], chi = misc[V2_CHI + c];
          const int q = misc[V2_Q + 7 + c];
          const int win_n = min(w + 64, chi - clo + 1);
          const int per = min((win_n + NW * 32 - 1) / (NW * 32), 8);
          const int o = warp * per * 32;
          bool got = false;
          int pos = 0;
          for (int u = 0; u < per; ++u) {  // per is 1 unless the run is longer than the CTA is wide
            const int off = o + 32 * u + lane;
            const int kk = side ? chi - off : clo + off;
            const int pv = u == 0 ? pv0[t] : (off < win_n ? rd(q, kk) : V2_NULL);
            const unsigned b = __ballot_sync(FULL, off < win_n && in_mat(pv, kk));
            if (b) { got = true; const int fo = o + 32 * u + __ffs((int)b) - 1; pos = side ? chi - fo : clo + fo; break; }
          }
          if (got && lane == 0) { if (side) atomicMax(misc + V2_CAND + t, pos); else atomicMin(misc + V2_CAND + t, pos); }
        }
        __syncthreads();
      }
      if (warp == 0) {
        {
          const int t = tr_t, j = tr_j, c = tr_c, side = tr_side;
          const bool en = tr_en, wide = tr_wide, ok = tr_ok;
          const int clo = tr_clo, chi = tr_chi, q = tr_q, k = tr_k, wprev = tr_wprev, v = tr_v;
          const unsigned hits = __ballot_sync(FULL, ok && in_mat(v, k));
          const unsigned mine = (hits >> (3 * t)) & 7u;
          int found = side ? -1 : 1;
          bool got = mine != 0u;
          int skip = 3;
          if (mine) found = side ? chi - (__ffs((int)mine) - 1) : clo + __ffs((int)mine) - 1;
          if (wide) {
            const int cand = misc[V2_CAND + t];
            got = side ? cand != INT_MIN : cand != INT_MAX;
            if (got) found = cand;
            const int win_n = min(wprev + 64, chi - clo + 1);
            skip = min(win_n, min((win_n + NW * 32 - 1) / (NW * 32), 8) * NW * 32);
          }
          unsigned unres = __ballot_sync(FULL, j == 0 && en && !got && chi - clo + 1 > skip);
          if (j == 0 && t < 2 * ncomp) {
            meta[5 * misc[V2_OSLOT + c] + side] = found;
            misc[V2_WIDE + t] = !en ? 0 : got ? (side ? chi - found : found - clo) : chi - clo + 1;
            misc[V2_CAND + t] = side ? INT_MIN : INT_MAX;
          }
          while (unres) {  // rare: warp-wide continuation of one scan
            const int src = __ffs((int)unres) - 1;
            unres &= unres - 1;
            const int c2 = __shfl_sync(FULL, c, src), side2 = __shfl_sync(FULL, side, src);
            const int lo2 = __shfl_sync(FULL, clo, src), hi2 = __shfl_sync(FULL, chi, src);
            const int q2 = __shfl_sync(FULL, q, src);
            const int skip2 = __shfl_sync(FULL, skip, src), t2 = __shfl_sync(FULL, t, src);
            int f2 = side2 ? -1 : 1;
            for (int base = side2 ? hi2 - skip2 : lo2 + skip2; side2 ? base >= lo2 : base <= hi2; base += side2 ? -256 : 256) {
              int pv8[8];
#pragma unroll
              for (int u = 0; u < 8; ++u) {
                const int kk = side2 ? base - 32 * u - lane : base + 32 * u + lane;
                pv8[u] = (kk >= lo2 && kk <= hi2) ? rd(q2, kk) : V2_NULL;
              }
              bool got2 = false;
#pragma unroll
              for (int u = 0; u < 8; ++u) {
                const int kk = side2 ? base - 32 * u - lane : base + 32 * u + lane;
                const unsigned b = __ballot_sync(FULL, kk >= lo2 && kk <= hi2 && in_mat(pv8[u], kk));
                if (b && !got2) { got2 = true; f2 = side2 ? base - 32 * u - (__ffs((int)b) - 1) : base + 32 * u + __ffs((int)b) - 1; }
              }
              if (got2) break;
            }
            if (lane == 0) {
              meta[5 * misc[V2_OSLOT + c2] + side2] = f2;
              if (side2 ? f2 >= lo2 : f2 <= hi2) misc[V2_WIDE + t2] = side2 ? hi2 - f2 : f2 - lo2;
            }
          }
          __syncwarp();
          const bool anyw = lane < 2 * ncomp && misc[V2_WIDE + min(lane, 9)] > 3;
          const unsigned aw = __ballot_sync(FULL, anyw);
          if (lane == 0) misc[V2_ANYW] = aw != 0u;
        }
        __syncwarp();
        if (crank == 0 && lane < V2_HIST) {  // history record of score s: what the backtrace needs to re-read its rows
          int v = 0;
          if (lane == 0) v = 1;
          else if (lane == 1) v = qM;
          else if (lane == 2) v = nup;
          else if (lane == 3) v = g_lo;
          else if (lane < 14) v = meta[5 * misc[V2_OSLOT + ((lane - 4) >> 1)] + ((lane - 4) & 1)];
          hist[(size_t)s * V2_HIST + lane] = v;
        }
        if (lane == 0) {
          D64[1] += (unsigned long long)(hi - lo + 1);
          D64[0] += (unsigned long long)misc[V2_NEED];
        }
        __syncwarp();
        V2_TICK(t_trim);
        plan(s + 1);
        V2_TICK(t_plan);
      }
    }

    V2_TICK(t_sync);
    if (crank == 0 && tid == 0) cells_acc += D64[1];
    // ---- backtrace (warp 0 of rank 0): decisions re-derived from the stored wavefronts, then the forward replay ----
    if (done) {
      csync();  // rows and history of every score visible to warp 0 of rank 0
      if (crank == 0 && warp == 0) {
        // stored value (biased) of component c at (score ss, diagonal kk); null outside the component's trimmed range
        auto get = [&](int c, int ss, int kk) -> int {
          if (ss < 0) return V2_NULL;
          const int* h = hist + (size_t)ss * V2_HIST;
          if (__ldcg(h) == 0) return V2_NULL;
          const int l = __ldcg(h + 4 + 2 * c), r = __ldcg(h + 5 + 2 * c);
          if (kk < l || kk > r) return V2_NULL;
          return rd(__ldcg(h + 1) + c * __ldcg(h + 2), kk);
        };
        const int end_off = rd(misc[V2_Q + 7], end_k) + V2_BIAS;
        int sc = s, k = end_k, comp = K2_C_M, nb = 0, bad = 0;
        int last_code = -1;
        uint32_t run = 0;
        auto push_n = [&](int code, uint32_t n) {  // lane 0 keeps the run-length list; every lane tracks nb / bad
          if (n == 0) return;
          if (code == last_code) { run += n; return; }
          if (run) { if ((size_t)nb < E_cap) { if (lane == 0) tmp_b[nb] = (run << 3) | (uint32_t)last_code; ++nb; } else bad = 1; }
          last_code = code; run = n;
        };
        // candidate table of an M step, lane t: component, score distance, diagonal shift, +1
        const int cb_comp = lane == 0 || lane == 1 || lane == 3 || lane == 5 || lane == 7 ? K2_C_M : lane == 2 ? K2_C_I1 : lane == 4 ? K2_C_I2 : lane == 6 ? K2_C_D1 : K2_C_D2;
        // lanes: 0 X, 1 I1 open, 2 I1 ext, 3 I2 open, 4 I2 ext, 5 D1 open, 6 D1 ext, 7 D2 open, 8 D2 ext
        const int cb_back = lane == 0 ? P.x : (lane == 1 || lane == 5) ? P.o1 + P.e1 : (lane == 2 || lane == 6) ? P.e1 : (lane == 3 || lane == 7) ? P.o2 + P.e2 : P.e2;
        const int cb_dk = lane == 0 ? 0 : lane <= 4 ? -1 : 1;
        int guard = 0;
        while (sc > 0 && !bad) {
          if (++guard > 4 * (int)E_cap + 64) { bad = 1; break; }
          if (comp == K2_C_M) {
            const int v = lane < 9 ? get(cb_comp, sc - cb_back, k + cb_dk) : V2_NULL;
            const int mx = __shfl_sync(FULL, v, 0);
            const int i1o = __shfl_sync(FULL, v, 1), i1e = __shfl_sync(FULL, v, 2), i2o = __shfl_sync(FULL, v, 3), i2e = __shfl_sync(FULL, v, 4);
            const int d1o = __shfl_sync(FULL, v, 5), d1e = __shfl_sync(FULL, v, 6), d2o = __shfl_sync(FULL, v, 7), d2e = __shfl_sync(FULL, v, 8);
            const int vi1 = max(i1o, i1e) + 1, vi2 = max(i2o, i2e) + 1, vd1 = max(d1o, d1e), vd2 = max(d2o, d2e), mis = mx + 1;
            int bb = vi1, sel = K2_SEL_I1;
            if (vi2 >= bb) { bb = vi2; sel = K2_SEL_I2; }
            if (vd1 >= bb) { bb = vd1; sel = K2_SEL_D1; }
            if (vd2 >= bb) { bb = vd2; sel = K2_SEL_D2; }
            if (mis >= bb) { bb = mis; sel = K2_SEL_X; }
            if (bb < V2_VMIN) { bad = 1; break; }  // a cell on the path always has a live predecessor
            if (sel == K2_SEL_X) { push_n(K2_E_X, 1); sc -= P.x; }
            else comp = sel == K2_SEL_I1 ? K2_C_I1 : sel == K2_SEL_I2 ? K2_C_I2 : sel == K2_SEL_D1 ? K2_C_D1 : K2_C_D2;
            continue;
          }
          // inside a gap component: lane j examines step j of the run assuming steps 0 .. j-1 were extensions
          const bool isI = comp == K2_C_I1 || comp == K2_C_I2, p1 = comp == K2_C_I1 || comp == K2_C_D1;
          const int e = p1 ? P.e1 : P.e2, oe = p1 ? P.o1 + P.e1 : P.o2 + P.e2, dk = isI ? -1 : 1;
          const int sj = sc - lane * e, kj = k + lane * dk;
          int vo = V2_NULL, ve = V2_NULL;
          if (sj > 0) { vo = get(K2_C_M, sj - oe, kj + dk); ve = get(comp, sj - e, kj + dk); }
          const bool is_ext = sj > 0 && ve >= vo && ve >= V2_VMIN;
          const bool dead = sj > 0 && max(vo, ve) < V2_VMIN;  // neither predecessor is live: inconsistent (only meaningful at the first non-extension)
          const unsigned extm = __ballot_sync(FULL, is_ext);
          const int J = extm == FULL ? 32 : __ffs((int)~extm) - 1;  // steps 0 .. J-1 are extensions
          push_n(isI ? K2_E_IE : K2_E_DE, (uint32_t)J);
          sc -= J * e; k += J * dk;
          if (J < 32) {  // step J opens the gap (or the score ran out)
            const int sJ = __shfl_sync(FULL, sj, J), deadJ = __shfl_sync(FULL, (int)dead, J);
            if (sJ <= 0 || deadJ) { bad = 1; break; }
            push_n(isI ? K2_E_IO : K2_E_DO, 1);
            sc -= oe; k += dk; comp = K2_C_M;
          }
        }
        if (run) { if ((size_t)nb < E_cap) { if (lane == 0) tmp_b[nb] = (run << 3) | (uint32_t)last_code; ++nb; } else bad = 1; }
        if (sc != 0 || comp != K2_C_M) bad = 1;
        const int k0 = k;
        __syncwarp();
        const int h_end = end_off, v_end = end_off - end_k;
        const int al_t = h_end - max(k0, 0), al_p = v_end - max(-k0, 0);
        res_score = (P.match == 0) ? -s : ((-P.match) * (al_p + al_t) - s) / 2;
        unsigned long long o_off = 0;
        uint32_t o_len = 0;
        if (P.want_cigar && !bad) {
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
    V2_TICK(t_bt);
    atomicAdd(P.prof + 0, (unsigned long long)t_cells); atomicAdd(P.prof + 1, (unsigned long long)t_sync);
    atomicAdd(P.prof + 2, (unsigned long long)t_bt); atomicAdd(P.prof + 3, (unsigned long long)t_setup);
    atomicAdd(P.prof + 4, (unsigned long long)t_trim); atomicAdd(P.prof + 5, (unsigned long long)t_plan);
    atomicMax(P.prof + 6, (unsigned long long)(t_cells + t_sync + t_bt + t_setup + t_trim + t_plan));
    atomicAdd(P.prof + 7, 1ull);
  }
#undef V2_TICK
}
#endif  // __CUDACC__

}  // namespace sfc
