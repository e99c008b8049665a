// sfc_k_score.cuh — K1: gap-linear ends-free wavefront alignment, SCORE ONLY.
//
// Replaces WFA2-lib's gap-linear aligner as sawfish's score_sv drives it:
// PairwiseAligner::new_consensus_aligner + align (reference src/wfa2_utils.rs:182-196,209-219),
// called per read x breakend x registration x allele at src/score_sv/mod.rs:837-901 where only the
// score is consumed (src/score_sv/mod.rs:702).
//
// Mapping: one group of WPP warps per pair (WPP = 1 for score_sv-sized flanks), persistent CTAs
// pulling pairs from a global work counter.  The M-wavefront ring (depth max(x',i')+1) and both
// sequences (4-bit, overlapped 8-byte windows) live in shared memory.  Each wavefront cell is one
// 32-bit word:  [31] null/sign | [30:16] offset h | [15:14] predecessor-type | [13:0] k0 + 8192
// so that ONE signed max over the three candidates reproduces WFA2's backtrace tie-break
// (mismatch > deletion > insertion on equal offsets) while carrying k0, the diagonal on which the
// winning path left wavefront 0 — needed because the reported score is taken over the non-free
// part of the alignment only (pinned by src/wfa2_utils.rs:294-308).
// Reads of source wavefronts are range-guarded from per-wavefront (lo,hi) metadata, so the ring
// never needs initialising or clearing; wavefront trimming (first/last in-matrix diagonal) is
// metadata only.  Integer-ALU bound; no tensor cores (integer DP).
#pragma once
#include "sfc_device.cuh"

namespace sfc {

constexpr int32_t K1_ONE = 1 << 16;
constexpr int32_t K1_TX = 2 << 14;  // mismatch predecessor   (priority 2)
constexpr int32_t K1_TD = 1 << 14;  // deletion predecessor   (priority 1); insertion = 0
constexpr int32_t K1_TMASK = 3 << 14;
constexpr int32_t K1_K0BIAS = 8192;
constexpr int K1_MAX_LEN = 32767;
constexpr int K1_MAX_FREE = 8191;

struct K1Params {
  const DevPair* pairs;
  const uint32_t* order;  // work order (pair indices, heaviest first)
  uint32_t n;
  uint32_t* counter;
  const uint32_t* nib;
  int32_t* scores;
  int32_t* status;
  int x, ip, match;  // adjusted mismatch / indel penalties, original match score (<= 0)
  int rd;            // ring depth = max(x, ip) + 1
  int wcap;          // cells per ring row (>= max(plen+tlen) + 5)
  int win_cap;       // uint2 windows reserved per sequence
  int group_bytes;   // shared bytes per group
  int max_score;
  unsigned long long* cells;
};

template <int WPP>
__device__ __forceinline__ void k1_gsync(int group) {
  if (WPP == 1) __syncwarp();
  else asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "r"(WPP * 32) : "memory");
}

// cnt = number of live diagonals of the source wavefront (0 when it does not exist)
__device__ __forceinline__ int32_t k1_guard(const int32_t* row, int k, int lo, int cnt) {
  return ((unsigned)(k - lo) < (unsigned)cnt) ? row[k] : WNULL;
}

template <int WPP>
__global__ void __launch_bounds__(256) k1_score(const K1Params P) {
  extern __shared__ __align__(16) unsigned char k1_smem[];
  constexpr int G = WPP * 32;
  const int group = threadIdx.x / G, g = threadIdx.x % G, lane = threadIdx.x & 31, gw = g >> 5;
  unsigned char* gs = k1_smem + (size_t)group * P.group_bytes;
  int32_t* ring = reinterpret_cast<int32_t*>(gs);
  uint2* wp = reinterpret_cast<uint2*>(gs + (size_t)P.rd * P.wcap * 4);
  uint2* wt = wp + P.win_cap;
  int* meta = reinterpret_cast<int*>(wt + P.win_cap);  // [rd][2] lo,hi
  int* part = meta + 2 * P.rd;                          // [2][WPP][3]
  int* slot_pair = part + 2 * WPP * 3;                  // [1]
  const int rd = P.rd, wcap = P.wcap;
  unsigned long long cells_acc = 0;

  for (;;) {
    if (g == 0) *reinterpret_cast<volatile int*>(slot_pair) = (int)atomicAdd(P.counter, 1u);
    k1_gsync<WPP>(group);
    const uint32_t wi = (uint32_t) * reinterpret_cast<volatile int*>(slot_pair);
    if (wi >= P.n) break;
    const uint32_t pi = P.order[wi];
    const DevPair pr = P.pairs[pi];
    const int plen = pr.plen, tlen = pr.tlen;
    {
      const int nwp = (plen >> 3) + 1, nwt = (tlen >> 3) + 1;
      const uint32_t* gp = P.nib + pr.p_woff;
      const uint32_t* gt = P.nib + pr.t_woff;
      for (int i = g; i < nwp; i += G) wp[i] = make_uint2(gp[i], gp[i + 1]);
      for (int i = g; i < nwt; i += G) wt[i] = make_uint2(gt[i], gt[i + 1]);
    }
    const int kbase = plen + 2;
    k1_gsync<WPP>(group);

    int s = 0, par = 0, res_status = 0, res_score = 0;
    for (;; ++s) {
      if (s > P.max_score) { res_status = -7; break; }
      int lo, hi, lox = 1, hix = -1, loi = 1, hii = -1;
      const int sx = s - P.x, si = s - P.ip;
      if (s == 0) {
        lo = -pr.pbf; hi = pr.tbf;
      } else {
        if (sx >= 0) { lox = meta[2 * (sx % rd)]; hix = meta[2 * (sx % rd) + 1]; }
        if (si >= 0) { loi = meta[2 * (si % rd)]; hii = meta[2 * (si % rd) + 1]; }
        const bool hasx = lox <= hix, hasi = loi <= hii;
        if (!hasx && !hasi) {  // wavefront s does not exist
          if (g == 0) { meta[2 * (s % rd)] = 1; meta[2 * (s % rd) + 1] = -1; }
          k1_gsync<WPP>(group);
          continue;
        }
        lo = min(hasx ? lox : INT_MAX, hasi ? loi - 1 : INT_MAX);
        hi = max(hasx ? hix : INT_MIN, hasi ? hii + 1 : INT_MIN);
      }
      // interior: every read of every source is in range (unguarded fast path)
      const int ilo = max(lox, loi + 1), ihi = min(hix, hii - 1);
      const int cntx = max(hix - lox + 1, 0), cnti = max(hii - loi + 1, 0);
      const int32_t* rowX = ring + (size_t)((sx >= 0 ? sx : 0) % rd) * wcap + kbase;
      const int32_t* rowI = ring + (size_t)((si >= 0 ? si : 0) % rd) * wcap + kbase;
      int32_t* rowO = ring + (size_t)(s % rd) * wcap + kbase;
      int term_k = INT_MAX, nn_lo = INT_MAX, nn_hi = INT_MIN;
      if (g == 0) cells_acc += (unsigned long long)(hi - lo + 1);

      for (int kb = lo + gw * 32; kb <= hi; kb += G) {
        const int k = kb + lane;
        const bool act = k <= hi;
        int32_t cell = WNULL;
        if (s == 0) {
          if (act) cell = (max(k, 0) << 16) | (k + K1_K0BIAS);
        } else {
          int32_t a, b, c;
          if (kb >= ilo && kb + 31 <= ihi) {
            a = rowX[k]; b = rowI[k - 1]; c = rowI[k + 1];
          } else {
            a = k1_guard(rowX, k, lox, cntx);
            b = k1_guard(rowI, k - 1, loi, cnti);
            c = k1_guard(rowI, k + 1, loi, cnti);
          }
          const int32_t mis = a + (K1_ONE | K1_TX), ins = b + K1_ONE, del = c | K1_TD;
          cell = max(max(mis, del), ins) & ~K1_TMASK;
          if (!act) cell = WNULL;
        }
        int off = cell >> 16;
        if ((unsigned)off > (unsigned)tlen || (unsigned)(off - k) > (unsigned)plen) cell = WNULL;
        bool pending = false;
        if (cell >= 0) {  // extend along matches: two private windows, then warp-cooperative
          const int v = off - k;
          uint32_t xw = window8(wp, v) ^ window8(wt, off);
          if (xw) off += lead_eq(xw);
          else {
            off += 8;
            xw = window8(wp, v + 8) ^ window8(wt, off);
            if (xw) off += lead_eq(xw);
            else { off += 8; pending = true; }
          }
        }
        unsigned pm = __ballot_sync(FULL, pending);
        while (pm) {
          const int src = __ffs((int)pm) - 1;
          pm &= pm - 1;
          const int ko = __shfl_sync(FULL, k, src), ho = __shfl_sync(FULL, off, src);
          const int adv = coop_extend(wp, wt, plen, tlen, ko, ho, lane);
          if (lane == src) off += adv;
        }
        if (cell >= 0) {
          cell = (cell & 0xFFFF) | (off << 16);
          const int v = off - k;
          if ((off >= tlen && plen - v <= pr.pef) || (v >= plen && tlen - off <= pr.tef)) term_k = min(term_k, k);
          nn_lo = min(nn_lo, k);
          nn_hi = max(nn_hi, k);
        }
        if (act) rowO[k] = cell;
      }
      term_k = warp_min(term_k);
      nn_lo = warp_min(nn_lo);
      nn_hi = warp_max(nn_hi);
      if (WPP > 1) {
        int* pp = part + par * WPP * 3;
        if (lane == 0) { pp[gw * 3] = term_k; pp[gw * 3 + 1] = nn_lo; pp[gw * 3 + 2] = nn_hi; }
        k1_gsync<WPP>(group);
#pragma unroll
        for (int w = 0; w < WPP; ++w) {
          term_k = min(term_k, pp[w * 3]);
          nn_lo = min(nn_lo, pp[w * 3 + 1]);
          nn_hi = max(nn_hi, pp[w * 3 + 2]);
        }
        par ^= 1;
        if (g == 0) { meta[2 * (s % rd)] = nn_lo <= nn_hi ? nn_lo : 1; meta[2 * (s % rd) + 1] = nn_lo <= nn_hi ? nn_hi : -1; }
        if (min(P.x, P.ip) < 2 || term_k != INT_MAX) k1_gsync<WPP>(group);
      } else {
        if (lane == 0) { meta[2 * (s % rd)] = nn_lo <= nn_hi ? nn_lo : 1; meta[2 * (s % rd) + 1] = nn_lo <= nn_hi ? nn_hi : -1; }
        __syncwarp();
      }
      if (term_k != INT_MAX) {
        const int32_t cell = rowO[term_k];
        const int off = cell >> 16, k0 = (cell & 0x3FFF) - K1_K0BIAS;
        const int aligned = off + (off - term_k) - abs(k0);
        res_score = (P.match == 0) ? -s : ((-P.match) * aligned - s) / 2;
        break;
      }
    }
    if (g == 0) { P.scores[pi] = res_score; P.status[pi] = res_status; }
    k1_gsync<WPP>(group);  // ring/windows are reused by the next pair
  }
  if (g == 0 && cells_acc) atomicAdd(P.cells, cells_acc);
}

}  // namespace sfc
