// sfc_k_score.cuh — K1: gap-linear ends-free wavefront alignment, SCORE ONLY.
//
// Replaces WFA2-lib's gap-linear aligner as sawfish's score_sv drives it:
// PairwiseAligner::new_consensus_aligner + align (reference src/wfa2_utils.rs:182-196,209-219),
// called per read x breakend x registration x allele at src/score_sv/mod.rs:837-901 where only the
// score is consumed (src/score_sv/mod.rs:702).
//
// Mapping: one group of WPP warps per pair, persistent groups pulling pairs from a global work counter.
// The M-wavefront ring (depth max(x',i')+1) and both sequences live in shared memory.  Sequences are kept as ONE
// 32-bit window PER BASE (the 8 bases from that position, first base in the top nibble, built from the 4-bit arena
// when the pair is fetched): the extension probe of a cell is two plain LDS.32 at (offset - k) and (offset), an XOR
// and a bfind — no window address arithmetic, no funnel shifts, no bit reversal.
// Each wavefront cell is one 32-bit word:
//     [31] null/sign | [30:16] offset h | [15:14] predecessor-type | [13:0] k0 + 8192
// so that ONE signed max over the three candidates reproduces WFA2's backtrace tie-break
// (mismatch > deletion > insertion on equal offsets) while carrying k0, the diagonal on which the
// winning path left wavefront 0 — needed because the reported score is taken over the non-free
// part of the alignment only (pinned by src/wfa2_utils.rs:294-308).
//
// Each thread owns 4 consecutive diagonals per iteration (one LDS.128 per source row, one STS.128 out) and
// runs recurrence, bounds and the first 8-base extend window of the four cells as straight-line code, so
// the LDS/XOR/FLO chains of four cells overlap.  Reads of source wavefronts are range-guarded
// from per-wavefront (lo,hi) metadata, so the ring never needs initialising or clearing.  With a single
// component, wavefront trimming only ever removes null cells, so it is not performed: the live range is the
// computed range clamped to the matrix diagonals.  That makes every wavefront's range a function of the two source
// ranges alone, never of cell values: ONE planner thread per group computes the plan of the next existing score
// (ranges, interior range, ring row offsets) while the group computes the current one, and after the score's single
// barrier (OR-reduction: "did a diagonal terminate") every thread reads the 12-word plan instead of re-deriving it;
// scores whose wavefront does not exist cost no barrier.
// Integer-ALU bound (alu pipe 54-70 % busy, issue ~67 %); no tensor cores (integer DP).
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
  int wcap;          // cells per ring row (multiple of 4, >= max(plen+tlen) + 12)
  int win_cap;       // uint2 windows reserved per sequence
  int group_bytes;   // shared bytes per group
  int max_score;
  unsigned long long* cells;
};

// Group synchronisation.  WPP == 1: the group is a warp.  WPP >= 2: the host launches exactly one group per
// block, so the group barrier is the block barrier (barrier 0) — a dynamic named-barrier id would make ptxas
// reserve all 16 barriers per block and cap residency at 4 blocks per SM.
template <int WPP>
__device__ __forceinline__ void k1_gsync(int) {
  if (WPP == 1) __syncwarp();
  else __syncthreads();
}

// group barrier + OR-reduction of a predicate
template <int WPP>
__device__ __forceinline__ bool k1_gsync_or(int, bool pred) {
  if (WPP == 1) { const bool r = __any_sync(FULL, pred); __syncwarp(); return r; }
  return __syncthreads_or(pred) != 0;
}

// count of leading zero bits; 0xFFFFFFFF for x == 0 (bfind.shiftamt = FLO.U32.SH, no fix-up for zero)
__device__ __forceinline__ uint32_t k1_clz(uint32_t x) {
  uint32_t d;
  asm("bfind.shiftamt.u32 %0, %1;" : "=r"(d) : "r"(x));
  return d;
}

// coop_extend (sfc_device.cuh) over the per-base windows
__device__ __forceinline__ int k1_coop_extend(const uint32_t* __restrict__ wp, const uint32_t* __restrict__ wt, int plen, int tlen, int k, int h0, int lane) {
  int adv = 0;
  for (;;) {
    const int h = h0 + adv + lane * 8, v = h - k;
    uint32_t x = ~0u;
    if (h <= tlen && v <= plen) x = wp[v] ^ wt[h];
    const unsigned full = __ballot_sync(FULL, x == 0u);
    if (full == FULL) { adv += 256; continue; }
    const int first = __ffs((int)~full) - 1;
    const int part = __shfl_sync(FULL, x ? (int)(k1_clz(x) >> 2) : 0, first);
    return adv + first * 8 + part;
  }
}

// Plan record of one score (written by the group's planner thread one score ahead, read by everybody after the barrier).
// Without trimming the diagonal range of every wavefront follows from the ranges of its two source wavefronts alone,
// never from cell values, so the planner runs ahead of the cells: no thread recomputes ranges, ring slots and row
// addresses per score, and scores whose wavefront does not exist cost no barrier.
enum { K1P_S = 0 /* score; -1: score cap reached */, K1P_LO, K1P_HI, K1P_ILO, K1P_IHI, K1P_LOX, K1P_CNTX, K1P_LOI, K1P_CNTI,
       K1P_ROWX /* ring row offsets (cells) */, K1P_ROWI, K1P_ROWO, K1P_N };

template <int WPP>
__global__ void __launch_bounds__(256) k1_score(const K1Params P) {
  extern __shared__ __align__(16) unsigned char k1_smem[];
  constexpr int G = WPP * 32;
  const int group = threadIdx.x / G, g = threadIdx.x % G, lane = threadIdx.x & 31, gw = g >> 5;
  unsigned char* gs = k1_smem + (size_t)group * P.group_bytes;
  int32_t* ring = reinterpret_cast<int32_t*>(gs);
  uint32_t* wbase = reinterpret_cast<uint32_t*>(gs + (size_t)P.rd * P.wcap * 4);  // wcap words: both sequences' windows
  int* plan = reinterpret_cast<int*>(wbase + P.wcap);  // [2][K1P_N], double buffered by score iteration
  int* meta = plan + 2 * K1P_N;                         // [rd][2] lo,hi of the wavefront in each ring slot (planner only)
  int* slot_pair = meta + 2 * P.rd;                     // [0] work index, [1] terminating diagonal
  const int rd = P.rd, wcap = P.wcap;
  const bool planner = g == G - 32;  // lane 0 of the group's last warp (the warp with the fewest chunks)
  unsigned long long cells_acc = 0;

  for (;;) {
    if (g == 0) { reinterpret_cast<volatile int*>(slot_pair)[0] = (int)atomicAdd(P.counter, 1u); slot_pair[1] = INT_MAX; }
    k1_gsync<WPP>(group);
    const uint32_t wi = (uint32_t) reinterpret_cast<volatile int*>(slot_pair)[0];
    if (wi >= P.n) break;
    const uint32_t pi = P.order[wi];
    const DevPair pr = P.pairs[pi];
    const int plen = pr.plen, tlen = pr.tlen;
    // One 32-bit window PER BASE: wp[i] = the 8 bases from pattern position i (first base in the top nibble; positions
    // past the end hold the end sentinel), so a cell's extension probe is two plain LDS.32, an XOR and a count of
    // leading zeros.  wt[-1] (= the word after the pattern's last window) is what a null cell (offset -1) reads.
    uint32_t* wp = wbase;
    uint32_t* wt = wbase + plen + 2;
    {
      const uint32_t* gp = P.nib + pr.p_woff;
      const uint32_t* gt = P.nib + pr.t_woff;
      for (int i = g; i <= plen; i += G) wp[i] = __brev(__funnelshift_r(gp[i >> 3], gp[(i >> 3) + 1], i << 2));
      for (int i = g; i <= tlen; i += G) wt[i] = __brev(__funnelshift_r(gt[i >> 3], gt[(i >> 3) + 1], i << 2));
    }
    const int kbase = plen + 4;  // element index e = k + kbase >= 3 for every diagonal that is ever read

    // planner state: the last score planned and its ring slot (s modulo rd, kept incrementally)
    int pl_s = 0, pl_slot = 0;
    auto plan_next = [&](int* out) {
      for (;;) {
        ++pl_s; pl_slot = pl_slot + 1 == rd ? 0 : pl_slot + 1;
        if (pl_s > P.max_score) { out[K1P_S] = -1; return; }
        const int sx = pl_s - P.x, si = pl_s - P.ip;
        const int slot_x = pl_slot - P.x + (pl_slot - P.x < 0 ? rd : 0), slot_i = pl_slot - P.ip + (pl_slot - P.ip < 0 ? rd : 0);
        int lox = 1, hix = -1, loi = 1, hii = -1;
        if (sx >= 0) { lox = meta[2 * slot_x]; hix = meta[2 * slot_x + 1]; }
        if (si >= 0) { loi = meta[2 * slot_i]; hii = meta[2 * slot_i + 1]; }
        const bool hasx = lox <= hix, hasi = loi <= hii;
        if (!hasx && !hasi) { meta[2 * pl_slot] = 1; meta[2 * pl_slot + 1] = -1; continue; }  // wavefront does not exist
        // computed range, clamped to the diagonals of the matrix (cells outside are always null)
        const int lo = max(min(hasx ? lox : INT_MAX, hasi ? loi - 1 : INT_MAX), -plen);
        const int hi = min(max(hasx ? hix : INT_MIN, hasi ? hii + 1 : INT_MIN), tlen);
        meta[2 * pl_slot] = lo; meta[2 * pl_slot + 1] = hi;
        out[K1P_S] = pl_s; out[K1P_LO] = lo; out[K1P_HI] = hi;
        out[K1P_ILO] = max(max(lox, loi + 1), lo); out[K1P_IHI] = min(min(hix, hii - 1), hi);  // every read in range, every cell active
        out[K1P_LOX] = lox; out[K1P_CNTX] = max(hix - lox + 1, 0); out[K1P_LOI] = loi; out[K1P_CNTI] = max(hii - loi + 1, 0);
        out[K1P_ROWX] = slot_x * wcap; out[K1P_ROWI] = slot_i * wcap; out[K1P_ROWO] = pl_slot * wcap;
        return;
      }
    };
    if (planner) {  // score 0: the free-begin diagonals
      int* out = plan;
      meta[0] = -pr.pbf; meta[1] = pr.tbf;
      out[K1P_S] = 0; out[K1P_LO] = -pr.pbf; out[K1P_HI] = pr.tbf; out[K1P_ILO] = 1; out[K1P_IHI] = 0;
      out[K1P_LOX] = 1; out[K1P_CNTX] = 0; out[K1P_LOI] = 1; out[K1P_CNTI] = 0; out[K1P_ROWX] = 0; out[K1P_ROWI] = 0; out[K1P_ROWO] = 0;
    }
    k1_gsync<WPP>(group);

    int s = 0, res_status = 0, res_score = 0;
    for (int it = 0;; ++it) {
      const int* pl = plan + (it & 1) * K1P_N;
      const int4 q0 = *reinterpret_cast<const int4*>(pl), q1 = *reinterpret_cast<const int4*>(pl + 4), q2 = *reinterpret_cast<const int4*>(pl + 8);
      s = q0.x;
      if (s < 0) { res_status = -7; break; }
      const int lo = q0.y, hi = q0.z, ilo = q0.w, ihi = q1.x, lox = q1.y, cntx = q1.z, loi = q1.w, cnti = q2.x;
      const int32_t* rowX = ring + q2.y;
      const int32_t* rowI = ring + q2.z;
      int32_t* rowO = ring + q2.w;
      if (planner) { cells_acc += (unsigned long long)(hi - lo + 1); plan_next(plan + ((it + 1) & 1) * K1P_N); }
      int term_k = INT_MAX;
      const int g_lo = (lo + kbase) >> 2, g_hi = (hi + kbase) >> 2;

      for (int gb = g_lo + gw * 32; gb <= g_hi; gb += G) {
        const int gi = gb + lane;
        const bool gact = gi <= g_hi;
        const int e0 = gi << 2, k0 = e0 - kbase;
        int32_t cell[4];
        if (s == 0) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int k = k0 + j;
            cell[j] = (k >= lo && k <= hi) ? ((max(k, 0) << 16) | (k + K1_K0BIAS)) : WNULL;
          }
        } else {
          // no guard for the lanes past the end of the wavefront: what they read lies inside this group's shared memory
          // (the next ring row or the windows), their chunk is a boundary chunk (every element range-checked below),
          // and they store nothing
          int4 a = *reinterpret_cast<const int4*>(rowX + e0);
          int4 b = *reinterpret_cast<const int4*>(rowI + e0);
          int32_t bl = rowI[e0 - 1];
          int32_t br = rowI[e0 + 4];
          const int kA = (gb << 2) - kbase;
          if (!(kA >= ilo && kA + 127 <= ihi)) {
            // boundary chunk: every element read is checked against its source row's live range — the four of the
            // mismatch row and the SIX distinct ones of the indel row (k0-1 .. k0+4; the k-1 and k+1 neighbours of the
            // four cells overlap).  No test against this wavefront's own [lo, hi]: a diagonal with a live source lies
            // inside the union of the source ranges by construction, and diagonals outside the matrix are nulled by the
            // bounds test below.
            const unsigned dx = (unsigned)(k0 - lox), di = (unsigned)(k0 - 1 - loi), ux = (unsigned)cntx, ui = (unsigned)cnti;
            a.x = dx < ux ? a.x : WNULL; a.y = dx + 1u < ux ? a.y : WNULL; a.z = dx + 2u < ux ? a.z : WNULL; a.w = dx + 3u < ux ? a.w : WNULL;
            bl = di < ui ? bl : WNULL;
            b.x = di + 1u < ui ? b.x : WNULL; b.y = di + 2u < ui ? b.y : WNULL; b.z = di + 3u < ui ? b.z : WNULL; b.w = di + 4u < ui ? b.w : WNULL;
            br = di + 5u < ui ? br : WNULL;
          }
          const int32_t av[4] = {a.x, a.y, a.z, a.w};
          const int32_t bm[4] = {bl, b.x, b.y, b.z};  // k-1
          const int32_t bp[4] = {b.y, b.z, b.w, br};  // k+1
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int32_t mis = av[j] + (K1_ONE | K1_TX), ins = bm[j] + K1_ONE, del = bp[j] | K1_TD;
            cell[j] = max(max(mis, del), ins) & ~K1_TMASK;
          }
        }
        // bounds + first 8-base window of the four cells (independent chains)
        uint32_t xw[4];
        int off[4];
        const unsigned char* wpk = reinterpret_cast<const unsigned char*>(wp) - 4 * k0;  // pattern position of (offset, diagonal k0 + j) = offset - k0 - j
        const unsigned char* wtb = reinterpret_cast<const unsigned char*>(wt);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int k = k0 + j;
          off[j] = cell[j] >> 16;
          if ((unsigned)off[j] > (unsigned)tlen || (unsigned)(off[j] - k) > (unsigned)plen) { cell[j] = WNULL; off[j] = -1; }
          // no address guard for null cells (off = -1): pattern index -1 - k lies between -(tlen + 1) and plen - 1 and text
          // index -1 just below 0, i.e. inside this group's own shared memory (the ring precedes the pattern windows, the
          // pattern windows precede the text windows); what is loaded there is made a mismatch by the sign of the offset.
          const int ob = off[j] << 2;
          xw[j] = (*reinterpret_cast<const uint32_t*>(wpk + ob - 4 * j) ^ *reinterpret_cast<const uint32_t*>(wtb + ob)) | (uint32_t)(off[j] >> 31);
        }
        int advmax = 0;  // 8 <=> some live cell matched its whole first window
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int a = (int)min(k1_clz(xw[j]) >> 2, 8u);  // equal bases in the window; 8 when xw == 0 (bfind gives ~0); 0 for a null cell
          off[j] += a;
          cell[j] += a << 16;  // the offset field moves with it
          advmax = max(advmax, a);
        }
        if (__any_sync(FULL, advmax == 8)) {  // rare: a run longer than 8 bases (the true diagonal)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int k = k0 + j;
            bool pending = false;
            if (off[j] >= 0 && xw[j] == 0u) {
              const uint32_t x2 = wp[off[j] - k] ^ wt[off[j]];  // 8 real bases matched, so both positions are inside the sequences
              if (x2) off[j] += (int)(k1_clz(x2) >> 2);
              else { off[j] += 8; pending = true; }
            }
            unsigned pm = __ballot_sync(FULL, pending);
            while (pm) {
              const int src = __ffs((int)pm) - 1;
              pm &= pm - 1;
              const int ko = __shfl_sync(FULL, k, src), ho = __shfl_sync(FULL, off[j], src);
              const int adv = k1_coop_extend(wp, wt, plen, tlen, ko, ho, lane);
              if (lane == src) off[j] += adv;
            }
            if (off[j] >= 0) cell[j] = (cell[j] & 0xFFFF) | (off[j] << 16);
          }
        }
        // a cell touches the end of a sequence <=> the largest text / pattern position of the four does
        const int hmax = max(max(off[0], off[1]), max(off[2], off[3]));
        const int vmax = max(max(off[0] - k0, off[1] - k0 - 1), max(off[2] - k0 - 2, off[3] - k0 - 3));
        if (hmax >= tlen || vmax >= plen) {  // a cell touches the end of a sequence: full ends-free termination test (rare)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int k = k0 + j, v = off[j] - k;
            if (off[j] >= 0 && ((off[j] >= tlen && plen - v <= pr.pef) || (v >= plen && tlen - off[j] <= pr.tef)))
              term_k = min(term_k, k);
          }
        }
        if (gact) *reinterpret_cast<int4*>(rowO + e0) = make_int4(cell[0], cell[1], cell[2], cell[3]);
      }
      // ---- one barrier per score: rows and the next plan complete + did any diagonal terminate ----
      term_k = warp_min(term_k);
      if (term_k != INT_MAX && lane == 0) atomicMin(slot_pair + 1, term_k);
      if (k1_gsync_or<WPP>(group, term_k != INT_MAX)) {
        if (WPP > 1) k1_gsync<WPP>(group);  // the atomicMin of every warp is visible
        const int tk = reinterpret_cast<volatile int*>(slot_pair)[1];
        const int32_t cellv = rowO[tk + kbase];
        const int offv = cellv >> 16, k0v = (cellv & 0x3FFF) - K1_K0BIAS;
        const int aligned = offv + (offv - tk) - abs(k0v);
        res_score = (P.match == 0) ? -s : ((-P.match) * aligned - s) / 2;
        break;
      }
    }
    if (g == 0) { P.scores[pi] = res_score; P.status[pi] = res_status; }
    k1_gsync<WPP>(group);  // ring/windows are reused by the next pair
  }
  if (planner && cells_acc) atomicAdd(P.cells, cells_acc);
}

}  // namespace sfc
