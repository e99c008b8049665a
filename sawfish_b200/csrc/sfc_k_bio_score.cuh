// sfc_k_bio_score.cuh — K5: rust-bio "custom" affine alignment with clipping, SCORE ONLY.
//
// The merge aligner of joint-call's haplotype de-duplication consumes nothing but the score
// (reference src/joint_call/merge_haplotypes/single_region/merge_pools.rs:253-272: norm score >= 0.97;
// multi_region/merge_pools.rs:57), and it is called O(haplotype pool^2) times per SV group.  K5 computes exactly the
// score K4 / oracle/bio_oracle.c compute (same three-layer recurrence over the whole matrix, same boundary rows and
// columns, same suffix-clip maxima and the final "insertion layer once more" pass) without a traceback matrix.
// Supported here: every scoring whose x-suffix clip is disabled and whose clip penalties are <= 0 (new_merge_aligner,
// new_glocal_*_aligner, new_pattern_prefix_aligner of src/bio_align_utils.rs:104-157); the host routes the rest to K4.
//
// What drops out of the score (proved from the recurrences, so the result stays bit-identical):
//   * the y-prefix-clip candidate yclip_prefix + go + ge*i never exceeds I[j][i] (I[j][1] >= S[j][0] + go + ge and
//     S[j][0] >= yclip_prefix), and every comparison is strict;
//   * with xclip_suffix disabled, S + xclip_suffix stays ~MIN_SCORE and never beats a real cell;
//   * the final passes reduce to  max_i ( max(S[n][i], max_j S[j][i] + yclip_suffix) + (m - i) * (go + ge) ).
//
// Mapping: one CTA of four warps per pair (persistent CTAs pull pairs, heaviest first).  The pattern x is cut into
// bands of 32*R rows (R = 4, 8 or 16 by pattern length, so that even a 500-base query gives every warp a band); inside
// a band lane l owns R consecutive rows in registers (S + go + ge of the previous column, D of the previous column,
// the row maximum for the y-suffix clip) and the warp sweeps the text as a systolic wavefront: at step t lane l does
// column t - l, taking S of the row above and its own first row's I from lane l - 1 by shuffle.  The last row of a
// band goes to a boundary array (8 bytes per column, L2 resident) that the warp doing the next band reads 32 columns
// at a time, one batch ahead, as soon as a progress word in shared memory says they are there: the four warps work on
// four consecutive bands at once, each 64 columns behind the previous one (band b reads buffer b % 5, writes buffer
// (b + 1) % 5; the warp that last read a buffer is the one that overwrites it, so producers never wait).
// 10 integer instructions per matrix cell: 6 on the alu pipe (4 x VIADDMNMX, 2 x VIMNMX), 4 on the fma pipe (the
// substitution score as dg + ma - min((x - y)^2, 1) * (ma - mi)); bound by the alu pipe (ncu: 86.6 % busy in the
// 7-alu-instruction version).  No tensor cores (integer DP).
#pragma once
#include "sfc_k_bio.cuh"

namespace sfc {

struct K5Params {
  const uint8_t* arena;
  const K4Pair* pairs;
  const uint32_t* order;  // pair indices, heaviest first
  uint32_t n;
  uint32_t* counter;
  K4Scoring sc;
  int2* bnd;                     // [CTAs][K5_NBUF][bnd_stride]
  unsigned long long bnd_stride; // entries per boundary buffer (>= max text length + 1)
  int32_t* scores;
  int32_t* status;
  unsigned long long* cells;
};

constexpr int K5_WARPS = 4, K5_NBUF = K5_WARPS + 1, K5_TAG_SHIFT = 20;  // progress word = (band tag << 20) | last column written

struct K5Batch { int s, i, y; };

// Boundary rows are written by lane 31 of ANOTHER warp of the CTA (the one doing the previous band, 64+ columns ahead):
// wait on that band's progress word, then read through L2 (never the non-coherent path; the producer's fence + progress
// store and the fence here order the data).
__device__ __forceinline__ K5Batch k5_load(const int2* in, const volatile int* prog_in, int in_tag,
                                           const uint8_t* __restrict__ y, int n, int base, int lane) {
  K5Batch b{K4_MIN_SCORE, K4_MIN_SCORE, 0};
  const int idx = base + lane;
  if (base <= n) {  // wait until the producing band has published columns base .. min(base + 31, n)
    const int need = (in_tag << K5_TAG_SHIFT) | min(base + 31, n);
    while (*prog_in < need) __nanosleep(32);
    __threadfence_block();
  }
  if (idx <= n) {
    const int2 e = __ldcg(in + idx);
    b.s = e.x; b.i = e.y;
    b.y = idx >= 1 ? (int)y[idx - 1] : 0;
  }
  return b;
}

// One band of 32*R rows (i0+1 ..), all n columns.  Returns this lane's best final-pass candidate of its rows.
template <int R, bool XP>
__device__ __forceinline__ int k5_band(const K4Scoring& sc, const uint8_t* __restrict__ x, const uint8_t* __restrict__ y,
                                       const int m, const int n, const int i0, const int2* in, int2* out,
                                       const volatile int* prog_in, volatile int* prog_out, const int tag,
                                       const int lane) {
  const int ge = sc.ge, goe = sc.go + sc.ge, dsub = sc.ma - sc.mi;
  int xb[R], SpG[R], Dp[R], SnG[R];
  const int top = i0 + lane * R;  // the row above this lane's first row
  auto col0_I = [&](int i) { return i == 1 ? goe : max(sc.go + ge * i, sc.xp + goe); };
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int i = top + r + 1;
    xb[r] = i <= m ? (int)x[i - 1] : 256;  // rows past the pattern never match; their values are not used
    const int s0 = max(col0_I(i), sc.xp);
    SpG[r] = s0 + goe; Dp[r] = K4_MIN_SCORE; SnG[r] = SpG[r];
  }
  K5Batch cur = k5_load(in, prog_in, tag, y, n, 0, lane), nxt = k5_load(in, prog_in, tag, y, n, 32, lane);
  int diag = __shfl_sync(FULL, cur.s, 0);                       // lane 0: S[0][i0] + goe
  if (lane) diag = max(col0_I(top), sc.xp) + goe;               // column 0 of the row above
  if (lane == 31) out[0] = make_int2(SpG[R - 1], K4_MIN_SCORE);  // only .x of entry 0 is ever read
  int outS = 0, outI = 0, yc = 0;
  const int t_end = n + 31;
  for (int t = 1; t <= t_end; ++t) {
    const int src = t & 31;
    if (src == 0) { cur = nxt; nxt = k5_load(in, prog_in, tag, y, n, t + 32, lane); }
    int upS = __shfl_up_sync(FULL, outS, 1), upI = __shfl_up_sync(FULL, outI, 1), ycn = __shfl_up_sync(FULL, yc, 1);
    const int injS = __shfl_sync(FULL, cur.s, src), injI = __shfl_sync(FULL, cur.i, src), injY = __shfl_sync(FULL, cur.y, src);
    if (lane == 0) { upS = injS; upI = injI; ycn = injY; }
    yc = ycn;
    const int j = t - lane;
    if (j >= 1 && j <= n) {
      int xclipjG = 0;
      if (XP) { const int gap_j = sc.go + ge * j; xclipjG = sc.xp + (sc.yp > gap_j ? sc.yp : gap_j) + goe; }
      // With G = go + ge added to every stored S:  T = max(diagonal, D) + G needs nothing from the row above, so the
      // only dependence that runs down the rows is  I(next row) = max(I + ge, T)  (I + G <= I + ge because go <= 0):
      // one VIADDMNMX per row on the critical path.
      int dgma = diag + sc.ma, I = upI, SG = upS;  // diagonal + match score; dsub = match - mismatch
#pragma unroll
      for (int r = 0; r < R; ++r) {
        // substitution score without compare + select (the alu pipe is the bound, 7 of 8 instructions per cell sat on it):
        // (x - y)^2 is 0 for equal symbols and >= 1 otherwise; the subtract, the square and the final multiply-add issue on
        // the fma pipe, only the min stays on the alu pipe
        const int dxy = xb[r] - yc;
        const int msG = dgma - min(dxy * dxy, 1) * dsub;
        const int D = max(Dp[r] + ge, SpG[r]);
        int T = max(D + goe, msG);
        if (XP) T = max(T, xclipjG);
        SG = max(I + goe, T);
        dgma = SpG[r] + sc.ma;
        SpG[r] = SG; Dp[r] = D; SnG[r] = max(SnG[r], SG);
        I = max(I + ge, T);
      }
      diag = upS; outS = SG; outI = I;  // S + G of the last row, and the I the row below it will use
      if (lane == 31) {
        out[j] = make_int2(outS, outI);
        if ((j & 31) == 31 || j == n) {  // publish: data first, then the progress word
          __threadfence_block();
          *prog_out = ((tag + 1) << K5_TAG_SHIFT) | j;
        }
      }
    }
  }
  int best = INT_MIN;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int i = top + r + 1;
    if (i <= m) {
      const int sfin = SpG[r] - goe, smax = SnG[r] - goe;
      best = max(best, max(sfin, smax + sc.ys) + (m - i) * goe);
    }
  }
  return best;
}

// rows of the band that starts at row i0: short patterns get low bands so that all four warps have one
__device__ __forceinline__ int k5_band_R(int m, int i0) {
  if (m <= 512) return 4;
  if (m <= 1024) return 8;
  const int rem = m - i0;
  return rem >= 384 ? 16 : rem > 128 ? 8 : 4;
}

template <bool XP>
__global__ void __launch_bounds__(K5_WARPS * 32) k5_bio_score(const K5Params P) {
  __shared__ int s_pair;
  __shared__ int s_prog[K5_NBUF];
  __shared__ int s_best[K5_WARPS];
  volatile int* prog = s_prog;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int2* bnd = P.bnd + (size_t)blockIdx.x * K5_NBUF * P.bnd_stride;
  const K4Scoring sc = P.sc;
  const int goe = sc.go + sc.ge;
  unsigned long long cells_acc = 0;
  for (;;) {
    __syncthreads();  // the previous pair is finished: progress words and s_best may be reused
    if (threadIdx.x == 0) {
      s_pair = (int)atomicAdd(P.counter, 1u);
      for (int k = 0; k < K5_NBUF; ++k) prog[k] = -1;
    }
    __syncthreads();
    const uint32_t wi = (uint32_t)s_pair;
    if (wi >= P.n) break;
    const uint32_t pi = P.order[wi];
    const K4Pair pr = P.pairs[pi];
    const int m = pr.m, n = pr.n;
    const uint8_t* x = P.arena + pr.x_off;
    const uint8_t* y = P.arena + pr.y_off;
    int row0 = 0;
    if (w == 0) {
      // ---- row 0 into buffer 0: S[j][0] + G and the I of row 1 (= S[j][0] + G, because I[j][0] = MIN); the
      // y-suffix clip of row 0 lands in column n ----
      int sn0 = sc.ys;  // column 0: S[0][0] + yclip_suffix
      for (int j = lane; j <= n; j += 32) {
        int s = 0;
        if (j >= 1) {
          const int d0 = j == 1 ? goe : max(sc.go + sc.ge * j, sc.yp + goe);
          s = max(d0, sc.yp);
          if (j < n) sn0 = max(sn0, s + sc.ys);
        }
        if (j < n) bnd[j] = make_int2(s + goe, s + goe);
      }
      sn0 = warp_max(sn0);
      const int d0 = n == 1 ? goe : max(sc.go + sc.ge * n, sc.yp + goe);
      row0 = max(max(d0, sc.yp), sn0);
      if (lane == 0) bnd[n] = make_int2(row0 + goe, row0 + goe);
      __threadfence_block();
      __syncwarp();
      if (lane == 0) { __threadfence_block(); prog[0] = n; }  // tag 0, all columns
    }
    int best = INT_MIN;
    int b = 0;
    for (int i0 = 0; i0 < m; ++b) {
      const int R = k5_band_R(m, i0);
      if ((b & (K5_WARPS - 1)) == w) {
        const int si = b % K5_NBUF, so = (b + 1) % K5_NBUF;
        const int2* in = bnd + (size_t)si * P.bnd_stride;
        int2* out = bnd + (size_t)so * P.bnd_stride;
        int v;
        if (R == 16) v = k5_band<16, XP>(sc, x, y, m, n, i0, in, out, prog + si, prog + so, b, lane);
        else if (R == 8) v = k5_band<8, XP>(sc, x, y, m, n, i0, in, out, prog + si, prog + so, b, lane);
        else v = k5_band<4, XP>(sc, x, y, m, n, i0, in, out, prog + si, prog + so, b, lane);
        best = max(best, v);
      }
      i0 += 32 * R;
    }
    best = warp_max(best);
    if (lane == 0) s_best[w] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
      int r = row0 + m * goe;
      for (int k = 0; k < K5_WARPS; ++k) r = max(r, s_best[k]);
      P.scores[pi] = r;
      P.status[pi] = 0;
      cells_acc += (unsigned long long)m * (unsigned long long)n;
    }
  }
  if (threadIdx.x == 0 && cells_acc) atomicAdd(P.cells, cells_acc);
}

}  // namespace sfc
