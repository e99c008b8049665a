// sfc_k_bio.cuh — K4: affine-gap alignment with prefix/suffix clipping (rust-bio "custom" semantics).
//
// Replaces bio::alignment::pairwise::banded::Aligner::custom_with_matches as sawfish's
// bio_align_utils::PairwiseAligner drives it (reference src/bio_align_utils.rs:97-188): the merge aligner of
// joint-call's haplotype de-duplication (src/joint_call/merge_haplotypes/single_region/merge_pools.rs:214-272, score
// only) and the glocal / overlap aligners of the banded pairwise assembler.  Same recurrences, comparison order and
// tie-breaks as oracle/bio_oracle.c (three layers S/I/D, strict '>' in the order diagonal, insertion, deletion,
// x-prefix clip, y-prefix clip; suffix clips as running maxima where the first maximum wins; 4-bit traceback code per
// layer).  The k-mer band of the reference is not applied: the whole matrix is filled, which yields the same
// alignment whenever the band contains the optimal path.
//
// Mapping: one CTA per pair (persistent CTAs pull pairs from a work counter); anti-diagonals d = i + j are swept with
// one block barrier each, a thread owning the rows i = tid, tid + T, ...  The three layers of the last two
// anti-diagonals and the traceback matrix (2 bytes per cell) live in the CTA's slab of device memory.  The running
// maximum of a row (y-suffix clip) stays with the row's thread; the running maximum of a column (x-suffix clip,
// only when that clip is enabled) is a 64-bit atomicMax on (score, rows left) so that the first maximum wins.
// Integer DP bound by barrier latency for short pairs and by L2 traffic of the traceback for long ones; it is a
// small share of joint-call (the reference runs it O(haplotype pool^2) per SV group).
#pragma once
#include "sfc_device.cuh"

namespace sfc {

constexpr int K4_MIN_SCORE = -858993459;  // rust-bio's MIN_SCORE
enum { K4_TB_START = 0, K4_TB_INS, K4_TB_DEL, K4_TB_SUBST, K4_TB_MATCH, K4_TB_XCLIP_PREFIX, K4_TB_XCLIP_SUFFIX,
       K4_TB_YCLIP_PREFIX, K4_TB_YCLIP_SUFFIX };
enum { K4_OP_MATCH = 0, K4_OP_SUBST, K4_OP_INS, K4_OP_DEL, K4_OP_XCLIP, K4_OP_YCLIP };

struct K4Scoring { int go, ge, ma, mi, xp, xs, yp, ys; };
struct K4Pair { unsigned long long x_off, y_off; int m, n; };

struct K4Params {
  const uint8_t* arena;
  const K4Pair* pairs;
  uint32_t n;
  uint32_t* counter;
  K4Scoring sc;
  unsigned char* slabs;
  unsigned long long slab_bytes;
  int32_t* scores;
  int32_t* status;
  int32_t* bounds;  // [n][4] xstart, xend, ystart, yend
  uint32_t* pool;   // run-length operations (len << 3 | code)
  unsigned long long pool_cap;
  unsigned long long* pool_used;
  unsigned long long* out_off;
  uint32_t* out_len;
};

__host__ __device__ inline size_t k4_align16(size_t v) { return (v + 15) & ~(size_t)15; }
// bytes of slab one pair needs (host and device agree through this function)
__host__ __device__ inline size_t k4_slab_need(int m, int n) {
  const size_t W = (size_t)m + 1, N = (size_t)n + 1;
  return k4_align16(2 * W * N) + 13 * k4_align16(4 * W) + k4_align16(4 * N) + k4_align16(8 * N) +
         k4_align16(4 * (2 * (W + N) + 16)) + 256;
}

__global__ void __launch_bounds__(256) k4_bio_align(const K4Params P) {
  __shared__ int s_pair;
  const int tid = threadIdx.x, NT = blockDim.x;
  unsigned char* slab = P.slabs + (size_t)blockIdx.x * P.slab_bytes;
  const K4Scoring sc = P.sc;
  const bool xs_on = sc.xs > K4_MIN_SCORE / 2;
  for (;;) {
    __syncthreads();
    if (tid == 0) s_pair = (int)atomicAdd(P.counter, 1u);
    __syncthreads();
    const uint32_t pi = (uint32_t)s_pair;
    if (pi >= P.n) break;
    const K4Pair pr = P.pairs[pi];
    const int m = pr.m, n = pr.n;
    const uint8_t* x = P.arena + pr.x_off;
    const uint8_t* y = P.arena + pr.y_off;
    const size_t W = (size_t)m + 1, N = (size_t)n + 1;
    if (k4_slab_need(m, n) > P.slab_bytes) {
      if (tid == 0) { P.scores[pi] = 0; P.status[pi] = -6; P.out_len[pi] = 0; P.out_off[pi] = 0; }
      continue;
    }
    // ---- carve the slab ----
    unsigned char* p = slab;
    uint16_t* tb = reinterpret_cast<uint16_t*>(p); p += k4_align16(2 * W * N);
    int32_t* Sb[3]; int32_t* Ib[2]; int32_t* Db[2]; int32_t* Tb[2];
    for (int k = 0; k < 3; ++k) { Sb[k] = reinterpret_cast<int32_t*>(p); p += k4_align16(4 * W); }
    for (int k = 0; k < 2; ++k) { Ib[k] = reinterpret_cast<int32_t*>(p); p += k4_align16(4 * W); }
    for (int k = 0; k < 2; ++k) { Db[k] = reinterpret_cast<int32_t*>(p); p += k4_align16(4 * W); }
    for (int k = 0; k < 2; ++k) { Tb[k] = reinterpret_cast<int32_t*>(p); p += k4_align16(4 * W); }  // S-layer code of the previous anti-diagonal
    int32_t* Sn = reinterpret_cast<int32_t*>(p); p += k4_align16(4 * W);
    int32_t* Ly = reinterpret_cast<int32_t*>(p); p += k4_align16(4 * W);
    int32_t* Scol = reinterpret_cast<int32_t*>(p); p += k4_align16(4 * W);  // last column, for the final fix-ups
    int32_t* Icol = reinterpret_cast<int32_t*>(p); p += k4_align16(4 * W);
    int32_t* Lx = reinterpret_cast<int32_t*>(p); p += k4_align16(4 * N);
    long long* XS = reinterpret_cast<long long*>(p); p += k4_align16(8 * N);  // (best S + xclip_suffix, rows left) per column
    uint32_t* tmp = reinterpret_cast<uint32_t*>(p);
    const int tmp_cap = 2 * (m + n) + 16;
    auto T = [&](int i, int j) -> uint16_t& { return tb[(size_t)j * W + (size_t)i]; };
    const long long xs_init = ((long long)K4_MIN_SCORE << 32) | 0xFFFFFFFFll;

    for (int i = tid; i <= m; i += NT) { Sn[i] = K4_MIN_SCORE; Ly[i] = 0; }
    for (int j = tid; j <= n; j += NT) { Lx[j] = 0; XS[j] = xs_init; }
    __syncthreads();
    if (tid == 0) { Sn[0] = sc.ys; Ly[0] = n; }
    __syncthreads();

    for (int d = 0; d <= m + n; ++d) {
      int32_t* Sd = Sb[d % 3]; const int32_t* S1 = Sb[(d + 2) % 3]; const int32_t* S2 = Sb[(d + 1) % 3];
      int32_t* Id = Ib[d & 1]; const int32_t* I1 = Ib[(d + 1) & 1];
      int32_t* Dd = Db[d & 1]; const int32_t* D1 = Db[(d + 1) & 1];
      int32_t* Td = Tb[d & 1]; const int32_t* T1 = Tb[(d + 1) & 1];
      const int ilo = max(0, d - n), ihi = min(m, d);
      for (int i = ilo + tid; i <= ihi; i += NT) {
        const int j = d - i;
        int S, I, D;
        int ts = K4_TB_START, ti = K4_TB_START, td = K4_TB_START;
        if (j == 0) {
          D = K4_MIN_SCORE;
          if (i == 0) { S = 0; I = K4_MIN_SCORE; }
          else {
            if (i == 1) { I = sc.go + sc.ge; }
            else {
              const int i_score = sc.go + sc.ge * i, c_score = sc.xp + sc.go + sc.ge;
              if (i_score > c_score) { I = i_score; ti = K4_TB_INS; } else { I = c_score; ti = K4_TB_XCLIP_PREFIX; }
            }
            S = K4_MIN_SCORE;
            if (i == m) { ts = K4_TB_XCLIP_SUFFIX; if (xs_on) { const long long key = XS[0]; S = (int)(key >> 32); Lx[0] = (unsigned)key == 0xFFFFFFFFu ? 0 : (int)(unsigned)key; } }
            if (I > S) { S = I; ts = K4_TB_INS; }
            if (sc.xp > S) { S = sc.xp; ts = K4_TB_XCLIP_PREFIX; }
            if (xs_on && i != m) atomicMax(XS + 0, ((long long)(S + sc.xs) << 32) | (unsigned)(m - i));
            if (S + sc.ys > Sn[i]) { Sn[i] = S + sc.ys; Ly[i] = n; }
          }
        } else if (i == 0) {
          I = K4_MIN_SCORE;
          if (j == 1) { D = sc.go + sc.ge; }
          else {
            const int d_score = sc.go + sc.ge * j, c_score = sc.yp + sc.go + sc.ge;
            if (d_score > c_score) { D = d_score; td = K4_TB_DEL; } else { D = c_score; td = K4_TB_YCLIP_PREFIX; }
          }
          if (D > sc.yp) { S = D; ts = K4_TB_DEL; } else { S = sc.yp; ts = K4_TB_YCLIP_PREFIX; }
          if (j == n && Sn[0] > S) { S = Sn[0]; ts = K4_TB_YCLIP_SUFFIX; }
          else if (S + sc.ys > Sn[0]) { Sn[0] = S + sc.ys; Ly[0] = n - j; }
        } else {
          const uint8_t pc = x[i - 1], qc = y[j - 1];
          const int m_score = S2[i - 1] + (pc == qc ? sc.ma : sc.mi);
          {
            const int i_score = I1[i - 1] + sc.ge, s_score = S1[i - 1] + sc.go + sc.ge;  // from (i-1, j)
            if (i_score > s_score) { I = i_score; ti = K4_TB_INS; } else { I = s_score; ti = T1[i - 1]; }
          }
          {
            const int d_score = D1[i] + sc.ge, s_score = S1[i] + sc.go + sc.ge;  // from (i, j-1)
            if (d_score > s_score) { D = d_score; td = K4_TB_DEL; } else { D = s_score; td = T1[i]; }
          }
          ts = K4_TB_XCLIP_SUFFIX;
          S = K4_MIN_SCORE;
          if (xs_on && i == m) { const long long key = XS[j]; S = (int)(key >> 32); Lx[j] = (unsigned)key == 0xFFFFFFFFu ? 0 : (int)(unsigned)key; }
          if (m_score > S) { S = m_score; ts = pc == qc ? K4_TB_MATCH : K4_TB_SUBST; }
          if (I > S) { S = I; ts = K4_TB_INS; }
          if (D > S) { S = D; ts = K4_TB_DEL; }
          const int gap_j = sc.go + sc.ge * j;
          const int xclip_score = sc.xp + (sc.yp > gap_j ? sc.yp : gap_j);
          if (xclip_score > S) { S = xclip_score; ts = K4_TB_XCLIP_PREFIX; }
          const int yclip_score = sc.yp + sc.go + sc.ge * i;
          if (yclip_score > S) { S = yclip_score; ts = K4_TB_YCLIP_PREFIX; }
          if (xs_on && i != m) atomicMax(XS + j, ((long long)(S + sc.xs) << 32) | (unsigned)(m - i));
          if (S + sc.ys > Sn[i]) { Sn[i] = S + sc.ys; Ly[i] = n - j; }
        }
        Sd[i] = S; Id[i] = I; Dd[i] = D; Td[i] = ts;
        T(i, j) = (uint16_t)(ts | (ti << 4) | (td << 8));
        if (j == n) { Scol[i] = S; Icol[i] = I; }
      }
      __syncthreads();
    }

    // ---- last column: suffix clips, then the insertion layer once more; traceback (one thread) ----
    if (tid == 0) {
      const int j = n;
      auto set_s = [&](int i, int v) { uint16_t& c = T(i, j); c = (uint16_t)((c & ~15) | v); };
      auto set_i = [&](int i, int v) { uint16_t& c = T(i, j); c = (uint16_t)((c & ~(15 << 4)) | (v << 4)); };
      for (int i = 0; i <= m; ++i) {
        if (Sn[i] > Scol[i]) { Scol[i] = Sn[i]; set_s(i, K4_TB_YCLIP_SUFFIX); }
        if (Scol[i] + sc.xs > Scol[m]) { Scol[m] = Scol[i] + sc.xs; Lx[j] = m - i; set_s(m, K4_TB_XCLIP_SUFFIX); }
      }
      for (int i = 1; i <= m; ++i) {
        const int s_score = Scol[i - 1] + sc.go + sc.ge;
        if (s_score > Icol[i]) { Icol[i] = s_score; set_i(i, T(i - 1, j) & 15); }
        if (s_score > Scol[i]) {
          Scol[i] = s_score; set_s(i, K4_TB_INS);
          if (Scol[i] + sc.xs > Scol[m]) { Scol[m] = Scol[i] + sc.xs; Lx[j] = m - i; set_s(m, K4_TB_XCLIP_SUFFIX); }
        }
      }
      int i = m, jj = n, cnt = 0, bad = 0;
      int xstart = 0, xend = m, ystart = 0, yend = n;
      int layer = T(i, jj) & 15;
      int run_code = -1;
      uint32_t run_len = 0;
      auto push = [&](int code, int len, bool mergeable) {
        if (mergeable && code == run_code) { run_len += (uint32_t)len; return; }
        if (run_code >= 0) { if (cnt < tmp_cap) tmp[cnt++] = (run_len << 3) | (uint32_t)run_code; else bad = 1; }
        run_code = code; run_len = (uint32_t)len;
        if (!mergeable) { if (cnt < tmp_cap) tmp[cnt++] = (run_len << 3) | (uint32_t)run_code; else bad = 1; run_code = -1; }
      };
      int guard = 2 * (m + n) + 8;
      while (layer != K4_TB_START && !bad) {
        int next = K4_TB_START;
        const uint16_t c = T(i, jj);
        switch (layer) {
          case K4_TB_INS: push(K4_OP_INS, 1, true); next = (c >> 4) & 15; i -= 1; break;
          case K4_TB_DEL: push(K4_OP_DEL, 1, true); next = (c >> 8) & 15; jj -= 1; break;
          case K4_TB_MATCH: push(K4_OP_MATCH, 1, true); next = T(i - 1, jj - 1) & 15; i -= 1; jj -= 1; break;
          case K4_TB_SUBST: push(K4_OP_SUBST, 1, true); next = T(i - 1, jj - 1) & 15; i -= 1; jj -= 1; break;
          case K4_TB_XCLIP_PREFIX: push(K4_OP_XCLIP, i, false); xstart = i; i = 0; next = T(0, jj) & 15; break;
          case K4_TB_XCLIP_SUFFIX: push(K4_OP_XCLIP, Lx[jj], false); i -= Lx[jj]; xend = i; next = T(i, jj) & 15; break;
          case K4_TB_YCLIP_PREFIX: push(K4_OP_YCLIP, jj, false); ystart = jj; jj = 0; next = T(i, 0) & 15; break;
          case K4_TB_YCLIP_SUFFIX: push(K4_OP_YCLIP, Ly[i], false); jj -= Ly[i]; yend = jj; next = T(i, jj) & 15; break;
          default: bad = 1; break;
        }
        layer = next;
        if (--guard < 0 || i < 0 || jj < 0) bad = 1;
      }
      if (run_code >= 0) { if (cnt < tmp_cap) tmp[cnt++] = (run_len << 3) | (uint32_t)run_code; else bad = 1; }
      int st = bad ? -9 : 0;
      unsigned long long base = 0;
      uint32_t olen = 0;
      if (!bad) {
        base = atomicAdd(P.pool_used, (unsigned long long)cnt);
        if (base + (unsigned long long)cnt > P.pool_cap) st = -10;  // SFC_STATUS_POOL (cannot happen: the host sizes the pool for every op)
        else {
          for (int a = 0; a < cnt; ++a) P.pool[base + a] = tmp[cnt - 1 - a];  // alignment order
          olen = (uint32_t)cnt;
        }
      }
      P.scores[pi] = Scol[m];
      P.status[pi] = st;
      P.bounds[4 * (size_t)pi + 0] = xstart; P.bounds[4 * (size_t)pi + 1] = xend;
      P.bounds[4 * (size_t)pi + 2] = ystart; P.bounds[4 * (size_t)pi + 3] = yend;
      P.out_off[pi] = base; P.out_len[pi] = olen;
    }
  }
}

}  // namespace sfc
