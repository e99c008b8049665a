/*
 * wfa_oracle.c — CPU oracle (TEST INFRASTRUCTURE ONLY; see wfa_oracle.h).
 *
 * Restates, from the published WFA2-lib algorithm, exactly what sawfish's
 * src/wfa2_utils.rs obtains from WFA2-lib:
 *   - penalty adjustment for match < 0 (Eizenga):       wfa2_utils.rs:159-168,183-189
 *   - ends-free / end-to-end wavefront alignment:        wfa2_utils.rs:106-140
 *   - backtrace with WFA2's tie-breaking priority:       wfa2_utils.rs:216-218 (score, cigar)
 *   - WFA ops -> SimpleAlignment translation:            wfa2_utils.rs:13-95
 *   - reversal of both sequences:                        wfa2_utils.rs:198-215
 *
 * Two backtrace engines share one wavefront computation:
 *   LITERAL  keeps every wavefront and backtraces by comparing offsets with the
 *            WFA2 "piggy-back type" priority (what MemoryHigh does);
 *   COMPACT  keeps a modular ring of wavefronts plus one choice byte per
 *            (score, diagonal) recorded at compute time with the same priority
 *            (what MemoryMed's piggy-back is designed to reproduce).
 * In LITERAL mode both are run and must agree (error -4 otherwise).
 *
 * Semantics restated (each a parity hazard):
 *  1. h = text pos, v = pattern pos, k = h - v, offset = h. 'I' consumes text, 'D' pattern.
 *  2. match m<0: x' = 2x-2m; linear i' = 2i-m; affine o' = 2o, e' = 2e-m.
 *  3. score = (-m*(aligned_plen+aligned_tlen) - wf_score)/2 over the non-free part
 *     (F1 of SURVEY.md §8c.3; pinned by wfa2_utils.rs:294-308: 4, not 5).
 *  4. ends-free init: M[0][k], k in [-pattern_begin_free, text_begin_free].
 *  5. recurrences; M nulls out-of-matrix offsets; every component's lo/hi is
 *     trimmed to its first/last in-matrix entry and reads outside [lo,hi] are NULL.
 *  6. extend in ascending k, terminate at the first k satisfying the ends-free test.
 *  7. backtrace priority X > D2e > D2o > D1e > D1o > I2e > I2o > I1e > I1o.
 */
#include "wfa_oracle.h"
#include <limits.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ONULL (INT32_MIN / 2)
#define MAXI(a, b) ((a) > (b) ? (a) : (b))
#define MINI(a, b) ((a) < (b) ? (a) : (b))

enum { C_M = 0, C_I1 = 1, C_D1 = 2, C_I2 = 3, C_D2 = 4, NCOMP = 5 };
enum { SEL_X = 0, SEL_I1 = 1, SEL_I2 = 2, SEL_D1 = 3, SEL_D2 = 4 };
#define BIT_I1E 0x08
#define BIT_I2E 0x10
#define BIT_D1E 0x20
#define BIT_D2E 0x40

/* WFA2 piggy-back backtrace type codes (tie-break priority, larger wins). */
enum {
  BT_I1_OPEN = 1, BT_I1_EXT = 2, BT_I2_OPEN = 3, BT_I2_EXT = 4,
  BT_D1_OPEN = 5, BT_D1_EXT = 6, BT_D2_OPEN = 7, BT_D2_EXT = 8, BT_M = 9
};
#define BT_SET(off, type) ((((int64_t)(off)) << 4) | (int64_t)(type))
#define BT_OFF(v) ((int)((v) >> 4))
#define BT_TYPE(v) ((int)((v)&0xF))

typedef struct {
  const int32_t* row;
  int lo, hi, ex;
} src_t;

typedef struct {
  int plen, tlen;
  uint8_t *p, *t; /* padded copies */
  int metric, ncomp;
  int x, o1, e1, o2, e2, match;
  int pbf, pef, tbf, tef;
  int mode;
  int Wp, kbase;
  int depth[NCOMP];
  int32_t** rows[NCOMP];
  int* lo[NCOMP];
  int* hi[NCOMP];
  char* ex[NCOMP];
  int32_t* nullrow;
  uint8_t** bt;
  uint8_t** bt_chunks; /* choice bytes live in 4 MiB arena chunks (one malloc per chunk, not per score) */
  int n_bt_chunks, cap_bt_chunks;
  size_t bt_chunk_used, bt_chunk_size;
  int* btlo;
  int S_cap;
  size_t mem_used, mem_cap;
  int oom;
  ora_stats st;
} wfa_t;

static void* xalloc(wfa_t* w, size_t n) {
  if (w->mem_used + n > w->mem_cap) { w->oom = 1; return NULL; }
  void* p = malloc(n ? n : 1);
  if (!p) { w->oom = 1; return NULL; }
  w->mem_used += n;
  return p;
}

static uint8_t* bt_alloc(wfa_t* w, size_t n) {
  if (w->n_bt_chunks == 0 || w->bt_chunk_used + n > w->bt_chunk_size) {
    const size_t sz = n > ((size_t)4 << 20) ? n : ((size_t)4 << 20);
    if (w->n_bt_chunks == w->cap_bt_chunks) {
      const int nc = w->cap_bt_chunks ? 2 * w->cap_bt_chunks : 16;
      uint8_t** p = (uint8_t**)realloc(w->bt_chunks, sizeof(uint8_t*) * (size_t)nc);
      if (!p) { w->oom = 1; return NULL; }
      w->bt_chunks = p;
      w->cap_bt_chunks = nc;
    }
    uint8_t* c = (uint8_t*)xalloc(w, sz);
    if (!c) return NULL;
    w->bt_chunks[w->n_bt_chunks++] = c;
    w->bt_chunk_used = 0;
    w->bt_chunk_size = sz;
  }
  uint8_t* r = w->bt_chunks[w->n_bt_chunks - 1] + w->bt_chunk_used;
  w->bt_chunk_used += n;
  return r;
}

static inline int slot_of(const wfa_t* w, int c, int s) {
  return w->mode == ORA_MODE_LITERAL ? s : s % w->depth[c];
}

static inline int in_matrix(const wfa_t* w, int32_t off, int k) {
  const uint32_t h = (uint32_t)off, v = (uint32_t)(off - k);
  return h <= (uint32_t)w->tlen && v <= (uint32_t)w->plen;
}

static src_t get_src(const wfa_t* w, int c, int s) {
  src_t r = {w->nullrow + w->kbase, 1, -1, 0};
  if (s < 0) return r;
  const int sl = slot_of(w, c, s);
  if (!w->ex[c][sl]) return r;
  r.row = w->rows[c][sl] + w->kbase;
  r.lo = w->lo[c][sl];
  r.hi = w->hi[c][sl];
  r.ex = 1;
  return r;
}

/* Output row for (c,s): NULL everywhere (fresh in literal mode, old live range cleared in compact). */
static int32_t* out_row(wfa_t* w, int c, int s) {
  const int sl = slot_of(w, c, s);
  int32_t* r = w->rows[c][sl];
  if (!r) {
    r = (int32_t*)xalloc(w, sizeof(int32_t) * (size_t)w->Wp);
    if (!r) return NULL;
    for (int i = 0; i < w->Wp; ++i) r[i] = ONULL;
    w->rows[c][sl] = r;
  } else if (w->ex[c][sl]) {
    for (int k = w->lo[c][sl]; k <= w->hi[c][sl]; ++k) r[k + w->kbase] = ONULL;
  }
  w->ex[c][sl] = 0;
  w->lo[c][sl] = 1;
  w->hi[c][sl] = -1;
  return r + w->kbase;
}

static void mark_absent(wfa_t* w, int c, int s) {
  const int sl = slot_of(w, c, s);
  if (w->rows[c][sl] && w->ex[c][sl]) {
    int32_t* r = w->rows[c][sl];
    for (int k = w->lo[c][sl]; k <= w->hi[c][sl]; ++k) r[k + w->kbase] = ONULL;
  }
  w->ex[c][sl] = 0;
  w->lo[c][sl] = 1;
  w->hi[c][sl] = -1;
}

/* lo/hi trimming to the first/last in-matrix entry (all components). */
static void trim(wfa_t* w, int c, int s, int32_t* row, int lo, int hi) {
  while (lo <= hi && !in_matrix(w, row[lo], lo)) { row[lo] = ONULL; ++lo; }
  while (hi >= lo && !in_matrix(w, row[hi], hi)) { row[hi] = ONULL; --hi; }
  const int sl = slot_of(w, c, s);
  w->ex[c][sl] = (lo <= hi);
  w->lo[c][sl] = lo;
  w->hi[c][sl] = hi;
}

static inline void widen(int* lo, int* hi, const src_t* s, int dlo, int dhi) {
  if (!s->ex) return;
  if (s->lo + dlo < *lo) *lo = s->lo + dlo;
  if (s->hi + dhi > *hi) *hi = s->hi + dhi;
}

/* One score step. Returns 0 ok / <0 error. */
static int compute(wfa_t* w, int s) {
  const int aff = (w->metric == 1);
  src_t MX = get_src(w, C_M, s - w->x);
  src_t MO1 = get_src(w, C_M, aff ? s - w->o1 - w->e1 : s - w->e1);
  src_t MO2 = get_src(w, C_M, aff ? s - w->o2 - w->e2 : -1);
  src_t I1E = get_src(w, C_I1, aff ? s - w->e1 : -1), D1E = get_src(w, C_D1, aff ? s - w->e1 : -1);
  src_t I2E = get_src(w, C_I2, aff ? s - w->e2 : -1), D2E = get_src(w, C_D2, aff ? s - w->e2 : -1);
  if (!(MX.ex | MO1.ex | MO2.ex | I1E.ex | D1E.ex | I2E.ex | D2E.ex)) {
    for (int c = 0; c < w->ncomp; ++c) mark_absent(w, c, s);
    w->bt[s] = NULL;
    return 0;
  }
  int lo = INT_MAX, hi = INT_MIN;
  widen(&lo, &hi, &MX, 0, 0);
  widen(&lo, &hi, &MO1, -1, +1);
  widen(&lo, &hi, &MO2, -1, +1);
  widen(&lo, &hi, &I1E, +1, +1);
  widen(&lo, &hi, &D1E, -1, -1);
  widen(&lo, &hi, &I2E, +1, +1);
  widen(&lo, &hi, &D2E, -1, -1);
  if (lo > hi) { /* possible when only I/D sources exist with a 1-wide range shifted out */
    for (int c = 0; c < w->ncomp; ++c) mark_absent(w, c, s);
    w->bt[s] = NULL;
    return 0;
  }
  int32_t* out[NCOMP] = {0, 0, 0, 0, 0};
  for (int c = 0; c < w->ncomp; ++c) {
    out[c] = out_row(w, c, s);
    if (!out[c]) return -3;
  }
  const int width = hi - lo + 1;
  uint8_t* bt = bt_alloc(w, (size_t)width);
  if (!bt) return -3;
  w->bt[s] = bt;
  w->btlo[s] = lo;
  w->st.cells += width;
  w->st.bt_bytes += width;
  const int plen = w->plen, tlen = w->tlen;
  if (!aff) {
    const int32_t *mx = MX.row, *mg = MO1.row;
    int32_t* om = out[C_M];
    uint8_t* __restrict__ btp = bt - lo;
#pragma omp simd
    for (int k = lo; k <= hi; ++k) {
      const int32_t ins = mg[k - 1] + 1, del = mg[k + 1], mis = mx[k] + 1;
      int32_t best = ins, sel = SEL_I1;
      sel = del >= best ? SEL_D1 : sel; best = del >= best ? del : best;
      sel = mis >= best ? SEL_X : sel;  best = mis >= best ? mis : best;
      const uint32_t h = (uint32_t)best, v = (uint32_t)(best - k);
      best = (h > (uint32_t)tlen || v > (uint32_t)plen) ? ONULL : best;
      om[k] = best;
      btp[k] = (uint8_t)sel;
    }
    trim(w, C_M, s, om, lo, hi);
    return 0;
  }
  const int32_t *__restrict__ mx = MX.row, *__restrict__ mo1 = MO1.row, *__restrict__ mo2 = MO2.row;
  const int32_t *__restrict__ i1e = I1E.row, *__restrict__ d1e = D1E.row, *__restrict__ i2e = I2E.row, *__restrict__ d2e = D2E.row;
  int32_t *__restrict__ om = out[C_M], *__restrict__ oi1 = out[C_I1], *__restrict__ od1 = out[C_D1], *__restrict__ oi2 = out[C_I2], *__restrict__ od2 = out[C_D2];
  /* branch-free body so the compiler can vectorise it (WFA2-lib's own kernels are auto-vectorised the same way) */
  uint8_t* __restrict__ btp = bt - lo;
#pragma omp simd
  for (int k = lo; k <= hi; ++k) {
    const int32_t i1o_v = mo1[k - 1], i1e_v = i1e[k - 1];
    const int32_t i2o_v = mo2[k - 1], i2e_v = i2e[k - 1];
    const int32_t d1o_v = mo1[k + 1], d1e_v = d1e[k + 1];
    const int32_t d2o_v = mo2[k + 1], d2e_v = d2e[k + 1];
    const int32_t ins1 = MAXI(i1o_v, i1e_v) + 1, ins2 = MAXI(i2o_v, i2e_v) + 1;
    const int32_t del1 = MAXI(d1o_v, d1e_v), del2 = MAXI(d2o_v, d2e_v);
    const int32_t mis = mx[k] + 1;
    int32_t c = (i1e_v >= i1o_v ? BIT_I1E : 0) | (i2e_v >= i2o_v ? BIT_I2E : 0) | (d1e_v >= d1o_v ? BIT_D1E : 0) |
                (d2e_v >= d2o_v ? BIT_D2E : 0);
    int32_t best = ins1, sel = SEL_I1;
    sel = ins2 >= best ? SEL_I2 : sel; best = ins2 >= best ? ins2 : best;
    sel = del1 >= best ? SEL_D1 : sel; best = del1 >= best ? del1 : best;
    sel = del2 >= best ? SEL_D2 : sel; best = del2 >= best ? del2 : best;
    sel = mis >= best ? SEL_X : sel;   best = mis >= best ? mis : best;
    const uint32_t h = (uint32_t)best, v = (uint32_t)(best - k);
    best = (h > (uint32_t)tlen || v > (uint32_t)plen) ? ONULL : best;
    om[k] = best;
    oi1[k] = ins1 < 0 ? ONULL : ins1;
    oi2[k] = ins2 < 0 ? ONULL : ins2;
    od1[k] = del1 < 0 ? ONULL : del1;
    od2[k] = del2 < 0 ? ONULL : del2;
    btp[k] = (uint8_t)(c | sel);
  }
  trim(w, C_M, s, om, lo, hi);
  trim(w, C_I1, s, oi1, lo, hi);
  trim(w, C_D1, s, od1, lo, hi);
  trim(w, C_I2, s, oi2, lo, hi);
  trim(w, C_D2, s, od2, lo, hi);
  return 0;
}

/* WFA2's packed extend: 8-byte blocks, sentinel-padded sequences. */
static inline int32_t extend_one(wfa_t* w, int k, int32_t off) {
  const uint8_t* pp = w->p + (off - k);
  const uint8_t* tp = w->t + off;
  int32_t n = 0;
  for (;;) {
    uint64_t a, b;
    memcpy(&a, pp + n, 8);
    memcpy(&b, tp + n, 8);
    const uint64_t cmp = a ^ b;
    if (cmp == 0) { n += 8; continue; }
    n += (int32_t)(__builtin_ctzll(cmp) >> 3);
    break;
  }
  w->st.ext_words += (n + 16) / 16;
  return off + n;
}

static inline int terminated(const wfa_t* w, int k, int32_t off) {
  const int h = off, v = off - k;
  if (h >= w->tlen && (w->plen - v) <= w->pef) return 1;
  if (v >= w->plen && (w->tlen - h) <= w->tef) return 1;
  return 0;
}

/* ---- literal backtrace: WFA2's offset-comparing walk over all stored wavefronts ---- */
static inline int64_t lit_get(const wfa_t* w, int c, int s, int k, int add, int type) {
  if (s < 0) return ONULL;
  if (!w->ex[c][s]) return ONULL;
  if (k < w->lo[c][s] || k > w->hi[c][s]) return ONULL;
  return BT_SET(w->rows[c][s][k + w->kbase] + add, type);
}

/* ops are written back-to-front into buf[cap]; returns begin index or -1. */
static int backtrace_literal(const wfa_t* w, int s_end, int k_end, int32_t off_end, uint8_t* buf, int cap) {
  int pos = cap;
#define PUSH(ch) do { if (pos <= 0) return -1; buf[--pos] = (ch); } while (0)
  const int aff = (w->metric == 1);
  int score = s_end, k = k_end;
  int32_t offset = off_end;
  int h = offset, v = offset - k;
  for (int i = w->plen - v; i > 0; --i) PUSH('D');
  for (int i = w->tlen - h; i > 0; --i) PUSH('I');
  int mt = C_M;
  while (v > 0 && h > 0 && score > 0) {
    const int mism = score - w->x;
    const int go1 = aff ? score - w->o1 - w->e1 : score - w->e1;
    const int ge1 = score - w->e1;
    const int go2 = score - w->o2 - w->e2, ge2 = score - w->e2;
    int64_t max_all;
    if (!aff) {
      const int64_t m = lit_get(w, C_M, mism, k, 1, BT_M);
      const int64_t in = lit_get(w, C_M, go1, k - 1, 1, BT_I1_OPEN);
      const int64_t de = lit_get(w, C_M, go1, k + 1, 0, BT_D1_OPEN);
      max_all = MAXI(m, MAXI(in, de));
    } else {
      switch (mt) {
        case C_M: {
          const int64_t m = lit_get(w, C_M, mism, k, 1, BT_M);
          const int64_t i1o = lit_get(w, C_M, go1, k - 1, 1, BT_I1_OPEN);
          const int64_t i1e = lit_get(w, C_I1, ge1, k - 1, 1, BT_I1_EXT);
          const int64_t i2o = lit_get(w, C_M, go2, k - 1, 1, BT_I2_OPEN);
          const int64_t i2e = lit_get(w, C_I2, ge2, k - 1, 1, BT_I2_EXT);
          const int64_t d1o = lit_get(w, C_M, go1, k + 1, 0, BT_D1_OPEN);
          const int64_t d1e = lit_get(w, C_D1, ge1, k + 1, 0, BT_D1_EXT);
          const int64_t d2o = lit_get(w, C_M, go2, k + 1, 0, BT_D2_OPEN);
          const int64_t d2e = lit_get(w, C_D2, ge2, k + 1, 0, BT_D2_EXT);
          const int64_t mi = MAXI(MAXI(i1o, i1e), MAXI(i2o, i2e));
          const int64_t md = MAXI(MAXI(d1o, d1e), MAXI(d2o, d2e));
          max_all = MAXI(m, MAXI(mi, md));
          break;
        }
        case C_I1: {
          const int64_t o = lit_get(w, C_M, go1, k - 1, 1, BT_I1_OPEN), e = lit_get(w, C_I1, ge1, k - 1, 1, BT_I1_EXT);
          max_all = MAXI(o, e);
          break;
        }
        case C_I2: {
          const int64_t o = lit_get(w, C_M, go2, k - 1, 1, BT_I2_OPEN), e = lit_get(w, C_I2, ge2, k - 1, 1, BT_I2_EXT);
          max_all = MAXI(o, e);
          break;
        }
        case C_D1: {
          const int64_t o = lit_get(w, C_M, go1, k + 1, 0, BT_D1_OPEN), e = lit_get(w, C_D1, ge1, k + 1, 0, BT_D1_EXT);
          max_all = MAXI(o, e);
          break;
        }
        default: {
          const int64_t o = lit_get(w, C_M, go2, k + 1, 0, BT_D2_OPEN), e = lit_get(w, C_D2, ge2, k + 1, 0, BT_D2_EXT);
          max_all = MAXI(o, e);
          break;
        }
      }
    }
    if (max_all < 0) break;
    if (mt == C_M) {
      const int max_off = BT_OFF(max_all);
      for (int i = offset - max_off; i > 0; --i) PUSH('M');
      offset = max_off;
      v = offset - k;
      h = offset;
      if (v <= 0 || h <= 0) break;
    }
    switch (BT_TYPE(max_all)) {
      case BT_M: score = mism; mt = C_M; PUSH('X'); --offset; break;
      case BT_I1_OPEN: score = go1; mt = C_M; PUSH('I'); --k; --offset; break;
      case BT_I1_EXT: score = ge1; mt = C_I1; PUSH('I'); --k; --offset; break;
      case BT_I2_OPEN: score = go2; mt = C_M; PUSH('I'); --k; --offset; break;
      case BT_I2_EXT: score = ge2; mt = C_I2; PUSH('I'); --k; --offset; break;
      case BT_D1_OPEN: score = go1; mt = C_M; PUSH('D'); ++k; break;
      case BT_D1_EXT: score = ge1; mt = C_D1; PUSH('D'); ++k; break;
      case BT_D2_OPEN: score = go2; mt = C_M; PUSH('D'); ++k; break;
      case BT_D2_EXT: score = ge2; mt = C_D2; PUSH('D'); ++k; break;
      default: return -1;
    }
    v = offset - k;
    h = offset;
  }
  if (v > 0 && h > 0) {
    const int n = MINI(v, h);
    for (int i = 0; i < n; ++i) PUSH('M');
    v -= n;
    h -= n;
  }
  while (v > 0) { PUSH('D'); --v; }
  while (h > 0) { PUSH('I'); --h; }
#undef PUSH
  return pos;
}

/* ---- compact backtrace: follow choice bytes back to score 0, then replay forward ---- */
static int backtrace_compact(wfa_t* w, int s_end, int k_end, int32_t off_end, uint8_t* buf, int cap, int* k0_out) {
  const int aff = (w->metric == 1);
  /* backward walk: edits (no matches) pushed in reverse order at the tail of buf.
   * NOTE the 'ext' bit of a gap step says how that step was REACHED: a step reached by extension
   * continues the gap opened by the step before it (forward order), so it is tagged lower-case. */
  int pos = cap, s = s_end, k = k_end, comp = C_M;
  while (s > 0) {
    if (!w->bt[s]) return -1;
    const uint8_t c = w->bt[s][k - w->btlo[s]];
    if (pos <= 0) return -1;
    if (!aff) {
      switch (c & 7) {
        case SEL_X: buf[--pos] = 'X'; s -= w->x; break;
        case SEL_I1: buf[--pos] = 'I'; s -= w->e1; k -= 1; break;
        default: buf[--pos] = 'D'; s -= w->e1; k += 1; break;
      }
      continue;
    }
    switch (comp) {
      case C_M:
        switch (c & 7) {
          case SEL_X: buf[--pos] = 'X'; s -= w->x; break;
          case SEL_I1: comp = C_I1; break;
          case SEL_I2: comp = C_I2; break;
          case SEL_D1: comp = C_D1; break;
          default: comp = C_D2; break;
        }
        break;
      case C_I1:
        if (c & BIT_I1E) { buf[--pos] = 'i'; s -= w->e1; } else { buf[--pos] = 'I'; s -= w->o1 + w->e1; comp = C_M; }
        k -= 1;
        break;
      case C_I2:
        if (c & BIT_I2E) { buf[--pos] = 'i'; s -= w->e2; } else { buf[--pos] = 'I'; s -= w->o2 + w->e2; comp = C_M; }
        k -= 1;
        break;
      case C_D1:
        if (c & BIT_D1E) { buf[--pos] = 'd'; s -= w->e1; } else { buf[--pos] = 'D'; s -= w->o1 + w->e1; comp = C_M; }
        k += 1;
        break;
      default:
        if (c & BIT_D2E) { buf[--pos] = 'd'; s -= w->e2; } else { buf[--pos] = 'D'; s -= w->o2 + w->e2; comp = C_M; }
        k += 1;
        break;
    }
  }
  if (s != 0 || comp != C_M) return -1;
  *k0_out = k;
  /* forward replay with greedy match extension; edits live in buf[pos..cap) in forward order */
  const int n_edits = cap - pos;
  int out = 0, h = k > 0 ? k : 0, v = k < 0 ? -k : 0;
  /* the output region [0,out) must never overrun the unread edits [pos+i,cap) */
#define EMIT(ch, rd) do { if (out >= (rd)) return -1; buf[out++] = (ch); } while (0)
  for (int i = 0; i < h; ++i) EMIT('I', pos);
  for (int i = 0; i < v; ++i) EMIT('D', pos);
  for (int i = 0; i <= n_edits; ++i) {
    const int rd = pos + i;
    /* lower-case = gap extension: the path is still inside the I/D component, no match extension there */
    const int in_gap = (i < n_edits) && (buf[rd] == 'i' || buf[rd] == 'd');
    if (!in_gap)
      while (h < w->tlen && v < w->plen && w->t[h] == w->p[v]) { EMIT('M', rd); ++h; ++v; }
    if (i == n_edits) break;
    const uint8_t op = (uint8_t)(buf[rd] & ~0x20);
    EMIT(op, rd + 1);
    if (op == 'X') { ++h; ++v; } else if (op == 'I') ++h; else ++v;
  }
  if (h != off_end || v != off_end - k_end) return -1;
  for (int i = w->tlen - h; i > 0; --i) EMIT('I', cap);
  for (int i = w->plen - v; i > 0; --i) EMIT('D', cap);
#undef EMIT
  return out;
}

static int gap_cost(const wfa_t* w, int L) {
  if (L <= 0) return 0;
  if (w->metric == 0) return w->e1 * L;
  const long a = (long)w->o1 + (long)w->e1 * L, b = (long)w->o2 + (long)w->e2 * L;
  const long m = a < b ? a : b;
  return m > INT_MAX / 4 ? INT_MAX / 4 : (int)m;
}

static void wfa_free(wfa_t* w) {
  for (int c = 0; c < NCOMP; ++c) {
    if (w->rows[c]) {
      for (int i = 0; i < w->depth[c]; ++i) free(w->rows[c][i]);
      free(w->rows[c]);
    }
    free(w->lo[c]);
    free(w->hi[c]);
    free(w->ex[c]);
  }
  for (int i = 0; i < w->n_bt_chunks; ++i) free(w->bt_chunks[i]);
  free(w->bt_chunks);
  free(w->bt);
  free(w->btlo);
  free(w->nullrow);
  free(w->p);
  free(w->t);
}

int ora_wfa_align(const ora_penalties* pen, const uint8_t* pattern, int plen, const uint8_t* text, int tlen,
                  int pbf, int pef, int tbf, int tef, int mode, int want_ops, int32_t* score_out, uint8_t* ops,
                  int ops_cap, ora_stats* st_out) {
  if (!pen || !pattern || !text || plen <= 0 || tlen <= 0) return -1;
  if (pen->metric != 0 && pen->metric != 1) return -1;
  if (pen->match > 0) return -1;
  if (pbf < 0 || pef < 0 || tbf < 0 || tef < 0 || pbf > plen || pef > plen || tbf > tlen || tef > tlen) return -1;
  wfa_t W;
  memset(&W, 0, sizeof(W));
  wfa_t* w = &W;
  w->plen = plen; w->tlen = tlen;
  w->metric = pen->metric;
  w->ncomp = pen->metric == 1 ? NCOMP : 1;
  w->match = pen->match;
  const int m = pen->match;
  if (m == 0) {
    w->x = pen->mismatch; w->o1 = pen->gap_open1; w->e1 = pen->gap_ext1; w->o2 = pen->gap_open2; w->e2 = pen->gap_ext2;
  } else {
    w->x = 2 * pen->mismatch - 2 * m;
    w->o1 = 2 * pen->gap_open1; w->e1 = 2 * pen->gap_ext1 - m;
    w->o2 = 2 * pen->gap_open2; w->e2 = 2 * pen->gap_ext2 - m;
  }
  if (w->metric == 0) { w->o1 = 0; w->o2 = 0; w->e2 = 0; }
  if (w->x <= 0 || w->e1 <= 0 || (w->metric == 1 && (w->e2 <= 0 || w->o1 < 0 || w->o2 < 0))) return -1;
  w->pbf = pbf; w->pef = pef; w->tbf = tbf; w->tef = tef;
  w->mode = mode;
  w->mem_cap = (size_t)6 << 30;
  w->Wp = plen + tlen + 5;
  w->kbase = plen + 2;
  {
    long cap = (long)gap_cost(w, plen) + (long)gap_cost(w, tlen) + 1;
    if (cap > INT_MAX / 4) cap = INT_MAX / 4;
    w->S_cap = (int)cap;
  }
  int rc = 0, s = 0, k_end = 0, found = 0;
  int32_t off_end = 0;
  uint8_t* work = NULL;
  /* padded sequence copies: distinct sentinels so extends stop at either end */
  w->p = (uint8_t*)malloc((size_t)plen + 24);
  w->t = (uint8_t*)malloc((size_t)tlen + 24);
  if (!w->p || !w->t) { rc = -3; goto done; }
  memcpy(w->p, pattern, (size_t)plen); memset(w->p + plen, '!', 24);
  memcpy(w->t, text, (size_t)tlen); memset(w->t + tlen, '?', 24);
  {
    const int lookM = w->metric == 1 ? MAXI(w->x, MAXI(w->o1 + w->e1, w->o2 + w->e2)) : MAXI(w->x, w->e1);
    const int full = w->S_cap + 1;
    w->depth[C_M] = mode == ORA_MODE_LITERAL ? full : lookM + 1;
    w->depth[C_I1] = w->depth[C_D1] = mode == ORA_MODE_LITERAL ? full : w->e1 + 1;
    w->depth[C_I2] = w->depth[C_D2] = mode == ORA_MODE_LITERAL ? full : w->e2 + 1;
  }
  for (int c = 0; c < w->ncomp; ++c) {
    const size_t d = (size_t)w->depth[c];
    w->rows[c] = (int32_t**)calloc(d, sizeof(int32_t*));
    w->lo[c] = (int*)calloc(d, sizeof(int));
    w->hi[c] = (int*)calloc(d, sizeof(int));
    w->ex[c] = (char*)calloc(d, 1);
    if (!w->rows[c] || !w->lo[c] || !w->hi[c] || !w->ex[c]) { rc = -3; goto done; }
  }
  w->bt = (uint8_t**)calloc((size_t)w->S_cap + 1, sizeof(uint8_t*));
  w->btlo = (int*)calloc((size_t)w->S_cap + 1, sizeof(int));
  w->nullrow = (int32_t*)malloc(sizeof(int32_t) * (size_t)w->Wp);
  if (!w->bt || !w->btlo || !w->nullrow) { rc = -3; goto done; }
  for (int i = 0; i < w->Wp; ++i) w->nullrow[i] = ONULL;
  /* wavefront 0 (ends-free init; end-to-end is the all-zero case) */
  {
    int32_t* r = out_row(w, C_M, 0);
    if (!r) { rc = -3; goto done; }
    for (int k = -pbf; k <= tbf; ++k) r[k] = k > 0 ? k : 0;
    const int sl = slot_of(w, C_M, 0);
    w->ex[C_M][sl] = 1; w->lo[C_M][sl] = -pbf; w->hi[C_M][sl] = tbf;
    w->st.cells += pbf + tbf + 1;
  }
  for (;;) {
    const int sl = slot_of(w, C_M, s);
    if (w->ex[C_M][sl]) {
      int32_t* r = w->rows[C_M][sl] + w->kbase;
      const int lo = w->lo[C_M][sl], hi = w->hi[C_M][sl];
      for (int k = lo; k <= hi; ++k) {
        int32_t off = r[k];
        if (off < 0) continue;
        off = extend_one(w, k, off);
        r[k] = off;
        if (terminated(w, k, off)) { found = 1; k_end = k; off_end = off; break; }
      }
      if (found) break;
    }
    ++s;
    if (s > w->S_cap) { rc = -5; goto done; }
    rc = compute(w, s);
    if (rc) goto done;
  }
  w->st.wf_score = s; w->st.end_k = k_end; w->st.end_off = off_end;
  {
    /* k0 and the op string */
    const int cap = plen + tlen + 8;
    work = (uint8_t*)malloc((size_t)cap * 2);
    if (!work) { rc = -3; goto done; }
    int k0 = 0;
    const int n = backtrace_compact(w, s, k_end, off_end, work, cap, &k0);
    if (n < 0) { rc = -4; goto done; }
    w->st.k0 = k0; w->st.n_ops = n;
    if (mode == ORA_MODE_LITERAL) {
      uint8_t* lit = work + cap;
      const int b = backtrace_literal(w, s, k_end, off_end, lit, cap);
      if (b < 0 || cap - b != n || memcmp(lit + b, work, (size_t)n) != 0) {
        if (getenv("ORA_DEBUG")) {
          fprintf(stderr, "compact: %.*s\nliteral: %.*s (b=%d)\n", n, work, b < 0 ? 0 : cap - b, b < 0 ? (uint8_t*)"" : lit + b, b);
        }
        rc = -4;
        goto done;
      }
    }
    const int h_end = off_end, v_end = off_end - k_end;
    const int al_t = h_end - (k0 > 0 ? k0 : 0), al_p = v_end - (k0 < 0 ? -k0 : 0);
    const int sc = (m == 0) ? -s : ((-m) * (al_p + al_t) - s) / 2;
    if (score_out) *score_out = sc;
    if (want_ops && ops) {
      if (n > ops_cap) { rc = -2; goto done; }
      memcpy(ops, work, (size_t)n);
    }
  }
done:
  free(work);
  if (w->oom && rc == 0) rc = -3;
  if (st_out) *st_out = w->st;
  wfa_free(w);
  return rc;
}

/* == process_wfa2_cigar_string, reference src/wfa2_utils.rs:13-95 */
int ora_translate_cigar(const uint8_t* wfa_ops, size_t n_ops, int64_t* ref_offset, uint32_t* cigar, size_t cigar_cap,
                        size_t* n_cigar, int* has_match) {
  uint8_t last = 0;
  uint32_t count = 0;
  int first_match = 0;
  int64_t roff = 0;
  size_t n = 0;
  for (size_t i = 0; i <= n_ops; ++i) {
    const uint8_t state = (i < n_ops) ? wfa_ops[n_ops - 1 - i] : 0;
    if (state != last) {
      const int final_state = (state == 0);
      if (count != 0) {
        int op = -1;
        switch (last) {
          case 'M': op = 7; break;
          case 'X': op = 8; break;
          case 'D': op = (!first_match || final_state) ? 4 : 1; break;
          case 'I':
            if (!first_match) roff += count;
            else if (!final_state) op = 2;
            break;
          default: return -1;
        }
        if (op >= 0) {
          if (n >= cigar_cap) return -2;
          cigar[n++] = (count << 4) | (uint32_t)op;
        }
      }
      count = 0;
    }
    if (state == 0) break;
    if (!first_match && (state == 'M' || state == 'X')) first_match = 1;
    ++count;
    last = state;
  }
  if (ref_offset) *ref_offset = roff;
  if (n_cigar) *n_cigar = n;
  if (has_match) *has_match = first_match;
  return 0;
}

int ora_sawfish_align(const ora_penalties* pen, const uint8_t* text, int tlen, const uint8_t* pattern, int plen,
                      double text_clip_frac, double pattern_clip_frac, int end_to_end, int mode, int32_t* score,
                      int64_t* ref_offset, uint32_t* cigar, size_t cigar_cap, size_t* n_cigar, int* has_match,
                      ora_stats* st) {
  if (plen <= 0 || tlen <= 0) return -1;
  uint8_t* rp = (uint8_t*)malloc((size_t)plen);
  uint8_t* rt = (uint8_t*)malloc((size_t)tlen);
  uint8_t* ops = (uint8_t*)malloc((size_t)plen + (size_t)tlen + 8);
  if (!rp || !rt || !ops) { free(rp); free(rt); free(ops); return -3; }
  for (int i = 0; i < plen; ++i) rp[i] = pattern[plen - 1 - i];
  for (int i = 0; i < tlen; ++i) rt[i] = text[tlen - 1 - i];
  int pc, tc;
  if (end_to_end) {
    pc = tc = 0;
  } else if (text_clip_frac < 0 || pattern_clip_frac < 0) {
    const int a = (int)((double)plen * 0.4), b = (int)((double)tlen * 0.4);
    pc = tc = a < b ? a : b;
  } else {
    pc = (int)((double)plen * pattern_clip_frac);
    tc = (int)((double)tlen * text_clip_frac);
  }
  ora_stats lst;
  int rc = ora_wfa_align(pen, rp, plen, rt, tlen, pc, pc, tc, tc, mode, 1, score, ops, plen + tlen + 8, &lst);
  if (rc == 0) rc = ora_translate_cigar(ops, (size_t)lst.n_ops, ref_offset, cigar, cigar_cap, n_cigar, has_match);
  if (st) *st = lst;
  free(rp); free(rt); free(ops);
  return rc;
}

int ora_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

int ora_align_batch(const ora_penalties* pen, const uint8_t* arena, const ora_pair* pairs, size_t n_pairs, int want_ops,
                    int mode, int n_threads, int32_t* scores, int32_t* status, ora_stats* stats, uint8_t* ops,
                    size_t ops_cap, uint64_t* ops_off) {
  /* ops layout: pair i owns [ops_off[i], ops_off[i+1]) with capacity plen+tlen each (prefix sums, fixed up front) */
  if (ops_off) {
    uint64_t acc = 0;
    for (size_t i = 0; i < n_pairs; ++i) { ops_off[i] = acc; acc += (uint64_t)pairs[i].pattern_len + pairs[i].text_len; }
    ops_off[n_pairs] = acc;
    if (want_ops && ops && acc > ops_cap) return -2;
  }
  if (n_threads <= 0) n_threads = ora_max_threads();
  long i;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(n_threads)
#endif
  for (i = 0; i < (long)n_pairs; ++i) {
    const ora_pair* pr = &pairs[i];
    const int plen = (int)pr->pattern_len, tlen = (int)pr->text_len;
    uint8_t* bp = (uint8_t*)malloc((size_t)plen + 1);
    uint8_t* btx = (uint8_t*)malloc((size_t)tlen + 1);
    int rc = -3;
    if (bp && btx) {
      const uint8_t* ps = arena + pr->pattern_off;
      const uint8_t* ts = arena + pr->text_off;
      if (pr->flags & 1u) {
        for (int j = 0; j < plen; ++j) bp[j] = ps[plen - 1 - j];
        for (int j = 0; j < tlen; ++j) btx[j] = ts[tlen - 1 - j];
      } else {
        memcpy(bp, ps, (size_t)plen);
        memcpy(btx, ts, (size_t)tlen);
      }
      ora_stats lst;
      uint8_t* o = (want_ops && ops && ops_off) ? ops + ops_off[i] : NULL;
      /* the reference always asks WFA2 for the full alignment, even when only the score is used
       * (src/score_sv/mod.rs:702; AlignmentScope::Alignment at src/wfa2_utils.rs:186) */
      rc = ora_wfa_align(pen, bp, plen, btx, tlen, pr->pattern_begin_free, pr->pattern_end_free, pr->text_begin_free,
                         pr->text_end_free, mode, o != NULL, &scores[i], o, plen + tlen, &lst);
      if (stats) stats[i] = lst;
    }
    free(bp);
    free(btx);
    status[i] = rc;
  }
  return 0;
}
