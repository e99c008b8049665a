/* bio_oracle.c — CPU restatement of rust-bio's pairwise "custom" aligner (affine gaps + prefix/suffix clipping),
 * the algorithm behind sawfish's bio_align_utils::PairwiseAligner (reference src/bio_align_utils.rs:97-188).
 *
 * TEST INFRASTRUCTURE ONLY (see wfa_oracle.h): imported by tests/ and bench.py's CPU legs, never by the product.
 *
 * Provenance: the arithmetic lives in the third-party crate `bio` ("bio = 2" in the reference's Cargo.toml, 2.3.x in
 * Cargo.lock), whose source is NOT under /root/reference.  This file restates bio::alignment::pairwise::Aligner::custom
 * from its published algorithm: three layers S/I/D filled column by column over the text y, every comparison strict
 * ('>') in the order diagonal, insertion, deletion, x-prefix clip, y-prefix clip; suffix clips tracked as running
 * maxima (Lx per column, Ly per row, first maximum wins); a traceback cell keeps one 4-bit code per layer.
 * sawfish calls the BANDED variant (banded::Aligner::custom_with_matches, k = 20, w = 10): cells outside a band built
 * from chained k-mer matches are skipped.  The band is not restated here — this oracle fills the whole matrix, which
 * gives the same alignment whenever the band contains the optimal path (rust-bio falls back to the full matrix when
 * there are no k-mer matches, which is the case for all of the reference's known answers).
 * PARITY: pinned by the reference's nine known answers (src/bio_align_utils.rs:291-403); everything else, and the
 * band, is "parity unpinned".
 *
 * x = pattern (rows i), y = text (columns j).  Ins consumes x, Del consumes y (rust-bio's naming).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define BIO_MIN_SCORE (-858993459) /* rust-bio's MIN_SCORE: "minus infinity" that survives a few additions */

typedef struct {
  int32_t gap_open, gap_extend, match, mismatch;
  int32_t xclip_prefix, xclip_suffix, yclip_prefix, yclip_suffix;
} bio_scoring;

enum { TB_START = 0, TB_INS = 1, TB_DEL = 2, TB_SUBST = 3, TB_MATCH = 4, TB_XCLIP_PREFIX = 5, TB_XCLIP_SUFFIX = 6,
       TB_YCLIP_PREFIX = 7, TB_YCLIP_SUFFIX = 8 };
enum { OP_MATCH = 0, OP_SUBST = 1, OP_INS = 2, OP_DEL = 3, OP_XCLIP = 4, OP_YCLIP = 5 };

/* traceback cell: s bits [3:0], i bits [7:4], d bits [11:8] */
#define TB_S(c) ((c) & 15)
#define TB_I(c) (((c) >> 4) & 15)
#define TB_D(c) (((c) >> 8) & 15)
static inline uint16_t set_s(uint16_t c, int v) { return (uint16_t)((c & ~15) | v); }
static inline uint16_t set_i(uint16_t c, int v) { return (uint16_t)((c & ~(15 << 4)) | (v << 4)); }
static inline uint16_t set_d(uint16_t c, int v) { return (uint16_t)((c & ~(15 << 8)) | (v << 8)); }
static inline uint16_t set_all(int v) { return (uint16_t)(v | (v << 4) | (v << 8)); }

typedef struct {
  int32_t score, xstart, xend, ystart, yend;
} bio_result;

/* ops_out: one word per operation, (len << 3) | code with len = 1 for Match/Subst/Ins/Del and the clipped length for
 * Xclip/Yclip, in alignment order.  Returns 0, or -1 when ops_cap is too small / allocation fails. */
int bio_custom_align(const bio_scoring* sc, const uint8_t* x, int m, const uint8_t* y, int n, bio_result* res,
                     uint32_t* ops_out, int ops_cap, int* n_ops) {
  const int go = sc->gap_open, ge = sc->gap_extend;
  const size_t W = (size_t)m + 1;
  uint16_t* tb = (uint16_t*)calloc(W * ((size_t)n + 1), sizeof(uint16_t)); /* tb[j * W + i] */
  int32_t* S[2], *I[2], *D[2];
  int32_t* buf = (int32_t*)malloc(sizeof(int32_t) * (6 * W + W + (size_t)n + 1 + W));
  if (!tb || !buf) { free(tb); free(buf); return -1; }
  for (int k = 0; k < 2; ++k) { S[k] = buf + (size_t)k * 3 * W; I[k] = S[k] + W; D[k] = I[k] + W; }
  int32_t* Sn = buf + 6 * W;            /* best score of "clip the rest of y after column j" per row */
  int32_t* Lx = Sn + W;                 /* [n+1] x-suffix clip length chosen in column j */
  int32_t* Ly = Lx + (size_t)n + 1;     /* [m+1] y-suffix clip length chosen in row i */
#define T(i, j) tb[(size_t)(j) * W + (size_t)(i)]

  /* column 0 (both buffers are initialised the same way; only k == 0 records traceback and trackers) */
  for (int k = 0; k < 2; ++k) {
    for (int i = 0; i <= m; ++i) { S[k][i] = I[k][i] = D[k][i] = BIO_MIN_SCORE; }
    S[k][0] = 0;
    if (k == 0) {
      T(0, 0) = set_all(TB_START);
      for (int j = 0; j <= n; ++j) Lx[j] = 0;
      for (int i = 0; i <= m; ++i) { Ly[i] = 0; Sn[i] = BIO_MIN_SCORE; }
      Sn[0] = sc->yclip_suffix;
      Ly[0] = n;
    }
    for (int i = 1; i <= m; ++i) {
      uint16_t c = set_all(TB_START);
      if (i == 1) {
        I[k][i] = go + ge;
        c = set_i(c, TB_START);
      } else {
        const int i_score = go + ge * i;                    /* insert all i characters */
        const int c_score = sc->xclip_prefix + go + ge;     /* clip, then insert one */
        if (i_score > c_score) { I[k][i] = i_score; c = set_i(c, TB_INS); }
        else { I[k][i] = c_score; c = set_i(c, TB_XCLIP_PREFIX); }
      }
      if (i == m) c = set_s(c, TB_XCLIP_SUFFIX);
      else S[k][i] = BIO_MIN_SCORE;
      if (I[k][i] > S[k][i]) { S[k][i] = I[k][i]; c = set_s(c, TB_INS); }
      if (sc->xclip_prefix > S[k][i]) { S[k][i] = sc->xclip_prefix; c = set_s(c, TB_XCLIP_PREFIX); }
      if (i != m && S[k][i] + sc->xclip_suffix > S[k][m]) { S[k][m] = S[k][i] + sc->xclip_suffix; Lx[0] = m - i; }
      if (k == 0) T(i, 0) = c;
      if (S[k][i] + sc->yclip_suffix > Sn[i]) { Sn[i] = S[k][i] + sc->yclip_suffix; Ly[i] = n; }
    }
  }

  for (int j = 1; j <= n; ++j) {
    const int cur = j & 1, prev = 1 - cur;
    {
      uint16_t c = 0;
      I[cur][0] = BIO_MIN_SCORE;
      if (j == 1) {
        D[cur][0] = go + ge;
        c = set_d(c, TB_START);
      } else {
        const int d_score = go + ge * j;
        const int c_score = sc->yclip_prefix + go + ge;
        if (d_score > c_score) { D[cur][0] = d_score; c = set_d(c, TB_DEL); }
        else { D[cur][0] = c_score; c = set_d(c, TB_YCLIP_PREFIX); }
      }
      if (D[cur][0] > sc->yclip_prefix) { S[cur][0] = D[cur][0]; c = set_s(c, TB_DEL); }
      else { S[cur][0] = sc->yclip_prefix; c = set_s(c, TB_YCLIP_PREFIX); }
      if (j == n && Sn[0] > S[cur][0]) { S[cur][0] = Sn[0]; c = set_s(c, TB_YCLIP_SUFFIX); }
      else if (S[cur][0] + sc->yclip_suffix > Sn[0]) { Sn[0] = S[cur][0] + sc->yclip_suffix; Ly[0] = n - j; }
      T(0, j) = c;
    }
    for (int i = 1; i <= m; ++i) S[cur][i] = BIO_MIN_SCORE;
    const uint8_t q = y[j - 1];
    const int gap_j = go + ge * j;
    const int xclip_score = sc->xclip_prefix + (sc->yclip_prefix > gap_j ? sc->yclip_prefix : gap_j);
    for (int i = 1; i <= m; ++i) {
      const uint8_t p = x[i - 1];
      uint16_t c = 0;
      const int m_score = S[prev][i - 1] + (p == q ? sc->match : sc->mismatch);
      int best_i, best_d;
      {
        const int i_score = I[cur][i - 1] + ge, s_score = S[cur][i - 1] + go + ge;
        if (i_score > s_score) { best_i = i_score; c = set_i(c, TB_INS); }
        else { best_i = s_score; c = set_i(c, TB_S(T(i - 1, j))); }
      }
      {
        const int d_score = D[prev][i] + ge, s_score = S[prev][i] + go + ge;
        if (d_score > s_score) { best_d = d_score; c = set_d(c, TB_DEL); }
        else { best_d = s_score; c = set_d(c, TB_S(T(i, j - 1))); }
      }
      c = set_s(c, TB_XCLIP_SUFFIX);
      int best = S[cur][i];
      if (m_score > best) { best = m_score; c = set_s(c, p == q ? TB_MATCH : TB_SUBST); }
      if (best_i > best) { best = best_i; c = set_s(c, TB_INS); }
      if (best_d > best) { best = best_d; c = set_s(c, TB_DEL); }
      if (xclip_score > best) { best = xclip_score; c = set_s(c, TB_XCLIP_PREFIX); }
      const int yclip_score = sc->yclip_prefix + go + ge * i;
      if (yclip_score > best) { best = yclip_score; c = set_s(c, TB_YCLIP_PREFIX); }
      S[cur][i] = best; I[cur][i] = best_i; D[cur][i] = best_d;
      if (S[cur][i] + sc->xclip_suffix > S[cur][m]) { S[cur][m] = S[cur][i] + sc->xclip_suffix; Lx[j] = m - i; }
      if (S[cur][i] + sc->yclip_suffix > Sn[i]) { Sn[i] = S[cur][i] + sc->yclip_suffix; Ly[i] = n - j; }
      T(i, j) = c;
    }
  }

  { /* suffix clipping in the last column, then the insertion layer of that column once more */
    const int j = n, cur = j & 1;
    for (int i = 0; i <= m; ++i) {
      if (Sn[i] > S[cur][i]) { S[cur][i] = Sn[i]; T(i, j) = set_s(T(i, j), TB_YCLIP_SUFFIX); }
      if (S[cur][i] + sc->xclip_suffix > S[cur][m]) {
        S[cur][m] = S[cur][i] + sc->xclip_suffix; Lx[j] = m - i; T(m, j) = set_s(T(m, j), TB_XCLIP_SUFFIX);
      }
    }
    for (int i = 1; i <= m; ++i) {
      const int s_score = S[cur][i - 1] + go + ge;
      if (s_score > I[cur][i]) { I[cur][i] = s_score; T(i, j) = set_i(T(i, j), TB_S(T(i - 1, j))); }
      if (s_score > S[cur][i]) {
        S[cur][i] = s_score; T(i, j) = set_s(T(i, j), TB_INS);
        if (S[cur][i] + sc->xclip_suffix > S[cur][m]) {
          S[cur][m] = S[cur][i] + sc->xclip_suffix; Lx[j] = m - i; T(m, j) = set_s(T(m, j), TB_XCLIP_SUFFIX);
        }
      }
    }
  }

  /* traceback */
  int i = m, j = n, cnt = 0, rc = 0;
  res->score = S[n & 1][m];
  res->xstart = 0; res->ystart = 0; res->xend = m; res->yend = n;
  int layer = TB_S(T(i, j));
  uint32_t* rev = ops_out; /* filled backwards, reversed below */
  while (layer != TB_START) {
    int next, code = 0, len = 1;
    switch (layer) {
      case TB_INS: code = OP_INS; next = TB_I(T(i, j)); i -= 1; break;
      case TB_DEL: code = OP_DEL; next = TB_D(T(i, j)); j -= 1; break;
      case TB_MATCH: code = OP_MATCH; next = TB_S(T(i - 1, j - 1)); i -= 1; j -= 1; break;
      case TB_SUBST: code = OP_SUBST; next = TB_S(T(i - 1, j - 1)); i -= 1; j -= 1; break;
      case TB_XCLIP_PREFIX: code = OP_XCLIP; len = i; res->xstart = i; i = 0; next = TB_S(T(0, j)); break;
      case TB_XCLIP_SUFFIX: code = OP_XCLIP; len = Lx[j]; i -= Lx[j]; res->xend = i; next = TB_S(T(i, j)); break;
      case TB_YCLIP_PREFIX: code = OP_YCLIP; len = j; res->ystart = j; j = 0; next = TB_S(T(i, 0)); break;
      case TB_YCLIP_SUFFIX: code = OP_YCLIP; len = Ly[i]; j -= Ly[i]; res->yend = j; next = TB_S(T(i, j)); break;
      default: rc = -2; next = TB_START; break;
    }
    if (cnt >= ops_cap) { rc = -1; break; }
    rev[cnt++] = ((uint32_t)len << 3) | (uint32_t)code;
    layer = next;
    if (cnt > 2 * (m + n) + 8) { rc = -2; break; }
  }
  for (int a = 0, b = cnt - 1; a < b; ++a, --b) { const uint32_t t = rev[a]; rev[a] = rev[b]; rev[b] = t; }
  *n_ops = cnt;
#undef T
  free(tb); free(buf);
  return rc;
}
