/*
 * dp_verify.c — independent full-matrix check used by tests (TEST INFRASTRUCTURE ONLY).
 *
 * Computes, with a plain O(|p|·|t|) Gotoh recurrence (nothing shared with the
 * wavefront code), the minimum ADJUSTED penalty over all ends-free alignments that
 * WFA2's termination rule admits.  The wavefront score found by the oracle / the CUDA
 * kernels must equal it (score optimality under the WFA objective).
 * Penalties here are already adjusted (x', o1', e1', o2', e2'); metric 0 uses e1 as
 * the linear indel cost.
 */
#include <stdint.h>
#include <stdlib.h>

#define INF (INT32_MAX / 4)
static inline int mn(int a, int b) { return a < b ? a : b; }

int dp_min_penalty(int metric, int x, int o1, int e1, int o2, int e2, const uint8_t* p, int plen, const uint8_t* t,
                   int tlen, int pbf, int pef, int tbf, int tef) {
  if (metric == 0) { o1 = 0; o2 = INF / 2; e2 = 0; }
  const int W = tlen + 1;
  int* H = (int*)malloc(sizeof(int) * (size_t)W * 2);
  int* F1 = (int*)malloc(sizeof(int) * (size_t)W * 2); /* gap consuming pattern ('D'), vertical */
  int* F2 = (int*)malloc(sizeof(int) * (size_t)W * 2);
  if (!H || !F1 || !F2) { free(H); free(F1); free(F2); return -1; }
  int best = INF;
  for (int i = 0; i <= plen; ++i) {
    int* h = H + (i & 1) * W;
    int* hp = H + ((i ^ 1) & 1) * W;
    int* f1 = F1 + (i & 1) * W;
    int* f1p = F1 + ((i ^ 1) & 1) * W;
    int* f2 = F2 + (i & 1) * W;
    int* f2p = F2 + ((i ^ 1) & 1) * W;
    int e1v = INF, e2v = INF; /* gap consuming text ('I'), horizontal */
    for (int j = 0; j <= tlen; ++j) {
      int v = INF;
      if ((j == 0 && i <= pbf) || (i == 0 && j <= tbf)) v = 0;
      if (i > 0) {
        f1[j] = mn(hp[j] + o1 + e1, f1p[j] + e1);
        f2[j] = metric == 0 ? INF : mn(hp[j] + o2 + e2, f2p[j] + e2);
      } else {
        f1[j] = INF;
        f2[j] = INF;
      }
      if (j > 0) {
        e1v = mn(h[j - 1] + o1 + e1, e1v + e1);
        e2v = metric == 0 ? INF : mn(h[j - 1] + o2 + e2, e2v + e2);
      } else {
        e1v = INF;
        e2v = INF;
      }
      if (i > 0 && j > 0) v = mn(v, hp[j - 1] + (p[i - 1] == t[j - 1] ? 0 : x));
      v = mn(v, mn(mn(f1[j], f2[j]), mn(e1v, e2v)));
      if (v > INF) v = INF;
      if (f1[j] > INF) f1[j] = INF;
      if (f2[j] > INF) f2[j] = INF;
      if (e1v > INF) e1v = INF;
      if (e2v > INF) e2v = INF;
      h[j] = v;
      if ((j == tlen && plen - i <= pef) || (i == plen && tlen - j <= tef)) best = mn(best, v);
    }
  }
  free(H); free(F1); free(F2);
  return best;
}
