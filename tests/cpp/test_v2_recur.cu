// Host-side check of the packed s16x2 recurrence of K2 v2 (sawfish_b200/csrc/sfc_k_align2.cuh, v2_recur8) against the
// scalar statement of the same recurrence (reference semantics: WFA2 gap-affine 2-piece compute, see oracle/wfa_oracle.c
// compute()).  Runs on the CPU: the DPX intrinsics are emulated by the header's host branch.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <random>
#include "../../sawfish_b200/csrc/sfc_k_align2.cuh"

using namespace sfc;

static int lo16(uint32_t w) { return (int)(int16_t)(w & 0xFFFFu); }
static int hi16(uint32_t w) { return (int)(int16_t)(w >> 16); }

int main() {
  std::mt19937 rng(12345);
  long checked = 0;
  for (int iter = 0; iter < 200000; ++iter) {
    // rows as element arrays e0-2 .. e0+9 (index 0 .. 11), values: null, crept null, valid, saturated
    auto val = [&]() -> int {
      const int r = (int)(rng() % 10);
      if (r == 0) return V2_NULL;
      if (r == 1) return V2_NULL + (int)(rng() % 700);
      if (r == 2) return V2_CAP - (int)(rng() % 3);
      return V2_VMIN + (int)(rng() % 64000);
    };
    int mx[12], mo1[12], i1e[12], d1e[12], i2e[12], d2e[12], mo2[12];
    for (int i = 0; i < 12; ++i) { mx[i] = val(); mo1[i] = val(); i1e[i] = val(); d1e[i] = val(); i2e[i] = val(); d2e[i] = val(); mo2[i] = val(); }
    // the mismatch row never holds saturated values (M is always in the matrix)
    for (int i = 0; i < 12; ++i) if (mx[i] > 32000) mx[i] = 32000;
    auto pk = [](const int* a, int i) { return ((uint32_t)a[i] & 0xFFFFu) | ((uint32_t)a[i + 1] << 16); };
    V2Src s;
    for (int i = 0; i < 4; ++i) {
      s.mx[i] = pk(mx, 2 + 2 * i); s.mo1[i] = pk(mo1, 2 + 2 * i); s.i1e[i] = pk(i1e, 2 + 2 * i); s.d1e[i] = pk(d1e, 2 + 2 * i);
      s.i2e[i] = pk(i2e, 2 + 2 * i); s.d2e[i] = pk(d2e, 2 + 2 * i); s.mo2[i] = pk(mo2, 2 + 2 * i);
    }
    s.mo1L = pk(mo1, 0); s.mo1R = pk(mo1, 10); s.i1eL = pk(i1e, 0); s.d1eR = pk(d1e, 10);
    s.i2eL = pk(i2e, 0); s.d2eR = pk(d2e, 10); s.mo2L = pk(mo2, 0); s.mo2R = pk(mo2, 10);
    for (int variant = 0; variant < 3; ++variant) {
      uint32_t best[4], ins1[4], del1[4], ins2[4], del2[4];
      if (variant == 0) v2_recur8<true, true>(s, best, ins1, del1, ins2, del2);
      else if (variant == 1) v2_recur8<true, false>(s, best, ins1, del1, ins2, del2);
      else v2_recur8<false, false>(s, best, ins1, del1, ins2, del2);
      for (int j = 0; j < 8; ++j) {
        const int e = 2 + j;  // index of diagonal e0 + j
        auto mx2 = [](int a, int b) { return a > b ? a : b; };
        auto mn2 = [](int a, int b) { return a < b ? a : b; };
        const int o2l = variant == 0 ? mo2[e - 1] : V2_NULL, o2r = variant == 0 ? mo2[e + 1] : V2_NULL;
        const int wi1 = mn2(mx2(mo1[e - 1], i1e[e - 1]) + 1, V2_CAP), wd1 = mx2(mo1[e + 1], d1e[e + 1]);
        int wi2 = mn2(mx2(o2l, i2e[e - 1]) + 1, V2_CAP), wd2 = mx2(o2r, d2e[e + 1]);
        if (variant == 2) { wi2 = V2_NULL; wd2 = V2_NULL; }
        const int wb = mx2(mx2(mx2(wi1, wi2), mx2(wd1, wd2)), mx[e] + 1);
        const int w = j >> 1;
        const int gi1 = (j & 1) ? hi16(ins1[w]) : lo16(ins1[w]), gd1 = (j & 1) ? hi16(del1[w]) : lo16(del1[w]);
        const int gi2 = (j & 1) ? hi16(ins2[w]) : lo16(ins2[w]), gd2 = (j & 1) ? hi16(del2[w]) : lo16(del2[w]);
        const int gb = (j & 1) ? hi16(best[w]) : lo16(best[w]);
        if (gi1 != wi1 || gd1 != wd1 || gi2 != wi2 || gd2 != wd2 || gb != wb) {
          printf("MISMATCH iter %d variant %d j %d: ins1 %d/%d del1 %d/%d ins2 %d/%d del2 %d/%d best %d/%d\n", iter, variant, j, gi1, wi1, gd1,
                 wd1, gi2, wi2, gd2, wd2, gb, wb);
          return 1;
        }
        ++checked;
      }
    }
  }
  printf("v2_recur8 ok: %ld cells checked\n", checked);
  return 0;
}
