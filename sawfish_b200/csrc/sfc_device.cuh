// sfc_device.cuh — shared device-side definitions for the sm_100a wavefront kernels.
//
// Coordinates follow WFA2-lib (what sawfish reaches through src/wfa2_utils.rs): h = text position,
// v = pattern position, diagonal k = h - v, wavefront offset = h.  'I' consumes text, 'D' pattern.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace sfc {

constexpr int32_t WNULL = INT32_MIN / 2;  // null offset (same value WFA2-lib uses: INT32_MIN/2)
constexpr unsigned FULL = 0xFFFFFFFFu;

// Sequence alphabet on the device: one nibble per base.  A,C,G,T,N (and lower case) get distinct
// codes so that byte equality of the reference (N==N matches, N!=base) is preserved exactly;
// 0xE / 0xF are the pattern / text end sentinels (never equal to anything on the other side).
constexpr uint32_t NIB_PAT_SENTINEL = 0xE;
constexpr uint32_t NIB_TXT_SENTINEL = 0xF;

struct DevPair {
  uint32_t p_woff, t_woff;  // word offsets into the packed nibble arena
  int32_t plen, tlen;
  int32_t pbf, pef, tbf, tef;  // ends-free allowances
};

struct SeqDesc {
  uint64_t src_off;  // byte offset in the ASCII arena
  uint32_t len;
  uint32_t flags;  // bit0 reverse, bit1 text (selects the sentinel)
};

// 8 consecutive bases (nibbles) starting at position pos, from the overlapped-window layout:
// win[i] = (word[i], word[i+1]) so that one 8-byte shared load + one funnel shift yields the window.
__device__ __forceinline__ uint32_t window8(const uint2* __restrict__ win, int pos) {
  const uint2 u = win[pos >> 3];
  return __funnelshift_r(u.x, u.y, pos << 2);  // the funnel shift uses its amount modulo 32: (pos << 2) & 31 == (pos & 7) << 2
}

// number of leading equal bases given xor of two windows (x != 0)
__device__ __forceinline__ int lead_eq(uint32_t x) { return (__ffs((int)x) - 1) >> 2; }

__device__ __forceinline__ int warp_min(int v) { return __reduce_min_sync(FULL, v); }
__device__ __forceinline__ int warp_max(int v) { return __reduce_max_sync(FULL, v); }

// Warp-cooperative match extension: all 32 lanes compare 8 bases each (256 per round) starting at
// (h0, v0 = h0 - k).  Returns the number of matching bases (uniform across the warp).
__device__ __forceinline__ int coop_extend(const uint2* __restrict__ wp, const uint2* __restrict__ wt, int plen,
                                           int tlen, int k, int h0, int lane) {
  int adv = 0;
  for (;;) {
    const int h = h0 + adv + lane * 8, v = h - k;
    uint32_t x = ~0u;
    if (h <= tlen && v <= plen) x = window8(wp, v) ^ window8(wt, h);
    const unsigned full = __ballot_sync(FULL, x == 0u);
    if (full == FULL) { adv += 256; continue; }
    const int first = __ffs((int)~full) - 1;
    const int part = __shfl_sync(FULL, x ? lead_eq(x) : 0, first);
    return adv + first * 8 + part;
  }
}

}  // namespace sfc
