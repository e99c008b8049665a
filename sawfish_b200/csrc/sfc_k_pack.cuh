// sfc_k_pack.cuh — K3: ASCII -> nibble packing with optional reversal.
//
// Replaces the `to_vec() + reverse()` of wfa2_utils::PairwiseAligner::{set_text,align}
// (reference src/wfa2_utils.rs:198-215) for a whole batch at once.  One thread produces one
// 32-bit word (8 bases); words past the end of a sequence are filled with the end sentinel so
// that wavefront extends stop at the sequence end without a bounds test.
// HBM-bound: 1 B read + 0.5 B written per base.
#pragma once
#include "sfc_device.cuh"

namespace sfc {

__constant__ uint8_t c_nib_lut[256];

// seq_woff: [n_seq+1] prefix of word offsets (each sequence owns (len+7)/8 + 2 words).
__global__ void __launch_bounds__(256) k3_pack(const uint8_t* __restrict__ arena, const SeqDesc* __restrict__ seqs,
                                              const uint32_t* __restrict__ seq_woff, uint32_t n_seq,
                                              uint32_t total_words, uint32_t* __restrict__ nib,
                                              uint32_t* __restrict__ bad_symbol) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t w = blockIdx.x * blockDim.x + threadIdx.x; w < total_words; w += stride) {
    // binary search: last sequence with seq_woff[i] <= w
    uint32_t lo = 0, hi = n_seq;
    while (hi - lo > 1) {
      const uint32_t mid = (lo + hi) >> 1;
      if (seq_woff[mid] <= w) lo = mid; else hi = mid;
    }
    const SeqDesc sd = seqs[lo];
    const uint32_t base = (w - seq_woff[lo]) * 8u;
    const uint32_t sent = (sd.flags & 2u) ? NIB_TXT_SENTINEL : NIB_PAT_SENTINEL;
    const uint8_t* src = arena + sd.src_off;
    uint32_t word = 0;
    bool bad = false;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint32_t pos = base + j;
      uint32_t code = sent;
      if (pos < sd.len) {
        const uint32_t sp = (sd.flags & 1u) ? (sd.len - 1u - pos) : pos;
        code = c_nib_lut[src[sp]];
        if (code > 9u) { bad = true; code = 0; }
      }
      word |= code << (4 * j);
    }
    nib[w] = word;
    if (bad) atomicOr(bad_symbol, 1u);
  }
}

}  // namespace sfc
