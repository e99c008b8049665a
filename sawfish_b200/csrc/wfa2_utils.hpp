// wfa2_utils.hpp — C++ host-side mirror of the reference's src/wfa2_utils.rs + src/simple_alignment.rs over the C ABI.
//
// Same names, argument meaning and error behaviour as the Rust originals:
//   wfa2_utils::PairwiseAligner::{new_large_indel_aligner,new_consensus_aligner,set_text,text,align,
//   align_clip_frac,align_end_to_end}   (src/wfa2_utils.rs:143-251)
//   process_wfa2_cigar_string            (src/wfa2_utils.rs:13-95)
//   SimpleAlignment{ref_offset,cigar} + get_reverse (src/simple_alignment.rs:6-28)
// Rust's assert!/panic (with panic = "abort", Cargo.toml:10-11) becomes std::abort() after a message.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <optional>
#include <string>
#include <utility>
#include <vector>

#include "../../include/sawfish_cuda.h"

namespace sawfish_b200 {

[[noreturn]] inline void panic(const char* what) {
  std::fprintf(stderr, "panic: %s\n", what);
  std::abort();
}
#define SFB_ASSERT(cond) do { if (!(cond)) ::sawfish_b200::panic(#cond); } while (0)

// htslib-style cigar element: op index into "MIDNSHP=X"
struct Cigar {
  char op;
  uint32_t len;
  bool operator==(const Cigar& o) const { return op == o.op && len == o.len; }
};
inline Cigar cigar_from_u32(uint32_t w) { return Cigar{"MIDNSHP=X"[w & 0xF], w >> 4}; }
inline std::string cigar_to_string(const std::vector<Cigar>& c) {
  std::string s;
  for (const auto& e : c) { s += std::to_string(e.len); s += e.op; }
  return s;
}
inline int64_t get_cigar_ref_offset(const std::vector<Cigar>& c) {
  int64_t r = 0;
  for (const auto& e : c) if (e.op == 'D' || e.op == 'N' || e.op == 'X' || e.op == '=' || e.op == 'M') r += e.len;
  return r;
}

struct SimpleAlignment {
  int64_t ref_offset = 0;
  std::vector<Cigar> cigar;
  SimpleAlignment get_reverse(int64_t ref_size) const {  // src/simple_alignment.rs:20-27
    const int64_t ref_end_offset = ref_offset + get_cigar_ref_offset(cigar);
    SFB_ASSERT(ref_end_offset <= ref_size);
    SimpleAlignment r;
    r.ref_offset = ref_size - ref_end_offset;
    r.cigar.assign(cigar.rbegin(), cigar.rend());
    return r;
  }
};

inline std::optional<SimpleAlignment> process_wfa2_cigar_string(const std::vector<uint8_t>& wfa2_cigar) {
  std::vector<uint32_t> cig(wfa2_cigar.size() + 1);
  int64_t roff = 0;
  size_t n = 0;
  int has_match = 0;
  const int rc = sfc_translate_cigar(wfa2_cigar.data(), wfa2_cigar.size(), &roff, cig.data(), cig.size(), &n, &has_match);
  if (rc != 0) panic("Unexpected wfa2 cigar string content");  // src/wfa2_utils.rs:62-64
  if (!has_match) return std::nullopt;
  SimpleAlignment al;
  al.ref_offset = roff;
  for (size_t i = 0; i < n; ++i) al.cigar.push_back(cigar_from_u32(cig[i]));
  return al;
}

class PairwiseAligner {
 public:
  static PairwiseAligner new_large_indel_aligner(sfc_ctx* ctx, bool is_long_contig) {
    PairwiseAligner a;
    SFB_ASSERT(sfc_aligner_new_large_indel(ctx, is_long_contig ? 1 : 0, &a.h_) == 0);
    return a;
  }
  static PairwiseAligner new_consensus_aligner(sfc_ctx* ctx) {
    PairwiseAligner a;
    SFB_ASSERT(sfc_aligner_new_consensus(ctx, &a.h_) == 0);
    return a;
  }
  PairwiseAligner(PairwiseAligner&& o) noexcept : h_(o.h_), text_(std::move(o.text_)) { o.h_ = nullptr; }
  PairwiseAligner(const PairwiseAligner&) = delete;
  ~PairwiseAligner() { if (h_) sfc_aligner_free(h_); }

  void set_text(const std::vector<uint8_t>& text) {
    SFB_ASSERT(!text.empty());  // src/wfa2_utils.rs:199
    text_.assign(text.rbegin(), text.rend());
    SFB_ASSERT(sfc_aligner_set_text(h_, text.data(), text.size()) == 0);
  }
  const std::vector<uint8_t>& text() const { return text_; }  // the reversed copy, as in the reference

  std::pair<int32_t, std::optional<SimpleAlignment>> align(const std::vector<uint8_t>& pattern) { return run(pattern, 0, 0, 0); }
  std::pair<int32_t, std::optional<SimpleAlignment>> align_clip_frac(const std::vector<uint8_t>& pattern, double text_clip_frac,
                                                                     double pattern_clip_frac) {
    return run(pattern, 1, text_clip_frac, pattern_clip_frac);
  }
  std::pair<int32_t, std::optional<SimpleAlignment>> align_end_to_end(const std::vector<uint8_t>& pattern) { return run(pattern, 2, 0, 0); }

 private:
  PairwiseAligner() = default;
  std::pair<int32_t, std::optional<SimpleAlignment>> run(const std::vector<uint8_t>& pattern, int mode, double tcf, double pcf) {
    SFB_ASSERT(!text_.empty());    // src/wfa2_utils.rs:210
    SFB_ASSERT(!pattern.empty());  // src/wfa2_utils.rs:211
    std::vector<uint32_t> cig(pattern.size() + text_.size() + 8);
    int32_t score = 0;
    int has = 0;
    int64_t roff = 0;
    size_t n = 0;
    const int rc = sfc_aligner_align(h_, pattern.data(), pattern.size(), mode, tcf, pcf, &score, &has, &roff, cig.data(), cig.size(), &n);
    SFB_ASSERT(rc == 0);  // == assert_eq!(status, AlignmentStatus::StatusAlgCompleted), src/wfa2_utils.rs:117
    std::optional<SimpleAlignment> al;
    if (has) {
      al.emplace();
      al->ref_offset = roff;
      for (size_t i = 0; i < n; ++i) al->cigar.push_back(cigar_from_u32(cig[i]));
    }
    return {score, al};
  }
  sfc_aligner* h_ = nullptr;
  std::vector<uint8_t> text_;
};

}  // namespace sawfish_b200
