// score_sv_batch.hpp — C++ host side of score_sv's two-phase batching over the C ABI.
//
// The reference scores one (read flank, allele flank) pair at a time inside its read loop
// (get_read_support_scores_from_segment, src/score_sv/mod.rs:758-905: a fresh consensus aligner per read x breakend x
// registration, set_text(read flank), then align(ref / alt / overlapping-haplotype flank) through
// get_read_flank_align_score_option, :721-745).  Whether a slot is scored depends only on geometry (flank present,
// truncations + 100 <= flank length), so the loop splits into:
//   phase 1  PairBatch::add_read_flank(...) for every read x breakend x registration — appends the read flank (text) once
//            and every allele flank (pattern, cut by the read's truncations) to one arena, one sfc_pair per slot with
//            SFC_PAIR_REVERSE and clip = min(int(0.4 |p|), int(0.4 |t|)) on all four ends (src/wfa2_utils.rs:106-118);
//   phase 2  PairBatch::score(ctx): ONE sfc_align_batch (consensus penalties, score only -> K1);
//   phase 3  PairBatch::norm_score(slot) = score / text.len() (:702-703), to be written into ReadSupportScores by the caller.
// Same rules as sawfish_b200/score_sv.py::PairBatchBuilder (the Python mirror the parity tests drive).
#pragma once
#include <cstdint>
#include <optional>
#include <string>
#include <vector>

#include "wfa2_utils.hpp"

namespace sawfish_b200 {

enum class AlleleType : int { Ref = 0, Alt = 1, Overlap = 2 };
constexpr size_t MIN_READ_SUPPORT_FLANK_SIZE = 100;  // ScoreSVSettings::new, src/score_sv/mod.rs:2047-2048

struct ReadFlank {           // a read's flank sequence + how much of the 500-base window the read ends cut off (ReadFlankInfo)
  const uint8_t* seq;
  size_t len;
  size_t start_truncation, end_truncation;
};

struct FlankSeq {            // Option<&[u8]>
  const uint8_t* seq = nullptr;
  size_t len = 0;
  bool some() const { return seq != nullptr; }
};

struct ScoreSlot {
  uint64_t tag;              // caller's key (SV group x sample)
  std::string qname;
  int breakend, registration;
  AlleleType allele;
  int overlap_index;         // index into the overlapping-haplotype list, -1 otherwise
  uint32_t text_len;
};

inline int32_t sawfish_clip(size_t pattern_len, size_t text_len) {  // src/wfa2_utils.rs:112-115 (f64 product truncated to i32)
  const int32_t a = (int32_t)((double)pattern_len * 0.4), b = (int32_t)((double)text_len * 0.4);
  return a < b ? a : b;
}

class PairBatch {
 public:
  // == the body of the registration loop at :829-901 minus the alignments themselves
  void add_read_flank(uint64_t tag, const std::string& qname, int breakend, int registration, const ReadFlank& flank,
                      FlankSeq ref_seq, FlankSeq alt_seq, const std::vector<FlankSeq>& overlap_seqs) {
    SFB_ASSERT(flank.len > 0);                                     // set_text: assert!(!text.is_empty()), :199
    int64_t t_off = -1;
    auto one = [&](AlleleType allele, int oi, FlankSeq s) {
      if (!s.some()) return;                                        // let seq = seq?;
      if (flank.start_truncation + flank.end_truncation + MIN_READ_SUPPORT_FLANK_SIZE > s.len) return;  // :732-736
      const uint8_t* p = s.seq + flank.start_truncation;
      const size_t plen = s.len - flank.start_truncation - flank.end_truncation;
      if (t_off < 0) t_off = (int64_t)append(flank.seq, flank.len);
      const uint64_t p_off = append(p, plen);
      const int32_t c = sawfish_clip(plen, flank.len);
      pairs_.push_back(sfc_pair{p_off, (uint64_t)t_off, (uint32_t)plen, (uint32_t)flank.len, c, c, c, c, SFC_PAIR_REVERSE});
      slots_.push_back(ScoreSlot{tag, qname, breakend, registration, allele, oi, (uint32_t)flank.len});
    };
    one(AlleleType::Ref, -1, ref_seq);
    one(AlleleType::Alt, -1, alt_seq);
    for (size_t i = 0; i < overlap_seqs.size(); ++i) one(AlleleType::Overlap, (int)i, overlap_seqs[i]);
  }

  size_t size() const { return pairs_.size(); }
  const std::vector<sfc_pair>& pairs() const { return pairs_; }
  const std::vector<uint8_t>& arena() const { return arena_; }
  const std::vector<ScoreSlot>& slots() const { return slots_; }

  // phase 2: one batch on the GPU.  Returns 0 or the library's error code; statuses must all be "completed", as the
  // reference asserts per alignment (src/wfa2_utils.rs:117).
  int score(sfc_ctx* ctx) {
    scores_.assign(pairs_.size(), 0);
    if (pairs_.empty()) return 0;
    std::vector<int32_t> status(pairs_.size(), 0);
    const sfc_penalties consensus{0, -1, 3, 0, 2, 0, 0};          // new_consensus_aligner, src/wfa2_utils.rs:183-189
    const int rc = sfc_align_batch(ctx, &consensus, arena_.data(), arena_.size(), pairs_.data(), pairs_.size(), 0, scores_.data(),
                                   status.data(), nullptr, 0, nullptr);
    if (rc) return rc;
    for (int32_t s : status) SFB_ASSERT(s == SFC_STATUS_COMPLETED);
    return 0;
  }

  // phase 3: norm_score = score as f64 / aligner.text().len() as f64 (:702-703)
  double norm_score(size_t slot) const { return (double)scores_[slot] / (double)slots_[slot].text_len; }
  int32_t raw_score(size_t slot) const { return scores_[slot]; }

 private:
  uint64_t append(const uint8_t* p, size_t n) {
    const uint64_t off = arena_.size();
    arena_.insert(arena_.end(), p, p + n);
    return off;
  }
  std::vector<uint8_t> arena_;
  std::vector<sfc_pair> pairs_;
  std::vector<ScoreSlot> slots_;
  std::vector<int32_t> scores_;
};

}  // namespace sawfish_b200
