// refine_sv_batch.hpp — C++ host side of refine_sv's batched contig alignment over the C ABI.
//
// The reference aligns one assembly contig at a time with a large-indel aligner built per cluster:
//   single region   align_single_ref_region_assemblies   src/refine_sv/align/mod.rs:373-451   text = reference window
//   two regions     align_multi_ref_region_assemblies    src/refine_sv/align/mod.rs:1975-2199  text = alt-haplotype model
//                   (segment 1 + 100 x 'N' + segment 2, construct_alt_hap_seq :708-737, flank rules :755-793)
// Here every contig of every cluster of a work unit goes into ONE sfc_align_batch (large-indel penalties, CIGAR):
//   phase 1  ContigBatch::add_text() once per cluster, add_contig() per assembly contig (clip 0.4 rule, sequences reversed)
//   phase 2  ContigBatch::align(ctx)
//   phase 3  ContigBatch::alignment(i) -> (score, optional SimpleAlignment) exactly as PairwiseAligner::align returns it;
//            first_passing() replays the reference's "first contig that survives the post-alignment filters" loop of a
//            two-region assembly (1500-flank contig, then the 2500-flank retry, `break` at :2176) against the batch —
//            the retry contigs are aligned speculatively and dropped when the first one passes.
// Same rules as sawfish_b200/two_region_alt_hap.py (MultiRegionBatch) and refine_sv_align.py (align_contigs_batch).
#pragma once
#include <cstdint>
#include <functional>
#include <optional>
#include <string>
#include <utility>
#include <vector>

#include "score_sv_batch.hpp"  // sawfish_clip
#include "wfa2_utils.hpp"

namespace sawfish_b200 {

inline std::vector<uint8_t> rev_comp(const std::vector<uint8_t>& s) {
  std::vector<uint8_t> r(s.rbegin(), s.rend());
  for (auto& c : r) {
    switch (c) {
      case 'A': c = 'T'; break; case 'C': c = 'G'; break; case 'G': c = 'C'; break; case 'T': c = 'A'; break;
      case 'a': c = 't'; break; case 'c': c = 'g'; break; case 'g': c = 'c'; break; case 't': c = 'a'; break;
      default: break;  // N stays N
    }
  }
  return r;
}

struct AltHapSeq {
  std::vector<uint8_t> seq;
  int64_t spacer_start, spacer_end;
};

// src/refine_sv/align/mod.rs:708-737
inline AltHapSeq construct_alt_hap_seq(const std::vector<uint8_t>& ref_segment1_seq, std::vector<uint8_t> ref_segment2_seq, size_t spacer_size,
                                       bool ref_segment2_revcomp, bool ref_segment2_after_ref_segment1) {
  if (ref_segment2_revcomp) ref_segment2_seq = rev_comp(ref_segment2_seq);
  const std::vector<uint8_t>& a = ref_segment2_after_ref_segment1 ? ref_segment1_seq : ref_segment2_seq;
  const std::vector<uint8_t>& b = ref_segment2_after_ref_segment1 ? ref_segment2_seq : ref_segment1_seq;
  AltHapSeq out;
  out.seq = a;
  out.spacer_start = (int64_t)a.size();
  out.spacer_end = out.spacer_start + (int64_t)spacer_size;
  out.seq.insert(out.seq.end(), spacer_size, (uint8_t)'N');
  out.seq.insert(out.seq.end(), b.begin(), b.end());
  return out;
}

// :776-793 — (left size, right size); the flank next to the spacer is capped at 2000
inline std::pair<int64_t, int64_t> get_left_right_ref_flank_size(int64_t max_ref_flank_size, bool is_segment_revcomped, bool is_segment_first) {
  const int64_t interior = max_ref_flank_size < 2000 ? max_ref_flank_size : 2000;
  return (is_segment_revcomped != is_segment_first) ? std::make_pair(max_ref_flank_size, interior) : std::make_pair(interior, max_ref_flank_size);
}

// :755-770 — breakends of opposite direction on one chromosome: the flank between the regions is half their distance
inline int64_t get_middle_flank_size(int64_t max_ref_flank_size, bool same_orientation, bool same_chrom, int64_t segment1_end, int64_t segment2_start) {
  if (!same_orientation && same_chrom) {
    int64_t f = segment2_start - segment1_end;
    if (f < 0) f = 0;
    f /= 2;
    return f < max_ref_flank_size ? f : max_ref_flank_size;
  }
  return max_ref_flank_size;
}

class ContigBatch {
 public:
  // set_text(): the text of every contig added until the next add_text()
  size_t add_text(const uint8_t* text, size_t len) {
    SFB_ASSERT(len > 0);  // assert!(!text.is_empty()), src/wfa2_utils.rs:199
    text_off_ = append(text, len);
    text_len_ = len;
    return texts_++;
  }
  // align(&contig): returns the pair index
  size_t add_contig(const uint8_t* contig, size_t len) {
    SFB_ASSERT(text_len_ > 0 && len > 0);  // :210-211
    const int32_t c = sawfish_clip(len, text_len_);
    pairs_.push_back(sfc_pair{append(contig, len), text_off_, (uint32_t)len, (uint32_t)text_len_, c, c, c, c, SFC_PAIR_REVERSE});
    return pairs_.size() - 1;
  }
  size_t size() const { return pairs_.size(); }

  int align(sfc_ctx* ctx) {
    const size_t n = pairs_.size();
    scores_.assign(n, 0);
    results_.assign(n, std::nullopt);
    if (!n) return 0;
    std::vector<int32_t> status(n, 0);
    size_t cap = 0;
    for (const auto& p : pairs_) cap += (size_t)p.pattern_len + p.text_len + 4;
    std::vector<uint32_t> ops(cap);
    std::vector<uint64_t> off(n + 1, 0);
    const sfc_penalties large_indel{1, -1, 3, 1, 2, 60, 0};  // new_large_indel_aligner, src/wfa2_utils.rs:159-168
    const int rc = sfc_align_batch(ctx, &large_indel, arena_.data(), arena_.size(), pairs_.data(), n, 1, scores_.data(), status.data(),
                                   ops.data(), ops.size(), off.data());
    if (rc) return rc;
    for (size_t i = 0; i < n; ++i) {
      SFB_ASSERT(status[i] == SFC_STATUS_COMPLETED);  // assert_eq!(status, StatusAlgCompleted), :117
      const size_t nr = (size_t)(off[i + 1] - off[i]);
      std::vector<uint32_t> cig(nr + 1);
      int64_t roff = 0;
      size_t nc = 0;
      int hm = 0;
      SFB_ASSERT(sfc_translate_rle(ops.data() + off[i], nr, &roff, cig.data(), cig.size(), &nc, &hm) == 0);
      if (hm) {
        SimpleAlignment a;
        a.ref_offset = roff;
        for (size_t j = 0; j < nc; ++j) a.cigar.push_back(cigar_from_u32(cig[j]));
        results_[i] = std::move(a);
      }
    }
    return 0;
  }

  int32_t score(size_t i) const { return scores_[i]; }
  const std::optional<SimpleAlignment>& alignment(size_t i) const { return results_[i]; }

  // the contig loop of one two-region assembly (:2022-2179): pair indices in the order the reference tries the contigs
  std::optional<size_t> first_passing(const std::vector<size_t>& contig_pairs,
                                      const std::function<bool(size_t, const SimpleAlignment&)>& accept) const {
    for (size_t k = 0; k < contig_pairs.size(); ++k) {
      const auto& a = results_[contig_pairs[k]];
      if (a && accept(k, *a)) return k;
    }
    return std::nullopt;
  }

 private:
  uint64_t append(const uint8_t* p, size_t n) {
    const uint64_t off = arena_.size();
    arena_.insert(arena_.end(), p, p + n);
    return off;
  }
  std::vector<uint8_t> arena_;
  std::vector<sfc_pair> pairs_;
  std::vector<int32_t> scores_;
  std::vector<std::optional<SimpleAlignment>> results_;
  uint64_t text_off_ = 0;
  size_t text_len_ = 0, texts_ = 0;
};

}  // namespace sawfish_b200
