// C++ replay of the reference's own unit tests for the hot path, through the C++ mirror (wfa2_utils.hpp) and the C ABI.
//   host  : test_process_wfa2_cigar_string (src/wfa2_utils.rs:280-292), test_get_reverse (src/simple_alignment.rs:78-109)
//   gpu   : test_consensus_aligner (src/wfa2_utils.rs:294-308)
// usage: test_wfa2_utils host|gpu
#include <cstring>
#include <iostream>

#include "../../sawfish_b200/csrc/wfa2_utils.hpp"
#include "../../sawfish_b200/csrc/score_sv_batch.hpp"
#include "../../sawfish_b200/csrc/refine_sv_batch.hpp"

using namespace sawfish_b200;

static std::vector<uint8_t> bytes(const char* s) { return std::vector<uint8_t>(s, s + std::strlen(s)); }
#define CHECK(c) do { if (!(c)) { std::cerr << "FAILED: " #c " (line " << __LINE__ << ")\n"; return 1; } } while (0)

static int test_process_wfa2_cigar_string() {
  auto wfa2_cigar = bytes("DDDDMMMXXMMMMIIIMMMM");
  std::vector<uint8_t> rev(wfa2_cigar.rbegin(), wfa2_cigar.rend());
  auto al = process_wfa2_cigar_string(rev);
  CHECK(al.has_value());
  CHECK(cigar_to_string(al->cigar) == "4S3=2X4=3D4=");
  CHECK(al->ref_offset == 0);
  CHECK(!process_wfa2_cigar_string(bytes("IIDD")).has_value());
  return 0;
}

static int test_get_reverse() {
  SimpleAlignment al{1, {{'M', 1}}};
  auto alr = al.get_reverse(4);
  CHECK(alr.ref_offset == 2 && alr.cigar == std::vector<Cigar>({{'M', 1}}));
  al = SimpleAlignment{1, {{'M', 2}}};
  alr = al.get_reverse(4);
  CHECK(alr.ref_offset == 1);
  al = SimpleAlignment{2, {{'M', 2}, {'D', 2}, {'X', 1}}};
  alr = al.get_reverse(20);
  CHECK(alr.ref_offset == 13 && alr.cigar == std::vector<Cigar>({{'X', 1}, {'D', 2}, {'M', 2}}));
  return 0;
}

static int test_consensus_aligner(sfc_ctx* ctx) {
  auto aligner = PairwiseAligner::new_consensus_aligner(ctx);
  aligner.set_text(bytes("AACTGTATAA"));
  auto [score, alignment] = aligner.align(bytes("ACTTAT"));
  // Expect:
  // AACTGTATAA
  // -ACT-TAT--
  CHECK(score == 4);
  CHECK(alignment.has_value());
  CHECK(alignment->ref_offset == 1);
  CHECK(cigar_to_string(alignment->cigar) == "3=1D3=");
  // same pair with the large-indel preset: one 1-base gap costs 1+2 = 3
  auto big = PairwiseAligner::new_large_indel_aligner(ctx, false);
  big.set_text(bytes("AACTGTATAA"));
  auto [s2, a2] = big.align(bytes("ACTTAT"));
  CHECK(s2 == 3 && a2.has_value() && cigar_to_string(a2->cigar) == "3=1D3=" && a2->ref_offset == 1);
  return 0;
}

// score_sv's two-phase batcher: the geometry rules of get_read_flank_align_score_option (src/score_sv/mod.rs:721-745)
static std::vector<uint8_t> rep(const char* unit, size_t n) {
  std::vector<uint8_t> v;
  const size_t u = std::strlen(unit);
  for (size_t i = 0; i < n; ++i) v.push_back((uint8_t)unit[i % u]);
  return v;
}

static int test_pair_batch_geometry() {
  PairBatch b;
  const auto read = rep("ACGTTGCA", 400), ref = rep("ACGTTGCA", 500), alt = rep("TTGACCAG", 500), tiny = rep("ACGT", 120);
  // untruncated read flank: ref + alt + one overlapping haplotype, the read flank is stored once
  b.add_read_flank(7, "read1", 0, 0, ReadFlank{read.data(), read.size(), 0, 0}, FlankSeq{ref.data(), ref.size()},
                   FlankSeq{alt.data(), alt.size()}, {FlankSeq{alt.data(), alt.size()}, FlankSeq{}});
  CHECK(b.size() == 3);
  CHECK(b.pairs()[0].text_off == b.pairs()[1].text_off && b.pairs()[1].text_off == b.pairs()[2].text_off);
  CHECK(b.pairs()[0].pattern_len == 500 && b.pairs()[0].text_len == 400 && b.pairs()[0].flags == SFC_PAIR_REVERSE);
  CHECK(b.pairs()[0].pattern_begin_free == 160 && b.pairs()[0].text_end_free == 160);   // min(int(0.4*500), int(0.4*400))
  CHECK(b.slots()[2].allele == AlleleType::Overlap && b.slots()[2].overlap_index == 0 && b.slots()[0].text_len == 400);
  // truncated read flank: the allele flank is cut by the same amounts; a flank shorter than truncations + 100 is skipped
  b.add_read_flank(7, "read2", 1, 1, ReadFlank{read.data(), 300, 150, 50}, FlankSeq{ref.data(), ref.size()}, FlankSeq{tiny.data(), tiny.size()}, {});
  CHECK(b.size() == 4);
  CHECK(b.pairs()[3].pattern_len == 300 && b.pairs()[3].text_len == 300 && b.pairs()[3].pattern_begin_free == 120);
  CHECK(std::memcmp(b.arena().data() + b.pairs()[3].pattern_off, ref.data() + 150, 300) == 0);
  CHECK(b.slots()[3].breakend == 1 && b.slots()[3].registration == 1 && b.slots()[3].allele == AlleleType::Ref);
  // nothing scorable: no pair, and the read flank is not stored either
  const size_t arena_before = b.arena().size();
  b.add_read_flank(8, "read3", 0, 0, ReadFlank{read.data(), 300, 200, 250}, FlankSeq{ref.data(), ref.size()}, FlankSeq{}, {});
  CHECK(b.size() == 4 && b.arena().size() == arena_before);
  CHECK(sawfish_clip(6, 10) == 2 && sawfish_clip(499, 1000) == 199);
  return 0;
}

static std::vector<uint8_t> lcg_seq(uint32_t seed, size_t n) {  // non-periodic ACGT (a periodic flank lets the ends-free
  std::vector<uint8_t> v(n);                                     // wavefront stop on an earlier, shorter diagonal)
  uint32_t x = seed;
  for (size_t i = 0; i < n; ++i) { x = x * 1664525u + 1013904223u; v[i] = (uint8_t)"ACGT"[(x >> 24) & 3]; }
  return v;
}

static int test_pair_batch_on_gpu(sfc_ctx* ctx) {
  PairBatch b;
  const auto same = lcg_seq(12345u, 500), other = lcg_seq(999u, 500);
  const std::vector<uint8_t> read(same.begin(), same.begin() + 480);       // the read flank is a prefix of this allele's flank
  b.add_read_flank(1, "r", 0, 0, ReadFlank{read.data(), read.size(), 0, 0}, FlankSeq{same.data(), same.size()}, FlankSeq{other.data(), other.size()}, {});
  CHECK(b.score(ctx) == 0);
  CHECK(b.raw_score(0) == 480 && b.norm_score(0) == 1.0);   // 480 matches, the 20 extra allele bases are a free end
  CHECK(b.norm_score(1) < 0.5);
  auto al = PairwiseAligner::new_consensus_aligner(ctx);     // the per-pair surface gives the same number as the batch
  al.set_text(read);
  CHECK(al.align(same).first == b.raw_score(0));
  return 0;
}

// refine_sv's two-region alt-haplotype text (src/refine_sv/align/mod.rs:708-793) and the speculative contig batch
static int test_alt_hap_text_rules() {
  const std::vector<uint8_t> s1 = {'A', 'A', 'A', 'A', 'C', 'C', 'C', 'C'}, s2 = {'G', 'G', 'T', 'T'};
  auto str = [](const std::vector<uint8_t>& v) { return std::string(v.begin(), v.end()); };
  AltHapSeq a = construct_alt_hap_seq(s1, s2, 3, false, true);
  CHECK(str(a.seq) == "AAAACCCCNNNGGTT" && a.spacer_start == 8 && a.spacer_end == 11);
  a = construct_alt_hap_seq(s1, s2, 3, false, false);
  CHECK(str(a.seq) == "GGTTNNNAAAACCCC" && a.spacer_start == 4 && a.spacer_end == 7);
  a = construct_alt_hap_seq(s1, s2, 3, true, true);
  CHECK(str(a.seq) == "AAAACCCCNNNAACC");
  CHECK(get_left_right_ref_flank_size(3500, false, true) == std::make_pair((int64_t)3500, (int64_t)2000));
  CHECK(get_left_right_ref_flank_size(3500, true, true) == std::make_pair((int64_t)2000, (int64_t)3500));
  CHECK(get_left_right_ref_flank_size(1500, true, false) == std::make_pair((int64_t)1500, (int64_t)1500));
  CHECK(get_middle_flank_size(3500, false, true, 1100, 2100) == 500 && get_middle_flank_size(3500, true, true, 1100, 2100) == 3500);
  return 0;
}

static int test_contig_batch_on_gpu(sfc_ctx* ctx) {
  // two clusters; the first contig of cluster 0 is junk (no '=' anchors across the spacer), its retry contig is clean
  const auto ref = lcg_seq(77u, 30000);
  auto sub = [&](size_t a, size_t b) { return std::vector<uint8_t>(ref.begin() + a, ref.begin() + b); };
  ContigBatch b;
  std::vector<std::vector<size_t>> asm_pairs;
  std::vector<AltHapSeq> texts;
  for (int c = 0; c < 2; ++c) {
    const size_t p1 = 6000 + 9000 * c, p2 = p1 + 4000;
    texts.push_back(construct_alt_hap_seq(sub(p1 - 3500, p1 + 2000), sub(p2 - 2000, p2 + 3500), 100, false, true));
    b.add_text(texts.back().seq.data(), texts.back().seq.size());
    std::vector<size_t> pairs;
    for (size_t flank : {1500u, 2500u}) {
      std::vector<uint8_t> contig = sub(p1 - flank, p1);
      const auto right = sub(p2, p2 + flank);
      contig.insert(contig.end(), right.begin(), right.end());
      if (c == 0 && flank == 1500) contig = lcg_seq(4242u, contig.size());
      pairs.push_back(b.add_contig(contig.data(), contig.size()));
    }
    asm_pairs.push_back(pairs);
  }
  CHECK(b.align(ctx) == 0);
  for (int c = 0; c < 2; ++c) {
    const AltHapSeq& t = texts[(size_t)c];
    auto crosses_spacer_with_deletion = [&](size_t, const SimpleAlignment& a) {
      int64_t pos = a.ref_offset;
      for (const auto& e : a.cigar) {
        if (e.op == 'D' && pos <= t.spacer_start && pos + (int64_t)e.len >= t.spacer_end) return true;
        if (e.op == 'D' || e.op == '=' || e.op == 'X') pos += e.len;
      }
      return false;
    };
    const auto k = b.first_passing(asm_pairs[(size_t)c], crosses_spacer_with_deletion);
    CHECK(k.has_value() && *k == (c == 0 ? 1u : 0u));
    // the batch equals the per-pair surface
    auto al = PairwiseAligner::new_large_indel_aligner(ctx, false);
    al.set_text(t.seq);
    const size_t pi = asm_pairs[(size_t)c][*k];
    std::vector<uint8_t> contig;  // rebuild the accepted contig
    const size_t p1 = 6000 + 9000 * c, p2 = p1 + 4000, flank = *k == 0 ? 1500 : 2500;
    contig = sub(p1 - flank, p1);
    const auto right = sub(p2, p2 + flank);
    contig.insert(contig.end(), right.begin(), right.end());
    const auto one = al.align(contig);
    CHECK(one.first == b.score(pi) && one.second.has_value() && one.second->ref_offset == b.alignment(pi)->ref_offset &&
          one.second->cigar == b.alignment(pi)->cigar);
  }
  return 0;
}

int main(int argc, char** argv) {
  const std::string mode = argc > 1 ? argv[1] : "host";
  if (test_process_wfa2_cigar_string() || test_get_reverse() || test_pair_batch_geometry() || test_alt_hap_text_rules()) return 1;
  if (mode == "gpu") {
    sfc_ctx* ctx = nullptr;
    if (sfc_create(0, &ctx) != 0) { std::cerr << "sfc_create failed: a B200 is required (no CPU fallback)\n"; return 2; }
    int rc = test_consensus_aligner(ctx);
    if (!rc) rc = test_pair_batch_on_gpu(ctx);
    if (!rc) rc = test_contig_batch_on_gpu(ctx);
    sfc_destroy(ctx);
    if (rc) return rc;
  }
  std::cout << "ok " << mode << "\n";
  return 0;
}
