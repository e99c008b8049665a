// C++ replay of the reference's own unit tests for the hot path, through the C++ mirror (wfa2_utils.hpp) and the C ABI.
//   host  : test_process_wfa2_cigar_string (src/wfa2_utils.rs:280-292), test_get_reverse (src/simple_alignment.rs:78-109)
//   gpu   : test_consensus_aligner (src/wfa2_utils.rs:294-308)
// usage: test_wfa2_utils host|gpu
#include <cstring>
#include <iostream>

#include "../../sawfish_b200/csrc/wfa2_utils.hpp"
#include "../../sawfish_b200/csrc/score_sv_batch.hpp"

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

int main(int argc, char** argv) {
  const std::string mode = argc > 1 ? argv[1] : "host";
  if (test_process_wfa2_cigar_string() || test_get_reverse() || test_pair_batch_geometry()) return 1;
  if (mode == "gpu") {
    sfc_ctx* ctx = nullptr;
    if (sfc_create(0, &ctx) != 0) { std::cerr << "sfc_create failed: a B200 is required (no CPU fallback)\n"; return 2; }
    int rc = test_consensus_aligner(ctx);
    if (!rc) rc = test_pair_batch_on_gpu(ctx);
    sfc_destroy(ctx);
    if (rc) return rc;
  }
  std::cout << "ok " << mode << "\n";
  return 0;
}
