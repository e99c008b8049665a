// C++ replay of the reference's own unit tests for the hot path, through the C++ mirror (wfa2_utils.hpp) and the C ABI.
//   host  : test_process_wfa2_cigar_string (src/wfa2_utils.rs:280-292), test_get_reverse (src/simple_alignment.rs:78-109)
//   gpu   : test_consensus_aligner (src/wfa2_utils.rs:294-308)
// usage: test_wfa2_utils host|gpu
#include <cstring>
#include <iostream>

#include "../../sawfish_b200/csrc/wfa2_utils.hpp"

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

int main(int argc, char** argv) {
  const std::string mode = argc > 1 ? argv[1] : "host";
  if (test_process_wfa2_cigar_string() || test_get_reverse()) return 1;
  if (mode == "gpu") {
    sfc_ctx* ctx = nullptr;
    if (sfc_create(0, &ctx) != 0) { std::cerr << "sfc_create failed: a B200 is required (no CPU fallback)\n"; return 2; }
    const int rc = test_consensus_aligner(ctx);
    sfc_destroy(ctx);
    if (rc) return rc;
  }
  std::cout << "ok " << mode << "\n";
  return 0;
}
