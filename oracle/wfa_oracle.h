/*
 * wfa_oracle.h — CPU oracle for the sawfish WFA alignment hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a from-scratch CPU restatement of the
 * semantics of WFA2-lib (C), which sawfish reaches through
 * rust-wfa2 @ 4342b3b06278656fa51c0a33b4eb0b67d53bfa8c (reference Cargo.lock:1731-1737)
 * from src/wfa2_utils.rs.  WFA2-lib is NOT vendored under /root/reference and
 * there is no network, so this file restates the published algorithm
 * (Marco-Sola et al., WFA / WFA2-lib v2.3.x) from recall; it is pinned by the
 * reference's own known-answer tests (src/wfa2_utils.rs:280-308) and by an
 * independent full-DP check of score optimality (oracle/dp_verify.c).
 * Everything beyond those two vectors is "parity unpinned" against the real
 * WFA2 fork.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may link or load this code.  The product path
 * (libsawfish_cuda.so) never does.
 */
#ifndef WFA_ORACLE_H
#define WFA_ORACLE_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Same layout as sfc_penalties (include/sawfish_cuda.h).
 * metric 0 = gap-linear (indel penalty = gap_ext1, gap_open1 ignored),
 * metric 1 = gap-affine 2-piece.  match <= 0 as in WFA2-lib. */
typedef struct {
  int32_t metric;
  int32_t match, mismatch, gap_open1, gap_ext1, gap_open2, gap_ext2;
} ora_penalties;

typedef struct {
  int64_t cells;      /* M-wavefront cells computed: sum over computed scores of (hi-lo+1) */
  int64_t ext_words;  /* 16-base compare words consumed by extends: sum ceil((matched+1)/16) */
  int64_t bt_bytes;   /* choice bytes recorded */
  int32_t wf_score;   /* final wavefront (adjusted-penalty) score */
  int32_t end_k, end_off, k0;
  int32_t n_ops;
} ora_stats;

enum { ORA_MODE_LITERAL = 0, ORA_MODE_COMPACT = 1 };

/* Core WFA (in the aligner's own coordinate space: no reversal here).
 * ops: WFA2 op bytes 'M','X','I','D' (I consumes text, D consumes pattern), begin->end.
 * Returns 0, or <0 on error (-1 bad args, -2 ops buffer too small, -3 memory cap,
 * -4 internal inconsistency between literal and compact backtrace, -5 score cap). */
int ora_wfa_align(const ora_penalties* pen,
                  const uint8_t* pattern, int plen, const uint8_t* text, int tlen,
                  int pattern_begin_free, int pattern_end_free,
                  int text_begin_free, int text_end_free,
                  int mode, int want_ops,
                  int32_t* score, uint8_t* ops, int ops_cap, ora_stats* st);

/* == process_wfa2_cigar_string (reference src/wfa2_utils.rs:13-95).
 * cigar: htslib encoding len<<4|op ('='=7, X=8, I=1, D=2, S=4). */
int ora_translate_cigar(const uint8_t* wfa_ops, size_t n_ops,
                        int64_t* ref_offset, uint32_t* cigar, size_t cigar_cap,
                        size_t* n_cigar, int* has_match);

/* == PairwiseAligner::align / align_clip_frac (reference src/wfa2_utils.rs:198-251):
 * reverses both sequences, ends-free with clip fractions, translates.
 * text_clip_frac/pattern_clip_frac < 0 selects the production rule
 * clip = min(floor(0.4|p|), floor(0.4|t|)) on all four ends (wfa2_utils.rs:106-118);
 * end_to_end != 0 selects align_end_to_end (wfa2_utils.rs:243-251). */
int ora_sawfish_align(const ora_penalties* pen,
                      const uint8_t* text, int tlen, const uint8_t* pattern, int plen,
                      double text_clip_frac, double pattern_clip_frac, int end_to_end, int mode,
                      int32_t* score, int64_t* ref_offset, uint32_t* cigar, size_t cigar_cap,
                      size_t* n_cigar, int* has_match, ora_stats* st);

/* Batch driver used as the CPU baseline: one fresh aligner state per pair (as
 * score_sv does, reference src/score_sv/mod.rs:837), OpenMP dynamic over pairs.
 * seqs are forward ASCII; reverse flag + clips are per pair, exactly the
 * information the C-ABI sfc_pair carries. */
typedef struct {
  uint64_t pattern_off, text_off;
  uint32_t pattern_len, text_len;
  int32_t pattern_begin_free, pattern_end_free, text_begin_free, text_end_free;
  uint32_t flags; /* bit0: reverse both sequences */
} ora_pair;

int ora_align_batch(const ora_penalties* pen, const uint8_t* arena,
                    const ora_pair* pairs, size_t n_pairs, int want_ops, int mode, int n_threads,
                    int32_t* scores, int32_t* status, ora_stats* stats /* [n] or NULL */,
                    uint8_t* ops /* may be NULL */, size_t ops_cap, uint64_t* ops_off /* [n+1] */);

int ora_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
