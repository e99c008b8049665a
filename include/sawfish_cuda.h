/*
 * sawfish_cuda.h — C ABI of libsawfish_cuda.so, the B200 (sm_100a) drop-in for the
 * pairwise-alignment hot path of PacificBiosciences/sawfish.
 *
 * What it replaces in the reference (paths relative to the sawfish repository):
 *   - rust_wfa2::aligner::{WFAligner, WFAlignerGapLinear, WFAlignerGapAffine2Pieces}
 *     as used by src/wfa2_utils.rs:5-8,116,130,135,159-169,183-190,216-218
 *     (the FFI into WFA2-lib that `wfa2_utils::PairwiseAligner` wraps);
 *   - process_wfa2_cigar_string                       src/wfa2_utils.rs:13-95
 *   - PairwiseAligner::{new_large_indel_aligner,new_consensus_aligner,set_text,align,
 *     align_clip_frac,align_end_to_end}               src/wfa2_utils.rs:143-251
 *   - the per-pair call loops that become batches:    src/score_sv/mod.rs:837-901,
 *     src/refine_sv/align/mod.rs:387-451,1991-2049
 *
 * Conventions: plain pointers and sizes only; every entry point returns 0 or a negative
 * sfc_error and never unwinds; the caller owns all host buffers, the library owns device
 * memory; one sfc_ctx per (host thread, GPU), no global state.  There is NO CPU fallback:
 * without a CUDA device sfc_create fails with SFC_ERR_NO_DEVICE.
 *
 * The reference-side binding (a `sawfish-cuda` crate: build.rs + extern "C" block) is in
 * INTEGRATION.md.
 */
#ifndef SAWFISH_CUDA_H
#define SAWFISH_CUDA_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SFC_ABI_VERSION 1

typedef struct sfc_ctx sfc_ctx; /* opaque: device buffers + stream for ONE GPU */

typedef enum {
  SFC_OK = 0,
  SFC_ERR_INVALID = -1,     /* bad argument (NULL, empty sequence, clip > length, match > 0 ...) */
  SFC_ERR_NO_DEVICE = -2,   /* no CUDA device / wrong architecture */
  SFC_ERR_CUDA = -3,        /* CUDA runtime error; see sfc_last_error */
  SFC_ERR_CAPACITY = -4,    /* caller's output buffer too small */
  SFC_ERR_UNSUPPORTED = -5, /* symbol outside {A,C,G,T,N,a,c,g,t,n}, or sequence too long for shared memory */
  SFC_ERR_OOM = -6,         /* device workspace exhausted even with one pair in flight */
  SFC_ERR_INTERNAL = -9     /* a per-pair SFC_STATUS_INTERNAL / SFC_STATUS_POOL / SFC_STATUS_MAX_SCORE surfaced through the per-pair API */
} sfc_error;

/* Per-pair status, mirrors rust_wfa2::aligner::AlignmentStatus so the caller can keep its
 * assert_eq!(status, StatusAlgCompleted) (src/wfa2_utils.rs:117). */
enum {
  SFC_STATUS_COMPLETED = 0,
  SFC_STATUS_UNSUPPORTED = -5, /* sequences do not fit the kernel's shared-memory windows */
  SFC_STATUS_OOM = -6,         /* device workspace too small even with one pair in flight */
  SFC_STATUS_MAX_SCORE = -7,   /* score cap reached */
  SFC_STATUS_INTERNAL = -9,    /* backtrace inconsistency (a bug: never expected) */
  SFC_STATUS_POOL = -10        /* run-length ops pool short; only seen transiently (the library grows the pool and re-runs) */
};

/* metric 0 = gap-linear (indel penalty = gap_ext1; gap_open1 ignored)   WFAlignerGapLinear::new_with_match
 * metric 1 = gap-affine 2-piece                                          WFAlignerGapAffine2Pieces::new_with_match
 * match <= 0, others are penalties (positive), exactly the integers sawfish passes
 * (src/wfa2_utils.rs:159-168: -1,3,1,2,60,0 and :183-189: -1,3,2). */
typedef struct {
  int32_t metric;
  int32_t match, mismatch, gap_open1, gap_ext1, gap_open2, gap_ext2;
} sfc_penalties;

/* One (pattern, text) pair.  Offsets are byte offsets into seq_arena (forward ASCII).
 * The four ends-free allowances are what rust-wfa2's align_ends_free takes
 * (src/wfa2_utils.rs:116,130); all zero == align_end_to_end (src/wfa2_utils.rs:135).
 * flags bit0: reverse both sequences before aligning (sawfish convention, wfa2_utils.rs:198-215). */
typedef struct {
  uint64_t pattern_off, text_off;
  uint32_t pattern_len, text_len;
  int32_t pattern_begin_free, pattern_end_free, text_begin_free, text_end_free;
  uint32_t flags;
} sfc_pair;
#define SFC_PAIR_REVERSE 1u

/* Run-length WFA ops: (run_length << 2) | code, in WFA2's cigar order (begin -> end of the
 * sequences as aligned, i.e. of the REVERSED sequences when SFC_PAIR_REVERSE is set).
 * 'I' consumes text, 'D' consumes pattern, as in WFA2-lib. */
enum { SFC_OP_M = 0, SFC_OP_X = 1, SFC_OP_I = 2, SFC_OP_D = 3 };

typedef struct {
  double h2d_ms, pack_ms, kernel_ms, d2h_ms; /* CUDA-event times of the last batch on the ctx stream */
  uint64_t h2d_bytes, d2h_bytes;
  uint64_t cells;            /* M-wavefront cells computed by the kernels (device counters) */
  uint32_t launches;         /* kernels launched for the last batch */
  uint32_t retries;          /* workspace-escalation passes */
  uint32_t n_score_kernel;   /* pairs served by the score-only kernel (K1) */
  uint32_t n_align_kernel;   /* pairs served by the backtrace kernel (K2) */
} sfc_batch_stats;

/* ---- context ---- */
int sfc_abi_version(void);
int sfc_device_count(void);
int sfc_create(int device_ordinal, sfc_ctx** out);
void sfc_destroy(sfc_ctx*);
const char* sfc_last_error(const sfc_ctx*);
/* Upper bound on device workspace the ctx may allocate (bytes); 0 = 80% of free memory. */
int sfc_set_workspace_limit(sfc_ctx*, size_t bytes);

/* ---- the batched hot path ---- */
/* One-shot: upload + pack + align + download.  scores/status: [n_pairs].
 * want_cigar == 0: ops_* may be NULL.  Otherwise ops_off is [n_pairs+1] and pair i's runs are
 * ops_rle[ops_off[i] .. ops_off[i+1]); SFC_ERR_CAPACITY if ops_cap (in uint32 entries) is too small
 * (ops_off[n_pairs] then holds the required size). */
int sfc_align_batch(sfc_ctx*, const sfc_penalties*, const uint8_t* seq_arena, size_t arena_len,
                    const sfc_pair* pairs, size_t n_pairs, int want_cigar,
                    int32_t* scores, int32_t* status,
                    uint32_t* ops_rle, size_t ops_cap, uint64_t* ops_off);

/* Staged form of the same call (what a pipelined submitter thread uses; also how bench.py separates
 * device-resident throughput from end-to-end throughput).
 * Buffer lifetime: seq_arena and pairs are only read inside sfc_batch_upload — when it returns (OK or error) the
 * copy of the arena to the device has completed and the pair table has been copied to library-owned memory, so
 * the caller may refill or free both immediately (a pinned arena overlaps its copy with the validation of the
 * pair table inside the call).  The output buffers of sfc_batch_download are written before it returns. */
int sfc_batch_upload(sfc_ctx*, const uint8_t* seq_arena, size_t arena_len, const sfc_pair* pairs, size_t n_pairs);
int sfc_batch_run(sfc_ctx*, const sfc_penalties*, int want_cigar);
int sfc_batch_download(sfc_ctx*, int32_t* scores, int32_t* status,
                       uint32_t* ops_rle, size_t ops_cap, uint64_t* ops_off);
int sfc_batch_get_stats(const sfc_ctx*, sfc_batch_stats* out);

/* ---- CIGAR helpers (host, exact restatements of the reference's host code) ---- */
/* Expand run-length ops to WFA2's byte string ('M','X','I','D'). Returns bytes written or <0. */
int64_t sfc_expand_ops(const uint32_t* ops_rle, size_t n_runs, uint8_t* wfa_ops, size_t cap);
/* == process_wfa2_cigar_string (src/wfa2_utils.rs:13-95): htslib-encoded cigar (len<<4|op). */
int sfc_translate_cigar(const uint8_t* wfa_ops, size_t n_ops, int64_t* ref_offset,
                        uint32_t* cigar, size_t cigar_cap, size_t* n_cigar, int* has_match);
/* Same, straight from run-length ops (no expansion). */
int sfc_translate_rle(const uint32_t* ops_rle, size_t n_runs, int64_t* ref_offset,
                      uint32_t* cigar, size_t cigar_cap, size_t* n_cigar, int* has_match);

/* ---- per-pair convenience mirroring wfa2_utils::PairwiseAligner (n = 1 batches) ---- */
typedef struct sfc_aligner sfc_aligner;
int sfc_aligner_new_consensus(sfc_ctx*, sfc_aligner** out);                       /* wfa2_utils.rs:182-196 */
int sfc_aligner_new_large_indel(sfc_ctx*, int is_long_contig, sfc_aligner** out); /* wfa2_utils.rs:149-175 */
void sfc_aligner_free(sfc_aligner*);
int sfc_aligner_set_text(sfc_aligner*, const uint8_t* text, size_t len);          /* wfa2_utils.rs:198-203 */
/* mode 0 = align (clip 0.4 rule, wfa2_utils.rs:106-118), 1 = align_clip_frac (:120-132), 2 = align_end_to_end (:134-140).
 * Outputs the post-translation result: score, has_alignment, ref_offset, cigar. */
int sfc_aligner_align(sfc_aligner*, const uint8_t* pattern, size_t len, int mode,
                      double text_clip_frac, double pattern_clip_frac,
                      int32_t* score, int* has_alignment, int64_t* ref_offset,
                      uint32_t* cigar, size_t cigar_cap, size_t* n_cigar);

/* Debug view of the LAST sfc_aligner_align call: gapped pattern, '|' under matches, gapped text (wfa2_utils.rs:253-273:
 * PairwiseAligner::matching / print_pairwise_alignment), in the order the sequences were aligned (both reversed).
 * *len = characters per line; SFC_ERR_CAPACITY (with *len set) when cap is too small. */
int sfc_aligner_matching(sfc_aligner*, char* pattern_line, char* mid_line, char* text_line, size_t cap, size_t* len);

/* ---- multi-GPU sharding helpers (host only; no collective: batches are independent) ----
 * The unit of sharding is the SV group / breakpoint cluster, the reference's own rayon task granularity
 * (src/joint_call/joint_call_all_samples.rs:127-142, src/refine_sv/mod.rs:861-887). */
/* Estimated alignment work of each pair (wavefront cells), the weight both the sharder and the in-batch scheduler use. */
int sfc_estimate_pair_work(const sfc_penalties*, const sfc_pair* pairs, size_t n_pairs, uint64_t* work);
/* Greedy longest-processing-time assignment of groups to shards (deterministic). shard_of_group: [n_groups]. */
int sfc_shard_plan(const uint64_t* group_work, size_t n_groups, int n_shards, uint32_t* shard_of_group);

/* ---- one process, several GPUs: the deployment shape of the reference (one rayon pool, no MPI) ----
 * sfc_multi owns submitter threads and contexts per device: two of each for score-only gap-linear batches (the
 * score_sv shape: one sub-shard's host staging overlaps the other's kernels on the same GPU; SFC_MULTI_SPLIT=1..4
 * changes the count), one otherwise.  A batch is split by SV group with sfc_shard_plan
 * (group_of_pair == NULL: every pair is its own group), each shard gets a compacted copy of just its sequences,
 * shards run concurrently on their GPUs, and results are gathered on the host in the caller's pair order.
 * No collective, no peer access: batches are independent (north_star). Replaces the fan-out / mpsc gather of
 * src/joint_call/joint_call_all_samples.rs:100-149 and src/refine_sv/mod.rs:827-896 for the alignment work. */
typedef struct sfc_multi sfc_multi;
int sfc_multi_create(const int* device_ordinals, int n_devices, sfc_multi** out);
void sfc_multi_destroy(sfc_multi*);
int sfc_multi_device_count(const sfc_multi*);
const char* sfc_multi_last_error(const sfc_multi*);
int sfc_multi_align_batch(sfc_multi*, const sfc_penalties*, const uint8_t* seq_arena, size_t arena_len,
                          const sfc_pair* pairs, size_t n_pairs, const uint32_t* group_of_pair,
                          int want_cigar, int32_t* scores, int32_t* status,
                          uint32_t* ops_rle, size_t ops_cap, uint64_t* ops_off);
/* kernel milliseconds of the last batch on each device, summed over the device's contexts (their kernels share the
 * GPU, so this bounds the time the device was busy from above). out: [n_devices] */
int sfc_multi_last_kernel_ms(const sfc_multi*, double* out);

/* ---- measurement helper: integer-ALU roofline denominator (IADD3/LOP3 issue rate) ---- */
/* Runs a dependent-chain-free INT32 micro-benchmark and returns giga-ops/s (thread-level ops). */
int sfc_measure_int_peak(sfc_ctx*, double* gops_per_s, double* sm_clock_mhz_est);

/* ---- affine-gap alignment with clipping: the rust-bio aligner behind bio_align_utils::PairwiseAligner ----
 * Replaces bio::alignment::pairwise::banded::Aligner::custom_with_matches as called at
 * reference src/bio_align_utils.rs:170-188 (PairwiseAligner::align / align_with_matches).  Scoring follows
 * bio::alignment::pairwise::Scoring: gap_open / gap_extend <= 0, match, mismatch, and the four clip penalties
 * (SFC_BIO_MIN_SCORE = clipping disabled); e.g. new_merge_aligner (:104-112) is {0,-2,1,-3, MIN,MIN,0,0}.
 * x = pattern (globally aligned unless x-clips are enabled), y = text.  The reference's k-mer band is not applied:
 * the whole matrix is filled (same result whenever the band contains the optimal path). */
#define SFC_BIO_MIN_SCORE (-858993459)
typedef struct {
  int32_t gap_open, gap_extend, match, mismatch;
  int32_t xclip_prefix, xclip_suffix, yclip_prefix, yclip_suffix;
} sfc_bio_scoring;
enum { SFC_BIO_MATCH = 0, SFC_BIO_SUBST = 1, SFC_BIO_INS = 2, SFC_BIO_DEL = 3, SFC_BIO_XCLIP = 4, SFC_BIO_YCLIP = 5 };
/* pairs: pattern_off/len = x, text_off/len = y (byte offsets into seq_arena; the free-end and flag fields are ignored).
 * scores[n], status[n] (0 = ok, -6 = pair does not fit the device workspace), bounds[4n] = xstart, xend, ystart, yend
 * (rust-bio Alignment fields), ops: run-length words (len << 3 | SFC_BIO_*) in alignment order, clip operations
 * carry the clipped length; ops_off[n+1] delimits each pair's words.  ops / ops_off may be NULL (score only). */
int sfc_bio_align_batch(sfc_ctx*, const sfc_bio_scoring*, const uint8_t* seq_arena, size_t arena_len, const sfc_pair* pairs,
                        size_t n_pairs, int32_t* scores, int32_t* status, int32_t* bounds, uint32_t* ops, size_t ops_cap,
                        uint64_t* ops_off);

/* Score-only form of the same aligner (K5: no traceback matrix, one warp per pair, rows in registers).  This is what
 * the merge aligner's callers consume: reference src/joint_call/merge_haplotypes/single_region/merge_pools.rs:253-257
 * and multi_region/merge_pools.rs:57 keep only `aligner.align(base, query)`'s score (norm score >= 0.97).
 * scores[i] is identical to what sfc_bio_align_batch returns for the pair.  Scorings with an enabled x-suffix clip
 * (new_pattern_suffix_aligner, src/bio_align_utils.rs:139-146) are served by the K4 kernel behind the same call.
 * sfc_batch_get_stats afterwards reports this call (cells = sum |x|*|y|). */
int sfc_bio_score_batch(sfc_ctx*, const sfc_bio_scoring*, const uint8_t* seq_arena, size_t arena_len, const sfc_pair* pairs,
                        size_t n_pairs, int32_t* scores, int32_t* status);

#ifdef __cplusplus
}
#endif
#endif
