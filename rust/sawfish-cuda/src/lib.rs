//! Thin FFI over `include/sawfish_cuda.h` (ABI version 1) plus the safe batch wrappers sawfish's callers use.
//!
//! Replaces, in the sawfish tree: `rust_wfa2::aligner::*` behind `src/wfa2_utils.rs` and the rust-bio aligner behind
//! `src/bio_align_utils.rs`.  `src/wfa2_utils_drop_in.rs` in this crate is the body swap for `PairwiseAligner`.
//! One `Ctx` per (host thread, GPU); the library keeps no global state.  Errors follow the reference's convention:
//! the callers `assert!` on the return code (`panic = "abort"`, sawfish Cargo.toml:10-11).
#![allow(clippy::too_many_arguments)]
use std::ffi::CStr;
use std::os::raw::c_char;
use std::ptr::null_mut;

pub mod wfa2_utils_drop_in;

#[repr(C)] pub struct SfcCtx { _p: [u8; 0] }
#[repr(C)] pub struct SfcAligner { _p: [u8; 0] }

/// metric 0 = gap-linear (WFAlignerGapLinear::new_with_match), 1 = gap-affine 2-piece (WFAlignerGapAffine2Pieces).
#[repr(C)] #[derive(Clone, Copy, Debug)]
pub struct SfcPenalties { pub metric: i32, pub match_: i32, pub mismatch: i32,
                          pub gap_open1: i32, pub gap_ext1: i32, pub gap_open2: i32, pub gap_ext2: i32 }
/// src/wfa2_utils.rs:183-189
pub const CONSENSUS: SfcPenalties = SfcPenalties { metric: 0, match_: -1, mismatch: 3, gap_open1: 0, gap_ext1: 2, gap_open2: 0, gap_ext2: 0 };
/// src/wfa2_utils.rs:159-168
pub const LARGE_INDEL: SfcPenalties = SfcPenalties { metric: 1, match_: -1, mismatch: 3, gap_open1: 1, gap_ext1: 2, gap_open2: 60, gap_ext2: 0 };

#[repr(C)] #[derive(Clone, Copy, Default, Debug)]
pub struct SfcPair { pub pattern_off: u64, pub text_off: u64, pub pattern_len: u32, pub text_len: u32,
                     pub pattern_begin_free: i32, pub pattern_end_free: i32,
                     pub text_begin_free: i32, pub text_end_free: i32, pub flags: u32 }
pub const SFC_PAIR_REVERSE: u32 = 1;

#[repr(C)] #[derive(Clone, Copy, Debug)]
pub struct SfcBioScoring { pub gap_open: i32, pub gap_extend: i32, pub match_: i32, pub mismatch: i32,
                           pub xclip_prefix: i32, pub xclip_suffix: i32, pub yclip_prefix: i32, pub yclip_suffix: i32 }
/// bio::alignment::pairwise::MIN_SCORE: clipping disabled
pub const SFC_BIO_MIN_SCORE: i32 = -858993459;
/// PairwiseAligner::new_merge_aligner, src/bio_align_utils.rs:104-112
pub const MERGE_SCORING: SfcBioScoring = SfcBioScoring { gap_open: 0, gap_extend: -2, match_: 1, mismatch: -3,
    xclip_prefix: SFC_BIO_MIN_SCORE, xclip_suffix: SFC_BIO_MIN_SCORE, yclip_prefix: 0, yclip_suffix: 0 };

extern "C" {
    pub fn sfc_abi_version() -> i32;
    pub fn sfc_device_count() -> i32;
    pub fn sfc_create(device_ordinal: i32, out: *mut *mut SfcCtx) -> i32;
    pub fn sfc_destroy(ctx: *mut SfcCtx);
    pub fn sfc_last_error(ctx: *const SfcCtx) -> *const c_char;
    pub fn sfc_align_batch(ctx: *mut SfcCtx, pen: *const SfcPenalties, arena: *const u8, arena_len: usize,
                           pairs: *const SfcPair, n_pairs: usize, want_cigar: i32, scores: *mut i32, status: *mut i32,
                           ops_rle: *mut u32, ops_cap: usize, ops_off: *mut u64) -> i32;
    pub fn sfc_batch_upload(ctx: *mut SfcCtx, arena: *const u8, arena_len: usize, pairs: *const SfcPair, n: usize) -> i32;
    pub fn sfc_batch_run(ctx: *mut SfcCtx, pen: *const SfcPenalties, want_cigar: i32) -> i32;
    pub fn sfc_batch_download(ctx: *mut SfcCtx, scores: *mut i32, status: *mut i32,
                              ops_rle: *mut u32, ops_cap: usize, ops_off: *mut u64) -> i32;
    pub fn sfc_translate_rle(ops: *const u32, n_runs: usize, ref_offset: *mut i64,
                             cigar: *mut u32, cigar_cap: usize, n_cigar: *mut usize, has_match: *mut i32) -> i32;
    pub fn sfc_aligner_new_consensus(ctx: *mut SfcCtx, out: *mut *mut SfcAligner) -> i32;
    pub fn sfc_aligner_new_large_indel(ctx: *mut SfcCtx, is_long_contig: i32, out: *mut *mut SfcAligner) -> i32;
    pub fn sfc_aligner_free(a: *mut SfcAligner);
    pub fn sfc_aligner_set_text(a: *mut SfcAligner, text: *const u8, len: usize) -> i32;
    pub fn sfc_aligner_align(a: *mut SfcAligner, pattern: *const u8, len: usize, mode: i32,
                             text_clip_frac: f64, pattern_clip_frac: f64, score: *mut i32, has_alignment: *mut i32,
                             ref_offset: *mut i64, cigar: *mut u32, cigar_cap: usize, n_cigar: *mut usize) -> i32;
    pub fn sfc_aligner_matching(a: *mut SfcAligner, pattern_line: *mut c_char, mid_line: *mut c_char, text_line: *mut c_char,
                                cap: usize, len: *mut usize) -> i32;
    pub fn sfc_estimate_pair_work(pen: *const SfcPenalties, pairs: *const SfcPair, n: usize, work: *mut u64) -> i32;
    pub fn sfc_shard_plan(group_work: *const u64, n_groups: usize, n_shards: i32, shard_of_group: *mut u32) -> i32;
    pub fn sfc_bio_align_batch(ctx: *mut SfcCtx, scoring: *const SfcBioScoring, arena: *const u8, arena_len: usize,
                               pairs: *const SfcPair, n: usize, scores: *mut i32, status: *mut i32, bounds: *mut i32,
                               ops: *mut u32, ops_cap: usize, ops_off: *mut u64) -> i32;
    pub fn sfc_bio_score_batch(ctx: *mut SfcCtx, scoring: *const SfcBioScoring, arena: *const u8, arena_len: usize,
                               pairs: *const SfcPair, n: usize, scores: *mut i32, status: *mut i32) -> i32;
}

/// Owns one `sfc_ctx` (device buffers + stream of one GPU).  Not `Sync`: one per worker thread.
pub struct Ctx { raw: *mut SfcCtx }
unsafe impl Send for Ctx {}

impl Ctx {
    pub fn new(device_ordinal: i32) -> Self {
        let mut raw = null_mut();
        let rc = unsafe { sfc_create(device_ordinal, &mut raw) };
        assert_eq!(rc, 0, "sfc_create({device_ordinal}) failed with {rc}: no B200, no fallback");
        Self { raw }
    }
    pub fn raw(&self) -> *mut SfcCtx { self.raw }
    pub fn last_error(&self) -> String {
        unsafe { CStr::from_ptr(sfc_last_error(self.raw)) }.to_string_lossy().into_owned()
    }
    fn check(&self, rc: i32) { assert_eq!(rc, 0, "libsawfish_cuda: {}", self.last_error()); }

    /// score_sv's batch (src/score_sv/mod.rs:837-901 turned into one call): consensus preset, scores only.
    pub fn score_batch(&mut self, arena: &[u8], pairs: &[SfcPair]) -> Vec<i32> {
        let (mut scores, mut status) = (vec![0i32; pairs.len()], vec![0i32; pairs.len()]);
        let rc = unsafe { sfc_align_batch(self.raw, &CONSENSUS, arena.as_ptr(), arena.len(), pairs.as_ptr(), pairs.len(), 0,
                                          scores.as_mut_ptr(), status.as_mut_ptr(), null_mut(), 0, null_mut()) };
        self.check(rc);
        assert!(status.iter().all(|s| *s == 0));  // == assert_eq!(status, StatusAlgCompleted), src/wfa2_utils.rs:117
        scores
    }

    /// refine_sv's batch (src/refine_sv/align/mod.rs:439-451,2027-2049): large-indel preset with CIGAR.
    /// Returns per pair (score, ref_offset, htslib-encoded cigar words) with `None` where the alignment has no match
    /// state (process_wfa2_cigar_string, src/wfa2_utils.rs:13-95).
    pub fn align_batch(&mut self, arena: &[u8], pairs: &[SfcPair]) -> Vec<(i32, Option<(i64, Vec<u32>)>)> {
        let n = pairs.len();
        let cap: usize = pairs.iter().map(|p| (p.pattern_len + p.text_len) as usize + 4).sum();
        let (mut scores, mut status) = (vec![0i32; n], vec![0i32; n]);
        let (mut ops, mut off) = (vec![0u32; cap], vec![0u64; n + 1]);
        let rc = unsafe { sfc_align_batch(self.raw, &LARGE_INDEL, arena.as_ptr(), arena.len(), pairs.as_ptr(), n, 1,
                                          scores.as_mut_ptr(), status.as_mut_ptr(), ops.as_mut_ptr(), cap, off.as_mut_ptr()) };
        self.check(rc);
        (0..n).map(|i| {
            assert_eq!(status[i], 0);
            let runs = &ops[off[i] as usize..off[i + 1] as usize];
            let mut cigar = vec![0u32; runs.len() + 2];
            let (mut roff, mut nc, mut has) = (0i64, 0usize, 0i32);
            let rc = unsafe { sfc_translate_rle(runs.as_ptr(), runs.len(), &mut roff, cigar.as_mut_ptr(), cigar.len(), &mut nc, &mut has) };
            assert_eq!(rc, 0);
            cigar.truncate(nc);
            (scores[i], (has != 0).then_some((roff, cigar)))
        }).collect()
    }

    /// Merge aligner scores for a pool of (base contig, query high-quality range) pairs
    /// (src/joint_call/merge_haplotypes/single_region/merge_pools.rs:253-257): x = query (pattern), y = base (text).
    pub fn merge_scores(&mut self, arena: &[u8], pairs: &[SfcPair]) -> Vec<i32> {
        let (mut scores, mut status) = (vec![0i32; pairs.len()], vec![0i32; pairs.len()]);
        let rc = unsafe { sfc_bio_score_batch(self.raw, &MERGE_SCORING, arena.as_ptr(), arena.len(), pairs.as_ptr(), pairs.len(),
                                              scores.as_mut_ptr(), status.as_mut_ptr()) };
        self.check(rc);
        assert!(status.iter().all(|s| *s == 0));
        scores
    }
}

impl Drop for Ctx {
    fn drop(&mut self) { unsafe { sfc_destroy(self.raw) } }
}

/// `clip = min((p.len() as f64 * 0.4) as i32, (t.len() as f64 * 0.4) as i32)`, src/wfa2_utils.rs:112-115
pub fn sawfish_clip(pattern_len: usize, text_len: usize) -> i32 {
    let frac = 0.4_f64;
    std::cmp::min((pattern_len as f64 * frac) as i32, (text_len as f64 * frac) as i32)
}

/// Greedy LPT assignment of SV groups to GPUs (src/joint_call/joint_call_all_samples.rs:127-142 fans groups out to a
/// rayon pool; here each group goes to one GPU's submitter thread).  `group_of_pair[i]` = group of pair i.
pub fn shard_groups(pairs: &[SfcPair], group_of_pair: &[u32], n_groups: usize, n_gpus: i32, pen: &SfcPenalties) -> Vec<u32> {
    let mut work = vec![0u64; pairs.len()];
    assert_eq!(unsafe { sfc_estimate_pair_work(pen, pairs.as_ptr(), pairs.len(), work.as_mut_ptr()) }, 0);
    let mut group_work = vec![0u64; n_groups];
    for (w, g) in work.iter().zip(group_of_pair) { group_work[*g as usize] += *w; }
    let mut shard = vec![0u32; n_groups];
    assert_eq!(unsafe { sfc_shard_plan(group_work.as_ptr(), n_groups, n_gpus, shard.as_mut_ptr()) }, 0);
    shard
}
