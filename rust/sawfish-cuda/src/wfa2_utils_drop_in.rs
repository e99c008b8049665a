//! Body swap for sawfish's `src/wfa2_utils.rs` `PairwiseAligner` (public API unchanged, :143-256).
//! In the sawfish tree `SimpleAlignment` / `Cigar` come from `crate::simple_alignment` / `rust_htslib`; here the
//! alignment is returned as (ref_offset, htslib-encoded cigar words `len << 4 | op`), which is what
//! `rust_htslib::bam::record::Cigar` values are built from.
use crate::*;
use std::ptr::null_mut;

pub struct PairwiseAligner { h: *mut SfcAligner, text: Vec<u8> }

impl PairwiseAligner {
    /// :149-175 (MemoryMed vs MemoryHigh is a WFA2-lib memory model; the GPU kernel has one)
    pub fn new_large_indel_aligner(ctx: &Ctx, is_long_contig: bool) -> Self {
        let mut h = null_mut();
        assert_eq!(unsafe { sfc_aligner_new_large_indel(ctx.raw(), is_long_contig as i32, &mut h) }, 0);
        Self { h, text: Vec::new() }
    }
    /// :182-196
    pub fn new_consensus_aligner(ctx: &Ctx) -> Self {
        let mut h = null_mut();
        assert_eq!(unsafe { sfc_aligner_new_consensus(ctx.raw(), &mut h) }, 0);
        Self { h, text: Vec::new() }
    }
    /// :198-203 — the reference stores the reversed text; `text()` keeps returning that
    pub fn set_text(&mut self, text: &[u8]) {
        assert!(!text.is_empty());
        self.text = text.iter().rev().copied().collect();
        assert_eq!(unsafe { sfc_aligner_set_text(self.h, text.as_ptr(), text.len()) }, 0);
    }
    /// :205-207
    pub fn text(&self) -> &[u8] { &self.text }

    fn run(&mut self, pattern: &[u8], mode: i32, text_clip: f64, pattern_clip: f64) -> (i32, Option<(i64, Vec<u32>)>) {
        assert!(!self.text.is_empty());
        assert!(!pattern.is_empty());
        let (mut score, mut has, mut roff, mut n) = (0i32, 0i32, 0i64, 0usize);
        let mut cigar = vec![0u32; pattern.len() + self.text.len() + 8];
        let rc = unsafe { sfc_aligner_align(self.h, pattern.as_ptr(), pattern.len(), mode, text_clip, pattern_clip,
                                            &mut score, &mut has, &mut roff, cigar.as_mut_ptr(), cigar.len(), &mut n) };
        assert_eq!(rc, 0);  // == assert_eq!(status, AlignmentStatus::StatusAlgCompleted), :117
        cigar.truncate(n);
        (score, (has != 0).then_some((roff, cigar)))
    }
    /// :209-219 (clip 0.4 rule of align_ends_free, :106-118)
    pub fn align(&mut self, pattern: &[u8]) -> (i32, Option<(i64, Vec<u32>)>) { self.run(pattern, 0, 0.0, 0.0) }
    /// :221-239
    pub fn align_clip_frac(&mut self, pattern: &[u8], text_clip_frac: f64, pattern_clip_frac: f64) -> (i32, Option<(i64, Vec<u32>)>) {
        self.run(pattern, 1, text_clip_frac, pattern_clip_frac)
    }
    /// :241-251
    pub fn align_end_to_end(&mut self, pattern: &[u8]) -> (i32, Option<(i64, Vec<u32>)>) { self.run(pattern, 2, 0.0, 0.0) }
}

impl Drop for PairwiseAligner {
    fn drop(&mut self) { unsafe { sfc_aligner_free(self.h) } }
}
