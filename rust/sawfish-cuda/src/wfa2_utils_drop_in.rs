//! Body swap for sawfish's `src/wfa2_utils.rs` `PairwiseAligner` (public API unchanged, :143-273): the same constructor
//! signatures (`new_large_indel_aligner(is_long_contig)`, `new_consensus_aligner()`), so the call sites at
//! `src/score_sv/mod.rs:837` and `src/refine_sv/align/mod.rs:392,1992` compile unchanged.  The GPU context the reference
//! has no notion of lives in a thread-local: one `Ctx` per rayon worker thread (the library's threading model), on the
//! GPU the worker was assigned round-robin at first use.
//! In the sawfish tree `SimpleAlignment` / `Cigar` come from `crate::simple_alignment` / `rust_htslib`; here the
//! alignment is returned as (ref_offset, htslib-encoded cigar words `len << 4 | op`), which is what
//! `rust_htslib::bam::record::Cigar` values are built from.
use crate::*;
use std::os::raw::c_char;
use std::ptr::null_mut;
use std::sync::atomic::{AtomicUsize, Ordering};

static NEXT_WORKER: AtomicUsize = AtomicUsize::new(0);

thread_local! {
    /// This worker thread's context: created at the first aligner construction on the thread.
    static THREAD_CTX: Ctx = {
        let n_dev = unsafe { sfc_device_count() }.max(1) as usize;
        Ctx::new((NEXT_WORKER.fetch_add(1, Ordering::Relaxed) % n_dev) as i32)
    };
}

pub struct PairwiseAligner { h: *mut SfcAligner, text: Vec<u8> }

impl PairwiseAligner {
    /// :149-175 (MemoryMed vs MemoryHigh is a WFA2-lib memory model; the GPU kernel has one)
    pub fn new_large_indel_aligner(is_long_contig: bool) -> Self {
        let mut h = null_mut();
        THREAD_CTX.with(|ctx| assert_eq!(unsafe { sfc_aligner_new_large_indel(ctx.raw(), is_long_contig as i32, &mut h) }, 0));
        Self { h, text: Vec::new() }
    }
    /// :182-196
    pub fn new_consensus_aligner() -> Self {
        let mut h = null_mut();
        THREAD_CTX.with(|ctx| assert_eq!(unsafe { sfc_aligner_new_consensus(ctx.raw(), &mut h) }, 0));
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

    /// :253-255 — the three display lines of the last alignment (gapped pattern, '|' under matches, gapped text)
    pub fn matching(&self, _text: &[u8], _pattern: &[u8]) -> (String, String, String) {
        let mut n = 0usize;
        unsafe { sfc_aligner_matching(self.h, null_mut(), null_mut(), null_mut(), 0, &mut n) };
        let mut lines = [vec![0u8; n], vec![0u8; n], vec![0u8; n]];
        if n > 0 {
            let rc = unsafe { sfc_aligner_matching(self.h, lines[0].as_mut_ptr() as *mut c_char, lines[1].as_mut_ptr() as *mut c_char,
                                                   lines[2].as_mut_ptr() as *mut c_char, n, &mut n) };
            assert_eq!(rc, 0);
        }
        let [a, b, c] = lines.map(|l| String::from_utf8(l).unwrap());
        (a, b, c)
    }
}

/// :258-273 — in the sawfish tree this calls `crate::utils::pairwise_alignment_printer(150, &[&l3, &l1])`
pub fn print_pairwise_alignment(aligner: &PairwiseAligner, pattern: &[u8]) {
    let mut pattern = pattern.to_vec();
    pattern.reverse();
    let (l1, _, l3) = aligner.matching(aligner.text(), &pattern);
    let (mut l1, mut l3) = (l1.into_bytes(), l3.into_bytes());
    l1.reverse();
    l3.reverse();
    for (start, chunk) in l3.chunks(150).enumerate().map(|(i, c)| (i * 150, c)) {
        eprintln!("{}", std::str::from_utf8(chunk).unwrap());
        eprintln!("{}", std::str::from_utf8(&l1[start..start + chunk.len()]).unwrap());
        eprintln!();
    }
}

impl Drop for PairwiseAligner {
    fn drop(&mut self) { unsafe { sfc_aligner_free(self.h) } }
}
