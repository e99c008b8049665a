//! Source patch for sawfish `src/score_sv/mod.rs`: `get_read_support_scores_from_segment` (:758-905) in two phases.
//! (Source only: no Rust toolchain in the build image.  It uses nothing but `sawfish_cuda::{Ctx, SfcPair}` and the
//! reference's own types; `tests/test_abi.py` checks that every `sfc_*` symbol the crate binds is declared by the header.)
//!
//! The reference builds a consensus aligner per read x breakend x registration and aligns the ref / alt / overlapping
//! haplotype flanks one by one.  Which slots get a score depends on geometry only: the flank exists and
//! `start_truncation + end_truncation + min_read_support_flank_size <= flank.len()` (:721-745); the early `continue` on
//! `is_any_condition_scored()` (:815-824) asks whether a slot of an earlier record of the same read was *filled*, which
//! the enumeration knows without any score.  So the record loop only ENUMERATES slots into a `ScoreBatch`; one
//! `Ctx::score_batch` call per SV group (or per chunk of groups) computes every score on the GPU; `scatter` writes
//! `score as f64 / text.len() as f64` (:702-703) into `LocusSupportScores`.  `add_guest_haplotypes` (:2293), the first
//! step that reads score VALUES, runs after the scatter exactly as before.
use sawfish_cuda::{Ctx, SfcPair, SFC_PAIR_REVERSE};

#[derive(Clone, Copy, PartialEq)]
pub enum SlotAllele { Ref, Alt, Overlap(usize) }

pub struct ScoreSlot {
    pub qname: String,
    pub breakend_index: usize,
    pub registration_index: usize,
    pub allele: SlotAllele,
    pub text_len: usize,
}

#[derive(Default)]
pub struct ScoreBatch {
    arena: Vec<u8>,
    pairs: Vec<SfcPair>,
    pub slots: Vec<ScoreSlot>,
}

fn sawfish_clip(pattern_len: usize, text_len: usize) -> i32 {
    // src/wfa2_utils.rs:112-115
    std::cmp::min((pattern_len as f64 * 0.4) as i32, (text_len as f64 * 0.4) as i32)
}

impl ScoreBatch {
    /// == the body of the registration loop (:829-901) minus the alignments.  Returns true if the ref slot will be
    /// scored (what `is_any_condition_scored()` reports for later records of the same read).
    #[allow(clippy::too_many_arguments)]
    pub fn add_read_flank(
        &mut self,
        qname: &str,
        breakend_index: usize,
        registration_index: usize,
        read_breakend_flank: &[u8],
        start_truncation: usize,
        end_truncation: usize,
        min_read_support_flank_size: usize,
        ref_flank: Option<&[u8]>,
        alt_flank: Option<&[u8]>,
        overlap_flanks: &[Option<&[u8]>],
    ) -> bool {
        assert!(!read_breakend_flank.is_empty()); // set_text, src/wfa2_utils.rs:199
        let mut text_off: Option<u64> = None;
        let mut ref_scored = false;
        let mut one = |this: &mut Self, allele: SlotAllele, seq: Option<&[u8]>| -> bool {
            let seq = match seq { Some(x) => x, None => return false };
            if start_truncation + end_truncation + min_read_support_flank_size > seq.len() { return false; } // :732-736
            let seq = &seq[start_truncation..seq.len() - end_truncation];
            let t_off = *text_off.get_or_insert_with(|| {
                let o = this.arena.len() as u64;
                this.arena.extend_from_slice(read_breakend_flank);
                o
            });
            let p_off = this.arena.len() as u64;
            this.arena.extend_from_slice(seq);
            let c = sawfish_clip(seq.len(), read_breakend_flank.len());
            this.pairs.push(SfcPair {
                pattern_off: p_off, text_off: t_off, pattern_len: seq.len() as u32, text_len: read_breakend_flank.len() as u32,
                pattern_begin_free: c, pattern_end_free: c, text_begin_free: c, text_end_free: c, flags: SFC_PAIR_REVERSE,
            });
            this.slots.push(ScoreSlot { qname: qname.to_string(), breakend_index, registration_index, allele, text_len: read_breakend_flank.len() });
            true
        };
        ref_scored |= one(self, SlotAllele::Ref, ref_flank);
        one(self, SlotAllele::Alt, alt_flank);
        for (i, f) in overlap_flanks.iter().enumerate() {
            one(self, SlotAllele::Overlap(i), *f);
        }
        ref_scored
    }

    /// Phase 2: every enumerated slot in one GPU batch (consensus preset, score only -> kernel K1).
    pub fn score(&self, ctx: &mut Ctx) -> Vec<i32> {
        if self.pairs.is_empty() { return Vec::new(); }
        ctx.score_batch(&self.arena, &self.pairs)
    }

    /// Phase 3: `norm_score` of slot i (:702-703); the caller writes it where the reference wrote `ref_score` /
    /// `alt_score` / `overlap_hap_score` (:848-899), keyed by `slots[i]`.
    pub fn norm_score(&self, scores: &[i32], i: usize) -> f64 {
        scores[i] as f64 / self.slots[i].text_len as f64
    }
}

// In get_read_support_scores_from_segment the record loop keeps its filters (:784-797) and get_read_flanks_for_breakpoint
// (:799-810) and replaces :838-901 by
//
//     let scored = batch.add_read_flank(&qname, breakend_index, registration_index, read_breakend_flank,
//         breakend_flank.start_truncation, breakend_flank.end_truncation, score_settings.min_read_support_flank_size,
//         ref_flank_seqs.breakend_flanks[breakend_index][standard_registration_index],
//         contig_flank_seqs.breakend_flanks[breakend_index][registration_index], &overlap_flanks);
//     will_be_scored[(qname.clone(), breakend_index)] |= scored;      // stands in for is_any_condition_scored() at :820
//
// and score_refined_sv_group (:2230-2330) calls `let s = batch.score(ctx);` once per group, scatters, then continues with
// add_guest_haplotypes / get_allele_depth as before.  joint_genotype_all_samples (src/joint_call/joint_call_all_samples.rs:
// 100-149) gives each rayon worker one `Ctx` (thread-local, see wfa2_utils_drop_in.rs) on the GPU its SV groups were
// assigned to by `sfc_shard_plan`.
