//! Source patch for sawfish `src/refine_sv/align/mod.rs`: the contig alignments of `align_single_ref_region_assemblies`
//! (:373-608) and `align_multi_ref_region_assemblies` (:1975-2199) as one GPU batch per chunk of breakpoint clusters.
//! (Source only, see score_sv_two_phase.rs.)
//!
//! `refine_sv_candidates` (src/refine_sv/mod.rs:781-896) runs assemble -> align per cluster inside a rayon task.  The
//! alignment of a contig needs nothing but the contig and the cluster's text (reference window, or the two-region
//! alt-haplotype model of get_two_region_alt_hap_info, :1125-1330), so a worker first assembles a chunk of clusters,
//! then enumerates every (text, contig) of the chunk into a `ContigBatch`, aligns them in ONE `Ctx::align_batch`
//! (large-indel preset, CIGAR), and only then runs the reference's post-alignment code per cluster, reading
//! `(score, Option<SimpleAlignment>)` from the batch where it used to call `aligner.align(&contig.seq)` (:439, :2029).
//! The two-region loop tries the 1500-flank contig of an assembly and falls back to the 2500-flank one when a
//! post-alignment filter rejects it (:2022-2179); both are enumerated, the retry is read only if needed.
use sawfish_cuda::{Ctx, SfcPair, SFC_PAIR_REVERSE};

#[derive(Default)]
pub struct ContigBatch {
    arena: Vec<u8>,
    pairs: Vec<SfcPair>,
    text: (u64, usize),
}

impl ContigBatch {
    /// aligner.set_text(..): the text of every contig added until the next call (:404, :2003)
    pub fn set_text(&mut self, text: &[u8]) {
        assert!(!text.is_empty());
        self.text = (self.arena.len() as u64, text.len());
        self.arena.extend_from_slice(text);
    }
    /// aligner.align(&contig): returns the index to read the result back with
    pub fn add_contig(&mut self, contig: &[u8]) -> usize {
        assert!(self.text.1 > 0 && !contig.is_empty());
        let c = std::cmp::min((contig.len() as f64 * 0.4) as i32, (self.text.1 as f64 * 0.4) as i32); // src/wfa2_utils.rs:112-115
        let p_off = self.arena.len() as u64;
        self.arena.extend_from_slice(contig);
        self.pairs.push(SfcPair {
            pattern_off: p_off, text_off: self.text.0, pattern_len: contig.len() as u32, text_len: self.text.1 as u32,
            pattern_begin_free: c, pattern_end_free: c, text_begin_free: c, text_end_free: c, flags: SFC_PAIR_REVERSE,
        });
        self.pairs.len() - 1
    }
    /// One batch: per contig (score, Some((ref_offset, htslib cigar words))) or None when the alignment has no match
    /// state — what `process_wfa2_alignment` returns (src/wfa2_utils.rs:97-104).
    pub fn align(&self, ctx: &mut Ctx) -> Vec<(i32, Option<(i64, Vec<u32>)>)> {
        if self.pairs.is_empty() { return Vec::new(); }
        ctx.align_batch(&self.arena, &self.pairs)
    }
}

// align_multi_ref_region_assemblies after the patch (control flow of :2011-2179 unchanged):
//
//     for (assembly_index, assembly) in assemblies.iter().enumerate() {
//         if assembly.supporting_read_count() < refine_settings.min_assembly_read_support { continue; }
//         for (contig_index, assembly_contig) in assembly.contigs.iter().enumerate() {
//             let (_score, mut alignment) = match results[pair_index[&(cluster_index, assembly_index, contig_index)]].clone() {
//                 (score, Some((ref_offset, cigar))) => (score, SimpleAlignment { ref_offset, cigar: cigar_words_to_htslib(&cigar) }),
//                 (_, None) => continue,
//             };
//             ... collapse_indels_at_breakpoint, transform_alt_hap_alignment_into_left_right_components, the quality and
//             clipping filters, get_two_region_refined_sv_candidate, update_two_ref_segment_contig_alignments ...
//             break;   // :2176
//         }
//     }
