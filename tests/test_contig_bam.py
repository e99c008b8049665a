"""contig.alignment.bam, the discover -> joint-call hand-off for assembled contigs (src/contig_output.rs:103-268 written,
src/joint_call/get_refined_svs.rs:38-195 read): round trip, BGZF/BAM structure, htslib-written files, and the hand-off
feeding the merge aligner's caller."""
import gzip
import os
import struct

import numpy as np
import pytest

from oracle import oracle as O
from sawfish_b200 import contig_bam as CB
from sawfish_b200 import merge_pools as MP

CHROMS = [("chr1", 248956422), ("chr20", 64444167)]


def _contigs(rng, n_clusters=5):
    acgt = np.frombuffer(b"ACGT", np.uint8)
    out = []
    for ci in range(n_clusters):
        for ai in range(int(rng.integers(1, 3))):
            L = int(rng.integers(301, 900))  # odd and even lengths: 4-bit packing
            seq = acgt[rng.integers(0, 4, L)].tobytes()
            hq = (int(rng.integers(0, 50)), L - int(rng.integers(0, 50)))
            nreads = int(rng.integers(2, 40))
            pos = int(rng.integers(1000, 60_000_000))
            if ci % 2 == 0:   # indel-like: one segment
                a = L // 2
                segs = [CB.ContigAlignmentSegment(1, pos, [("=", a), ("I", 37), ("=", L - a - 37)])]
            else:             # breakpoint: two segments, the second on the reverse strand for every other cluster
                a = L // 2
                segs = [CB.ContigAlignmentSegment(1, pos, [("=", a), ("S", L - a)]),
                        CB.ContigAlignmentSegment(0, pos + 5000, [("S", a), ("=", L - a)], is_fwd_strand=(ci % 4 == 1))]
            for sid in range(len(segs)):
                out.append(CB.ContigAlignmentInfo(0, ci, ai, seq, sid, segs, nreads, hq))
    return out


def test_round_trip_and_joint_call_view(tmp_path):
    rng = np.random.default_rng(7)
    contigs = _contigs(rng)
    path = str(tmp_path / CB.CONTIG_ALIGNMENT_FILENAME)
    CB.write_contig_alignments(path, CHROMS, contigs, cmdline="sawfish discover --bam x.bam")
    chroms, recs = CB.read_bam(path)
    assert chroms == CHROMS and len(recs) == len(contigs)
    keys = [(r.tid, r.pos) for r in recs]
    assert keys == sorted(keys)                                            # coordinate-sorted (:228-233)
    by_name = {}
    for r in recs:
        by_name.setdefault(r.qname, []).append(r)
    for ca in contigs:
        r = [x for x in by_name[f"sawfish:0:{ca.cluster_index}:{ca.assembly_index}"]
             if bool(x.flag & CB.BAM_FSUPPLEMENTARY) == (ca.segment_id != 0)][0]
        seg = ca.get_segment()
        assert (r.tid, r.pos, r.mapq, r.cigar, r.seq) == (seg.tid, seg.pos, 20, seg.cigar, ca.seq)
        assert bool(r.flag & CB.BAM_FREVERSE) == (not seg.is_fwd_strand)
        assert CB.get_supporting_read_count(r) == ca.supporting_read_count
        assert CB.get_high_quality_contig_range(r) == ca.high_quality_contig_range
        assert (CB.SA_AUX_TAG in r.aux) == (len(ca.segments) > 1)
    # joint-call's view: primary records only, breakend 2 rebuilt from the SA tag
    assemblies = CB.get_candidate_sv_contig_info(path)
    for ca in contigs:
        if ca.segment_id != 0:
            continue
        got = assemblies[ca.cluster_index][ca.assembly_index]
        assert got.supporting_read_count == ca.supporting_read_count
        assert len(got.contig_alignments) == len(ca.segments)
        p = got.contig_alignments[0]
        assert (p.contig_seq, p.chrom_index, p.contig_alignment.ref_offset, p.contig_alignment.cigar, p.is_fwd_strand) == \
               (ca.seq, ca.segments[0].tid, ca.segments[0].pos, ca.segments[0].cigar, True)
        assert p.high_quality_contig_range == ca.high_quality_contig_range
        if len(ca.segments) > 1:
            s, seg = got.contig_alignments[1], ca.segments[1]
            assert (s.chrom_index, s.contig_alignment.ref_offset, s.contig_alignment.cigar, s.is_fwd_strand) == \
                   (seg.tid, seg.pos, seg.cigar, seg.is_fwd_strand)
            if seg.is_fwd_strand:
                assert s.contig_seq == ca.seq and s.high_quality_contig_range == ca.high_quality_contig_range
            else:
                lo, hi = ca.high_quality_contig_range
                assert s.contig_seq == CB.rev_comp(ca.seq)
                assert s.high_quality_contig_range == (len(ca.seq) - hi, len(ca.seq) - lo)


def test_bgzf_structure(tmp_path):
    rng = np.random.default_rng(8)
    path = str(tmp_path / "c.bam")
    CB.write_contig_alignments(path, CHROMS, _contigs(rng, n_clusters=120))    # several BGZF blocks
    raw = open(path, "rb").read()
    assert raw.endswith(CB._BGZF_EOF)
    o, blocks, total = 0, 0, 0
    while o < len(raw):
        assert raw[o:o + 4] == b"\x1f\x8b\x08\x04" and raw[o + 12:o + 16] == b"BC\x02\x00"
        (bsize,) = struct.unpack_from("<H", raw, o + 16)
        (isize,) = struct.unpack_from("<I", raw, o + bsize + 1 - 4)
        assert isize <= 65536
        total += isize
        o += bsize + 1
        blocks += 1
    assert o == len(raw) and blocks > 3
    assert len(gzip.open(path, "rb").read()) == total
    assert CB.reg2bin(0, 1) == 4681 and CB.reg2bin(1 << 14, (1 << 14) + 1) == 4682 and CB.reg2bin(0, 1 << 29) == 0


def test_reads_htslib_written_bams():
    """The reference's own test BAMs (written by htslib) parse with the same reader.  They live under /root/reference,
    which exists only in the build container: skipped elsewhere."""
    d = "/root/reference/test_data"
    if not os.path.exists(os.path.join(d, "header_only.bam")):
        pytest.skip("reference test data not present")
    chroms, recs = CB.read_bam(os.path.join(d, "header_only.bam"))
    assert len(chroms) > 0 and recs == []
    assert all(isinstance(n, str) and ln > 0 for n, ln in chroms)
    chroms2, recs2 = CB.read_bam(os.path.join(d, "minimal.bam"))
    assert recs2 == []


def _discover_dirs(rng, tmp_path, n_samples):
    """One locus, a few true alleles, every sample writes its assembled contigs to its own discover dir."""
    acgt = np.frombuffer(b"ACGT", np.uint8)
    left, right = acgt[rng.integers(0, 4, 250)], acgt[rng.integers(0, 4, 250)]
    alleles = [np.concatenate([left, acgt[rng.integers(0, 4, int(rng.integers(40, 140)))], right]) for _ in range(3)]
    dirs, truth = [], []
    for s in range(n_samples):
        picks = rng.choice(3, size=int(rng.integers(1, 3)), replace=False)
        contigs = []
        for ai, a in enumerate(picks):
            seq = alleles[int(a)].copy()
            for _e in range(int(rng.integers(0, 3))):
                seq[int(rng.integers(0, len(seq)))] = acgt[int(rng.integers(0, 4))]
            ins = len(seq) - 500
            segs = [CB.ContigAlignmentSegment(1, 40_000 + int(rng.choice([0, 2, 40])), [("=", 250), ("I", ins), ("=", 250)])]
            contigs.append(CB.ContigAlignmentInfo(s, 0, ai, seq.tobytes(), 0, segs, int(rng.integers(3, 30)),
                                                  (int(rng.integers(0, 20)), len(seq) - int(rng.integers(0, 20)))))
            truth.append(int(a))
        d = tmp_path / f"sample{s}"
        d.mkdir()
        CB.write_contig_alignments(str(d / CB.CONTIG_ALIGNMENT_FILENAME), CHROMS, contigs)
        dirs.append(str(d))
    return dirs, truth


def _haplotypes_from_dirs(dirs):
    haps, sample_lists = [], []
    for d in dirs:
        lst = []
        for cluster in CB.get_candidate_sv_contig_info(os.path.join(d, CB.CONTIG_ALIGNMENT_FILENAME)):
            for asm in cluster:
                if not asm.contig_alignments:
                    continue
                a = asm.contig_alignments[0]
                ins = [n for op, n in a.contig_alignment.cigar if op == "I"]
                pos = a.contig_alignment.ref_offset + a.contig_alignment.cigar[0][1]
                bp = MP.Breakpoint(MP.Breakend(a.chrom_index, pos), MP.Breakend(a.chrom_index, pos), ins[0] if ins else 0, 0)
                lst.append(len(haps))
                haps.append(MP.MergeHaplotype(a.contig_seq, a.high_quality_contig_range, asm.supporting_read_count, [bp]))
        sample_lists.append(lst)
    return haps, sample_lists


def test_hand_off_feeds_the_merge_pools(tmp_path):
    rng = np.random.default_rng(9)
    dirs, truth = _discover_dirs(rng, tmp_path, n_samples=5)
    haps, sample_lists = _haplotypes_from_dirs(dirs)
    assert len(haps) == len(truth)
    merge = O.bio_scoring(0, -2, 1, -3, yclip_prefix=0, yclip_suffix=0)
    pools, _ = MP.get_haplotype_merge_pools(haps, sample_lists, lambda pairs: [O.bio_custom_align(merge, q, b)[0] for b, q in pairs])
    for p in pools:
        assert len({truth[h] for h in p}) <= 1        # only copies of the same allele were merged
    assert sum(1 for p in pools if p) == len(set(truth))


@pytest.mark.gpu
def test_hand_off_feeds_the_merge_pools_gpu(tmp_path):
    from sawfish_b200.batch import AlignContext
    ctx = AlignContext(0)
    rng = np.random.default_rng(10)
    dirs, truth = _discover_dirs(rng, tmp_path, n_samples=8)
    haps, sample_lists = _haplotypes_from_dirs(dirs)
    merge = O.bio_scoring(0, -2, 1, -3, yclip_prefix=0, yclip_suffix=0)
    want = MP.get_haplotype_merge_pools_sequential(haps, sample_lists, lambda b, q: O.bio_custom_align(merge, q, b)[0])
    got, _ = MP.get_haplotype_merge_pools(haps, sample_lists, MP.gpu_merge_score_fn(ctx))
    assert got == want
    for p in got:
        assert len({truth[h] for h in p}) <= 1
    ctx.close()


def test_parse_bnd_alt_allele_reference_vectors():
    """The reference's two unit tests (src/joint_call/get_refined_svs.rs:926-976), plus the mirrored orientations."""
    from sawfish_b200.candidate_sv import LEFT_ANCHOR, RIGHT_ANCHOR, parse_bnd_alt_allele
    for chrom in ("chr1", "HLA-DRB1*10:01:01"):
        info = parse_bnd_alt_allele({chrom: 0}, f"TCT]{chrom}:500]".encode(), 10)
        assert (info.breakend1_direction, info.breakend2_direction) == (LEFT_ANCHOR, LEFT_ANCHOR)
        assert (info.breakend2_chrom_index, info.breakend2_pos, info.insert_seq) == (0, 489, b"CT")
    info = parse_bnd_alt_allele({"chr1": 0}, b"T[chr1:500[", 10)      # left anchor joined to a right-anchored mate
    assert (info.breakend1_direction, info.breakend2_direction, info.breakend2_pos, info.insert_seq) == (LEFT_ANCHOR, RIGHT_ANCHOR, 498, b"")
    info = parse_bnd_alt_allele({"chr1": 0}, b"[chr1:500[GAT", 10)    # both right-anchored: homlen shift and the -1
    assert (info.breakend1_direction, info.breakend2_direction, info.breakend2_pos, info.insert_seq) == (RIGHT_ANCHOR, RIGHT_ANCHOR, 488, b"GA")
    with pytest.raises(ValueError):
        parse_bnd_alt_allele({"chr1": 0}, b"TCT", 0)


def test_candidate_records_to_refined_svs():
    """process_bcf_record_to_anno_refine_sv (get_refined_svs.rs:334-552) on the record content a candidate VCF carries."""
    from sawfish_b200 import candidate_sv as CS
    l2i = {"chr1": 0, "chr20": 1}
    sv_id, be = CS.get_sv_id_from_label("sawfish:0:12:1:0")
    assert (sv_id.sample_index, sv_id.cluster_index, sv_id.assembly_index, sv_id.alignment_index, be) == (0, 12, 1, 0, None)
    assert CS.get_sv_id_from_label("sawfish:0:12:1:0:1")[1] is False and CS.get_sv_id_from_label("sawfish:0:12:1:0:0")[1] is True
    lines = [
        # symbolic deletion with 3 bases of breakend homology and a small insertion at the junction
        "chr20\t1001\tsawfish:0:3:0:0\tA\t<DEL>\t.\t.\tSVTYPE=DEL;END=1501;SVLEN=500;HOMLEN=3;HOMSEQ=CAG;INSSEQ=TT;CONTIG_POS=1499",
        # sequence-resolved insertion, contig shared with assembly 1
        "chr20\t2001\tsawfish:0:4:0:0\tG\tGACGTACGT\t.\t.\tSVTYPE=INS;END=2001;SVLEN=8;CONTIG_POS=1500;OVERLAP_ASM=0,1",
        # breakend pair: only breakend 1 is converted
        "chr1\t5001\tsawfish:0:5:0:0:0\tT\tTCT]chr20:9000]\t.\t.\tSVTYPE=BND;HOMLEN=2;HOMSEQ=AC;CONTIG_POS=1498;BND0_NEIGHBOR=7:1",
        "chr20\t9000\tsawfish:0:5:0:0:1\tC\tC]chr1:5001]\t.\t.\tSVTYPE=BND;HOMLEN=2;HOMSEQ=AC;CONTIG_POS=1498",
    ]
    recs = [CS.parse_vcf_record_line(x, l2i) for x in lines]
    d = CS.process_record_to_anno_refined_sv(l2i, recs[0])
    assert (d.bp.breakend1, d.bp.breakend2) == (CS.RecBreakend(1, 1000, 1004, CS.LEFT_ANCHOR), CS.RecBreakend(1, 1500, 1504, CS.RIGHT_ANCHOR))
    assert (d.bp.insert_seq, d.breakend1_homology_seq, d.contig_pos_before_breakend1, d.contig_map) == (b"TT", b"CAG", 1499, [0])
    assert d.bp.is_indel() and d.bp.deletion_size() == 500
    i = CS.process_record_to_anno_refined_sv(l2i, recs[1])
    assert (i.bp.breakend1.start, i.bp.breakend2.start, i.bp.insert_seq, i.contig_map) == (2000, 2000, b"ACGTACGT", [0, 1])
    assert i.bp.deletion_size() == 0
    b = CS.process_record_to_anno_refined_sv(l2i, recs[2])
    assert b.bp.breakend1 == CS.RecBreakend(0, 5000, 5003, CS.LEFT_ANCHOR)
    assert b.bp.breakend2 == CS.RecBreakend(1, 8997, 9000, CS.LEFT_ANCHOR)        # same orientation: shifted left by homlen
    assert (b.bp.insert_seq, b.bp.breakend1_neighbor, b.bp.breakend2_neighbor) == (b"CT", (7, 1), None)
    assert not b.bp.is_indel() and b.bp.deletion_size() is None
    assert CS.process_record_to_anno_refined_sv(l2i, recs[3]) is None


def test_candidate_bcf_round_trip_and_layout(tmp_path):
    """BCF2 subset used by candidate.sv.bcf: encoder / decoder round trip, typed-value layout per the VCF specification,
    and the decoded records feed the same record -> breakpoint logic as the VCF text form."""
    from sawfish_b200 import candidate_sv as CS
    chroms = [("chr1", 248956422), ("chr20", 64444167)]
    l2i = {"chr20": 0, "chr1": 1}          # the caller's chromosome order differs from the file's: mapped by name
    types = {"SVTYPE": "String", "END": "Integer", "SVLEN": "Integer", "HOMLEN": "Integer", "HOMSEQ": "String", "INSSEQ": "String",
             "CONTIG_POS": "Integer", "OVERLAP_ASM": "Integer", "BND0_NEIGHBOR": "String", "IMPRECISE": "Flag"}
    file_index = {n: i for i, (n, _l) in enumerate(chroms)}
    lines = [
        "chr20\t1001\tsawfish:0:3:0:0\tA\t<DEL>\t.\t.\tSVTYPE=DEL;END=1501;SVLEN=500;HOMLEN=3;HOMSEQ=CAG;INSSEQ=TT;CONTIG_POS=1499",
        "chr20\t2001\tsawfish:0:4:0:0\tG\tG" + "ACGT" * 40 + "\t.\t.\tSVTYPE=INS;END=2001;SVLEN=160;CONTIG_POS=70000;OVERLAP_ASM=0,1;IMPRECISE",
        "chr1\t5001\tsawfish:0:5:0:0:0\tT\tTCT]chr20:9000]\t.\t.\tSVTYPE=BND;HOMLEN=2;HOMSEQ=AC;CONTIG_POS=1498;BND0_NEIGHBOR=7:1",
    ]
    recs_file_order = [CS.parse_vcf_record_line(x, file_index) for x in lines]
    path = str(tmp_path / CS.CANDIDATE_SV_FILENAME)
    CS.write_candidate_bcf(path, chroms, recs_file_order, types)
    got = CS.read_candidate_bcf(path, l2i)
    want = [CS.parse_vcf_record_line(x, l2i) for x in lines]
    assert len(got) == 3
    for g, w in zip(got, want):
        assert (g.chrom_index, g.pos, g.id, g.alleles) == (w.chrom_index, w.pos, w.id, w.alleles)
        assert {k: [str(x) for x in v] for k, v in g.info.items()} == {k: [str(x) for x in v] for k, v in w.info.items()}
        a, b = CS.process_record_to_anno_refined_sv(l2i, g), CS.process_record_to_anno_refined_sv(l2i, w)
        assert (a.bp, a.contig_map, a.contig_pos_before_breakend1) == (b.bp, b.contig_map, b.contig_pos_before_breakend1)
    # layout spot checks on the uncompressed stream
    import gzip
    raw = gzip.open(path, "rb").read()
    assert raw[:5] == b"BCF\x02\x02"
    (l_text,) = struct.unpack_from("<I", raw, 5)
    assert raw[9 + l_text - 1] == 0 and b"##contig=<ID=chr20" in raw[9:9 + l_text]
    o = 9 + l_text
    l_shared, l_indiv = struct.unpack_from("<II", raw, o)
    chrom, pos, rlen, qual_bits, nai, nfs = struct.unpack_from("<iiiIII", raw, o + 8)
    assert (l_indiv, chrom, pos, rlen, qual_bits, nai >> 16, nai & 0xFFFF, nfs) == (0, 1, 1000, 501, 0x7F800001, 2, 7, 0)
    assert raw[o + 32] == 0xF7 and raw[o + 33] == 0x11 and raw[o + 34] == 15          # ID: 15 chars -> overflow count form
    assert CS._enc_ints([1499]) == bytes([0x12]) + struct.pack("<h", 1499) and CS._enc_ints([70000])[0] == 0x13
    assert CS._dec_typed(bytes([0x21, 5, 0x81]), 0) == ([5], 3)                         # end-of-vector padding is dropped


def _write_discover_dirs(rng, tmp_path, n_samples):
    """Per sample: contig.alignment.bam + candidate.sv.bcf for one insertion locus with a few true alleles."""
    from sawfish_b200 import candidate_sv as CS
    acgt = np.frombuffer(b"ACGT", np.uint8)
    left, right = acgt[rng.integers(0, 4, 250)], acgt[rng.integers(0, 4, 250)]
    alleles = [acgt[rng.integers(0, 4, int(rng.integers(40, 140)))] for _ in range(3)]
    types = {"SVTYPE": "String", "END": "Integer", "SVLEN": "Integer", "CONTIG_POS": "Integer"}
    dirs, truth = [], []
    for s in range(n_samples):
        picks = rng.choice(3, size=int(rng.integers(1, 3)), replace=False)
        contigs, recs = [], []
        for ai, a in enumerate(picks):
            ins = alleles[int(a)]
            seq = np.concatenate([left, ins, right])
            for _e in range(int(rng.integers(0, 3))):
                p = int(rng.integers(0, len(seq)))
                seq[p] = acgt[int(rng.integers(0, 4))]
            pos = 40_000 + int(rng.choice([0, 2, 40]))          # leftmost aligned base of the contig
            segs = [CB.ContigAlignmentSegment(1, pos, [("=", 250), ("I", len(ins)), ("=", 250)])]
            contigs.append(CB.ContigAlignmentInfo(s, 0, ai, seq.tobytes(), 0, segs, int(rng.integers(3, 30)),
                                                  (int(rng.integers(0, 20)), len(seq) - int(rng.integers(0, 20)))))
            anchor = pos + 249                                   # 0-based last reference base before the insertion
            ref_base = chr(seq[249]).encode()
            recs.append(CS.CandidateRecord(1, anchor, f"sawfish:{s}:0:{ai}:0", [ref_base, ref_base + seq[250:250 + len(ins)].tobytes()],
                                           {"SVTYPE": ["INS"], "END": [anchor + 1], "SVLEN": [len(ins)], "CONTIG_POS": [249]}))
            truth.append(int(a))
        d = tmp_path / f"discover{s}"
        d.mkdir()
        CB.write_contig_alignments(str(d / CB.CONTIG_ALIGNMENT_FILENAME), CHROMS, contigs)
        CS.write_candidate_bcf(str(d / CS.CANDIDATE_SV_FILENAME), CHROMS, recs, types)
        dirs.append(str(d))
    return dirs, truth


def test_discover_dirs_to_merge_pools(tmp_path):
    """Both files of the hand-off, per sample, through load_discover_dir into the duplicate-haplotype pooling."""
    from sawfish_b200 import discover_dir as DD
    rng = np.random.default_rng(12)
    dirs, truth = _write_discover_dirs(rng, tmp_path, n_samples=5)
    l2i = {n: i for i, (n, _l) in enumerate(CHROMS)}
    samples = [DD.load_discover_dir(d, l2i) for d in dirs]
    haps, sample_lists = DD.merge_haplotypes_of_samples(samples)
    assert len(haps) == len(truth) and all(len(h.breakpoints) == 1 for h in haps)
    assert all(h.breakpoints[0].deletion_size == 0 and h.breakpoints[0].insert_size >= 40 for h in haps)
    merge = O.bio_scoring(0, -2, 1, -3, yclip_prefix=0, yclip_suffix=0)
    pools, _ = MP.get_haplotype_merge_pools(haps, sample_lists, lambda pairs: [O.bio_custom_align(merge, q, b)[0] for b, q in pairs])
    for p in pools:
        assert len({truth[h] for h in p}) <= 1
    assert sum(1 for p in pools if p) == len(set(truth))
    with pytest.raises(FileNotFoundError):
        DD.load_discover_dir(str(tmp_path / "missing"), l2i)


def test_refined_svs_to_sv_groups_index_remapping():
    """process_rsv_cluster_to_sv_group / group_single_sample_refined_svs (get_refined_svs.rs:606-846)."""
    from sawfish_b200 import candidate_sv as CS
    from sawfish_b200 import discover_dir as DD

    def asm(tag):
        return CB.ClusterAssembly([], tag)               # supporting_read_count doubles as a label here

    def sv(cluster, assembly, alignment, contig_map):
        be = CS.RecBreakend(0, 100, 101, CS.LEFT_ANCHOR)
        return CS.AnnotatedRefinedSV(CS.SVUniqueIdData(0, cluster, assembly, alignment), CS.RecBreakpoint(be, be, b""), b"", 0, contig_map)

    contigs = [[asm(10), asm(11), asm(12)], [asm(20)], [asm(30), asm(31)]]
    svs = [sv(2, 1, 0, [0, 1]), sv(0, 2, 0, [2, 0]), sv(0, 0, 1, [2, 0]), sv(0, 0, 0, [2, 0]), sv(0, 1, 0, [2, 0]), sv(1, 0, 0, [0]),
           sv(2, 0, 0, [0, 1])]
    groups = DD.group_single_sample_refined_svs(contigs, svs)
    assert [g.cluster_index for g in groups] == [0, 1, 2]
    g0 = groups[0]
    # cluster 0: the sample's haplotypes are assemblies [2, 0] (first SV in sorted order is (0,0,0) with contig map [2, 0]);
    # assembly 1 is dropped with its SV, the others are re-indexed in contig-map order: 2 -> 0, 0 -> 1
    # (the reference keeps the surviving haplotypes in their ORIGINAL order while numbering them in contig-map order, :712-733;
    #  the two agree because discover writes ascending contig maps — the descending map here pins the mirrored behaviour)
    assert [h.supporting_read_count for h in g0.group_haplotypes] == [10, 12]      # kept in original order
    assert g0.sample_haplotype_list == [[0, 1]]
    assert [(s.id.assembly_index, s.id.alignment_index) for s in g0.refined_svs] == [(0, 0), (0, 1), (2, 0)]
    assert g0.sv_haplotype_map == [1, 1, 0]
    assert groups[1].sample_haplotype_list == [[0]] and groups[1].sv_haplotype_map == [0]
    assert groups[2].sample_haplotype_list == [[0, 1]] and groups[2].sv_haplotype_map == [0, 1]
    # haploid / multi-region loci keep one haplotype: the SV on the other assembly goes away with it
    g2 = DD.group_single_sample_refined_svs(contigs, svs, lambda c: 1)[2]
    assert len(g2.group_haplotypes) == 1 and [s.id.assembly_index for s in g2.refined_svs] == [0] and g2.sv_haplotype_map == [0]
    # a cluster whose only SVs sit on dropped haplotypes yields no group
    only = [sv(0, 1, 0, [2])]
    assert DD.group_single_sample_refined_svs(contigs, only, lambda c: 1) == []
