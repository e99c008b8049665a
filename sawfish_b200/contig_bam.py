"""The on-disk hand-off between `discover` and `joint-call` for assembled contigs: `contig.alignment.bam`.

Writer side = reference src/contig_output.rs:103-268 (convert_contig_alignment_to_bam_record, get_contig_alignment_header,
write_contig_alignments); reader side = src/joint_call/get_refined_svs.rs:38-195 (get_supporting_read_count,
get_high_quality_contig_range, update_sample_assemblies_from_bam_record, get_candidate_sv_contig_info).  The reference
does both through rust-htslib; htslib is not in this image, so BGZF + BAM are written / parsed here directly (SAM spec
v1.6 sections 4.1-4.2): what matters for the hot path is that contig sequence, per-segment alignment, high-quality range
and supporting-read count survive the file exactly, because joint-call rebuilds its haplotypes from them — the inputs of
the merge aligner (sawfish_b200/merge_pools.py -> K5) and of score_sv's allele flanks.
The CSI index the reference also writes (:248-266) is not produced: nothing on this path reads it.
"""
from __future__ import annotations

import gzip
import struct
import zlib
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

from .simple_alignment import Cigar, SimpleAlignment, cigar_from_string, cigar_from_u32, cigar_to_string, cigar_to_u32

SA_AUX_TAG = b"SA"       # src/contig_output.rs:14
CONTIG_AUX_TAG = b"sf"   # :15
CONTIG_MAPQ = 20         # :108
CONTIG_ALIGNMENT_FILENAME = "contig.alignment.bam"  # src/filenames.rs
BAM_FREVERSE, BAM_FSUPPLEMENTARY = 0x10, 0x800
_SEQ_CODES = "=ACMGRSVTWYHKDBN"
_ENC = {c: i for i, c in enumerate(_SEQ_CODES)}
_COMPLEMENT = bytes.maketrans(b"ACGTNacgtn", b"TGCANtgcan")


@dataclass
class ContigAlignmentSegment:  # src/contig_output.rs:17-29
    tid: int
    pos: int
    cigar: List[Cigar]
    is_fwd_strand: bool = True


@dataclass
class ContigAlignmentInfo:  # :47-86 (one record per segment: segment_id selects it)
    sample_index: int
    cluster_index: int
    assembly_index: int
    seq: bytes
    segment_id: int
    segments: List[ContigAlignmentSegment]
    supporting_read_count: int
    high_quality_contig_range: Tuple[int, int]

    def get_segment(self) -> ContigAlignmentSegment:
        return self.segments[self.segment_id]


@dataclass
class ClusterAssemblyAlignment:  # what update_sample_assemblies_from_bam_record builds per segment
    contig_seq: bytes
    chrom_index: int
    contig_alignment: SimpleAlignment
    high_quality_contig_range: Tuple[int, int]
    is_fwd_strand: bool


@dataclass
class ClusterAssembly:
    contig_alignments: List[ClusterAssemblyAlignment] = field(default_factory=list)
    supporting_read_count: int = 0


# ------------------------------------------------------------------------------------------------ BGZF
_BGZF_EOF = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


def _bgzf_block(data: bytes) -> bytes:
    comp = zlib.compressobj(6, zlib.DEFLATED, -15)
    body = comp.compress(data) + comp.flush()
    bsize = 12 + 6 + len(body) + 8 - 1  # header (12) + extra (6) + deflate + crc/isize, minus one
    assert bsize < 65536
    return (b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", bsize) + body +
            struct.pack("<II", zlib.crc32(data) & 0xFFFFFFFF, len(data)))


def _bgzf_write(path: str, payload: bytes) -> None:
    with open(path, "wb") as f:
        for o in range(0, len(payload), 0xFF00):
            f.write(_bgzf_block(payload[o:o + 0xFF00]))
        f.write(_BGZF_EOF)


# ------------------------------------------------------------------------------------------------ helpers
def reg2bin(beg: int, end: int) -> int:
    """SAM spec 5.3 (rust_vc_utils::bam_reg2bin)."""
    end -= 1
    for shift, off in ((14, 4681), (17, 585), (20, 73), (23, 9), (26, 1)):
        if beg >> shift == end >> shift:
            return off + (beg >> shift)
    return 0


def _ref_len(cigar: Sequence[Cigar]) -> int:
    return sum(n for op, n in cigar if op in "MDN=X")


def rev_comp(seq: bytes) -> bytes:
    return seq.translate(_COMPLEMENT)[::-1]


def get_sa_tag_segment(chrom_labels: Sequence[str], seg: ContigAlignmentSegment, mapq: int) -> str:  # :88-101
    return f"{chrom_labels[seg.tid]},{seg.pos + 1},{'+' if seg.is_fwd_strand else '-'},{cigar_to_string(seg.cigar)},{mapq},0;"


# ------------------------------------------------------------------------------------------------ writer
def _record_bytes(chrom_labels: Sequence[str], pkg_name: str, ca: ContigAlignmentInfo) -> bytes:
    """convert_contig_alignment_to_bam_record (:103-173)."""
    qname = f"{pkg_name}:{ca.sample_index}:{ca.cluster_index}:{ca.assembly_index}".encode() + b"\0"
    seg = ca.get_segment()
    flag = (0 if seg.is_fwd_strand else BAM_FREVERSE) | (BAM_FSUPPLEMENTARY if ca.segment_id != 0 else 0)
    cig = cigar_to_u32(seg.cigar)
    l_seq = len(ca.seq)
    codes = [_ENC.get(chr(b).upper(), 15) for b in ca.seq]
    if l_seq & 1:
        codes.append(0)
    packed = bytes((codes[i] << 4) | codes[i + 1] for i in range(0, len(codes), 2))
    aux = b""
    if len(ca.segments) > 1:
        sa = ""
        for si, s in enumerate(ca.segments):
            if si == ca.segment_id:
                continue
            assert any(op in "M=X" for op, _n in s.cigar), f"invalid CIGAR string in contig SA output for contig qname {qname!r}"
            sa += get_sa_tag_segment(chrom_labels, s, CONTIG_MAPQ)
        aux += SA_AUX_TAG + b"Z" + sa.encode() + b"\0"
    lo, hi = ca.high_quality_contig_range
    aux += CONTIG_AUX_TAG + b"Z" + f"n_reads:{ca.supporting_read_count};hq_range:{lo}-{hi};".encode() + b"\0"
    end = seg.pos + _ref_len(seg.cigar)
    core = struct.pack("<iiBBHHHIiii", seg.tid, seg.pos, len(qname), CONTIG_MAPQ, reg2bin(seg.pos, max(end, seg.pos + 1)),
                       len(cig), flag, l_seq, -1, -1, 0)
    body = core + qname + struct.pack(f"<{len(cig)}I", *cig) + packed + b"\xff" * l_seq + aux
    return struct.pack("<i", len(body)) + body


def write_contig_alignments(path: str, chroms: Sequence[Tuple[str, int]], contig_alignments: Sequence[ContigAlignmentInfo],
                            pkg_name: str = "sawfish", version: str = "b200", cmdline: str = "") -> None:
    """write_contig_alignments (:220-247) without the index: sort by (tid, pos) of each record's own segment, coordinate-
    sorted header with @SQ per chromosome and one @PG line (:178-203)."""
    recs = sorted(contig_alignments, key=lambda a: (a.get_segment().tid, a.get_segment().pos))  # stable, like sort_by
    text = "@HD\tVN:1.6\tSO:coordinate\n" + "".join(f"@SQ\tSN:{n}\tLN:{ln}\n" for n, ln in chroms) + \
           f"@PG\tPN:{pkg_name}\tID:{pkg_name}-{version}\tVN:{version}\tCL:{cmdline}\n"
    out = [b"BAM\x01", struct.pack("<i", len(text)), text.encode(), struct.pack("<i", len(chroms))]
    for n, ln in chroms:
        nb = n.encode() + b"\0"
        out += [struct.pack("<i", len(nb)), nb, struct.pack("<i", ln)]
    labels = [n for n, _ in chroms]
    out += [_record_bytes(labels, pkg_name, ca) for ca in recs]
    _bgzf_write(path, b"".join(out))


# ------------------------------------------------------------------------------------------------ reader
@dataclass
class BamRecord:
    qname: str
    flag: int
    tid: int
    pos: int
    mapq: int
    cigar: List[Cigar]
    seq: bytes
    aux: Dict[bytes, str]


def read_bam(path: str) -> Tuple[List[Tuple[str, int]], List[BamRecord]]:
    with gzip.open(path, "rb") as f:  # BGZF is a series of gzip members
        data = f.read()
    assert data[:4] == b"BAM\x01", "not a BAM file"
    (l_text,) = struct.unpack_from("<i", data, 4)
    o = 8 + l_text
    (n_ref,) = struct.unpack_from("<i", data, o)
    o += 4
    chroms = []
    for _ in range(n_ref):
        (l_name,) = struct.unpack_from("<i", data, o)
        name = data[o + 4:o + 4 + l_name - 1].decode()
        (l_ref,) = struct.unpack_from("<i", data, o + 4 + l_name)
        chroms.append((name, l_ref))
        o += 8 + l_name
    recs = []
    while o < len(data):
        (bs,) = struct.unpack_from("<i", data, o)
        tid, pos, l_qn, mapq, _bin, n_cig, flag, l_seq, _nt, _np, _tl = struct.unpack_from("<iiBBHHHIiii", data, o + 4)
        p = o + 36
        qname = data[p:p + l_qn - 1].decode()
        p += l_qn
        cigar = cigar_from_u32(struct.unpack_from(f"<{n_cig}I", data, p))
        p += 4 * n_cig
        nb = (l_seq + 1) // 2
        seq = "".join(_SEQ_CODES[b >> 4] + _SEQ_CODES[b & 15] for b in data[p:p + nb])[:l_seq].encode()
        p += nb + l_seq
        aux: Dict[bytes, str] = {}
        end = o + 4 + bs
        while p < end:
            tag, typ = data[p:p + 2], data[p + 2:p + 3]
            assert typ == b"Z", f"unexpected aux type {typ!r} in contig BAM"  # the reference writes only string tags here
            z = data.index(b"\0", p + 3)
            aux[tag] = data[p + 3:z].decode()
            p = z + 1
        recs.append(BamRecord(qname, flag, tid, pos, mapq, cigar, seq, aux))
        o = end
    return chroms, recs


def _contig_tag_value(rec: BamRecord, key: str) -> Optional[str]:
    for word in rec.aux[CONTIG_AUX_TAG].split(";"):
        if not word:
            continue
        kv = word.split(":")
        assert len(kv) == 2
        if kv[0] == key:
            return kv[1]
    return None


def get_supporting_read_count(rec: BamRecord) -> int:  # get_refined_svs.rs:38-54
    v = _contig_tag_value(rec, "n_reads")
    return int(v) if v is not None else 0


def get_high_quality_contig_range(rec: BamRecord) -> Optional[Tuple[int, int]]:  # :56-77
    v = _contig_tag_value(rec, "hq_range")
    if v is None:
        return None
    a, b = v.split("-")
    return int(a), int(b)


def get_simplified_dna_seq(seq: bytes) -> bytes:  # src/bam_utils.rs:326-335
    return bytes(b if b in b"ACGTN" else ord("N") for b in seq)


def get_candidate_sv_contig_info(path: str) -> List[List[ClusterAssembly]]:
    """get_candidate_sv_contig_info + update_sample_assemblies_from_bam_record (:85-195): sample_assemblies[cluster_index]
    [assembly_index].  Only primary records are read; breakend 2 comes from the primary record's SA tag."""
    chroms, recs = read_bam(path)
    chrom_index = {n: i for i, (n, _l) in enumerate(chroms)}
    sample_assemblies: List[List[ClusterAssembly]] = []
    for rec in recs:
        if rec.flag & BAM_FSUPPLEMENTARY:
            continue
        words = rec.qname.split(":")
        assert len(words) >= 4
        cluster_index, assembly_index = int(words[2]), int(words[3])
        while len(sample_assemblies) <= cluster_index:
            sample_assemblies.append([ClusterAssembly()])
        ca_list = sample_assemblies[cluster_index]
        while len(ca_list) <= assembly_index:
            ca_list.append(ClusterAssembly())
        ca = ca_list[assembly_index]
        ca.supporting_read_count = get_supporting_read_count(rec)
        hq = get_high_quality_contig_range(rec)
        assert hq is not None
        contig_seq = get_simplified_dna_seq(rec.seq)
        ca.contig_alignments.append(ClusterAssemblyAlignment(contig_seq, rec.tid, SimpleAlignment(rec.pos, list(rec.cigar)), hq, True))
        if SA_AUX_TAG in rec.aux:
            # every other segment of the contig, in read (sequencing) order — get_seq_order_read_split_segments
            from .read_flanks import get_seq_order_read_split_segments
            segs = get_seq_order_read_split_segments(chrom_index, rec.tid, rec.pos, bool(rec.flag & BAM_FREVERSE), rec.cigar, rec.mapq,
                                                     rec.aux[SA_AUX_TAG])
            for seg in segs:
                if seg.from_primary_bam_record:
                    continue
                seq2, hq2 = contig_seq, hq
                if not seg.is_fwd_strand:  # rev_comp_in_place + IntRange::reverse(len): [len - end, len - start)
                    seq2 = rev_comp(contig_seq)
                    hq2 = (len(contig_seq) - hq[1], len(contig_seq) - hq[0])
                ca.contig_alignments.append(ClusterAssemblyAlignment(seq2, seg.chrom_index, SimpleAlignment(seg.pos, list(seg.cigar)), hq2,
                                                                     seg.is_fwd_strand))
    return sample_assemblies
