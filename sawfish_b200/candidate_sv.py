"""Pieces of the `candidate.sv.bcf` side of the discover -> joint-call hand-off that do not need a BCF parser.

Reference: src/joint_call/get_refined_svs.rs.  Built: `parse_bnd_alt_allele` (:206-287), pinned by the reference's two unit
tests (:926-976); `get_sv_id_from_label` (src/sv_id.rs:52-76); `process_record_to_anno_refined_sv`, the record -> breakpoint
logic of `process_bcf_record_to_anno_refine_sv` (:334-552), fed from VCF text lines.  Not built: the BCF2 binary decoding
(typed INFO values, string / contig dictionaries) the reference gets from htslib.  `sawfish_b200/contig_bam.py` covers the
contig half of the hand-off.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict

LEFT_ANCHOR, RIGHT_ANCHOR = "LeftAnchor", "RightAnchor"  # crate::breakpoint::BreakendDirection


@dataclass
class BndAltInfo:  # :196-204
    breakend1_direction: str
    breakend2_chrom_index: int
    breakend2_pos: int
    breakend2_direction: str
    insert_seq: bytes


def parse_bnd_alt_allele(label_to_index: Dict[str, int], alt_allele: bytes, homlen: int) -> BndAltInfo:
    """Parse a VCF breakend alt allele such as b"TCT]chr1:500]" (:206-287).  `homlen` corrects the remote breakend to
    its left-shifted position when both breakends have the same orientation (inverted breakpoint)."""
    s = alt_allele.decode()
    if "]" in s:
        breakend2_direction, split_char = LEFT_ANCHOR, "]"
    elif "[" in s:
        breakend2_direction, split_char = RIGHT_ANCHOR, "["
    else:
        raise ValueError(f"Invalid BND alt alt allele '{s}'")
    words = s.split(split_char)
    assert len(words) == 3, f"Unexpected BND alt allele format `{s}`"
    mate = words[1].rsplit(":", 1)  # chromosome names may contain ':' (HLA contigs): split at the last one
    assert len(mate) == 2, f"Unexpected BND alt allele format `{s}`"
    breakend2_chrom_index = label_to_index[mate[0]]
    breakend2_raw_pos = int(mate[1]) - 1
    if words[0] and not words[2]:
        breakend1_direction = LEFT_ANCHOR
    elif not words[0] and words[2]:
        breakend1_direction = RIGHT_ANCHOR
    else:
        raise ValueError(f"Invalid BND alt alt allele '{s}'")
    insert_seq = words[0].encode()[1:] if breakend1_direction == LEFT_ANCHOR else words[2].encode()[:-1]
    pos = breakend2_raw_pos
    if breakend1_direction == breakend2_direction:
        pos -= homlen
    if breakend2_direction == RIGHT_ANCHOR:
        pos -= 1
    return BndAltInfo(breakend1_direction, breakend2_chrom_index, pos, breakend2_direction, insert_seq)


# ---------------------------------------------------------------------------------------------- records -> refined SVs
# process_bcf_record_to_anno_refine_sv (:334-552) on a parsed record.  The reference reads `candidate.sv.bcf` through
# htslib; the record CONTENT is what this mirrors (the same CHROM / POS / ID / ALT / INFO a `bcftools view` of that file
# shows), fed here from VCF text lines.
from typing import List, Optional, Tuple, Union  # noqa: E402

CONTIG_POS_INFO_KEY, OVERLAP_ASM_INFO_KEY = "CONTIG_POS", "OVERLAP_ASM"        # src/large_variant_output.rs:23-24
BND0_NEIGHBOR_INFO_KEY, BND1_NEIGHBOR_INFO_KEY = "BND0_NEIGHBOR", "BND1_NEIGHBOR"  # :25-26


@dataclass(frozen=True)
class SVUniqueIdData:  # src/sv_id.rs
    sample_index: int
    cluster_index: int
    assembly_index: int
    alignment_index: int


def get_sv_id_from_label(label: str) -> Tuple[SVUniqueIdData, Optional[bool]]:
    """src/sv_id.rs:52-76: "pkg:sample:cluster:assembly:alignment[:breakend]" -> (id, is_breakend1)."""
    parts = label.split(":")
    assert 5 <= len(parts) <= 6
    sv_id = SVUniqueIdData(int(parts[1]), int(parts[2]), int(parts[3]), int(parts[4]))
    return sv_id, (int(parts[5]) == 0 if len(parts) > 5 else None)


@dataclass
class CandidateRecord:
    chrom_index: int
    pos: int                       # 0-based, like bcf::Record::pos()
    id: str
    alleles: List[bytes]
    info: Dict[str, List[Union[int, str]]]


def parse_vcf_record_line(line: str, label_to_index: Dict[str, int]) -> CandidateRecord:
    f = line.rstrip("\n").split("\t")
    info: Dict[str, List[Union[int, str]]] = {}
    for item in f[7].split(";"):
        if not item or item == ".":
            continue
        if "=" not in item:
            info[item] = []
            continue
        key, val = item.split("=", 1)
        info[key] = [int(v) if v.lstrip("-").isdigit() else v for v in val.split(",")]
    return CandidateRecord(label_to_index[f[0]], int(f[1]) - 1, f[2], [f[3].encode()] + [a.encode() for a in f[4].split(",")], info)


@dataclass(frozen=True)
class RecBreakend:  # Breakend { segment: GenomeSegment { chrom_index, range }, dir }
    chrom_index: int
    start: int
    end: int
    dir: str


@dataclass
class RecBreakpoint:
    breakend1: RecBreakend
    breakend2: RecBreakend
    insert_seq: bytes                      # InsertInfo::Seq / NoInsert
    breakend1_neighbor: Optional[Tuple[int, int]] = None   # (cluster_index, breakend_index)
    breakend2_neighbor: Optional[Tuple[int, int]] = None

    def is_indel(self) -> bool:            # src/breakpoint.rs:251-262
        return self.breakend1.chrom_index == self.breakend2.chrom_index and \
            self.breakend1.dir == LEFT_ANCHOR and self.breakend2.dir == RIGHT_ANCHOR

    def deletion_size(self) -> Optional[int]:  # :264-273
        return max(self.breakend2.start - self.breakend1.start, 0) if self.is_indel() else None


@dataclass
class AnnotatedRefinedSV:
    id: SVUniqueIdData
    bp: RecBreakpoint
    breakend1_homology_seq: bytes
    contig_pos_before_breakend1: int
    contig_map: List[int]


def _first(info, key, default=None):
    v = info.get(key)
    return v[0] if v else default


def _neighbor(info, key) -> Optional[Tuple[int, int]]:  # parse_breakend_neighbor_info (:300-329): "cluster:breakend"
    v = _first(info, key)
    if v is None:
        return None
    a, b = str(v).split(":")
    return int(a), int(b)


def process_record_to_anno_refined_sv(label_to_index: Dict[str, int], rec: CandidateRecord) -> Optional[AnnotatedRefinedSV]:
    """:334-552.  DEL / INS records give an indel breakpoint (breakend ranges widened by the homology length); of a BND pair
    only the breakend-1 record is used, its mate comes from the alt allele."""
    sv_id, is_breakend1 = get_sv_id_from_label(rec.id)
    if len(rec.alleles) != 2:
        raise ValueError(f"Unexpected allele count in sample candidate SV input {len(rec.alleles)}")
    alt = rec.alleles[1]
    svtype = str(_first(rec.info, "SVTYPE"))
    homlen = int(_first(rec.info, "HOMLEN", 0))
    homseq = str(_first(rec.info, "HOMSEQ", "")).encode()
    assert homlen == len(homseq)
    contig_pos = int(_first(rec.info, CONTIG_POS_INFO_KEY))
    overlapping = [int(x) for x in rec.info.get(OVERLAP_ASM_INFO_KEY, [])]
    contig_map = overlapping if overlapping else [sv_id.assembly_index]
    if svtype in ("DEL", "INS"):
        assert is_breakend1 is None
        end = int(_first(rec.info, "END")) - 1
        be1 = RecBreakend(rec.chrom_index, rec.pos, rec.pos + homlen + 1, LEFT_ANCHOR)
        be2 = RecBreakend(rec.chrom_index, end, end + homlen + 1, RIGHT_ANCHOR)
        insert_seq = str(_first(rec.info, "INSSEQ", "")).encode() if alt == b"<DEL>" else alt[1:]
        return AnnotatedRefinedSV(sv_id, RecBreakpoint(be1, be2, insert_seq), homseq, contig_pos, contig_map)
    if svtype == "BND":
        assert is_breakend1 is not None
        if not is_breakend1:
            return None
        b = parse_bnd_alt_allele(label_to_index, alt, homlen)
        p1 = rec.pos if b.breakend1_direction == LEFT_ANCHOR else rec.pos - 1
        be1 = RecBreakend(rec.chrom_index, p1, p1 + homlen + 1, b.breakend1_direction)
        be2 = RecBreakend(b.breakend2_chrom_index, b.breakend2_pos, b.breakend2_pos + homlen + 1, b.breakend2_direction)
        bp = RecBreakpoint(be1, be2, b.insert_seq, _neighbor(rec.info, BND0_NEIGHBOR_INFO_KEY), _neighbor(rec.info, BND1_NEIGHBOR_INFO_KEY))
        return AnnotatedRefinedSV(sv_id, bp, homseq, contig_pos, contig_map)
    raise ValueError(f"Unexpected SV type: {svtype}")


# ---------------------------------------------------------------------------------------------- BCF2 (VCF spec v4.3, section 6)
# `candidate.sv.bcf` is written and read through htslib in the reference (get_anno_refined_svs, :554-601).  htslib is not in
# this image: the subset of BCF2 those records use (site-only records: CHROM, POS, ID, REF/ALT, FILTER, INFO with Integer /
# String / Flag values, no FORMAT block) is decoded / encoded here directly.  Parity of the decoder with htslib-written files
# is UNPINNED (no candidate.sv.bcf ships with the reference); it is checked against this module's own encoder and against the
# byte layout of the specification.
import gzip as _gzip  # noqa: E402
import struct as _struct  # noqa: E402
import re as _re  # noqa: E402

CANDIDATE_SV_FILENAME = "candidate.sv.bcf"  # src/filenames.rs
_BCF_INT = {1: ("b", 1, -128), 2: ("h", 2, -32768), 3: ("i", 4, -2147483648)}  # type code -> (struct fmt, size, missing)


def _typed_int_type(vals) -> int:
    lo, hi = (min(vals), max(vals)) if vals else (0, 0)
    if -120 <= lo and hi <= 127:
        return 1
    if -32760 <= lo and hi <= 32767:
        return 2
    return 3


def _enc_desc(count: int, typ: int) -> bytes:
    if count < 15:
        return bytes([(count << 4) | typ])
    t = _typed_int_type([count])
    return bytes([0xF0 | typ, (1 << 4) | t]) + _struct.pack("<" + _BCF_INT[t][0], count)


def _enc_ints(vals) -> bytes:
    t = _typed_int_type(vals)
    return _enc_desc(len(vals), t) + _struct.pack(f"<{len(vals)}{_BCF_INT[t][0]}", *vals)


def _enc_str(s: bytes) -> bytes:
    return _enc_desc(len(s), 7) + s if s else bytes([0x07])


def write_candidate_bcf(path: str, chroms, records: List[CandidateRecord], info_types: Dict[str, str]) -> None:
    """records' chrom_index refers to `chroms` [(name, length)]; info_types: key -> "Integer" | "String" | "Flag"."""
    from .contig_bam import _bgzf_write
    keys = list(info_types)
    text = "##fileformat=VCFv4.2\n##FILTER=<ID=PASS,Description=\"All filters passed\">\n"
    for k in keys:
        number = "0" if info_types[k] == "Flag" else "."
        text += f"##INFO=<ID={k},Number={number},Type={info_types[k]},Description=\"{k}\">\n"
    for name, ln in chroms:
        text += f"##contig=<ID={name},length={ln}>\n"
    text += "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\n"
    tb = text.encode() + b"\0"
    out = [b"BCF\x02\x02", _struct.pack("<I", len(tb)), tb]
    dict_index = {"PASS": 0, **{k: i + 1 for i, k in enumerate(keys)}}
    for r in records:
        rlen = len(r.alleles[0])
        end = r.info.get("END")
        if end:
            rlen = int(end[0]) - r.pos
        body = _struct.pack("<iiifII", r.chrom_index, r.pos, rlen, float("nan"), (len(r.alleles) << 16) | len(r.info), 0)
        body = body[:12] + _struct.pack("<I", 0x7F800001) + body[16:]  # QUAL missing
        body += _enc_str(r.id.encode())
        for a in r.alleles:
            body += _enc_str(a)
        body += bytes([0x00])  # FILTER: empty vector
        for k, v in r.info.items():
            body += _enc_ints([dict_index[k]])
            if info_types[k] == "Flag":
                body += bytes([0x00])
            elif info_types[k] == "Integer":
                body += _enc_ints([int(x) for x in v])
            else:
                body += _enc_str(",".join(str(x) for x in v).encode())
        out.append(_struct.pack("<II", len(body), 0) + body)
    _bgzf_write(path, b"".join(out))


def _dec_typed(data: bytes, o: int):
    """-> (values, new offset): list of ints / floats, bytes for char vectors, [] for MISSING (flags)."""
    d = data[o]
    o += 1
    count, typ = d >> 4, d & 15
    if count == 15:
        (n,), o = _dec_typed(data, o)
        count = n
    if typ == 0:
        return [], o
    if typ == 7:
        return data[o:o + count], o + count
    if typ == 5:
        return list(_struct.unpack_from(f"<{count}f", data, o)), o + 4 * count
    fmt, size, missing = _BCF_INT[typ]
    vals = [v for v in _struct.unpack_from(f"<{count}{fmt}", data, o) if v != missing and v != missing + 1]  # missing / end-of-vector
    return vals, o + size * count


def read_candidate_bcf(path: str, label_to_index: Dict[str, int]) -> List[CandidateRecord]:
    """Site records of a BCF2 file as CandidateRecords; chromosomes are mapped by NAME to the caller's chromosome list
    (bcf_chrom_index_map, :566)."""
    with _gzip.open(path, "rb") as f:
        data = f.read()
    assert data[:5] == b"BCF\x02\x02", "not a BCF2 file"
    (l_text,) = _struct.unpack_from("<I", data, 5)
    text = data[9:9 + l_text].rstrip(b"\0").decode()
    # dictionaries: strings of FILTER / INFO / FORMAT lines in order of first appearance, PASS first; IDX= wins when present
    names: Dict[int, str] = {0: "PASS"}
    seen = {"PASS": 0}
    contigs: List[str] = []
    types: Dict[str, str] = {}
    for line in text.split("\n"):
        m = _re.match(r"##(FILTER|INFO|FORMAT)=<(.*)>", line)
        if m:
            attrs = dict(_re.findall(r"(\w+)=(\"[^\"]*\"|[^,]*)", m.group(2)))
            name = attrs["ID"]
            if m.group(1) == "INFO":
                types[name] = attrs.get("Type", "String")
            if name not in seen:
                idx = int(attrs["IDX"]) if "IDX" in attrs else len(seen)
                seen[name] = idx
                names[idx] = name
            continue
        m = _re.match(r"##contig=<(.*)>", line)
        if m:
            attrs = dict(_re.findall(r"(\w+)=(\"[^\"]*\"|[^,]*)", m.group(1)))
            contigs.append(attrs["ID"])
    recs: List[CandidateRecord] = []
    o = 9 + l_text
    while o < len(data):
        l_shared, l_indiv = _struct.unpack_from("<II", data, o)
        p = o + 8
        chrom, pos, _rlen, _qual, n_allele_info, _n_fmt_sample = _struct.unpack_from("<iiifII", data, p)
        p += 24
        n_allele, n_info = n_allele_info >> 16, n_allele_info & 0xFFFF
        rid, p = _dec_typed(data, p)
        alleles = []
        for _ in range(n_allele):
            a, p = _dec_typed(data, p)
            alleles.append(bytes(a))
        _flt, p = _dec_typed(data, p)
        info: Dict[str, List[Union[int, str]]] = {}
        for _ in range(n_info):
            (k,), p = _dec_typed(data, p)
            v, p = _dec_typed(data, p)
            key = names[k]
            if isinstance(v, (bytes, bytearray)):
                info[key] = [x for x in bytes(v).decode().split(",")] if v else []
            else:
                info[key] = list(v)
        recs.append(CandidateRecord(label_to_index[contigs[chrom]], pos, bytes(rid).decode(), alleles, info))
        o += 8 + l_shared + l_indiv
    return recs
