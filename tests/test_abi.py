"""CPU tests of the C-ABI boundary: the library loads, exports every symbol the header declares, and its
host-only entry points (CIGAR translation / expansion) agree with the oracle.  No GPU compute here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from sawfish_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "sawfish_cuda.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(sfc_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.EXPORTS)
    L = C.CDLL(_lib.SO_PATH)
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} not exported"
    assert L.sfc_abi_version() == 1


def test_struct_layouts_match_header():
    from sawfish_b200.batch import PAIR_DTYPE
    from sawfish_b200._lib import Penalties
    assert PAIR_DTYPE.itemsize == 48 and C.sizeof(Penalties) == 28
    assert [PAIR_DTYPE.fields[n][1] for n in PAIR_DTYPE.names] == [0, 8, 16, 20, 24, 28, 32, 36, 40]


def test_no_gpu_means_loud_failure():
    """Without a CUDA device there is no fallback: context creation fails with SFC_ERR_NO_DEVICE."""
    from sawfish_b200 import _lib, SfcError
    from sawfish_b200.batch import AlignContext
    if _lib.lib().sfc_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(SfcError) as ei:
        AlignContext(0)
    assert ei.value.code == -2


def test_product_package_does_not_import_oracle():
    import sawfish_b200
    pkg = os.path.dirname(sawfish_b200.__file__)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "wfa_oracle" not in src and "from oracle" not in src and "import oracle" not in src, f


def test_translate_cigar_matches_reference_vector_and_oracle(oracle):
    from sawfish_b200.batch import translate_cigar, translate_rle, expand_ops
    from sawfish_b200.simple_alignment import cigar_from_u32, cigar_to_string
    from helpers import rle_of_bytes
    roff, cig, hm = translate_cigar(b"DDDDMMMXXMMMMIIIMMMM"[::-1])   # reference src/wfa2_utils.rs:280-292
    assert hm and roff == 0 and cigar_to_string(cigar_from_u32(cig)) == "4S3=2X4=3D4="
    rng = np.random.default_rng(3)
    for _ in range(500):
        n = int(rng.integers(0, 40))
        ops = bytes(rng.choice(np.frombuffer(b"MXIDMMMM", np.uint8), n))
        want = oracle.translate_cigar(ops)
        got = translate_cigar(ops)
        assert got[0] == want[0] or not want[2]
        assert list(got[1]) == want[1] and got[2] == want[2]
        rle = rle_of_bytes(ops)
        assert expand_ops(rle) == ops
        got2 = translate_rle(rle)
        assert got2[0] == got[0] and list(got2[1]) == list(got[1]) and got2[2] == got[2]


def test_process_wfa2_cigar_string_mirror():
    from sawfish_b200.wfa2_utils import process_wfa2_cigar_string
    from sawfish_b200.simple_alignment import cigar_to_string
    al = process_wfa2_cigar_string(b"DDDDMMMXXMMMMIIIMMMM"[::-1])
    assert al is not None and al.ref_offset == 0 and cigar_to_string(al.cigar) == "4S3=2X4=3D4="
    assert process_wfa2_cigar_string(b"DDII") is None


def test_rust_crate_binds_only_declared_symbols_with_matching_arity():
    """rust/sawfish-cuda (source only: no Rust toolchain here) must bind symbols the header declares, with the same
    number of parameters, and its #[repr(C)] structs must list the header's fields in the header's order."""
    hdr = open(os.path.join(ROOT, "include", "sawfish_cuda.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    rs = open(os.path.join(ROOT, "rust", "sawfish-cuda", "src", "lib.rs")).read()
    rs = re.sub(r"//[^\n]*", "", rs)

    def arity(text, name, opener):
        m = re.search(r"\b" + name + r"\s*\(", text)
        assert m, name
        depth, i = 1, m.end()
        start = i
        while depth:
            depth += {"(": 1, ")": -1}.get(text[i], 0)
            i += 1
        args = text[start:i - 1].strip()
        return 0 if args in ("", "void") else args.count(",") + 1

    block = rs[rs.index('extern "C" {'):]
    block = block[:block.index("\n}")]
    bound = re.findall(r"pub fn (sfc_[a-z0-9_]+)", block)
    assert len(bound) >= 15
    for name in bound:
        assert arity(block, name, "(") == arity(hdr, name, "("), name

    def c_fields(struct):
        body = dict((n, b) for b, n in re.findall(r"typedef struct \{([^{}]*)\}\s*([a-z0-9_]+);", hdr))[struct]
        names = []
        for decl in body.split(";"):
            decl = decl.strip()
            if decl:
                names += [n.strip().rstrip("_") for n in re.sub(r"^[a-z0-9_]+\s+", "", decl).split(",")]
        return names

    def rs_fields(struct):
        body = re.search(r"pub struct " + struct + r"\s*\{(.*?)\}", rs, flags=re.S).group(1)
        return [n.rstrip("_") for n in re.findall(r"pub ([a-z0-9_]+):", body)]

    assert rs_fields("SfcPenalties") == c_fields("sfc_penalties")
    assert rs_fields("SfcPair") == c_fields("sfc_pair")
    assert rs_fields("SfcBioScoring") == c_fields("sfc_bio_scoring")


def test_rust_patches_use_only_the_crate_surface():
    """rust/sawfish-cuda/patches/*.rs (the two-phase score_sv / refine_sv batchers, source only) may use the crate's public
    items only: no raw `sfc_*` FFI symbol, and every crate item they name exists in lib.rs."""
    pdir = os.path.join(ROOT, "rust", "sawfish-cuda", "patches")
    lib_rs = open(os.path.join(ROOT, "rust", "sawfish-cuda", "src", "lib.rs")).read()
    files = sorted(f for f in os.listdir(pdir) if f.endswith(".rs"))
    assert len(files) >= 2
    for f in files:
        src = re.sub(r"//[^\n]*", "", open(os.path.join(pdir, f)).read())
        assert not re.findall(r"\bsfc_[a-z_]+\s*\(", src), f
        for item in re.findall(r"use sawfish_cuda::\{([^}]*)\}", src):
            for name in [x.strip() for x in item.split(",")]:
                assert re.search(r"pub (struct|const|fn) " + name + r"\b", lib_rs), (f, name)
        for meth in re.findall(r"ctx\.([a-z_]+)\(", src):
            assert re.search(r"pub fn " + meth + r"\b", lib_rs), (f, meth)
