"""Generates tests/golden/wfa_golden.json.

The reference (Rust + WFA2-lib) cannot be run in this container, so apart from the two known answers copied
from the reference's own unit tests (src/wfa2_utils.rs:280-308) these vectors are REGRESSION fixtures produced
by the oracle's LITERAL mode (oracle/wfa_oracle.c), each also checked against the independent full-DP
optimum.  They pin the oracle and the CUDA path to each other and to this committed history.

Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from sawfish_b200 import synth  # noqa: E402


def main():
    out = {"reference_known_answers": [
        {"source": "src/wfa2_utils.rs:294-308", "preset": "consensus", "text": "AACTGTATAA", "pattern": "ACTTAT",
         "score": 4, "ref_offset": 1, "cigar": "3=1D3="},
        {"source": "src/wfa2_utils.rs:280-292", "wfa_ops_reversed": "DDDDMMMXXMMMMIIIMMMM", "ref_offset": 0,
         "cigar": "4S3=2X4=3D4="},
    ], "oracle_vectors": []}
    cases = []
    arena, pairs = synth.random_pairs(120, seed=424242, max_len=90)
    cases.append(("random", arena, pairs))
    arena, pairs, _ = synth.config2b_score_sv_flanks(12, seed=77)
    cases.append(("score_sv_flanks", arena, pairs))
    arena, pairs, _ = synth.config2_read_vs_haplotypes(3, seed=78, len_lo=600, len_hi=1200, sv_hi=300)
    cases.append(("reads_vs_haps", arena, pairs))
    for name, arena, pairs in cases:
        for preset, pen in (("consensus", O.CONSENSUS), ("large_indel", O.LARGE_INDEL)):
            for pr in pairs:
                p = arena[int(pr["pattern_off"]):int(pr["pattern_off"]) + int(pr["pattern_len"])].tobytes()
                t = arena[int(pr["text_off"]):int(pr["text_off"]) + int(pr["text_len"])].tobytes()
                if pr["flags"] & 1:
                    p, t = p[::-1], t[::-1]
                ends = [int(pr[k]) for k in ("pattern_begin_free", "pattern_end_free", "text_begin_free", "text_end_free")]
                r = O.wfa_align(pen, p, t, *ends, mode=O.MODE_LITERAL)
                dp = O.dp_min_penalty(pen, p, t, *ends)
                assert dp == r.stats["wf_score"], (name, preset)
                roff, cig, hm = O.translate_cigar(r.ops)
                out["oracle_vectors"].append({
                    "case": name, "preset": preset, "pattern": p.decode(), "text": t.decode(), "ends_free": ends,
                    "wf_score": r.stats["wf_score"], "score": r.score, "ops": r.ops.decode(), "k0": r.stats["k0"],
                    "ref_offset": roff, "cigar": O.cigar_to_string(cig) if hm else None})
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "wfa_golden.json")
    with open(path, "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print(path, len(out["oracle_vectors"]), "vectors", os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
