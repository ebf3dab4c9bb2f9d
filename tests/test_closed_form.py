"""CPU: the closed form the library uses for Evaluate(s, s) (csrc/gp_api.cu closed_form_self,
include/gappadder_b200.h gp_closed_form_stats) against the oracle, the reference build when present, and
the golden vectors: score m, ends (m, m), nclip 0, walk ends in the corner, bcontained -- whenever
mismatch <= 1 and indel <= 0."""
import json
import os
import random

import pytest

import _oracle
from _oracle import oracle_evaluate

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _closed(m):
    return (m, m, m, 0, 0, 0, 1)       # score, row_end, col_end, nclip, tb_row, tb_col, bcontained


def _oracle_tuple(s, mismatch, indel, clip, full):
    o = oracle_evaluate(s, s, mismatch, indel, clip, full=full)
    return (o.score, o.row_end, o.col_end, o.nclip, o.tb_row, o.tb_col, o.bcontained)


@pytest.mark.parametrize("mismatch,indel,clip", [(-2, -2, 50), (-1, -1, 0), (1, 0, 50), (0, 0, 3), (-3, -1, 200), (1, -1, 7), (-20, -30, 7)])
def test_self_pair_closed_form_matches_oracle(mismatch, indel, clip):
    rng = random.Random(1000 + clip)
    seqs = [b"A", b"N", b"AC", b"AAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAA", b"ACACACACACACACACACACACACAC",
            b"ACGTNNNNACGT", b"NNNNNNNNNN"]
    for alpha in (b"A", b"AC", b"ACGT", b"ACGTN", b"ACGTNRY"):
        for _ in range(12):
            seqs.append(bytes(rng.choice(alpha) for _ in range(rng.randint(1, 400))))
    for s in seqs:
        assert _oracle_tuple(s, mismatch, indel, clip, full=len(s) <= 120) == _closed(len(s)), (s[:40], mismatch, indel, clip)


def test_self_pair_closed_form_matches_reference_build():
    if _oracle.ref_lib() is None:
        pytest.skip("oracle/_ref not built here")
    rng = random.Random(5)
    for _ in range(20):
        s = bytes(rng.choice(b"ACGT") for _ in range(rng.randint(12, 600)))
        d = _oracle.ref_evaluate(s, s, relax=True)
        m = len(s)
        assert (d["score"], d["row_end"], d["col_end"], d["nclip"], d["bcontained"]) == (m, m, m, 0, 1)


def test_self_pairs_in_golden_vectors():
    data = json.load(open(os.path.join(ROOT, "tests", "golden", "evaluate.json")))
    n = 0
    for c in data["cases"]:
        if c["s1"] == c["s2"] and c["bcontained"] >= 0 and len(c["s1"]) >= 1:
            m = len(c["s1"])
            assert (c["score"], c["row_end"], c["col_end"], c["nclip"], c["bcontained"]) == (m, m, m, 0, 1)
            n += 1
    assert n >= 1
