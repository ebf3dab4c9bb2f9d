"""CPU: the oracle (oracle/overlap_oracle.c) against golden vectors produced by the reference's own
code (tests/golden/make_golden.py -> oracle/_ref/libcm_ref.so), and against the live reference
library when it is present.  This is what pins the oracle."""
import json
import os
import random

import pytest

import _oracle
from _oracle import DPResult

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    with open(os.path.join(G, name)) as f:
        return json.load(f)


def test_evaluate_golden():
    data = _load("evaluate.json")
    lib = _oracle.oracle_lib()
    t = _oracle.gappadder_thresholds()
    n_checked = 0
    for c in data["cases"]:
        a, b = c["s1"].encode(), c["s2"].encode()
        for full in (True, False):
            o = _oracle.oracle_evaluate(a, b, -2, -2, 50, full=full)
            assert (o.score, o.row_end, o.col_end, o.nclip) == (c["score"], c["row_end"], c["col_end"], c["nclip"]), (c, full)
            sig = lib.gpo_is_score_significant(t, o.score, len(a), len(b), o.row_end, o.col_end, o.nclip)
            res = 2 if c["relax"] else sig
            assert res == c["res"]
            if c["bcontained"] >= 0:
                assert o.bcontained == c["bcontained"], (c, full)
                merged = _oracle.oracle_merged(a, b, o)
                assert merged.decode() == c["merged"]
                assert lib.gpo_is_containment(len(a), len(b), o.row_end, o.col_end, o.nclip, o.bcontained) == c["is_containment"]
                assert lib.gpo_overlap_size(len(a), len(b), o.nclip, len(merged)) == c["overlap"]
        n_checked += 1
    assert n_checked >= 500


def test_significant_golden():
    data = _load("significant.json")
    lib = _oracle.oracle_lib()
    t = _oracle.gappadder_thresholds()
    for c in data["cases"]:
        assert lib.gpo_is_score_significant(t, c["score"], c["l1"], c["l2"], c["row"], c["col"], c["nclip"]) == c["res"], c


def test_revcomp_golden():
    for c in _load("revcomp.json"):
        assert _oracle.oracle_revcomp(c["s"].encode()).decode() == c["rc"]


def test_quickcheck_golden():
    lib = _oracle.oracle_lib()
    for c in _load("quickcheck.json"):
        a, b = c["si"].encode(), c["sj"].encode()
        assert lib.gpo_quickcheck(a, len(a), b, len(b), c["k"]) == c["feasible"], c


def test_full_equals_rolling_on_ties():
    rng = random.Random(3)
    for _ in range(400):
        alpha = rng.choice([b"A", b"AC", b"ACG"])
        a = bytes(rng.choice(alpha) for _ in range(rng.randint(1, 60)))
        b = bytes(rng.choice(alpha) for _ in range(rng.randint(1, 60)))
        clip = rng.choice([0, 1, 5, 50])
        f = _oracle.oracle_evaluate(a, b, -2, -2, clip, full=True)
        r = _oracle.oracle_evaluate(a, b, -2, -2, clip, full=False)
        assert f.key() == r.key()
        assert (f.tb_row == 0, f.tb_col == 0) == (r.tb_row == 0, r.tb_col == 0)


@pytest.mark.skipif(_oracle.ref_lib() is None, reason="oracle/_ref not built (needs /root/reference)")
def test_oracle_vs_live_reference():
    rng = random.Random(99)
    lib = _oracle.oracle_lib()
    t = _oracle.gappadder_thresholds()
    for _ in range(800):
        alpha = rng.choice(["ACGT", "AC", "ACGTN", "A"])
        m, n = rng.randint(1, 90), rng.randint(1, 90)
        a = "".join(rng.choice(alpha) for _ in range(m)).encode()
        if rng.random() < 0.6:
            k = rng.randint(1, m)
            b = a[-k:] + "".join(rng.choice(alpha) for _ in range(max(0, n - k))).encode()
        else:
            b = "".join(rng.choice(alpha) for _ in range(n)).encode()
        b = b or b"A"
        relax = rng.random() < 0.5
        ref = _oracle.ref_evaluate(a, b, relax)
        o = _oracle.oracle_evaluate(a, b, full=rng.random() < 0.5)
        assert (o.score, o.row_end, o.col_end, o.nclip) == (ref["score"], ref["row_end"], ref["col_end"], ref["nclip"])
        sig = lib.gpo_is_score_significant(t, o.score, len(a), len(b), o.row_end, o.col_end, o.nclip)
        assert (2 if relax else sig) == ref["res"]
        if ref["bcontained"] >= 0:
            assert o.bcontained == ref["bcontained"]
            assert _oracle.oracle_merged(a, b, o) == ref["merged"]
        assert lib.gpo_quickcheck(a * 3 + b"ACGTACGTACGTACGTACGTACGTACGTACGT", 3 * len(a) + 32, b * 3 + b"TTGACCATGCATGCCGATTAGCAGGATCAT", 3 * len(b) + 30, 10) == \
            _oracle.ref_lib().cmref_quickcheck(a * 3 + b"ACGTACGTACGTACGTACGTACGTACGTACGT", b * 3 + b"TTGACCATGCATGCCGATTAGCAGGATCAT", 10)
