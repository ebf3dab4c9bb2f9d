"""CPU: the oracle's restatement of TERefiner's affine local aligner (oracle/local_affine_oracle.c) against the golden
vectors made by the reference's own code (tests/golden/local_affine.json, tests/golden/make_golden_affine.py) and, when
oracle/_ref/libla_ref.so is present, against that code live -- including LocalAlignment::optAlign itself."""
import ctypes as C
import json
import os
import random

import pytest

import _oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_oracle_against_golden_vectors():
    with open(os.path.join(ROOT, "tests", "golden", "local_affine.json")) as f:
        recs = json.load(f)["cases"]
    assert len(recs) > 300
    for rec in recs:
        a, b = rec["s1"].encode(), rec["s2"].encode()
        got = _oracle.oracle_local_affine(a, b)
        if rec["forward_score"] < 1:
            assert got is None
        else:
            assert got == (rec["score"], rec["start1"], rec["end1"], rec["start2"], rec["end2"]), (len(a), len(b))


def test_oracle_against_live_reference():
    ref = _oracle.la_ref_lib()
    if ref is None:
        pytest.skip("oracle/_ref/libla_ref.so not built (needs /root/reference)")
    rng = random.Random(123)
    for it in range(300):
        n = rng.randrange(1, 400)
        a = bytes(rng.choice(b"ACGTacgtN") for _ in range(n))
        cut = rng.randrange(n)
        b = bytearray(a[cut:] + a[:cut] if it % 3 == 0 else a)
        for _ in range(rng.randrange(0, 1 + n // 8)):
            k = rng.randrange(len(b))
            if rng.random() < 0.5:
                b[k] = rng.choice(b"ACGT")
            elif rng.random() < 0.5:
                del b[k]
            else:
                b.insert(k, rng.choice(b"ACGT"))
            if not b:
                b = bytearray(b"A")
        b = bytes(b)
        want = _oracle.ref_local_affine(a, b)
        assert _oracle.oracle_local_affine(a, b) == want, (it, len(a), len(b))
        if want is not None:                     # the public method TERefiner's callers use (main.cpp:209-212)
            o = (C.c_int32 * 4)()
            ref.laref_opt_align(a, b, o)
            assert tuple(o) == want[1:]
