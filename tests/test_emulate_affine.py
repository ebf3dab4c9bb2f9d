"""CPU: the affine local aligner's device functions (gappadder_b200/csrc/affine_local.cuh) run on the host by
tests/emulate_affine.cu -- the forward kernel's per-lane step with 32 lanes in lock step, the start recovery both in the
reference's order of evaluation (aff_epilogue) and as one warp computes it on the device (aff_epilogue_warp, every phase
looped over the lanes) -- against the golden vectors made by the reference's own
TERefiner/algorithms/local_alignment.cpp (tests/golden/local_affine.json) and, when oracle/_ref is present, against that
code live on fresh random pairs."""
import ctypes as C
import json
import os
import random
import shutil
import subprocess

import pytest

import _oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# gp_local_affine_batch's letter classes (aln_nt4_table): a/A c/C g/G t/T are the bases, everything else one class
CODE = bytes([{65: 0, 67: 1, 71: 2, 84: 3, 97: 0, 99: 1, 103: 2, 116: 3}.get(b, 4) for b in range(256)])
BLAST = (1, -3, -2, 5, 2, 50)


@pytest.fixture(scope="module")
def emu():
    if shutil.which("nvcc") is None:
        pytest.skip("nvcc not available")
    os.makedirs(os.path.join(ROOT, "build"), exist_ok=True)
    so = os.path.join(ROOT, "build", "libemulate_affine.so")
    srcs = [os.path.join(ROOT, "tests", "emulate_affine.cu"), os.path.join(ROOT, "gappadder_b200", "csrc", "affine_local.cuh"),
            os.path.join(ROOT, "gappadder_b200", "csrc", "common.cuh")]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["nvcc", "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "-shared",
                               "-o", so, srcs[0]])
    L = C.CDLL(so)
    L.aff_emulate_forward.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int32)]
    L.aff_host_epilogue.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_int), C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int32)]
    L.aff_host_epilogue_warp.argtypes = L.aff_host_epilogue.argtypes

    def run(a, b, params=BLAST):
        """-> (score, start1, end1, start2, end2, flags) or None when nothing aligns."""
        P = (C.c_int * 6)(*params)
        ca, cb = a.translate(CODE), b.translate(CODE)
        f = (C.c_int32 * 3)()
        assert L.aff_emulate_forward(ca, len(a), cb, len(b), P, f) == 0
        if f[0] <= 0:
            return None
        o = (C.c_int32 * 6)()
        assert L.aff_host_epilogue(ca, len(a), cb, len(b), P, f[0], f[1], f[2], o) == 0
        assert (o[2], o[4]) == (f[1], f[2])
        ow = (C.c_int32 * 6)()                  # what the device runs: one warp per pair
        assert L.aff_host_epilogue_warp(ca, len(a), cb, len(b), P, f[0], f[1], f[2], ow) == 0
        assert tuple(ow) == tuple(o), (len(a), len(b), tuple(o), tuple(ow))
        return tuple(o)
    return run


def golden():
    with open(os.path.join(ROOT, "tests", "golden", "local_affine.json")) as f:
        return json.load(f)["cases"]


def test_golden_vectors(emu):
    n = 0
    for rec in golden():
        a, b = rec["s1"].encode(), rec["s2"].encode()
        got = emu(a, b)
        if rec["forward_score"] < 1:
            assert got is None
            continue
        want = (rec["score"], rec["start1"], rec["end1"], rec["start2"], rec["end2"])
        assert got is not None and got[:5] == want, (len(a), len(b), want, got)
        n += 1
    assert n > 300


def test_known_answer(emu):
    # 11 = 9 matches + a 1-base gap (5 + 2) ... checked by hand against the reference's printed alignment
    assert emu(b"ACGTACGTTTGACCAGTAGGATCCA", b"TTTTTGTACGTTTGACAGTAGGTTTT")[:5] == (11, 3, 13, 6, 16)
    assert emu(b"ACGT", b"ACGT")[:5] == (4, 1, 4, 1, 4)
    assert emu(b"NNNN", b"ACGT") is None
    assert emu(b"acgtacgt", b"ACGTACGT")[:5] == (8, 1, 8, 1, 8)


def _mutate(s, rate, rng):
    out = bytearray()
    for ch in s:
        x = rng.random()
        if x < rate / 3:
            continue
        if x < 2 * rate / 3:
            out.append(rng.choice(b"ACGT"))
        if x < rate:
            out.append(rng.choice(b"ACGT"))
            continue
        out.append(ch)
    return bytes(out)


def test_live_reference_random_pairs(emu):
    if _oracle.la_ref_lib() is None:
        pytest.skip("oracle/_ref/libla_ref.so not built (needs /root/reference)")
    rng = random.Random(77)
    n = 0
    for it in range(400):
        kind = it % 4
        if kind == 0:
            a = bytes(rng.choice(b"ACGT") for _ in range(rng.randrange(1, 150)))
            b = bytes(rng.choice(b"ACGT") for _ in range(rng.randrange(1, 150)))
        elif kind == 1:
            core = bytes(rng.choice(b"ACGT") for _ in range(rng.randrange(20, 700)))
            a = bytes(rng.choice(b"ACGT") for _ in range(rng.randrange(0, 200))) + core
            b = _mutate(core, rng.choice([0.0, 0.03, 0.1]), rng) + bytes(rng.choice(b"ACGT") for _ in range(rng.randrange(0, 200)))
        elif kind == 2:
            unit = bytes(rng.choice(b"ACGT") for _ in range(rng.randrange(1, 10)))
            a, b = _mutate(unit * rng.randrange(4, 60), 0.04, rng), _mutate(unit * rng.randrange(4, 60), 0.04, rng)
        else:
            core = bytes(rng.choice(b"ACGTNacgt") for _ in range(rng.randrange(10, 400)))
            a, b = core, _mutate(core, 0.06, rng)
        if not a or not b:
            continue
        want = _oracle.ref_local_affine(a, b)
        got = emu(a, b)
        assert (got[:5] if got else None) == want, (kind, len(a), len(b), want, got)
        n += 1
    assert n > 380
