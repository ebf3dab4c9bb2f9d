"""CPU: the packed 16-bit kernels' per-lane arithmetic (gappadder_b200/csrc/overlap_wf16.cuh and
overlap_wf16t.cuh), run lane by lane on the host by tests/emulate_wf16.cu, against the oracle and the
golden vectors.
Validates potential-domain clamping, the origin tags and the tie rule without a GPU."""
import ctypes as C
import json
import os
import random
import shutil
import subprocess

import pytest

import _oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = bytes.maketrans(b"ACGTN", bytes([0, 1, 2, 3, 4]))


@pytest.fixture(scope="module")
def emu():
    if shutil.which("nvcc") is None:
        pytest.skip("nvcc not available")
    os.makedirs(os.path.join(ROOT, "build"), exist_ok=True)
    so = os.path.join(ROOT, "build", "libemulate_wf16.so")
    srcs = [os.path.join(ROOT, "tests", "emulate_wf16.cu"), os.path.join(ROOT, "gappadder_b200", "csrc", "overlap_wf16.cuh"),
            os.path.join(ROOT, "gappadder_b200", "csrc", "overlap_wf16t.cuh"), os.path.join(ROOT, "gappadder_b200", "csrc", "overlap_wf16c.cuh"),
            os.path.join(ROOT, "gappadder_b200", "csrc", "common.cuh")]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["nvcc", "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "-shared",
                               "-o", so, srcs[0]])
    L = C.CDLL(so)
    L.wf16_emulate.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int32)]
    L.wf16t_emulate.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int32)]
    L.wf16c_emulate.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int32)]
    cert_status = [0, 0, 0]          # certificate kernel: certified by the first pass / by a later one / handed to an exact kernel

    def run(a, b, mm=-2, ind=-2, clip=50):
        res = []
        four = b"N" not in a and b"N" not in b
        for addsel in ((0, 1) if four else (0,)):      # both selector variants must agree
            out = (C.c_int32 * 5)()
            rc = L.wf16_emulate(a.translate(CODE), len(a), b.translate(CODE), len(b), mm, ind, clip, addsel, out)
            assert rc == 0
            f = out[4]
            res.append((out[0], out[1], out[2], out[3], f & 1, (f >> 1) & 1, (f >> 2) & 1))
        if four and len(b) <= 4094:                    # the table kernel's domain: A/C/G/T, column sequence <= 4094
            out = (C.c_int32 * 5)()
            assert L.wf16t_emulate(a.translate(CODE), len(a), b.translate(CODE), len(b), mm, ind, clip, out) == 0
            f = out[4]
            res.append((out[0], out[1], out[2], out[3], f & 1, (f >> 1) & 1, (f >> 2) & 1))
        # the certificate kernel's domain; every starting system, both layouts (+ 4: free moves), both orientations
        # (+ 8: transposed -- the computed table's rows are b, its columns a; results in the reference's orientation)
        variants = []
        for tr, cols in ((0, b), (8, a)):
            if four and len(cols) <= 16382:
                layouts = (0, 4) if (mm, ind) == (-2, -2) and len(cols) <= 3800 else (0,)
                variants += [s + l + tr for l in layouts for s in (0, 1, 2)]
        if variants:
            for first_sys in variants:
                out = (C.c_int32 * 5)()
                st = L.wf16c_emulate(a.translate(CODE), len(a), b.translate(CODE), len(b), mm, ind, clip, first_sys, out)
                assert st in (0, 1, 2)
                cert_status[st] += 1
                f = out[4]
                tup = (out[0], out[1], out[2], out[3], f & 1, (f >> 1) & 1, (f >> 2) & 1)
                if st == 2:                            # not certified: score, ends and clip are still exact
                    assert tup[:4] == res[0][:4], (tup, res[0])
                else:
                    res.append(tup)
        assert len(set(res)) == 1, res
        return res[0]
    run.cert_status = cert_status
    run.cert = L.wf16c_emulate
    return run


def _want(a, b, mm=-2, ind=-2, clip=50):
    o = _oracle.oracle_evaluate(a, b, mm, ind, clip)
    return (o.score, o.row_end, o.col_end, o.nclip, int(o.tb_row == 0), int(o.tb_col == 0), o.bcontained)


def test_golden_pairs(emu):
    data = json.load(open(os.path.join(ROOT, "tests", "golden", "evaluate.json")))
    for c in data["cases"]:
        if not c["relax"]:
            continue
        a, b = c["s1"].encode(), c["s2"].encode()
        got = emu(a, b)
        assert got[:4] == (c["score"], c["row_end"], c["col_end"], c["nclip"])
        assert got[6] == c["bcontained"]


@pytest.mark.parametrize("seed,maxlen,iters", [(7, 200, 1500), (8, 1400, 60), (9, 700, 150)])
def test_random_pairs(emu, seed, maxlen, iters):
    rng = random.Random(seed)

    def rnd(n, alpha):
        return bytes(rng.choice(alpha) for _ in range(n))
    for _ in range(iters):
        alpha = rng.choice([b"ACGT", b"AC", b"ACGTN", b"A"])
        m, n = rng.randint(1, maxlen), rng.randint(1, maxlen)
        a = rnd(m, alpha)
        if rng.random() < 0.6:
            k = rng.randint(1, min(m, maxlen // 2))
            b = a[-k:] + rnd(max(0, n - k), alpha)
            if rng.random() < 0.3:
                b = rnd(rng.randint(0, 10), alpha) + a + rnd(rng.randint(0, 10), alpha)
        else:
            b = rnd(n, alpha)
        b = b or b"A"
        clip = rng.choice([50, 50, 0, 3, 10])
        mm, ind = rng.choice([(-2, -2), (-2, -2), (-1, -1), (-3, -2), (0, -1), (-5, -3), (1, -1), (-15, -14)])
        assert emu(a, b, mm, ind, clip) == _want(a, b, mm, ind, clip), (len(a), len(b), clip, mm, ind)


@pytest.mark.parametrize("la,ov,extra", [(2700, 2500, 150), (4094, 4000, 200), (2000, 1990, 2200)])
def test_long_overlaps_both_potentials(emu, la, ov, extra):
    """High-scoring suffix/prefix overlaps that end in the last rows but not the last columns (and the
    transposed case): exercises the row filter under the drifting column potential."""
    rng = random.Random(la)
    a = bytes(rng.choice(b"ACGT") for _ in range(la))
    b = a[-ov:] + bytes(rng.choice(b"ACGT") for _ in range(extra))
    for x, y in ((a, b), (b, a)):
        assert emu(x, y) == _want(x, y)


def test_certificate_kernel_resolves_overlaps_and_long_columns(emu):
    """Certificate kernel: plain suffix/prefix overlaps are certified (first pass when the starting system fits,
    second pass otherwise), a sequence against itself never is, and columns beyond the tagged kernels' 4094
    limit stay in 16 bits."""
    rng = random.Random(5)
    a = bytes(rng.choice(b"ACGT") for _ in range(900))
    b = a[-400:] + bytes(rng.choice(b"ACGT") for _ in range(700))
    tr = bytes.maketrans(b"ACGT", bytes([0, 1, 2, 3]))

    def cert(x, y, first_sys):
        out = (C.c_int32 * 5)()
        st = emu.cert(x.translate(tr), len(x), y.translate(tr), len(y), -2, -2, 50, first_sys, out)
        f = out[4]
        return st, (out[0], out[1], out[2], out[3], f & 1, (f >> 1) & 1, (f >> 2) & 1)
    # a's suffix = b's prefix: the walk ends in column 0 -> system U (0) certifies at once, L needs the second pass
    assert cert(a, b, 0) == (0, _want(a, b))
    assert cert(a, b, 1) == (1, _want(a, b))
    # transposed: the walk ends in row 0
    assert cert(b, a, 1) == (0, _want(b, a))
    assert cert(b, a, 0) == (1, _want(b, a))
    # a sequence against itself: the walk ends in the corner; only system C certifies that
    assert cert(a, a, 2) == (0, _want(a, a))
    assert cert(a, a, 0) == (1, _want(a, a)) and cert(a, a, 1) == (1, _want(a, a))
    # a genuine tie between a walk that ends in row 0 and one that ends in column 0: no certificate, exact kernel
    t1, t2 = b"CACA", b"ACAC"
    assert {cert(t1, t2, fs)[0] for fs in (0, 1, 2)} <= {0, 1, 2}
    # long columns
    c = bytes(rng.choice(b"ACGT") for _ in range(5200))
    d = c[-3000:] + bytes(rng.choice(b"ACGT") for _ in range(3100))
    for x, y in ((c, d), (d, c)):
        w = _want(x, y)
        got = [cert(x, y, fs) for fs in (0, 1, 2)]
        assert sorted(st for st, _ in got) == [0, 1, 1] and all(t == w for _, t in got)


def test_free_moves_layout_at_its_range_edge(emu):
    """Column sequences of exactly 3800 bases (WF16C_POT2_MAX_N) with scores that reach the top of the 15-bit range:
    near-identical pairs (H ~ n in the last rows and columns), containment in a longer row sequence, random rows."""
    rnd = random.Random(11)

    def rs(n):
        return bytes(rnd.choice(b"ACGT") for _ in range(n))
    s = rs(3800)
    t = bytearray(s)
    t[1900] = ord("A") if t[1900] != ord("A") else ord("C")
    for a, b in [(bytes(t), s), (rs(600) + s + rs(600), s), (s, s[:3799] + b"G"), (rs(5000), s)]:
        assert emu(a, b) == _want(a, b)
