"""GPU: TERefiner's affine local aligner (gp_local_affine_batch, gappadder_b200/csrc/affine_local.cuh) through the C ABI
against the reference's own code: the golden vectors made by TERefiner/algorithms/local_alignment.cpp
(tests/golden/local_affine.json) and, when oracle/_ref/libla_ref.so travelled with the snapshot, that code live on
cfg1-shaped contig pairs.  Parity is PINNED here (unlike the BWA-defined flank placement): score and all four
coordinates must be identical."""
import json
import os
import random
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

import gappadder_b200 as g
import synth_gaps
import _oracle

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _tuple(r):
    return (int(r["score"]), int(r["start1"]), int(r["end1"]), int(r["start2"]), int(r["end2"]))


def _golden():
    with open(os.path.join(ROOT, "tests", "golden", "local_affine.json")) as f:
        return json.load(f)["cases"]


def test_golden_vectors_in_one_batch(ctx):
    recs = _golden()
    seqs, pairs = [], []
    for rec in recs:
        seqs += [rec["s1"].encode(), rec["s2"].encode()]
        pairs.append((len(seqs) - 2, len(seqs) - 1))
    res = ctx.local_affine_batch(seqs, pairs)
    assert len(res) == len(recs)
    bad = []
    for rec, r in zip(recs, res):
        if rec["forward_score"] < 1:
            ok = int(r["flags"]) & 1 and int(r["score"]) == 0
        else:
            ok = _tuple(r) == (rec["score"], rec["start1"], rec["end1"], rec["start2"], rec["end2"]) and int(r["flags"]) == 0
        if not ok:
            bad.append((len(rec["s1"]), len(rec["s2"]), rec.get("score"), rec.get("start1"), rec.get("end1"), rec.get("start2"), rec.get("end2"), r))
    assert not bad, "%d of %d differ; first: %r" % (len(bad), len(recs), bad[:5])
    st = ctx.local_affine_stats()
    assert st["cells"] == sum(len(a["s1"]) * len(a["s2"]) for a in recs) and st["forward_ms"] > 0 and st["epilogue_ms"] > 0


def test_split_form_relaunch_and_pair_order(ctx):
    """Upload once, launch twice, fetch: same results; results come back in the caller's pair order, not the kernel's."""
    # the split form works on the caller's table, where only upper-case A C G T are bases (the one-call form folds case)
    recs = [r for r in _golden() if r["forward_score"] >= 1 and r["s1"].isupper() and r["s2"].isupper()][:80]
    seqs = []
    for rec in recs:
        seqs += [rec["s1"].encode(), rec["s2"].encode()]
    order = list(range(len(recs)))
    random.Random(5).shuffle(order)
    pairs = np.zeros(len(order), dtype=g.capi.PAIR_DTYPE)
    pairs["row_seq"] = [2 * k for k in order]
    pairs["col_seq"] = [2 * k + 1 for k in order]
    ctx.upload_host_sequences(g.capi.HostBatch(seqs, []))
    ctx.local_affine_upload_pairs(pairs)
    ctx.local_affine_launch()
    ctx.local_affine_launch()
    res = ctx.local_affine_fetch()
    for k, r in zip(order, res):
        rec = recs[k]
        assert _tuple(r) == (rec["score"], rec["start1"], rec["end1"], rec["start2"], rec["end2"])


def test_empty_lowercase_and_other_letters(ctx):
    seqs = [b"", b"ACGT", b"acgtacgtac", b"ACGTACGTAC", b"ACGTNNNNACGTACGTAAC", b"ACGTRYKMACGTACGTAAC", b"NNNN"]
    pairs = [(0, 1), (1, 0), (2, 3), (4, 5), (6, 1), (1, 1)]
    res = ctx.local_affine_batch(seqs, pairs)
    assert int(res[0]["score"]) == -1 and int(res[0]["flags"]) == 1 and int(res[1]["score"]) == -1      # aln_local_core :545
    assert _tuple(res[2]) == (10, 1, 10, 1, 10)                                                         # case folds (aln_nt4_table)
    assert int(res[4]["score"]) == 0 and int(res[4]["flags"]) == 1                                      # N against anything: -2
    assert _tuple(res[5]) == (4, 1, 4, 1, 4)
    if _oracle.la_ref_lib() is not None:
        assert _tuple(res[3]) == _oracle.ref_local_affine(seqs[4], seqs[5])
    assert len(ctx.local_affine_batch(seqs, [])) == 0


def test_range_and_parameter_errors(ctx):
    long_a = b"A" * 40000
    with pytest.raises(g.GpError) as e:
        ctx.local_affine_batch([long_a, long_a], [(0, 1)])
    assert e.value.code == -5
    ok = ctx.local_affine_batch([long_a, b"ACGTAAAAAAAAAAAAAAAAAAAAAAAAAAAAA"], [(0, 1)])       # the shorter side bounds the score
    assert int(ok[0]["score"]) == 29
    with pytest.raises(g.GpError) as e:
        ctx.local_affine_batch([b"ACGT", b"ACGT"], [(0, 1)], g.AffineParams(0, -3, -2, 5, 2, 50))
    assert e.value.code == -1
    with pytest.raises(g.GpError):
        ctx.local_affine_batch([b"ACGT", b"ACGT"], [(0, 5)])


def test_other_parameters_against_the_host_compiled_device_functions(ctx):
    """Parameters other than aln_param_blast have no reference caller; the same __host__ __device__ functions compiled for
    the host (tests/emulate_affine.cu) must give what the device gives."""
    import ctypes as C
    import subprocess
    so = os.path.join(ROOT, "build", "libemulate_affine.so")
    if not os.path.exists(so):
        pytest.skip("build/libemulate_affine.so not built (CPU tests build it)")
    L = C.CDLL(so)
    L.aff_emulate_forward.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int32)]
    L.aff_host_epilogue.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_int), C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int32)]
    code = bytes([{65: 0, 67: 1, 71: 2, 84: 3}.get(b, 4) for b in range(256)])
    rng = random.Random(9)
    seqs, pairs = [], []
    for _ in range(60):
        core = bytes(rng.choice(b"ACGT") for _ in range(rng.randrange(20, 900)))
        seqs += [bytes(rng.choice(b"ACGT") for _ in range(rng.randrange(0, 200))) + core,
                 bytes(ch if rng.random() > 0.05 else rng.choice(b"ACGT") for ch in core) + bytes(rng.choice(b"ACGT") for _ in range(rng.randrange(0, 200)))]
        pairs.append((len(seqs) - 2, len(seqs) - 1))
    for params in ((2, -3, -1, 4, 1, 20), (1, -1, -1, 11, 1, 50), (5, -4, -2, 10, 3, 8)):
        res = ctx.local_affine_batch(seqs, pairs, g.AffineParams(*params))
        P = (C.c_int * 6)(*params)
        for (a, b), r in zip(pairs, res):
            ca, cb = seqs[a].translate(code), seqs[b].translate(code)
            f = (C.c_int32 * 3)()
            assert L.aff_emulate_forward(ca, len(ca), cb, len(cb), P, f) == 0
            o = (C.c_int32 * 6)()
            assert L.aff_host_epilogue(ca, len(ca), cb, len(cb), P, f[0], f[1], f[2], o) == 0
            assert _tuple(r) == tuple(o[:5]) and int(r["flags"]) == o[5], (params, len(ca), len(cb), tuple(o), r)


def test_cfg1_contig_pairs_against_the_live_reference(ctx):
    """Two cfg1 gaps (BASELINE configs[0] shape): every candidate pair of the quick check, aligned by the reference's own
    aln_stdaln on the host cores and by the kernels."""
    if _oracle.la_ref_lib() is None:
        pytest.skip("oracle/_ref/libla_ref.so not present")
    seqs, pairs = [], []
    for seed in (11, 12):
        nodes = []
        for _, s in synth_gaps.make_gap(seed, synth_gaps.CONFIGS["cfg1"]):
            nodes += [s, g.revcomp(s)]
        base = len(seqs)
        seqs += nodes
        pairs += [(base + a, base + b) for a, b in g.candidate_pairs(nodes, 10)][:700]

    def one(ab):
        return _oracle.ref_local_affine(seqs[ab[0]], seqs[ab[1]])
    with ThreadPoolExecutor(max_workers=min(32, os.cpu_count() or 1)) as ex:
        want = list(ex.map(one, pairs))
    res = ctx.local_affine_batch(seqs, pairs)
    bad = [(ab, len(seqs[ab[0]]), len(seqs[ab[1]]), w, _tuple(r), int(r["flags"])) for ab, r, w in zip(pairs, res, want)
           if (w is None and not int(r["flags"]) & 1) or (w is not None and w != _tuple(r))]
    assert not bad, "%d of %d differ; first: %r" % (len(bad), len(pairs), bad[:5])
    assert sum(w is not None and w[0] > 200 for w in want) > 20          # real overlaps are in the sample
