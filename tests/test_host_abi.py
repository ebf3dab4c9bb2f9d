"""CPU: the C-ABI library loads, exports every symbol include/gappadder_b200.h declares, and its
host-side half (packing, candidate filter, significance, merged strings, reverse complement)
matches the golden vectors / the oracle.  No compute calls that need a GPU."""
import json
import ctypes as C
import os
import re

import numpy as np
import pytest

import gappadder_b200 as g
from gappadder_b200 import capi
import _oracle
import synth_gaps

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden")


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "gappadder_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(gp_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(capi.EXPORTS), declared ^ set(capi.EXPORTS)
    L = g.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert L.gp_abi_version() == 5


def test_affine_defaults_are_the_reference_parameters():
    """gp_affine_params_terefiner = aln_param_blast (TERefiner/algorithms/local_alignment.cpp:193-206); layouts of the two
    affine structs as the header declares them."""
    p = capi.AffineParams()
    g.lib().gp_affine_params_terefiner(C.byref(p))
    assert (p.match, p.mismatch, p.n_score, p.gap_open, p.gap_ext, p.band_width) == (1, -3, -2, 5, 2, 50)
    t = capi.TEREFINER_AFFINE
    assert (t.match, t.mismatch, t.n_score, t.gap_open, t.gap_ext, t.band_width) == (1, -3, -2, 5, 2, 50)
    assert C.sizeof(capi.AffineParams) == 24 and capi.LOCAL_DTYPE.itemsize == 24
    assert capi.LOCAL_DTYPE.names == ("score", "start1", "end1", "start2", "end2", "flags")


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(g.GpError):
        g.Context(0)


def test_pack_layout_and_codes():
    seqs = [b"ACGTN", b"", b"AAAAAAAAAC", b"RRYA"]
    packed, off, lens, nsym = g.pack_sequences(seqs)
    assert list(lens) == [5, 0, 10, 4]
    assert all(o % 4 == 0 for o in off)           # 16-byte aligned starts
    def code(s, i):
        return (int(packed[off[s] + i // 8]) >> (4 * (i % 8))) & 15
    assert [code(0, i) for i in range(5)] == [0, 1, 2, 3, 4]
    assert [code(2, i) for i in range(10)] == [0] * 9 + [1]
    r, y = code(3, 0), code(3, 2)
    assert r == code(3, 1) and r >= 5 and y >= 5 and r != y and code(3, 3) == 0
    assert nsym == 7
    with pytest.raises(g.GpError):
        g.pack_sequences([bytes(range(65, 65 + 20))])   # 20 distinct symbols


def test_revcomp_golden():
    for c in json.load(open(os.path.join(G, "revcomp.json"))):
        assert g.revcomp(c["s"].encode()).decode() == c["rc"]


def test_significance_golden():
    data = json.load(open(os.path.join(G, "significant.json")))
    t = g.gappadder_thresholds()
    for c in data["cases"]:
        assert g.is_score_significant(t, c["score"], c["l1"], c["l2"], c["row"], c["col"], c["nclip"]) == c["res"], c


def test_epilogue_golden():
    """merged string / IsContainment / GetOverlapSize from the reference's DP outputs."""
    data = json.load(open(os.path.join(G, "evaluate.json")))
    n = 0
    for c in data["cases"]:
        if c["bcontained"] < 0:
            continue
        a, b = c["s1"].encode(), c["s2"].encode()
        r = dict(score=c["score"], row_end=c["row_end"], col_end=c["col_end"], nclip=c["nclip"],
                 flags=capi.FLAG_CONTAINED if c["bcontained"] else 0)
        assert g.merged_concat(a, b, r).decode() == c["merged"]
        assert int(capi.is_containment(len(a), len(b), r)) == c["is_containment"]
        assert capi.overlap_size(len(a), len(b), r) == c["overlap"]
        n += 1
    assert n > 300


def test_candidate_pairs_golden_and_oracle():
    for c in json.load(open(os.path.join(G, "quickcheck.json"))):
        a, b = c["si"].encode(), c["sj"].encode()
        if len(a) < 30:
            continue
        got = g.candidate_pairs([a, b], c["k"])
        pairs = {(int(p["row_seq"]), int(p["col_seq"])) for p in got}
        assert ((0, 1) in pairs) == bool(c["feasible"]), c
    for cfg, seed in (("tiny", 1), ("small", 3), ("noisy", 2), ("cfg1", 1)):
        recs = synth_gaps.make_gap(seed, synth_gaps.CONFIGS[cfg])
        nodes = []
        for _, s in recs:
            nodes += [s, g.revcomp(s)]
        got = [(int(p["row_seq"]), int(p["col_seq"])) for p in g.candidate_pairs(nodes, 10)]
        want = _oracle.oracle_candidate_pairs(nodes, 10)
        assert got == want          # same pairs, same (-t 1) order
        assert got == sorted(got)


def _reference_pack(seqs):
    """Plain restatement of gp_pack_sequences' layout: 4-bit codes (A C G T N = 0..4, other bytes in order of first
    appearance), eight per word, every sequence padded to a multiple of 4 words and at least one block."""
    code = {ord("A"): 0, ord("C"): 1, ord("G"): 2, ord("T"): 3, ord("N"): 4}
    words, offs = [], []
    for s in seqs:
        offs.append(len(words))
        nw = ((len(s) + 7) // 8 + 3) & ~3
        w = [0] * (nw or 4)
        for i, ch in enumerate(s):
            if ch not in code:
                code[ch] = len(code)
            w[i // 8] |= (code[ch] & 15) << (4 * (i % 8))
        words += w
    return np.array(words, dtype=np.uint32), offs


def test_pack_matches_reference_layout_all_lengths_and_alphabets():
    """The packer's vector path (16 bases at a time) and its table path must agree with the plain layout for every
    length around the 8/16-base boundaries, with N, with other letters, and on the threaded large-batch path."""
    import random
    rnd = random.Random(5)
    for _ in range(150):
        alph = rnd.choice([b"ACGT", b"ACGTN", b"ACGTNRY", b"AC"])
        seqs = [bytes(rnd.choice(alph) for _ in range(rnd.choice([0, 1, 7, 8, 15, 16, 17, 31, 32, 33, 100, 257])))
                for _ in range(rnd.randint(1, 6))]
        packed, off, lens, nsym = g.pack_sequences(seqs)
        want, woff = _reference_pack(seqs)
        assert list(off) == woff
        assert np.array_equal(packed, want)
        assert (nsym <= 4) == all(ch in b"ACGT" for s in seqs for ch in s)
    seqs = [bytes(rnd.choice(b"ACGT") for _ in range(3000)) for _ in range(1500)]      # > 4 Mbases: several threads
    seqs[777] = seqs[777][:100] + b"N" + seqs[777][101:]
    packed, off, lens, nsym = g.pack_sequences(seqs)
    want, woff = _reference_pack(seqs)
    assert np.array_equal(packed, want) and nsym == 5
