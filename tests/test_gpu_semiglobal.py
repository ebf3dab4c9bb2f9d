"""GPU: flank placement (gp_semiglobal_batch, gappadder_b200/csrc/flank_place.cuh) through the C ABI against the oracle's
builder-written semi-global definition (gpo_semiglobal).  BWA parity is UNPINNED at this boundary: bit-exact
here means against that definition, which tests/test_semiglobal_oracle.py checks against a brute-force statement."""
import os
import random
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

import gappadder_b200 as g
import synth_gaps
from _oracle import oracle_semiglobal, oracle_revcomp

pytestmark = pytest.mark.gpu


def _check(ctx, seqs, pairs, params=None):
    params = params or g.GAPPADDER_DP

    def one(ab):
        o = oracle_semiglobal(seqs[ab[0]], seqs[ab[1]], params.mismatch, params.indel)
        return (o.score, o.col_start, o.col_end)
    with ThreadPoolExecutor(max_workers=min(32, os.cpu_count() or 1)) as ex:
        want = list(ex.map(one, pairs))
    res = ctx.semiglobal_batch(seqs, pairs, params)
    assert len(res) == len(pairs)
    bad = [(ab, len(seqs[ab[0]]), len(seqs[ab[1]]), w, (int(r["score"]), int(r["col_start"]), int(r["col_end"])))
           for ab, r, w in zip(pairs, res, want) if w != (int(r["score"]), int(r["col_start"]), int(r["col_end"]))]
    assert not bad, "first mismatches (pair, m, n, oracle, gpu): %r" % bad[:5]
    return res


def _rand(rng, n, alpha=b"ACGT"):
    return bytes(rng.choice(alpha) for _ in range(n))


def test_known_answer_and_empty(ctx):
    seqs = [b"ACGTACGT", b"TTTTACGTACGTTTT", b"", b"A", b"ACG", b"GGACGTACGTACGTCC", b"ACGTTCGT"]
    pairs = [(i, j) for i in range(len(seqs)) for j in range(len(seqs))]
    res = _check(ctx, seqs, pairs)
    r = res[pairs.index((0, 1))]
    assert (int(r["score"]), int(r["col_start"]), int(r["col_end"])) == (8, 4, 12)
    assert len(ctx.semiglobal_batch(seqs, [])) == 0


def test_cfg2_shape_flanks_against_contigs_and_reverse_complements(ctx):
    """BASELINE configs[1]: the two 995-base flanks of a gap against its 40 contigs and their reverse complements."""
    seqs, pairs = [], []
    for seed in (1, 2, 3):
        spec = synth_gaps.CONFIGS["cfg1"]
        base = len(seqs)
        flanks = [s for _, s in synth_gaps.make_flanks(seed, spec)]
        seqs += flanks
        for _, s in synth_gaps.make_gap(seed, spec):
            seqs += [s, g.revcomp(s)]
            pairs += [(base, len(seqs) - 2), (base, len(seqs) - 1), (base + 1, len(seqs) - 2), (base + 1, len(seqs) - 1)]
    res = _check(ctx, seqs, pairs)
    assert (res["flags"] == 1).all()                       # pure A/C/G/T: the shared-memory-table kernel
    assert int((res["score"] > 900).sum()) >= 1             # some contig contains a whole flank
    st = ctx.semiglobal_stats()
    assert st["table_pairs"] == len(pairs) and st["cells"] == sum(len(seqs[a]) * len(seqs[b]) for a, b in pairs)


def test_strip_boundaries_ties_and_other_letters(ctx):
    rng = random.Random(17)
    base = _rand(rng, 6000)
    seqs = []
    for L in (1, 15, 16, 17, 511, 512, 513, 1023, 1024, 1025, 1600):       # flank lengths around the 512-row strips
        st = rng.randrange(0, len(base) - L)
        s = bytearray(base[st:st + L])
        for p in range(L):
            if rng.random() < 0.01:
                s[p] = rng.choice(b"ACGT")
        seqs.append(bytes(s))
    nf = len(seqs)
    for L in (1, 31, 32, 33, 64, 700, 3000, 6000):
        st = rng.randrange(0, len(base) - L + 1)
        seqs.append(base[st:st + L])
    seqs.append(oracle_revcomp(base[:2000]))
    pairs = [(i, j) for i in range(nf) for j in range(nf, len(seqs))]
    _check(ctx, seqs, pairs)
    # tie-heavy small alphabets, and N / IUPAC letters (compare-per-cell kernel; N == N is a match, as in Evaluate)
    seqs = []
    for _ in range(60):
        seqs.append(_rand(rng, rng.randint(1, 120), rng.choice([b"A", b"AC", b"ACG", b"ACGTN", b"ACGTNRY"])))
    pairs = [(rng.randrange(60), rng.randrange(60)) for _ in range(900)]
    res = _check(ctx, seqs, pairs)
    assert (res["flags"] == 0).any() and (res["flags"] == 1).any()


@pytest.mark.parametrize("mismatch,indel", [(-1, -1), (-3, -2), (-2, -5), (0, -1), (-4, 0)])
def test_other_scoring_parameters(ctx, mismatch, indel):
    rng = random.Random(200 + mismatch * 10 + indel)
    base = _rand(rng, 900)
    seqs = []
    for _ in range(30):
        L = rng.randint(10, 400)
        st = rng.randrange(0, len(base) - L)
        s = bytearray(base[st:st + L])
        for p in range(L):
            if rng.random() < 0.04:
                s[p] = rng.choice(b"ACGT")
        seqs.append(bytes(s))
    pairs = [(rng.randrange(30), rng.randrange(30)) for _ in range(300)]
    _check(ctx, seqs, pairs, g.DpParams(mismatch, indel, 0))


def test_long_contigs_and_range_limits(ctx):
    rng = random.Random(23)
    big = _rand(rng, 16383)
    flank = bytearray(big[9000:9995])
    flank[400] = ord("A") if flank[400] != ord("A") else ord("C")
    seqs = [bytes(flank), big, _rand(rng, 995), big[:10000], _rand(rng, 16384)]
    _check(ctx, seqs, [(0, 1), (2, 1), (0, 3), (2, 3)])
    with pytest.raises(g.GpError):
        ctx.semiglobal_batch(seqs, [(0, 4)])                # contig beyond the 14-bit start field
