"""GPU parity: the CUDA path, called through the C ABI, against the oracle on the same inputs.

Bit-exact (integer work): (score, row_end, col_end, nclip, row0/col0/contained flags) per pair.
"""
import os
import random
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

import gappadder_b200 as g
from gappadder_b200.capi import (FLAG_CLOSED, FLAG_COL0, FLAG_CONTAINED, FLAG_ROW0, KERNEL_ALL, KERNEL_CERT16, KERNEL_DP_ALL,
                                 KERNEL_PRMT16, KERNEL_TABLE16)

KERNEL_TAGGED = KERNEL_TABLE16 | KERNEL_PRMT16      # everything but the certificate kernel
from _oracle import oracle_evaluate, oracle_revcomp
import synth_gaps

pytestmark = pytest.mark.gpu


def _check(ctx, seqs, pairs, params=None, full=False, masks=(KERNEL_ALL, KERNEL_DP_ALL, KERNEL_TAGGED, KERNEL_PRMT16)):
    """Every kernel the library can route these pairs to (default routing = closed form for a sequence against
    itself, then the certificate kernel -- run with the probe and with either starting system forced, one warp
    per pair and one CTA per pair --, then the table kernel, then the PRMT kernel / general kernel; KERNEL_DP_ALL
    is the same without the closed form) against the oracle.  Returns the default routing's results."""
    params = params or g.GAPPADDER_DP

    def one(ab):
        o = oracle_evaluate(seqs[ab[0]], seqs[ab[1]], params.mismatch, params.indel, params.max_clip, full=full)
        return (o.score, o.row_end, o.col_end, o.nclip, int(o.tb_row == 0), int(o.tb_col == 0), o.bcontained)
    if len(pairs) >= 64:                    # realistic-size gaps: the oracle (ctypes, GIL released) on every host core
        with ThreadPoolExecutor(max_workers=min(32, os.cpu_count() or 1)) as ex:
            want = list(ex.map(one, pairs))
    else:
        want = [one(ab) for ab in pairs]
    first = None
    try:
        runs = []        # (kernel mask, certificate system to start with, team mode, certificate kernel layout, orientation)
        for mask in masks:
            if mask & KERNEL_CERT16:      # layout 0: free moves whenever the launch allows it, 1: column potential only
                # orientation 0: the longer sequence becomes the row sequence; 1 never transposed; 2 every pair transposed that can be
                runs += [(mask, 0, 0, 0, 0), (mask, 1, 1, 0, 0), (mask, 2, 2, 0, 0), (mask, 3, 0, 0, 0), (mask, 0, 2, 0, 0), (mask, 0, 3, 0, 0), (mask, 2, 3, 1, 0),
                         (mask, 0, 0, 1, 0), (mask, 1, 2, 1, 0), (mask, 2, 1, 1, 0), (mask, 3, 0, 1, 0),
                         (mask, 0, 0, 0, 1), (mask, 0, 0, 0, 2), (mask, 1, 2, 1, 2), (mask, 2, 0, 1, 2), (mask, 3, 3, 0, 2), (mask, 0, 2, 0, 2)]
            else:
                runs += [(mask, 0, 0, 0, 0)]
        for mask, system, team, layout, orient in runs:
            ctx.set_kernel_mask(mask)
            ctx.set_cert_system(system)
            ctx.set_team_mode(team)
            ctx.set_cert_layout(layout)
            ctx.set_orientation(orient)
            res = ctx.overlap_batch(seqs, pairs, params)
            assert len(res) == len(pairs)
            bad = []
            for (a, b), r, w in zip(pairs, res, want):
                got = (int(r["score"]), int(r["row_end"]), int(r["col_end"]), int(r["nclip"]),
                       int(bool(r["flags"] & FLAG_ROW0)), int(bool(r["flags"] & FLAG_COL0)), int(bool(r["flags"] & FLAG_CONTAINED)))
                if w != got:
                    bad.append(((a, b), len(seqs[a]), len(seqs[b]), w, got))
            assert not bad, "kernel mask %d system %d team %d layout %d orientation %d (free moves used: %d, transposed pairs: %d): first mismatches (pair, m, n, oracle, gpu): %r" % (
                mask, system, team, layout, orient, ctx.last_layout, ctx.transposed_pairs, bad[:5])
            if first is None:
                first = res
    finally:
        ctx.set_kernel_mask(KERNEL_ALL)
        ctx.set_cert_system(0)
        ctx.set_team_mode(0)
        ctx.set_cert_layout(0)
        ctx.set_orientation(0)
    return first


def _rand(rng, n, alpha=b"ACGT"):
    return bytes(rng.choice(alpha) for _ in range(n))


def test_known_answer(ctx):
    # the two-contig case whose reference output is NEW_CONTIG_MERGE_1 = a b
    a = b"ACGTACGTAGCTAGCTAGCTAGCATCGATCGATCGATCAGCTAGCTAGCATCGATCAGCTACGACTAGC"
    b = b"GATCGATCAGCTAGCTAGCATCGATCAGCTACGACTAGCTTTTGGGGCCCCAAAATTTTGGGCCCAATTGGCCAATT"
    seqs = [a, oracle_revcomp(a), b, oracle_revcomp(b)]
    pairs = [(i, j) for i in range(4) for j in range(i, 4)]
    res = _check(ctx, seqs, pairs, full=True)
    r = res[pairs.index((0, 2))]
    assert g.merged_concat(a, b, r) == a + b[39:]


def test_tie_heavy_small_alphabets(ctx):
    rng = random.Random(11)
    seqs = []
    for _ in range(120):
        alpha = rng.choice([b"A", b"AC", b"ACG", b"ACGT", b"ACGTN"])
        seqs.append(_rand(rng, rng.randint(1, 90), alpha))
    pairs = [(rng.randrange(len(seqs)), rng.randrange(len(seqs))) for _ in range(1500)]
    _check(ctx, seqs, pairs, full=True, masks=(KERNEL_ALL, KERNEL_TAGGED, KERNEL_PRMT16, 0))


def test_empty_and_tiny(ctx):
    seqs = [b"", b"A", b"C", b"AC", b"ACGTACGT", b"N", b"NN"]
    pairs = [(i, j) for i in range(len(seqs)) for j in range(len(seqs))]
    _check(ctx, seqs, pairs, full=True, masks=(KERNEL_ALL, KERNEL_CERT16, KERNEL_TAGGED, KERNEL_PRMT16, KERNEL_TABLE16, 0))
    assert len(ctx.overlap_batch(seqs, [])) == 0


@pytest.mark.parametrize("config,seed", [("tiny", 1), ("small", 2), ("noisy", 3)])
def test_synthetic_gap_all_pairs(ctx, config, seed):
    recs = synth_gaps.make_gap(seed, synth_gaps.CONFIGS[config])
    nodes = []
    for _, s in recs:
        nodes += [s, oracle_revcomp(s)]
    pairs = [(i, j) for i in range(len(nodes)) for j in range(i, len(nodes))]
    _check(ctx, nodes, pairs)


def test_strip_boundaries_and_clip_edges(ctx):
    # lengths around the strip heights (32*R rows) and around max_clip
    rng = random.Random(5)
    base = _rand(rng, 1400)
    seqs = []
    for L in (49, 50, 51, 52, 255, 256, 257, 511, 512, 513, 1023, 1025):
        st = rng.randrange(0, len(base) - L)
        seqs.append(base[st:st + L])
    pairs = [(i, j) for i in range(len(seqs)) for j in range(len(seqs))]
    _check(ctx, seqs, pairs)


@pytest.mark.parametrize("mismatch,indel,clip", [(-1, -1, 0), (-3, -2, 10), (-2, -5, 50), (0, -1, 3), (-20, -30, 7)])
def test_other_scoring_parameters(ctx, mismatch, indel, clip):
    rng = random.Random(100 + clip)
    base = _rand(rng, 600)
    seqs = []
    for _ in range(24):
        L = rng.randint(20, 300)
        st = rng.randrange(0, len(base) - L)
        s = bytearray(base[st:st + L])
        for p in range(L):
            if rng.random() < 0.03:
                s[p] = rng.choice(b"ACGT")
        seqs.append(bytes(s))
    pairs = [(rng.randrange(24), rng.randrange(24)) for _ in range(300)]
    _check(ctx, seqs, pairs, g.DpParams(mismatch, indel, clip))


def test_wide_alphabet_goes_through_general_kernel(ctx):
    rng = random.Random(9)
    alpha = b"ACGTNRYKMSWB"
    seqs = [_rand(rng, rng.randint(30, 200), alpha) for _ in range(20)]
    pairs = [(i, j) for i in range(20) for j in range(20)]
    _check(ctx, seqs, pairs)


def test_cfg1_gap_candidates_full_size(ctx):
    """One BASELINE cfg1 gap (40 contigs, 300-3000 bp): every candidate pair of the pairwise phase."""
    recs = synth_gaps.make_gap(1, synth_gaps.CONFIGS["cfg1"])
    nodes = []
    for _, s in recs:
        nodes += [s, g.revcomp(s)]
    cand = g.candidate_pairs(nodes, 10)
    pairs = [(int(p["row_seq"]), int(p["col_seq"])) for p in cand]
    assert len(pairs) > 100
    _check(ctx, nodes, pairs)
    # certificate kernel bookkeeping on realistic pairs: a sequence against itself is routed to an exact kernel by
    # the host; of the rest few need the second pass and hardly any the exact kernel
    ctx.overlap_batch(nodes, pairs)
    st = ctx.cert_stats()
    assert st["cert16"] == len(pairs) - sum(1 for a, b in pairs if a == b)
    assert st["second_passes"] <= 0.3 * st["cert16"] and st["exact_retries"] <= 0.02 * st["cert16"], st


def test_long_overlaps_both_potentials(ctx):
    """Suffix/prefix overlaps of 1400-4000 bases ending in the last rows but not the last columns, and
    transposed; lengths up to the 16-bit kernel's limit (min(m,n) = 4094) and just beyond it."""
    rng = random.Random(77)
    seqs = []
    for la, ov, extra in [(2700, 2500, 150), (3000, 2900, 60), (4094, 4000, 200), (2000, 1990, 2200), (1500, 1400, 3000), (4200, 4100, 300)]:
        a = _rand(rng, la)
        seqs += [a, a[-ov:] + _rand(rng, extra)]
    pairs = []
    for k in range(0, len(seqs), 2):
        pairs += [(k, k + 1), (k + 1, k), (k, k), (k + 1, k + 1)]
    res = _check(ctx, seqs, pairs)
    from gappadder_b200.capi import FLAG_KERNEL16
    # the 12 overlap pairs go through the certificate kernel (columns <= 16382), the 12 self pairs have the closed form
    assert int((res["flags"] & FLAG_KERNEL16 != 0).sum()) == 12 and int((res["flags"] & FLAG_CLOSED != 0).sum()) == 12
    ctx.overlap_batch(seqs, pairs)
    assert ctx.cert_stats() == dict(cert16=12, second_passes=0, exact_retries=0)
    assert ctx.closed_form_stats() == dict(pairs=12, cells=sum(len(seqs[a]) ** 2 for a, b in pairs if a == b))
    # without the closed form: 7 of the 12 self pairs go through the table kernel (<= 4094 columns), 5 through the general one
    try:
        ctx.set_kernel_mask(KERNEL_DP_ALL)
        res = ctx.overlap_batch(seqs, pairs)
        assert int((res["flags"] & FLAG_KERNEL16 != 0).sum()) == 19 and not (res["flags"] & FLAG_CLOSED).any()
    finally:
        ctx.set_kernel_mask(KERNEL_ALL)


def test_kernel_routing(ctx):
    """A/C/G/T pairs with <= 16382 columns take the certificate kernel; without it those with <= 4094 columns take
    the table kernel, pairs with N (or longer columns but short rows) the PRMT kernel, the rest the general kernel;
    masks move pairs down that list and never change results."""
    rng = random.Random(3)
    seqs = [_rand(rng, 300), _rand(rng, 700), _rand(rng, 500, b"ACGTN"), _rand(rng, 4500), _rand(rng, 4300)]
    pairs = [(0, 1), (1, 0), (0, 2), (2, 1), (0, 3), (3, 0), (3, 4)]
    packed, off, lens, nsym = g.pack_sequences(seqs)
    ctx.set_sequences(packed, off, lens, nsym)
    ctx.upload_pairs(np.array(pairs, dtype=np.uint32).view(g.capi.PAIR_DTYPE).reshape(-1))
    assert ctx.pair_split() == dict(table16=0, prmt16=2, wide32=0) and ctx.cert_stats()["cert16"] == 5
    try:
        ctx.set_kernel_mask(KERNEL_TAGGED)
        ctx.upload_pairs(np.array(pairs, dtype=np.uint32).view(g.capi.PAIR_DTYPE).reshape(-1))
        assert ctx.pair_split() == dict(table16=3, prmt16=3, wide32=1)   # (3,0): 4500 rows x 300 columns -> table
        assert ctx.cert_stats()["cert16"] == 0
        ctx.set_kernel_mask(KERNEL_PRMT16)
        ctx.upload_pairs(np.array(pairs, dtype=np.uint32).view(g.capi.PAIR_DTYPE).reshape(-1))
        assert ctx.pair_split() == dict(table16=0, prmt16=6, wide32=1)
    finally:
        ctx.set_kernel_mask(KERNEL_ALL)
    _check(ctx, seqs, pairs, masks=(KERNEL_ALL, KERNEL_TAGGED, KERNEL_PRMT16, 0))


def test_certificate_kernel_long_columns_and_fallbacks(ctx):
    """Certificate kernel beyond the tagged kernels' range (columns up to 16382), pairs whose walks end in the
    corner (identical sequences under different indices: the corner system certifies them; when the run is forced
    to start with another system it takes sub-table passes), and junk pairs whose best cell sits in a corner."""
    rng = random.Random(21)
    a = _rand(rng, 9000)
    b = a[-5000:] + _rand(rng, 6000)          # 11000 columns
    c = _rand(rng, 16382)
    d = c[:700]
    e = _rand(rng, 5000)
    seqs = [a, b, c, d, e, bytes(e), a[:3000], bytes(a[:3000])]
    pairs = [(0, 1), (1, 0), (3, 2), (2, 3), (0, 2), (4, 5), (5, 4), (6, 7), (3, 4), (4, 3), (1, 2)]
    _check(ctx, seqs, pairs, masks=(KERNEL_ALL,))
    ctx.overlap_batch(seqs, pairs)
    st = ctx.cert_stats()
    assert st["cert16"] == len(pairs) and st["exact_retries"] == 0, st      # (4,5), (5,4), (6,7) end in the corner: system C
    try:
        ctx.set_cert_system(1)
        ctx.overlap_batch(seqs, pairs)
        assert ctx.cert_stats()["second_passes"] >= 6 and ctx.cert_stats()["exact_retries"] == 0      # U, then L, then C for those three
    finally:
        ctx.set_cert_system(0)


def test_team_mode_many_long_pairs(ctx):
    """The CTA-per-pair form of the certificate kernel under load: more pairs than CTAs, 1 to 16 strips per pair,
    all four warps of a CTA pipelined through one boundary line.  Same results as one warp per pair on every pair,
    and as the oracle on a sample."""
    rng = random.Random(314)
    base = _rand(rng, 12000)
    seqs = []
    for _ in range(90):
        L = rng.choice([64, 300, 513, 1100, 2049, 3000, 5000, 8000])
        st = rng.randrange(0, len(base) - L)
        s = bytearray(base[st:st + L])
        for p in range(L):
            if rng.random() < 0.004:
                s[p] = rng.choice(b"ACGT")
        s = bytes(s)
        seqs.append(oracle_revcomp(s) if rng.random() < 0.3 else s)
    pairs = []
    while len(pairs) < 1400:
        a, b = rng.randrange(len(seqs)), rng.randrange(len(seqs))
        if a != b and len(seqs[b]) <= 5000:
            pairs.append((a, b))
    try:
        ctx.set_team_mode(1)
        warp = ctx.overlap_batch(seqs, pairs)
        assert not ctx.last_team
        ctx.set_team_mode(2)
        team = ctx.overlap_batch(seqs, pairs)
        assert ctx.last_team
        for layout in (0, 1):                        # eight warps per pair, both value layouts
            ctx.set_cert_layout(layout)
            ctx.set_team_mode(3)
            big = ctx.overlap_batch(seqs, pairs)
            assert ctx.last_team and (big == warp).all()
        ctx.set_cert_layout(0)
        ctx.set_team_mode(0)
        ctx.overlap_batch(seqs, pairs[:40])          # few long pairs: the library picks the team form itself
        assert ctx.last_team
    finally:
        ctx.set_team_mode(0)
        ctx.set_cert_layout(0)
    assert (warp == team).all()
    for k in range(0, len(pairs), 23):
        a, b = pairs[k]
        o = oracle_evaluate(seqs[a], seqs[b])
        r = team[k]
        assert (o.score, o.row_end, o.col_end, o.nclip, int(o.tb_row == 0), int(o.tb_col == 0), o.bcontained) == \
            (int(r["score"]), int(r["row_end"]), int(r["col_end"]), int(r["nclip"]), int(bool(r["flags"] & FLAG_ROW0)),
             int(bool(r["flags"] & FLAG_COL0)), int(bool(r["flags"] & FLAG_CONTAINED))), (a, b)


def test_free_moves_layout_range_edge_and_fallback(ctx):
    """The certificate kernel's free-moves layout is used when every column sequence of the launch has at most 3800
    bases; exactly 3800 with near-identical sequences drives the values to the top of the 15-bit range.  One base more
    and the launch takes the column-potential layout.  Results are the oracle's either way."""
    rng = random.Random(12)
    s = _rand(rng, 3800)
    t = bytearray(s)
    t[1900] = ord("A") if t[1900] != ord("A") else ord("C")
    seqs = [bytes(t), s, _rand(rng, 600) + s + _rand(rng, 600), s[:3799] + b"G", _rand(rng, 5000)]
    pairs = [(0, 1), (2, 1), (1, 3), (4, 1), (3, 0)]
    _check(ctx, seqs, pairs, masks=(KERNEL_ALL,))
    ctx.overlap_batch(seqs, pairs)
    assert ctx.last_layout == 1
    seqs.append(s + b"A")                                    # 3801 bases as a column sequence ...
    pairs.append((0, 5))                                     # ... of a 3800-base row sequence: computed transposed (the
    _check(ctx, seqs, pairs, masks=(KERNEL_ALL,))            # longer sequence as rows), the columns still fit the layout
    ctx.overlap_batch(seqs, pairs)
    assert ctx.last_layout == 1
    pairs.append((4, 5))                                     # 5000 rows x 3801 columns: no orientation keeps the layout
    _check(ctx, seqs, pairs, masks=(KERNEL_ALL,))
    ctx.overlap_batch(seqs, pairs)
    assert ctx.last_layout == 0
    ctx.set_orientation(1)                                   # never transpose: (0, 5) alone already forces the fallback
    try:
        ctx.overlap_batch(seqs, pairs[:-1])
        assert ctx.last_layout == 0
    finally:
        ctx.set_orientation(0)


@pytest.mark.parametrize("config,n_gaps", [("cfg1", 6), ("cfg3", 8), ("noisy", 12), ("tiny", 5)])
def test_quick_check_on_the_device_equals_the_host_filter(ctx, config, n_gaps):
    """gp_quick_check_device on the packed table vs gp_candidate_pairs on the ASCII nodes (itself pinned to the
    reference's candidate lists in tests/test_oracle_golden.py), gap by gap: same pairs in the same order.  `noisy`
    has N bases (k-mer letter A), `tiny` has nodes around the 30-base window."""
    seqs, gap_first, want = [], [0], []
    for gi in range(n_gaps):
        nodes = []
        for _, s in synth_gaps.make_gap(100 + gi, synth_gaps.CONFIGS[config]):
            nodes.append(s)
            nodes.append(g.revcomp(s))
        if gi == 1:
            nodes += [b"ACGTACGTAC", b"ACGTACGTACG" * 2, b"ACG", b""]        # shorter than the window / than k / empty
        seqs += nodes
        gap_first.append(len(seqs))
        want.append(g.candidate_pairs(nodes, 10))
    packed, off, lens, nsym = g.pack_sequences(seqs)
    ctx.set_sequences(packed, off, lens, nsym)
    got = ctx.quick_check_device(gap_first, 10)
    assert len(got) == n_gaps
    for gi in range(n_gaps):
        assert np.array_equal(got[gi]["row_seq"], want[gi]["row_seq"]) and np.array_equal(got[gi]["col_seq"], want[gi]["col_seq"]), (config, gi)


def _gap_nodes_and_candidates(config, seed):
    nodes = []
    for _, s in synth_gaps.make_gap(seed, synth_gaps.CONFIGS[config]):
        nodes += [s, g.revcomp(s)]
    cand = g.candidate_pairs(nodes, 10)
    return nodes, [(int(p["row_seq"]), int(p["col_seq"])) for p in cand]


def test_cfg3_gap_80_contigs_every_candidate_pair(ctx):
    """BASELINE cfg3/cfg4 shape at its upper end: the 80-contig gap whose whole-binary output is a committed reference
    golden (tests/golden/big/cfg3_s15): every candidate pair, every layout / team / starting system."""
    nodes, pairs = _gap_nodes_and_candidates("cfg3", 15)
    assert len(nodes) == 160 and len(pairs) > 2000
    _check(ctx, nodes, pairs, masks=(KERNEL_ALL, KERNEL_TAGGED))


def test_cfg5_reduced_repeat_rich_long_columns(ctx):
    """The tie-heaviest input: 8 kb contigs of a repeat-rich locus (reduced cfg5, golden tests/golden/big/cfg5r_s1).
    Columns beyond 3800 bases: the column-potential layout, second passes, exact retries if any; every layout / team /
    starting system on every candidate pair; and two full-size cfg5 contig pairs (10 kb x 10 kb)."""
    nodes, pairs = _gap_nodes_and_candidates("cfg5r", 1)
    assert len(nodes) == 28 and max(len(s) for s in nodes) == 8000 and len(pairs) > 100
    _check(ctx, nodes, pairs, masks=(KERNEL_ALL,))
    ctx.overlap_batch(nodes, pairs)
    assert ctx.last_layout == 0                       # 8000 columns: not the free-moves layout
    big, bp = _gap_nodes_and_candidates("cfg5", 3)
    bp = [ab for ab in bp if ab[0] != ab[1]][:400:100]
    assert len(bp) >= 2 and len(big[bp[0][0]]) == 10000
    _check(ctx, big, bp, masks=(KERNEL_ALL,))


@pytest.mark.parametrize("k", [1, 2, 3, 5, 9])
def test_quick_check_on_the_device_small_k(ctx, k):
    """k = 1, 2 (a k-mer set smaller than one word), and other k below GAPPadder's 10."""
    seqs, gap_first, want = [], [0], []
    for gi in range(4):
        nodes = []
        for _, s in synth_gaps.make_gap(300 + gi, synth_gaps.CONFIGS["noisy" if gi % 2 else "small"]):
            nodes += [s, g.revcomp(s)]
        seqs += nodes
        gap_first.append(len(seqs))
        want.append(g.candidate_pairs(nodes, k))
    packed, off, lens, nsym = g.pack_sequences(seqs)
    ctx.set_sequences(packed, off, lens, nsym)
    got = ctx.quick_check_device(gap_first, k)
    for gi in range(4):
        assert np.array_equal(got[gi]["row_seq"], want[gi]["row_seq"]) and np.array_equal(got[gi]["col_seq"], want[gi]["col_seq"]), (k, gi)


def test_quick_check_on_the_device_big_gaps_and_many_items(ctx):
    """A full cfg5 gap (400 nodes of 10 kb: more nodes than one CTA's old limit, cut into dozens of work items), next to
    small gaps and an empty one: same candidate lists as the host filter."""
    seqs, gap_first, want = [], [0], []
    for config, seed in (("cfg5", 1), ("tiny", 7), ("cfg3", 15), ("tiny", 8)):
        nodes = []
        for _, s in synth_gaps.make_gap(seed, synth_gaps.CONFIGS[config]):
            nodes += [s, g.revcomp(s)]
        seqs += nodes
        gap_first.append(len(seqs))
        want.append(g.candidate_pairs(nodes, 10))
    gap_first.append(len(seqs))                      # an empty gap
    want.append(g.candidate_pairs([], 10))
    packed, off, lens, nsym = g.pack_sequences(seqs)
    ctx.set_sequences(packed, off, lens, nsym)
    got = ctx.quick_check_device(gap_first, 10)
    st = ctx.quick_check_stats()
    assert st["bases"] == sum(len(s) for s in seqs) and st["items"] > 8
    for gi in range(len(want)):
        assert np.array_equal(got[gi]["row_seq"], want[gi]["row_seq"]) and np.array_equal(got[gi]["col_seq"], want[gi]["col_seq"]), gi
