"""GPU: the dedup stage through the real binary -- device quick check over every ordered contig pair,
the overlap DP kernels, the reference's removal rules -- against oracle/dedup_oracle.py (rules pinned to the reference's
TERefiner_1, records builder-defined: BWA parity unpinned), single set and batch, and the full-matrix quick check against
the host filter."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

import gappadder_b200 as g
from _dedupcases import FLAGS, expected, make_set

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "build", "ContigsMerger_b200")


@pytest.mark.parametrize("seed,config,cutoff,contained", [(1, "tiny", 0.99, True), (2, "tiny", 0.9, False), (3, "small", 0.95, True), (4, "small", 0.85, False),
                                                          (5, "noisy", 0.9, True), (6, "noisy", 0.85, False), (7, "cfg1", 0.99, True), (8, "cfg3", 0.95, False)])
def test_single_set(seed, config, cutoff, contained):
    text = make_set(seed, config)
    want, removed = expected(text, cutoff, contained)
    assert removed
    with tempfile.TemporaryDirectory() as td:
        fa, out = os.path.join(td, "c.fa"), os.path.join(td, "o.fa")
        open(fa, "wb").write(text)
        p = subprocess.run([BIN] + FLAGS + ["--dedup", fa, out, "--cutoff", str(cutoff)] + (["--contained"] if contained else []), capture_output=True)
        assert p.returncode == 0, p.stderr
        assert open(out, "rb").read() == want


def test_batch_of_sets_equals_the_oracle_and_a_second_pass_is_stable():
    sets = [(make_set(20 + k, ("tiny", "small", "noisy")[k % 3]), (0.99, 0.9, 0.85)[k % 3], k % 2 == 0) for k in range(9)]
    with tempfile.TemporaryDirectory() as td:
        lst = os.path.join(td, "list.tsv")
        with open(lst, "w") as f:
            for k, (text, cutoff, contained) in enumerate(sets):
                open(os.path.join(td, "s%d.fa" % k), "wb").write(text)
                f.write("%s\t%s\t%s\t%s\n" % (os.path.join(td, "s%d.fa" % k), os.path.join(td, "s%d.out" % k), cutoff, "g" if contained else "p"))
        p = subprocess.run([BIN] + FLAGS + ["--dedup-batch", lst, "--stats"], capture_output=True)
        assert p.returncode == 0, p.stderr
        outs = [open(os.path.join(td, "s%d.out" % k), "rb").read() for k in range(len(sets))]
        for k, (text, cutoff, contained) in enumerate(sets):
            assert outs[k] == expected(text, cutoff, contained)[0], k
        # the duplicate rule keeps the smaller name of every group: its output has no duplicates left under the same rule
        for k, (text, cutoff, contained) in enumerate(sets):
            if contained:
                continue
            fa2, out2 = os.path.join(td, "again%d.fa" % k), os.path.join(td, "again%d.out" % k)
            open(fa2, "wb").write(outs[k])
            p = subprocess.run([BIN] + FLAGS + ["--dedup", fa2, out2, "--cutoff", str(cutoff)], capture_output=True)
            assert p.returncode == 0, p.stderr
            assert open(out2, "rb").read() == expected(outs[k], cutoff, False)[0]


def test_full_matrix_quick_check_equals_the_host_filter_on_ordered_pairs():
    import synth_gaps
    nodes = []
    for _, s in synth_gaps.make_gap(31, synth_gaps.CONFIGS["small"]):
        nodes += [bytes(s), g.revcomp(bytes(s))]
    nodes += [b"ACGTACGTAC", b"ACG"]
    with g.Context(0) as ctx:
        ctx.set_sequences(*g.pack_sequences(nodes))
        m = ctx.quick_check_matrix([0, len(nodes)], 10)[0]
    for i in range(len(nodes)):
        for j in range(len(nodes)):
            if i == j:
                want = len(g.candidate_pairs([nodes[i]], 10)) == 1
            else:                                   # pair (0, 1) of the two-node list: the ends of nodes[j] occur in nodes[i]
                want = any(int(p["row_seq"]) == 0 and int(p["col_seq"]) == 1 for p in g.candidate_pairs([nodes[i], nodes[j]], 10))
            assert bool(m[i, j]) == want, (i, j, len(nodes[i]), len(nodes[j]))
