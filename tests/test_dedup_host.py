"""CPU: the dedup stage's host code (gappadder_b200/host/dedup.cpp: raw FASTA records, unique names, ordered candidate
pairs, record synthesis, the reference's removal rules, output) through the drop-in binary built on the oracle shim,
against oracle/dedup_oracle.py.  tests/test_gpu_dedup.py runs the real binary on the GPU."""
import os
import subprocess
import tempfile

import pytest

from _dedupcases import FLAGS, expected, make_set
from test_contigsmerger_host import hosttest_binary  # noqa: F401  (fixture)


@pytest.mark.parametrize("seed,config,cutoff,contained", [(1, "tiny", 0.99, True), (2, "tiny", 0.9, False), (3, "small", 0.95, True), (4, "small", 0.85, False)])
def test_single_set(hosttest_binary, seed, config, cutoff, contained):  # noqa: F811
    text = make_set(seed, config)
    want, removed = expected(text, cutoff, contained)
    with tempfile.TemporaryDirectory() as td:
        fa, out = os.path.join(td, "c.fa"), os.path.join(td, "o.fa")
        open(fa, "wb").write(text)
        p = subprocess.run([hosttest_binary] + FLAGS + ["--dedup", fa, out, "--cutoff", str(cutoff)] + (["--contained"] if contained else []), capture_output=True)
        assert p.returncode == 0, p.stderr
        assert open(out, "rb").read() == want
    assert want != text                      # the repeated name at least is gone


def test_batch_equals_single_and_nothing_to_remove_copies_the_file(hosttest_binary):  # noqa: F811
    sets = [(make_set(11, "tiny"), 0.99, True), (make_set(12, "tiny"), 0.9, False), (b">a\nACGTACGTTTGACCA\nGGA\n>b\nTTTTTTTTTTGGGGGGGGGGCCCCCCCCCCAAAAAAAAAAGTGT\n", 0.99, True),
            (b"", 0.99, False)]
    with tempfile.TemporaryDirectory() as td:
        lst = os.path.join(td, "list.tsv")
        with open(lst, "w") as f:
            for k, (text, cutoff, contained) in enumerate(sets):
                open(os.path.join(td, "s%d.fa" % k), "wb").write(text)
                f.write("%s\t%s\t%s\t%s\n" % (os.path.join(td, "s%d.fa" % k), os.path.join(td, "s%d.out" % k), cutoff, "g" if contained else "p"))
        p = subprocess.run([hosttest_binary] + FLAGS + ["--dedup-batch", lst, "--gpus", "2", "--stats"], capture_output=True)
        assert p.returncode == 0, p.stderr
        for k, (text, cutoff, contained) in enumerate(sets):
            assert open(os.path.join(td, "s%d.out" % k), "rb").read() == expected(text, cutoff, contained)[0], k
        assert open(os.path.join(td, "s2.out"), "rb").read() == sets[2][0]      # verbatim copy
