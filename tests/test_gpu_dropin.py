"""GPU: the real drop-in binary (build/ContigsMerger_b200, DP on the B200 through the C ABI) against
the reference's whole-binary golden outputs, in single-gap and in batch form, and against the
reference binary itself (oracle/_ref/ContigsMerger, when it travelled with the repo) on fresh gaps."""
import os
import subprocess
import tempfile

import pytest

import synth_gaps

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "contigsmerger")
BIN = os.path.join(ROOT, "build", "ContigsMerger_b200")
REF = os.path.join(ROOT, "oracle", "_ref", "ContigsMerger")
FLAGS = "-s 0.4 -i1 -2.0 -i2 -2.0 -x 12 -y 50 -k 10 -t 1 -m 1".split()
CASES = sorted(f[:-3] for f in os.listdir(GOLD) if f.endswith(".fa"))


def _run(binary, fasta, flags=FLAGS):
    with tempfile.TemporaryDirectory() as td:
        info = os.path.join(td, "x.info")
        p = subprocess.run([binary] + flags + ["-o", info, fasta], cwd=td, capture_output=True)
        info_b = open(info, "rb").read() if os.path.exists(info) else b""
        gml = os.path.join(td, "tmp.gml")
        gml_b = open(gml, "rb").read() if os.path.exists(gml) else b""
        return p.returncode, p.stdout, info_b, gml_b


def test_binary_exists():
    assert os.path.exists(BIN), "run `make` (or __graft_entry__.build()) first"


@pytest.mark.parametrize("case", ["ka", "single", "empty", "tiny1", "small1", "noisy1"])
def test_single_gap_golden(case):
    rc, out, info, gml = _run(BIN, os.path.join(GOLD, case + ".fa"))
    assert rc == int(open(os.path.join(GOLD, case + ".rc")).read())
    assert out == open(os.path.join(GOLD, case + ".stdout"), "rb").read()
    assert info == open(os.path.join(GOLD, case + ".info"), "rb").read()
    assert gml == open(os.path.join(GOLD, case + ".gml"), "rb").read()


@pytest.mark.parametrize("streams", [1, 3])
def test_batch_golden(streams):
    """--batch: many gaps through one process; with --streams 3 the gaps are split over three workers (host thread +
    context + stream each) on the same GPU -- the outputs must not depend on the split."""
    with tempfile.TemporaryDirectory() as td:
        lst = os.path.join(td, "list.tsv")
        with open(lst, "w") as f:
            for c in CASES:
                f.write("%s\t%s\t%s\n" % (os.path.join(GOLD, c + ".fa"), os.path.join(td, c + ".out"), os.path.join(td, c + ".info")))
        p = subprocess.run([BIN] + FLAGS + ["--batch", lst, "--streams", str(streams)], cwd=td, capture_output=True)
        assert p.returncode == 0, p.stderr
        for c in CASES:
            assert open(os.path.join(td, c + ".out"), "rb").read() == open(os.path.join(GOLD, c + ".stdout"), "rb").read(), c
            assert open(os.path.join(td, c + ".info"), "rb").read() == open(os.path.join(GOLD, c + ".info"), "rb").read(), c
            assert open(os.path.join(td, c + ".out.gml"), "rb").read() == open(os.path.join(GOLD, c + ".gml"), "rb").read(), c


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/ContigsMerger did not travel with the repo")
def test_against_reference_binary_fresh_gaps():
    """GAPPadder's own command line (-t 5) on fresh seeded gaps: reference binary vs drop-in, bytes."""
    flags = [f if not (i == FLAGS.index("-t") + 1) else "5" for i, f in enumerate(FLAGS)]
    with tempfile.TemporaryDirectory() as td:
        for cfg, seed in (("small", 21), ("noisy", 22), ("small", 23)):
            fa = os.path.join(td, "g_%s_%d.fa" % (cfg, seed))
            synth_gaps.write_fasta(fa, synth_gaps.make_gap(seed, synth_gaps.CONFIGS[cfg]))
            assert _run(BIN, fa, flags) == _run(REF, fa, flags), (cfg, seed)
