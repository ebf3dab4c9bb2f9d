"""GPU: build/ContigsMerger_b200 (DP, quick check and relax chains on the B200) against whole-binary REFERENCE goldens at
realistic size (tests/golden/big, tests/golden/make_golden_big.py): full cfg1 gaps (32-deep relax chains, merged rows
to 8 kb, team kernels chosen by the library, prefix sharing), cfg3 gaps of 30 and 80 contigs, a reduced cfg5 gap (8 kb
repeat-rich contigs), fan-shaped graphs with more than 21 paths per root, contigs with IUPAC letters.  Single-gap form,
--batch, --streams and, when the box has them, --gpus 2."""
import os
import subprocess
import tempfile

import pytest

import _bigcases as B

pytestmark = pytest.mark.gpu

BIN = os.path.join(B.ROOT, "build", "ContigsMerger_b200")
EXACT = [c for c in B.CASES if c != "fan2"]


def _n_gpus():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True).stdout
        return sum(1 for ln in out.splitlines() if ln.startswith("GPU "))
    except Exception:
        return 0


def test_cases_present():
    assert os.path.exists(BIN), "run `make` (or __graft_entry__.build()) first"
    assert {"cfg1_s1", "cfg1_s2", "cfg3_s15", "cfg3_s43", "cfg5r_s1", "fan1", "fan2", "fan3", "iupac1", "iupac2"} <= set(B.CASES)


@pytest.mark.parametrize("case", EXACT)
def test_single_gap_reference_bytes(case):
    with tempfile.TemporaryDirectory() as td:
        B.check_single(BIN, case, td)


def test_fan2_modulo_allocator_order():
    with tempfile.TemporaryDirectory() as td:
        rc, out, info, gml, err = B.run_single(BIN, "fan2", td)
        assert rc == 0, err[-300:]
        B.check_paths_modulo_ties(info, out, gml, "fan2")


@pytest.mark.parametrize("extra", [(), ("--streams", "2"), ("--host-quick-check",), ("--host-relax",),
                                   # the chunk pipeline at its most concurrent: one gap per chunk, three mergers on the GPU, a relax
                                   # kernel of one chunk running beside the pairwise kernel of the next (host/device_gate.hpp)
                                   ("--chunk-gaps", "1", "--streams", "3"), ("--chunk-gaps", "2", "--streams", "2", "--no-gml"),
                                   ("--chunk-gaps", "1", "--streams", "3", "--host-relax")])
def test_batch_reference_bytes(extra):
    """All cases in one process (one sequence table, one pairwise launch, relax chains of all gaps together); the
    outputs must not depend on the batch, on the worker split or on where the quick check runs."""
    with tempfile.TemporaryDirectory() as td:
        B.check_batch(BIN, EXACT, td, extra)


@pytest.mark.skipif(_n_gpus() < 2, reason="needs two GPUs")
def test_batch_two_gpus_reference_bytes():
    """The product's multi-GPU path: gp_partition_gaps over two contexts on two devices, results gathered by the host."""
    with tempfile.TemporaryDirectory() as td:
        p = B.check_batch(BIN, EXACT, td, ("--gpus", "2", "--stats"))
        assert b'"gpus": 2' in p.stderr


def test_too_many_letters_fails_that_gap_only():
    with tempfile.TemporaryDirectory() as td:
        bad = os.path.join(td, "bad.fa")
        with open(bad, "w") as f:
            f.write(">x\nACGTNBDEFHIJKLMOPQRSUVWXYZACGTACGTACGTACGTAGCATCGATCGATCGACTAGCTAGCTAGCATCG\n"
                    ">y\nGATCGACTAGCTAGCTAGCATCGTTTTGGGGCCCCAAAATTTTGGGCCCAATTGGCCAATTACGATCGACTAGC\n")
        good = B.write_input("fan1", td)
        lst = os.path.join(td, "l.tsv")
        with open(lst, "w") as f:
            f.write("%s\t%s\t%s\n" % (bad, os.path.join(td, "bad.out"), os.path.join(td, "bad.info")))
            f.write("%s\t%s\t%s\n" % (good, os.path.join(td, "good.out"), os.path.join(td, "good.info")))
        p = subprocess.run([BIN] + B.FLAGS + ["--batch", lst], cwd=td, capture_output=True)
        assert p.returncode == 3 and b"distinct sequence letters" in p.stderr
        assert not os.path.exists(os.path.join(td, "bad.out"))
        assert open(os.path.join(td, "good.out"), "rb").read() == B.golden("fan1", "stdout")
