"""CPU: the host half of the drop-in (graph, path search and its truncation rule, relax chains, per-gap letter renaming,
output) against REFERENCE whole-binary goldens at realistic size (tests/golden/big); the DP is served by the oracle
through tests/shim_gp_oracle.cpp.  tests/test_gpu_dropin_big.py runs the same cases (and the larger ones) on the GPU."""
import os
import subprocess
import tempfile

import pytest

import _bigcases as B
from test_contigsmerger_host import hosttest_binary  # noqa: F401  (fixture)


@pytest.mark.parametrize("case", ["fan1", "fan3", "iupac1", "iupac2", "cfg1_s1", "cfg3_s43"])
def test_reference_bytes(hosttest_binary, case):  # noqa: F811
    with tempfile.TemporaryDirectory() as td:
        B.check_single(hosttest_binary, case, td)


def test_fan2_modulo_allocator_order(hosttest_binary):  # noqa: F811
    """More than 21 equal-length paths from one root: which survive depends on the C library's allocator in the
    reference (see merge_graph.cpp find_paths); everything else must match."""
    with tempfile.TemporaryDirectory() as td:
        rc, out, info, gml, err = B.run_single(hosttest_binary, "fan2", td)
        assert rc == 0
        B.check_paths_modulo_ties(info, out, gml, "fan2")


def test_batch_with_22_letters_over_two_gaps(hosttest_binary):  # noqa: F811
    """iupac1 and iupac2 hold 11 letters besides A C G T N each, 22 together: one packed table has 16 codes, so the
    letters are renamed per gap before packing (merger.cpp GapState::dp_seq); outputs keep the original letters."""
    with tempfile.TemporaryDirectory() as td:
        B.check_batch(hosttest_binary, ["iupac1", "iupac2", "fan1"], td)


def test_gap_with_too_many_letters_fails_alone(hosttest_binary):  # noqa: F811
    with tempfile.TemporaryDirectory() as td:
        bad = os.path.join(td, "bad.fa")
        with open(bad, "w") as f:
            f.write(">x\nACGTNBDEFHIJKLMOPQRSUVWXYZACGTACGTACGTACGTAGCATCGATCGATCGACTAGCTAGCTAGCATCG\n"
                    ">y\nGATCGACTAGCTAGCTAGCATCGTTTTGGGGCCCCAAAATTTTGGGCCCAATTGGCCAATTACGATCGACTAGC\n")
        p = subprocess.run([hosttest_binary] + B.FLAGS + ["-o", os.path.join(td, "i"), bad], cwd=td, capture_output=True)
        assert p.returncode == 3 and p.stdout == b"" and b"distinct sequence letters" in p.stderr
        good = B.write_input("fan1", td)
        lst = os.path.join(td, "l.tsv")
        with open(lst, "w") as f:
            f.write("%s\t%s\t%s\n" % (bad, os.path.join(td, "bad.out"), os.path.join(td, "bad.info")))
            f.write("%s\t%s\t%s\n" % (good, os.path.join(td, "good.out"), os.path.join(td, "good.info")))
        p = subprocess.run([hosttest_binary] + B.FLAGS + ["--batch", lst], cwd=td, capture_output=True)
        assert p.returncode == 3 and not os.path.exists(os.path.join(td, "bad.out"))
        assert open(os.path.join(td, "good.out"), "rb").read() == B.golden("fan1", "stdout")


def test_step_by_step_relax_path(hosttest_binary, monkeypatch):  # noqa: F811
    """The merger's fallback for gaps outside gp_relax_chains' domain (and for unresolved steps): the relax chains step by
    step through gp_overlap_batch, with prefix sharing -- same bytes."""
    monkeypatch.setenv("GP_SHIM_NO_RELAX", "1")
    with tempfile.TemporaryDirectory() as td:
        B.check_single(hosttest_binary, "cfg3_s43", td)
    with tempfile.TemporaryDirectory() as td:
        B.check_batch(hosttest_binary, ["fan1", "fan3", "iupac1"], td, ("--host-relax",))
