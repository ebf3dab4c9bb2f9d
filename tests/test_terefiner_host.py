"""CPU: the host half of TERefiner_b200 (gappadder_b200/host/terefiner_main.cpp, local_alignment.cpp: command line, batch
list, LocalAlignment's rest alignments, validateRepeats, output text) against golden outputs of the PREBUILT reference
binary, `TERefiner_1 -M` and `TERefiner_1 -A` (tests/golden/terefiner_modes.json, tests/golden/make_golden_terefiner.py).
The alignments themselves are served by the oracle through tests/shim_gp_affine_oracle.cpp here;
tests/test_gpu_terefiner.py runs the real binary on the GPU."""
import json
import os
import subprocess
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def golden():
    with open(os.path.join(ROOT, "tests", "golden", "terefiner_modes.json")) as f:
        return json.load(f)["cases"]


@pytest.fixture(scope="module")
def hosttest_binary():
    out = os.path.join(ROOT, "build", "TERefiner_hosttest")
    host = os.path.join(ROOT, "gappadder_b200", "host")
    srcs = [os.path.join(host, "terefiner_main.cpp"), os.path.join(host, "local_alignment.cpp"), os.path.join(ROOT, "tests", "shim_gp_affine_oracle.cpp")]
    oracle_c = os.path.join(ROOT, "oracle", "local_affine_oracle.c")
    deps = srcs + [os.path.join(host, "local_alignment.hpp"), oracle_c, os.path.join(ROOT, "include", "gappadder_b200.h")]
    os.makedirs(os.path.join(ROOT, "build"), exist_ok=True)
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(d) for d in deps):
        obj = os.path.join(ROOT, "build", "local_affine_oracle.o")
        subprocess.check_call(["gcc", "-O2", "-c", oracle_c, "-o", obj])
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-I" + os.path.join(ROOT, "include"), "-o", out] + srcs + [obj])
    return out


def run_batch(binary, mode, recs, extra=()):
    with tempfile.TemporaryDirectory() as td:
        lst = os.path.join(td, "pairs.tsv")
        with open(lst, "w") as f:
            for r in recs:
                f.write(r["s1"] + "\t" + r["s2"] + "\n")
        p = subprocess.run([binary, "-" + mode, "--batch", lst] + list(extra), capture_output=True)
    return p.returncode, p.stdout.decode(), p.stderr.decode()


@pytest.mark.parametrize("mode", ["M", "A"])
def test_batch_list_equals_the_prebuilt_binary(hosttest_binary, mode):
    recs = [r for r in golden() if mode in r]
    assert len(recs) > 100
    rc, out, err = run_batch(hosttest_binary, mode, recs)
    assert rc == 0, err
    assert out == "".join(r[mode] for r in recs)


@pytest.mark.parametrize("mode", ["M", "A"])
def test_pairs_dealt_to_three_gpus_come_back_in_list_order(hosttest_binary, mode):
    """--gpus N: one host thread and one context per GPU, pairs dealt longest first, no exchange; same bytes."""
    recs = [r for r in golden() if mode in r]
    rc, out, err = run_batch(hosttest_binary, mode, recs, ["--gpus", "3"])
    assert rc == 0, err
    assert out == "".join(r[mode] for r in recs)


def test_one_pair_per_process_as_the_reference_is_called(hosttest_binary):
    for r in golden()[:12]:
        for mode in "MA":
            if mode in r:
                p = subprocess.run([hosttest_binary, "-" + mode, "-r", r["s1"], "-s", r["s2"]], capture_output=True)
                assert p.returncode == 0 and p.stdout.decode() == r[mode]


def test_align_with_the_alignments_before_and_after(hosttest_binary):
    """LocalAlignment::align (TERefiner/algorithms/local_alignment.cpp:1053-1090) through the host mirror's batch form."""
    recs = [r for r in golden() if "align" in r]
    assert len(recs) > 90
    for r in recs[::3]:
        p = subprocess.run([hosttest_binary, "--align", "-r", r["s1"], "-s", r["s2"]], capture_output=True)
        assert p.returncode == 0 and [int(x) for x in p.stdout.split()] == r["align"]


def test_nothing_aligns_and_usage_errors(hosttest_binary):
    p = subprocess.run([hosttest_binary, "-M", "-r", "AAAA", "-s", "CCCC"], capture_output=True)
    assert p.returncode == 0 and p.stdout == b"0 0 0 0\n"           # the reference reads path[-1] here
    p = subprocess.run([hosttest_binary, "-A", "-r", "AAAA", "-s", "CCCC"], capture_output=True)
    assert p.returncode == 0 and p.stdout == b"0\n"
    assert subprocess.run([hosttest_binary, "-M", "-r", "ACGT"], capture_output=True).returncode == 2
    assert subprocess.run([hosttest_binary, "-U", "-r", "x.fa", "-o", "y.fa"], capture_output=True).returncode == 2
    assert subprocess.run([hosttest_binary, "-M", "--batch", "/nonexistent"], capture_output=True).returncode == 2
