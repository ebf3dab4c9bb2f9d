"""GPU: build/TERefiner_b200 (the real binary: libgappadder_b200.so, affine kernels) against golden outputs of the PREBUILT
reference binary, `TERefiner_1 -M` (LocalAlignment::optAlign) and `TERefiner_1 -A` (RepeatsClassifier::validateRepeats):
tests/golden/terefiner_modes.json."""
import json
import os
import subprocess
import tempfile

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "build", "TERefiner_b200")


def _n_gpus():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True).stdout
        return sum(1 for l in out.splitlines() if l.startswith("GPU "))
    except Exception:
        return 0


def golden():
    with open(os.path.join(ROOT, "tests", "golden", "terefiner_modes.json")) as f:
        return json.load(f)["cases"]


@pytest.mark.parametrize("mode", ["M", "A"])
def test_batch_list_equals_the_prebuilt_binary(mode):
    if not os.path.exists(BIN):
        pytest.skip("build/TERefiner_b200 not built")
    recs = [r for r in golden() if mode in r]
    with tempfile.TemporaryDirectory() as td:
        lst = os.path.join(td, "pairs.tsv")
        with open(lst, "w") as f:
            for r in recs:
                f.write(r["s1"] + "\t" + r["s2"] + "\n")
        p = subprocess.run([BIN, "-" + mode, "--batch", lst], capture_output=True, timeout=300)
    assert p.returncode == 0, p.stderr.decode()
    assert p.stdout.decode() == "".join(r[mode] for r in recs)


def test_one_pair_per_process():
    if not os.path.exists(BIN):
        pytest.skip("build/TERefiner_b200 not built")
    r = golden()[0]
    p = subprocess.run([BIN, "-M", "-r", r["s1"], "-s", r["s2"]], capture_output=True, timeout=300)
    assert p.returncode == 0 and p.stdout.decode() == r["M"]


@pytest.mark.skipif(_n_gpus() < 2, reason="needs two GPUs")
def test_pairs_dealt_to_two_gpus():
    if not os.path.exists(BIN):
        pytest.skip("build/TERefiner_b200 not built")
    recs = [r for r in golden() if "M" in r]
    with tempfile.TemporaryDirectory() as td:
        lst = os.path.join(td, "pairs.tsv")
        with open(lst, "w") as f:
            for r in recs:
                f.write(r["s1"] + "\t" + r["s2"] + "\n")
        p = subprocess.run([BIN, "-M", "--batch", lst, "--gpus", "2"], capture_output=True, timeout=300)
    assert p.returncode == 0, p.stderr.decode()
    assert p.stdout.decode() == "".join(r["M"] for r in recs)
