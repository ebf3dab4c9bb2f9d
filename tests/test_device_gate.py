"""CPU: the DeviceGate that several mergers of one GPU share (gappadder_b200/host/device_gate.hpp) -- its invariants under four
threads taking every exit (relax launch, cancelled relax, no relax), and no deadlock (tests/gate_test.cpp)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_gate_invariants_and_no_deadlock():
    out = os.path.join(ROOT, "build", "gate_test")
    os.makedirs(os.path.join(ROOT, "build"), exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", out, os.path.join(ROOT, "tests", "gate_test.cpp"), "-lpthread"])
    p = subprocess.run([out], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "violations 0" in p.stdout
