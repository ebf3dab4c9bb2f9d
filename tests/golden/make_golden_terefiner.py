#!/usr/bin/env python3
"""Golden outputs of the PREBUILT reference binary for the two TERefiner_1 modes that are the affine local aligner
(tests/golden/terefiner_modes.json):

    TERefiner_1 -M -r SEQ1 -s SEQ2      LocalAlignment::optAlign           (TERefiner/main.cpp:207-213)
    TERefiner_1 -A -r SEQ1 -s SEQ2      RepeatsClassifier::validateRepeats (TERefiner/main.cpp:202-206)

The binary is /root/reference/TERefiner/TERefiner_1 (or /root/reference/TERefiner_1), run from an executable copy in a
temporary directory.  Pairs on which one of the alignments involved finds nothing are left out: the reference reads
path[-1] there (local_alignment.cpp:611-614, :817) and prints whatever the heap holds.  Which pairs those are is decided
with the reference's own forward pass (oracle/_ref/libla_ref.so).  "align" holds LocalAlignment::align's twelve outputs (-1 where
the method leaves them untouched) from the reference class compiled into that library: no TERefiner_1 mode prints them.
Needs /root/reference (this container only).

    python tests/golden/make_golden_terefiner.py
"""
import ctypes as C
import json
import os
import random
import shutil
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
COMP = bytes.maketrans(b"ACGTacgt", b"TGCATGCA")


def supplementary(s):
    return bytes(c if chr(c) in "ACGTacgt" else ord("N") for c in s[::-1]).translate(COMP)


def rand(rng, n, alpha=b"ACGT"):
    return bytes(rng.choice(alpha) for _ in range(n))


def mutate(s, rate, rng):
    out = bytearray()
    for ch in s:
        x = rng.random()
        if x < rate / 3:
            continue
        if x < 2 * rate / 3:
            out.append(rng.choice(b"ACGT"))
        if x < rate:
            out.append(rng.choice(b"ACGT"))
            continue
        out.append(ch)
    return bytes(out)


def cases():
    rng = random.Random(4242)
    out = [(b"ACGTACGTTTGACCAGTAGGATCCA", b"TTTTTGTACGTTTGACAGTAGGTTTT"),
           (b"ACGTACGTTTGACCAGTAGGATCCAGGATTTACCCA", b"CAGTAGGATCCAGGATTTACCCAACGTACGTTTGAC")]
    for _ in range(40):                                  # a shared core inside flanks
        core = rand(rng, rng.randrange(20, 600))
        out.append((rand(rng, rng.randrange(5, 150)) + core + rand(rng, rng.randrange(5, 150)),
                    rand(rng, rng.randrange(5, 150)) + mutate(core, rng.choice([0, 0.02, 0.06]), rng) + rand(rng, rng.randrange(5, 150))))
    for _ in range(40):                                  # tandem repeats, rotated, either strand (what -A is for)
        unit = rand(rng, rng.randrange(3, 60))
        a = mutate(unit * rng.randrange(2, 12), 0.02, rng)
        rot = rng.randrange(len(unit))
        b = mutate((unit[rot:] + unit[:rot]) * rng.randrange(2, 12), 0.02, rng)
        out.append((a, supplementary(b) if rng.random() < 0.5 else b))
    for _ in range(20):                                  # unrelated
        out.append((rand(rng, rng.randrange(30, 300)), rand(rng, rng.randrange(30, 300))))
    for _ in range(10):                                  # N and lower case
        core = rand(rng, rng.randrange(40, 300), b"ACGTNacgt")
        out.append((core, mutate(core, 0.04, rng)))
    return [(a, b) for a, b in out if a and b]


def main():
    L = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libla_ref.so"))
    L.laref_forward_score.restype = C.c_int32

    def aligned(a, b):
        return bool(a) and bool(b) and L.laref_forward_score(a, b) >= 1

    def local(a, b):
        o = (C.c_int32 * 6)()
        L.laref_stdaln_local(a, b, o)
        return o[1], o[2], o[3], o[4]

    def rest_defined(a, b):
        """Every alignment optAlignWithRestSecondOpt(a, b) makes finds something (or its input is empty, which is defined)."""
        if not aligned(a, b):
            return False
        s1, e1, s2, e2 = local(a, b)
        ra, rb = a[:s1 - 1] + a[e1:], b[:s2 - 1] + b[e2:]
        return (not ra or not rb) or aligned(ra, rb)

    src = next(p for p in ("/root/reference/TERefiner/TERefiner_1", "/root/reference/TERefiner_1") if os.path.exists(p))
    recs = []
    with tempfile.TemporaryDirectory() as td:
        exe = os.path.join(td, "TERefiner_1")
        shutil.copy(src, exe)
        os.chmod(exe, 0o755)
        for a, b in cases():
            rec = {"s1": a.decode(), "s2": b.decode()}
            if aligned(a, b):
                rec["M"] = subprocess.run([exe, "-M", "-r", a, "-s", b], capture_output=True, check=True).stdout.decode()
                s1, e1, s2, e2 = local(a, b)         # LocalAlignment::align (no TERefiner_1 mode prints it): the compiled reference class
                if ((not (s1 > 1 and s2 > 1)) or aligned(a[:s1 - 1], b[:s2 - 1])) and ((not (e1 < len(a) and e2 < len(b))) or aligned(a[e1:], b[e2:])):
                    w = (C.c_int32 * 12)()
                    L.laref_align(a, b, w)
                    rec["align"] = list(w)
            if rest_defined(a, b) and rest_defined(a, supplementary(b)):
                rec["A"] = subprocess.run([exe, "-A", "-r", a, "-s", b], capture_output=True, check=True).stdout.decode()
            recs.append(rec)
    path = os.path.join(HERE, "terefiner_modes.json")
    with open(path, "w") as f:
        json.dump({"generator": "tests/golden/make_golden_terefiner.py", "binary": src, "cases": recs}, f, separators=(",", ":"))
    print(len(recs), "pairs,", sum("M" in r for r in recs), "with -M,", sum("A" in r for r in recs), "with -A,", sum("align" in r for r in recs),
          "with align ->", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
