#!/usr/bin/env python3
"""Golden vectors for TERefiner's affine local aligner (tests/golden/local_affine.json).

Runs the REFERENCE's own code -- TERefiner/algorithms/local_alignment.cpp compiled as it lies by oracle/build_ref.sh into
oracle/_ref/libla_ref.so -- on seeded sequence pairs and records what LocalAlignment::optAlign's call returns
(aln_stdaln(ref, sgmt, &aln_param_blast, ALN_TYPE_LOCAL, 1): score, start1, end1, start2, end2).  Pairs on which nothing
aligns (forward score < 1) are recorded with the forward score only: the reference reads path[-1] for them.
Needs /root/reference (this container only); the vectors travel, the reference does not.

    python tests/golden/make_golden_affine.py
"""
import ctypes as C
import json
import os
import random

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def rand(rng, n, alpha=b"ACGT"):
    return bytes(rng.choice(alpha) for _ in range(n))


def mutate(s, rate, rng, alpha=b"ACGT"):
    out = bytearray()
    for ch in s:
        x = rng.random()
        if x < rate / 3:
            continue                                  # deletion
        if x < 2 * rate / 3:
            out.append(rng.choice(alpha))             # insertion before the letter
        if x < rate:
            out.append(rng.choice(alpha))             # substitution
            continue
        out.append(ch)
    return bytes(out)


def cases():
    rng = random.Random(20261018)
    out = []
    # hand-picked corners
    out += [(b"A", b"A"), (b"A", b"C"), (b"ACGT", b"ACGT"), (b"ACGTACGT", b"TTTTACGTACGTTTT"), (b"N", b"N"), (b"NNNN", b"ACGT"),
            (b"ACGTACGTTTGACCAGTAGGATCCA", b"TTTTTGTACGTTTGACAGTAGGTTTT"), (b"acgtacgtacgt", b"ACGTACGTACGT"), (b"AAAAAAAAAA", b"AAAAA"),
            (b"ACGTRYACGTACGTAAC", b"ACGTACACGTACGTAAC"), (b"GATTACA" * 20, b"GATTACA" * 7), (b"A" * 300, b"A" * 299 + b"C")]
    for _ in range(60):                                # unrelated random pairs: short local hits, first-maximum ties
        out.append((rand(rng, rng.randrange(1, 250)), rand(rng, rng.randrange(1, 250))))
    for _ in range(120):                               # a shared core with substitutions and indels inside random flanks
        core = rand(rng, rng.randrange(20, 1200))
        a = rand(rng, rng.randrange(0, 400)) + core + rand(rng, rng.randrange(0, 400))
        b = rand(rng, rng.randrange(0, 400)) + mutate(core, rng.choice([0, 0.01, 0.03, 0.06, 0.12, 0.2]), rng) + rand(rng, rng.randrange(0, 400))
        out.append((a, b))
    for _ in range(50):                                # tandem repeats and rotations (RepeatsClassifier.cpp:50-57's use)
        unit = rand(rng, rng.randrange(1, 40))
        a = mutate(unit * rng.randrange(3, 60), 0.03, rng)
        rot = rng.randrange(len(unit))
        b = mutate((unit[rot:] + unit[:rot]) * rng.randrange(3, 60), 0.03, rng)
        out.append((a, b))
    for _ in range(30):                                # N runs and other letters (aln_nt4_table: one class, score -2)
        core = rand(rng, rng.randrange(30, 600), b"ACGTN")
        out.append((core, mutate(core, 0.05, rng, b"ACGTNRY")))
    for _ in range(30):                                # two-letter alphabet: many equal maxima
        core = rand(rng, rng.randrange(5, 500), b"AC")
        out.append((mutate(core, 0.1, rng, b"AC"), mutate(core, 0.1, rng, b"AC")))
    for _ in range(24):                                # several strips of the forward kernel, long gaps, long alignments
        core = rand(rng, rng.randrange(1500, 4000))
        b = mutate(core, rng.choice([0.0, 0.02, 0.08]), rng)
        if rng.random() < 0.5:                          # a long insertion in the middle
            cut = rng.randrange(len(b))
            b = b[:cut] + rand(rng, rng.randrange(1, 120)) + b[cut:]
        out.append((rand(rng, rng.randrange(0, 600)) + core, b + rand(rng, rng.randrange(0, 600))))
    for _ in range(12):                                # contig overlaps: suffix of one = prefix of the other (scaffolding.cpp:103-105)
        a = rand(rng, rng.randrange(300, 2500))
        ov = rng.randrange(20, min(len(a), 800))
        out.append((a, mutate(a[-ov:], 0.02, rng) + rand(rng, rng.randrange(100, 1500))))
    return out


def main():
    L = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libla_ref.so"))
    recs = []
    for a, b in cases():
        fwd = L.laref_forward_score(a, b)
        rec = {"s1": a.decode(), "s2": b.decode(), "forward_score": fwd}
        if fwd >= 1:
            o = (C.c_int32 * 6)()
            L.laref_stdaln_local(a, b, o)
            rec.update(score=o[0], start1=o[1], end1=o[2], start2=o[3], end2=o[4])
        recs.append(rec)
    path = os.path.join(HERE, "local_affine.json")
    with open(path, "w") as f:
        json.dump({"generator": "tests/golden/make_golden_affine.py", "reference": "TERefiner/algorithms/local_alignment.cpp (aln_param_blast)",
                   "cases": recs}, f, separators=(",", ":"))
    print(len(recs), "cases ->", path, os.path.getsize(path), "bytes;", sum(r["forward_score"] < 1 for r in recs), "without a match")


if __name__ == "__main__":
    main()
