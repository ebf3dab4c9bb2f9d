#!/usr/bin/env python
"""Generates tests/golden/*.json from the REFERENCE's own code (oracle/_ref, built by
oracle/build_ref.sh from /root/reference).  Run in the build container only; the vectors are
committed so that machines without /root/reference can still pin the oracle and the product.

  evaluate.json      per-pair outputs of ContigsCompactor::Evaluate (relax and non-relax)
  quickcheck.json    QuickCheckerContigsMatch::IsMatchFeasible on node pairs
  significant.json   IsScoreSignificant on a grid of inputs
  revcomp.json       FastaSequence::RevsereComplement
  contigsmerger/     whole-binary runs: input FASTA, stdout, .merge.info (GAPPadder's flags, -t 1)
"""
import json
import os
import random
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")]
import _oracle  # noqa: E402
import synth_gaps  # noqa: E402

FLAGS = "-s 0.4 -i1 -2.0 -i2 -2.0 -x 12 -y 50 -k 10 -t 1 -m 1".split()


def rnd(rng, n, alpha):
    return "".join(rng.choice(alpha) for _ in range(n))


def main():
    assert _oracle.ref_lib() is not None, "build oracle/_ref first (oracle/build_ref.sh)"
    rng = random.Random(20261017)
    # ---- Evaluate -------------------------------------------------------------------------
    ev = []
    pairs = []
    for _ in range(260):
        alpha = rng.choice(["ACGT", "ACGT", "AC", "ACGTN", "A"])
        m, n = rng.randint(1, 120), rng.randint(1, 120)
        a = rnd(rng, m, alpha)
        mode = rng.random()
        if mode < 0.45:
            k = rng.randint(1, m)
            b = a[-k:] + rnd(rng, max(0, n - k), alpha)
        elif mode < 0.6:
            b = rnd(rng, rng.randint(0, 12), alpha) + a + rnd(rng, rng.randint(0, 12), alpha)
        elif mode < 0.7:
            k = rng.randint(1, m)
            b = rnd(rng, max(0, n - k), alpha) + a[:k]
        else:
            b = rnd(rng, n, alpha)
        b = b or "A"
        if rng.random() < 0.3:          # sprinkle mutations
            bl = list(b)
            for p in range(len(bl)):
                if rng.random() < 0.05:
                    bl[p] = rng.choice(alpha)
            b = "".join(bl)
        pairs.append((a, b))
    # a few real-size contig pairs from a synthetic gap
    recs = synth_gaps.make_gap(5, synth_gaps.CONFIGS["small"])
    nodes = []
    for _, s in recs:
        nodes += [s.decode(), _oracle.oracle_revcomp(s).decode()]
    for _ in range(40):
        pairs.append((rng.choice(nodes), rng.choice(nodes)))
    for a, b in pairs:
        for relax in (False, True):
            r = _oracle.ref_evaluate(a.encode(), b.encode(), relax)
            m = r.pop("merged")
            r["merged"] = m.decode() if m is not None else None
            ev.append(dict(s1=a, s2=b, relax=int(relax), **r))
    json.dump(dict(params=_oracle.GAPPADDER_PARAMS, cases=ev), open(os.path.join(HERE, "evaluate.json"), "w"))
    # ---- quick check ----------------------------------------------------------------------
    qc = []
    lib = _oracle.ref_lib()
    for _ in range(150):
        L = rng.randint(30, 200)
        a = rnd(rng, rng.randint(30, 300), "ACGT" if rng.random() < 0.8 else "ACGTN")
        if rng.random() < 0.5:
            st = rng.randint(0, max(0, len(a) - 30))
            piece = a[st:st + rng.randint(10, 30)]
            b = rnd(rng, L, "ACGT")
            pos = rng.choice([0, max(0, len(b) - len(piece)), rng.randint(0, max(0, len(b) - len(piece)))])
            b = b[:pos] + piece + b[pos + len(piece):]
        else:
            b = rnd(rng, L, "ACGT")
        b = b[:max(30, len(b))]
        if len(b) < 30:
            continue
        for k in (10, 6):
            qc.append(dict(si=a, sj=b, k=k, feasible=int(lib.cmref_quickcheck(a.encode(), b.encode(), k))))
    json.dump(qc, open(os.path.join(HERE, "quickcheck.json"), "w"))
    # ---- IsScoreSignificant ---------------------------------------------------------------
    sig = []
    for _ in range(600):
        l1, l2 = rng.randint(1, 400), rng.randint(1, 400)
        nclip = rng.randint(0, 50)
        if rng.random() < 0.5:
            row, col = l1 - nclip, rng.randint(0, l2)
        else:
            row, col = rng.randint(0, l1), l2 - nclip
        if row < 0 or col < 0:
            continue
        score = rng.randint(-5, min(l1, l2))
        sig.append(dict(score=score, l1=l1, l2=l2, row=row, col=col, nclip=nclip,
                        res=int(lib.cmref_is_score_significant(score, l1, l2, row, col, nclip))))
    json.dump(dict(params=_oracle.GAPPADDER_PARAMS, cases=sig), open(os.path.join(HERE, "significant.json"), "w"))
    # ---- reverse complement ---------------------------------------------------------------
    import ctypes as C
    rc = []
    for _ in range(40):
        s = rnd(rng, rng.randint(0, 60), "ACGTNacgtnRYKMXB")
        out = C.create_string_buffer(len(s) + 1)
        lib.cmref_revcomp(s.encode(), out)
        rc.append(dict(s=s, rc=out.value.decode()))
    json.dump(rc, open(os.path.join(HERE, "revcomp.json"), "w"))
    # ---- whole-binary runs ----------------------------------------------------------------
    outdir = os.path.join(HERE, "contigsmerger")
    os.makedirs(outdir, exist_ok=True)
    cases = [("ka", None)] + [("tiny%d" % s, ("tiny", s)) for s in (1, 2, 3, 4)] + [("small%d" % s, ("small", s)) for s in (1, 2)] \
        + [("noisy%d" % s, ("noisy", s)) for s in (1, 2)] + [("single", "single"), ("empty", "empty")]
    for name, spec in cases:
        fa = os.path.join(outdir, name + ".fa")
        if spec is None:
            open(fa, "w").write(">a\nACGTACGTAGCTAGCTAGCTAGCATCGATCGATCGATCAGCTAGCTAGCATCGATCAGCTACGACTAGC\n"
                                ">b\nGATCGATCAGCTAGCTAGCATCGATCAGCTACGACTAGCTTTTGGGGCCCCAAAATTTTGGGCCCAATTGGCCAATT\n")
        elif spec == "single":
            open(fa, "w").write(">only\nACGTACGTAGCTAGCTAGCTAGCATCGATCGATCGATCAGCTAGCTAGC\n")
        elif spec == "empty":
            open(fa, "w").write("")
        else:
            synth_gaps.write_fasta(fa, synth_gaps.make_gap(spec[1], synth_gaps.CONFIGS[spec[0]]))
        with tempfile.TemporaryDirectory() as td:
            info = os.path.join(td, "x.info")
            p = subprocess.run([_oracle.ref_binary()] + FLAGS + ["-o", info, fa], cwd=td, capture_output=True)
            open(os.path.join(outdir, name + ".stdout"), "wb").write(p.stdout)
            open(os.path.join(outdir, name + ".info"), "wb").write(open(info, "rb").read() if os.path.exists(info) else b"")
            gml = os.path.join(td, "tmp.gml")
            open(os.path.join(outdir, name + ".gml"), "wb").write(open(gml, "rb").read() if os.path.exists(gml) else b"")
            open(os.path.join(outdir, name + ".rc"), "w").write(str(p.returncode))
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
