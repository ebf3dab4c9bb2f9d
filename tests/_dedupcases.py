"""Synthetic contig sets for the dedup stage tests: a synth_gaps gap plus the things the stage exists to remove --
contigs contained in another one (either strand), near-duplicates of similar length, exact duplicates under another
name, records with a repeated name -- and its expected output from oracle/dedup_oracle.py."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import synth_gaps  # noqa: E402
import dedup_oracle  # noqa: E402
from _oracle import oracle_candidate_pairs, oracle_evaluate, oracle_revcomp  # noqa: E402

FLAGS = "-s 0.4 -i1 -2.0 -i2 -2.0 -x 12 -y 50 -k 10 -t 1 -m 1".split()


def make_set(seed: int, config: str = "small") -> bytes:
    rng = np.random.default_rng(seed)
    contigs = [(n, bytes(s)) for n, s in synth_gaps.make_gap(seed, synth_gaps.CONFIGS[config])]
    extra = []
    pick = lambda: contigs[int(rng.integers(0, len(contigs)))]          # noqa: E731
    for k in range(3):                                                   # contained: a substring, forward or reverse complement
        name, s = pick()
        a = int(rng.integers(0, max(1, len(s) // 3)))
        b = len(s) - int(rng.integers(0, max(1, len(s) // 3)))
        sub = s[a:b]
        if k % 2:
            sub = oracle_revcomp(sub)
        extra.append((b"Z_sub%d_of_%s" % (k, name.encode() if isinstance(name, str) else name), sub))
    for k in range(2):                                                   # near-duplicate of similar length: a few substitutions, ends trimmed
        name, s = pick()
        t = bytearray(s[int(rng.integers(0, 4)):len(s) - int(rng.integers(0, 4))])
        for _ in range(int(rng.integers(0, 3))):
            p = int(rng.integers(0, len(t)))
            t[p] = b"ACGT"[(b"ACGT".index(t[p]) + 1) % 4] if t[p] in b"ACGT" else t[p]
        extra.append((b"Y_dup%d" % k, bytes(t) if k else oracle_revcomp(bytes(t))))
    name, s = pick()
    extra.append((b"X_same", s))                                          # exact duplicate, other name
    extra.append((contigs[0][0].encode() if isinstance(contigs[0][0], str) else contigs[0][0], b"ACGTTGCAAGGCTTAACCGGTTAATTCCGGAGAGTCTCAGAGTTTGCA"))   # repeated name
    allc = [(n.encode() if isinstance(n, str) else n, s) for n, s in contigs] + extra
    order = rng.permutation(len(allc))
    out = b""
    for i in order:
        n, s = allc[int(i)]
        if n == allc[-1][0] and int(i) == len(allc) - 1:
            pass
        out += b">" + n + (b" extra words" if int(i) % 5 == 0 else b"") + b"\n"
        for p in range(0, len(s), 70):
            out += s[p:p + 70] + b"\n"
    return out


def expected(text: bytes, cutoff: float, contained: bool):
    return dedup_oracle.dedup(text, cutoff, contained, oracle_evaluate, oracle_candidate_pairs, oracle_revcomp, k=10, frac_loss=0.4)
