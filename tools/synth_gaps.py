"""Seeded synthetic gap sets for tests and bench.py.

One "gap" = the FASTA ContigsMerger receives for one scaffold gap: Velvet-style contigs that are
noisy substrings (or reverse complements) of one random locus.  Names follow the prefix
assemble_gaps.py:133 gives Velvet contigs: <k>_<subk>_NODE_<i>_length_<L>_cov_<x>.

Deterministic for a given (config, seed): numpy's PCG64 Generator and integer draws only.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import List, Tuple

import numpy as np

_COMP = np.zeros(256, dtype=np.uint8)
for a, b in zip(b"ACGTN", b"TGCAN"):
    _COMP[a] = b
_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


@dataclass
class GapSpec:
    n_contigs: Tuple[int, int] = (40, 40)     # inclusive range of contigs per gap
    length: Tuple[int, int] = (300, 3000)     # inclusive range of contig length
    locus: int = 8000                         # locus length the contigs are sampled from
    sub_rate: float = 0.002                   # substitution rate
    indel_rate: float = 0.0                   # per-base chance of a 1-base indel
    rc_frac: float = 0.5                      # fraction of contigs emitted as reverse complement
    n_rate: float = 0.0                       # per-base chance of an 'N'
    repeats: Tuple[int, int, int, int] = (0, 0, 0, 0)  # (min copies, max copies, min len, max len) of planted repeats
    repeat_div: float = 0.01                  # divergence between repeat copies
    ksets: Tuple[int, ...] = (30,)            # velvet k values cycled through for names


CONFIGS = {
    # BASELINE.json configs[0]/[1]: 200 gaps x 40 contigs, 300-3000 bp, 8 kb locus
    "cfg1": GapSpec(),
    # configs[2]/[3]: contigs/gap ~ U[10,80], six k-mer sets (configuration.json:42-66)
    "cfg3": GapSpec(n_contigs=(10, 80), ksets=(30, 36, 42, 48, 54, 60)),
    # configs[4]: long-contig stress, 200 contigs x 10 kb on a 40 kb repeat-rich locus
    "cfg5": GapSpec(n_contigs=(200, 200), length=(10000, 10000), locus=40000,
                    sub_rate=0.002, repeats=(2, 4, 1000, 3000)),
    # reduced cfg5 the reference binary can finish (golden outputs): 14 contigs x 8 kb on a repeat-rich 20 kb locus
    "cfg5r": GapSpec(n_contigs=(14, 14), length=(8000, 8000), locus=20000, sub_rate=0.002, repeats=(2, 4, 1000, 3000)),
    # small shapes for unit tests
    "tiny": GapSpec(n_contigs=(6, 6), length=(60, 220), locus=500, sub_rate=0.01),
    "small": GapSpec(n_contigs=(12, 12), length=(100, 600), locus=1500, sub_rate=0.005, indel_rate=0.001),
    "noisy": GapSpec(n_contigs=(10, 10), length=(80, 400), locus=900, sub_rate=0.02, indel_rate=0.005, n_rate=0.003),
}


def revcomp_bytes(a: np.ndarray) -> np.ndarray:
    return _COMP[a[::-1]]


def make_locus(rng: np.random.Generator, spec: GapSpec) -> np.ndarray:
    locus = _ACGT[rng.integers(0, 4, size=spec.locus)]
    lo, hi, rl_lo, rl_hi = spec.repeats
    if hi > 0:
        n_fam = int(rng.integers(1, 4))
        for _ in range(n_fam):
            rl = int(rng.integers(rl_lo, rl_hi + 1))
            unit = _ACGT[rng.integers(0, 4, size=rl)]
            for _ in range(int(rng.integers(lo, hi + 1))):
                cp = unit.copy()
                mut = rng.random(rl) < spec.repeat_div
                cp[mut] = _ACGT[rng.integers(0, 4, size=int(mut.sum()))]
                pos = int(rng.integers(0, spec.locus - rl))
                locus[pos:pos + rl] = cp
    return locus


FLANK_LENGTH = 995     # flank_length - 5 with GAPPadder's flank_length = 1000 (gnrt_pos_true_seqs.py:93-99)


def make_flanks(seed: int, spec: GapSpec) -> List[Tuple[str, bytes]]:
    """The two flanks of the gap make_gap(seed, spec) draws its contigs from (BASELINE cfg2): the first and the last
    995 bases of the same locus, named <scaffold>_<gap>_left / _right like gnrt_pos_true_seqs.py:93-99."""
    rng = np.random.default_rng(int(seed))
    locus = make_locus(rng, spec)
    L = min(FLANK_LENGTH, spec.locus)
    return [("0_%d_left" % seed, locus[:L].tobytes()), ("0_%d_right" % seed, locus[spec.locus - L:].tobytes())]


def make_gap(seed: int, spec: GapSpec) -> List[Tuple[str, bytes]]:
    """Returns [(name, sequence bytes)] for one gap."""
    rng = np.random.default_rng(int(seed))
    locus = make_locus(rng, spec)
    n = int(rng.integers(spec.n_contigs[0], spec.n_contigs[1] + 1))
    out = []
    for i in range(n):
        L = int(rng.integers(spec.length[0], spec.length[1] + 1))
        L = min(L, spec.locus)
        start = int(rng.integers(0, spec.locus - L + 1))
        s = locus[start:start + L].copy()
        if spec.sub_rate > 0:
            mut = rng.random(L) < spec.sub_rate
            s[mut] = _ACGT[rng.integers(0, 4, size=int(mut.sum()))]
        if spec.indel_rate > 0:
            ev = rng.random(L) < spec.indel_rate
            kind = rng.integers(0, 2, size=L)
            ins = _ACGT[rng.integers(0, 4, size=L)]
            pieces = []
            for p in range(L):
                if ev[p]:
                    if kind[p] == 0:
                        continue            # deletion
                    pieces.append(ins[p])   # insertion before the base
                pieces.append(s[p])
            s = np.array(pieces, dtype=np.uint8)
        if spec.n_rate > 0:
            nm = rng.random(len(s)) < spec.n_rate
            s[nm] = ord("N")
        if rng.random() < spec.rc_frac:
            s = revcomp_bytes(s)
        k = spec.ksets[i % len(spec.ksets)]
        cov = 5 + int(rng.integers(0, 4000)) / 100.0
        name = "%d_%d_NODE_%d_length_%d_cov_%.6f" % (k, k - 5, i + 1, len(s), cov)
        out.append((name, s.tobytes()))
    return out


def write_fasta(path: str, records: List[Tuple[str, bytes]], width: int = 60) -> None:
    with open(path, "wb") as f:
        for name, seq in records:
            f.write(b">" + name.encode() + b"\n")
            for o in range(0, len(seq), width):
                f.write(seq[o:o + width] + b"\n")


def make_gap_set(config: str, n_gaps: int, first_seed: int = 1):
    spec = CONFIGS[config]
    return [make_gap(first_seed + g, spec) for g in range(n_gaps)]


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser(description=__doc__)
    ap.add_argument("--config", default="cfg1", choices=sorted(CONFIGS))
    ap.add_argument("--gaps", type=int, default=4)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--out", required=True, help="output directory; one gap_<seed>.fa per gap")
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    for g in range(a.gaps):
        write_fasta(os.path.join(a.out, "gap_%d.fa" % (a.seed + g)), make_gap(a.seed + g, CONFIGS[a.config]))
