#!/usr/bin/env python
"""Generates tests/golden/big/ from the REFERENCE binary (oracle/_ref/ContigsMerger, built by
oracle/build_ref.sh from /root/reference): whole-binary runs at realistic size, GAPPadder's flags, -t 1.
Run in the build container only (takes 10-20 minutes of CPU); the outputs are committed gzip-compressed.

  cfg1_s1, cfg1_s2   two full BASELINE cfg1 gaps (40 contigs, 300-3000 bp): 32-deep relax chains, merged rows to 8 kb
  cfg3_s15, cfg3_s43 BASELINE cfg3/cfg4 shape: 80 contigs and 30 contigs
  cfg5r_s1           reduced cfg5 (long-contig stress): 14 x 8 kb contigs on a repeat-rich 20 kb locus
  fan1, fan2, fan3   overlap graphs with more than max_per_root + 1 = 21 equal-length paths from one root and a
                     multi-node strongly connected component: the truncation of AbstractGraph::FindSimplePathsTopSort
                     (GraphUtils.cpp:719-753).  WHICH of the equal-length paths survive depends on heap addresses in the
                     reference (a std::set of pointers), i.e. on the C library's allocator: fan1 and fan3 are cases in which
                     glibc's order here equals the drop-in's deterministic rule (reverse insertion order) and are compared
                     byte for byte; fan2 is one in which it does not (the reference keeps end_00 and end_10..29) and is
                     compared modulo that choice (tests/_bigcases.py: same roots, same number and lengths of paths)
  iupac1, iupac2     contigs with IUPAC / other letters: 11 letters besides A C G T N each, 22 over both (a batch of both
                     exceeds the packed table's 16 codes unless letters are renamed per gap)
"""
import gzip
import os
import random
import subprocess
import sys
import tempfile
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")]
import _oracle  # noqa: E402
import synth_gaps  # noqa: E402

FLAGS = "-s 0.4 -i1 -2.0 -i2 -2.0 -x 12 -y 50 -k 10 -t 1 -m 1".split()
OUT = os.path.join(HERE, "big")


def rnd(rng, n):
    return "".join(rng.choice("ACGT") for _ in range(n))


def fan_case(seed, n_ends, two_level):
    """One root whose tail every end starts with (root -> end edges, all paths the same length), optionally a middle
    layer (root -> mid -> ends), plus a three-node cycle from a circular locus (a multi-node SCC)."""
    rng = random.Random(seed)
    recs = []
    root = rnd(rng, 300)
    recs.append(("root", root))
    if two_level:
        mid = root[-80:] + rnd(rng, 220)
        recs.append(("mid", mid))
        for k in range(n_ends):
            recs.append(("end_%02d" % k, mid[-70:] + rnd(rng, 150 + 7 * k)))
        for k in range(6):
            recs.append(("short_%d" % k, root[-60:] + rnd(rng, 140 + 11 * k)))
    else:
        for k in range(n_ends):
            recs.append(("end_%02d" % k, root[-60:] + rnd(rng, 200 + 3 * k)))
    circ = rnd(rng, 600)
    recs.append(("cyc_a", circ[0:300]))
    recs.append(("cyc_b", circ[200:500]))
    recs.append(("cyc_c", circ[400:600] + circ[0:100]))
    recs.append(("cyc_tail", circ[250:300] + rnd(rng, 180)))
    return [(n, s.encode()) for n, s in recs]


def iupac_case(seed, letters):
    """A small noisy gap whose contigs carry other letters: each of `letters` replaces about 0.4 % of the bases."""
    rng = random.Random(seed)
    out = []
    for name, seq in synth_gaps.make_gap(seed, synth_gaps.CONFIGS["noisy"]):
        b = bytearray(seq)
        for p in range(len(b)):
            if rng.random() < 0.004 * len(letters):
                b[p] = ord(rng.choice(letters))
        out.append((name, bytes(b)))
    return out


CASES = {
    "cfg1_s1": lambda: synth_gaps.make_gap(1, synth_gaps.CONFIGS["cfg1"]),
    "cfg1_s2": lambda: synth_gaps.make_gap(2, synth_gaps.CONFIGS["cfg1"]),
    "cfg3_s15": lambda: synth_gaps.make_gap(15, synth_gaps.CONFIGS["cfg3"]),
    "cfg3_s43": lambda: synth_gaps.make_gap(43, synth_gaps.CONFIGS["cfg3"]),
    "cfg5r_s1": lambda: synth_gaps.make_gap(1, synth_gaps.CONFIGS["cfg5r"]),
    "fan1": lambda: fan_case(101, 26, False),
    "fan2": lambda: fan_case(102, 30, True),
    "fan3": lambda: fan_case(33, 26, True),
    "iupac1": lambda: iupac_case(201, "RYKMSWBDHVX"),
    "iupac2": lambda: iupac_case(202, "UEFIJLOPQZX"),
}


def run_case(name):
    fa = os.path.join(OUT, name + ".fa")
    synth_gaps.write_fasta(fa, CASES[name]())
    with tempfile.TemporaryDirectory() as td:
        info = os.path.join(td, "x.info")
        p = subprocess.run([_oracle.ref_binary()] + FLAGS + ["-o", info, fa], cwd=td, capture_output=True)
        outs = {"stdout": p.stdout, "info": open(info, "rb").read() if os.path.exists(info) else b"",
                "gml": open(os.path.join(td, "tmp.gml"), "rb").read() if os.path.exists(os.path.join(td, "tmp.gml")) else b""}
    for k, v in outs.items():
        with open(os.path.join(OUT, name + "." + k + ".gz"), "wb") as f:
            f.write(gzip.compress(v, 9, mtime=0))
    with open(fa, "rb") as f:
        data = f.read()
    with open(fa + ".gz", "wb") as f:
        f.write(gzip.compress(data, 9, mtime=0))
    os.remove(fa)
    open(os.path.join(OUT, name + ".rc"), "w").write(str(p.returncode))
    return name, p.returncode, len(outs["stdout"]), len(outs["info"])


def main():
    os.makedirs(OUT, exist_ok=True)
    names = sys.argv[1:] or sorted(CASES)
    with ThreadPoolExecutor(max_workers=6) as ex:
        for r in ex.map(run_case, names):
            print(*r, flush=True)


if __name__ == "__main__":
    main()
