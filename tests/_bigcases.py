"""Realistic-size whole-binary goldens (tests/golden/big, made by tests/golden/make_golden_big.py from the reference
binary): helpers shared by the CPU host test and the GPU drop-in test."""
import gzip
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIG = os.path.join(ROOT, "tests", "golden", "big")
FLAGS = "-s 0.4 -i1 -2.0 -i2 -2.0 -x 12 -y 50 -k 10 -t 1 -m 1".split()
CASES = sorted(f[:-len(".fa.gz")] for f in os.listdir(BIG) if f.endswith(".fa.gz")) if os.path.isdir(BIG) else []


def golden(case, what):
    with open(os.path.join(BIG, "%s.%s.gz" % (case, what)), "rb") as f:
        return gzip.decompress(f.read())


def write_input(case, directory):
    fa = os.path.join(directory, case + ".fa")
    with open(fa, "wb") as f:
        f.write(golden(case, "fa"))
    return fa


def run_single(binary, case, td):
    fa = write_input(case, td)
    info = os.path.join(td, case + ".info")
    p = subprocess.run([binary] + FLAGS + ["-o", info, fa], cwd=td, capture_output=True)
    info_b = open(info, "rb").read() if os.path.exists(info) else b""
    gml = os.path.join(td, "tmp.gml")
    gml_b = open(gml, "rb").read() if os.path.exists(gml) else b""
    return p.returncode, p.stdout, info_b, gml_b, p.stderr


def check_single(binary, case, td):
    rc, out, info, gml, err = run_single(binary, case, td)
    assert rc == int(open(os.path.join(BIG, case + ".rc")).read()), err[-400:]
    assert out == golden(case, "stdout"), "%s: stdout differs" % case
    assert info == golden(case, "info"), "%s: info differs" % case
    assert gml == golden(case, "gml"), "%s: gml differs" % case


def check_batch(binary, cases, td, extra=()):
    lst = os.path.join(td, "list.tsv")
    with open(lst, "w") as f:
        for c in cases:
            f.write("%s\t%s\t%s\n" % (write_input(c, td), os.path.join(td, c + ".out"), os.path.join(td, c + ".binfo")))
    p = subprocess.run([binary] + FLAGS + ["--batch", lst] + list(extra), cwd=td, capture_output=True)
    assert p.returncode == 0, p.stderr[-400:]
    for c in cases:
        assert open(os.path.join(td, c + ".out"), "rb").read() == golden(c, "stdout"), "%s: stdout differs" % c
        assert open(os.path.join(td, c + ".binfo"), "rb").read() == golden(c, "info"), "%s: info differs" % c
        if "--no-gml" in extra:
            assert not os.path.exists(os.path.join(td, c + ".out.gml"))
        else:
            assert open(os.path.join(td, c + ".out.gml"), "rb").read() == golden(c, "gml"), "%s: gml differs" % c
    return p



def _rc_name(n):
    return n[:-2] if n.endswith("_R") else n + "_R"


def _path_lines(info):
    return [ln.split()[1:] for ln in info.decode().splitlines()]


def check_paths_modulo_ties(info, out, gml, case, max_per_root=20):
    """For graphs with more than max_per_root + 1 equal-length paths from one root the reference's choice among the
    equals follows heap addresses (GraphUtils.cpp:719-753) and is not reproducible.  Everything that does not depend
    on that choice is compared: the graph (gml bytes); every path whose root -- and whose reverse-complement path's
    root (RemoveDupRevCompPaths drops the later of the two) -- keeps fewer paths than the cap, exactly; for the
    others their number and lengths; and the original-contig part of stdout."""
    assert gml == golden(case, "gml")
    got, want = _path_lines(info), _path_lines(golden(case, "info"))
    n_by_root = {}
    for p in want:
        n_by_root[p[0]] = n_by_root.get(p[0], 0) + 1
    full = {r for r, c in n_by_root.items() if c > max_per_root}
    assert full, "case has no truncated root"

    def affected(p):
        return p[0] in full or _rc_name(p[-1]) in full

    assert sorted(p for p in got if not affected(p)) == sorted(p for p in want if not affected(p))
    assert sorted(len(p) for p in got if affected(p)) == sorted(len(p) for p in want if affected(p))
    assert sum(1 for p in got if p[0] in full) == sum(1 for p in want if p[0] in full)

    def originals(text):            # the input contigs echoed after the merged records
        recs = text.decode().split(">")
        return [r for r in recs if r and not r.startswith("NEW_CONTIG_MERGE_")]
    assert originals(out) == originals(golden(case, "stdout"))
