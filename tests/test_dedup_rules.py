"""CPU: the dedup rules of the reference's TERefiner_1 (MergeContigs.py:15-70 -> `-U`, `-P -c cutoff [-g]`;
TERefiner/refiner.cpp:660-801,1045-1140, Alignment.cpp:397-437) as restated in gp_dedup_unique_names / gp_dedup_decide,
against golden vectors produced by the reference's own prebuilt binary on hand-written BAM files
(tests/golden/make_dedup_golden.py -> tests/golden/dedup/dedup_rules.json)."""
import json
import os

import numpy as np

import gappadder_b200 as g

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dedup", "dedup_rules.json")))


def _summary(cig):
    m = sum(l for o, l in cig if o == "M")
    other = sum(l for o, l in cig if o in "SHI")
    return (1 if len(cig) == 1 and cig[0][0] == "M" else 0, m, other)


def test_rules_against_the_reference_binary():
    removed_somewhere = 0
    for case in GOLD["rules"]:
        names = [c[0].encode() for c in case["contigs"]]
        lens = [c[1] for c in case["contigs"]]
        recs = np.zeros(len(case["records"]), dtype=g.DEDUP_RECORD_DTYPE)
        for k, (q, r, cig) in enumerate(case["records"]):
            recs[k] = (q, r) + _summary(cig)
        removed = g.dedup_decide(recs, names, lens, case["cutoff"], bool(case["g"]))
        kept = [c[0] for c, rm in zip(case["contigs"], removed) if not rm]
        assert kept == case["kept"], case
        removed_somewhere += int(removed.any())
    assert removed_somewhere > 50          # the vectors do exercise the rules


def test_unique_names_against_the_reference_binary():
    for case in GOLD["unique"]:
        names = [c[0].encode() for c in case["contigs"]]
        keep = g.dedup_unique_names(names)
        assert [c[0] for c, k in zip(case["contigs"], keep) if k] == case["kept"], case
        if case["kept_lens"] is not None:      # of equally named records the FIRST one stays
            assert [c[1] for c, k in zip(case["contigs"], keep) if k] == case["kept_lens"], case


def test_python_oracle_rules_against_the_reference_binary():
    """oracle/dedup_oracle.py (the checker of the GPU dedup tests) is pinned to the same vectors."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import dedup_oracle as do
    for case in GOLD["rules"]:
        names = [c[0].encode() for c in case["contigs"]]
        lens = [c[1] for c in case["contigs"]]
        removed = do.decide([(q, r, [tuple(x) for x in cig]) for q, r, cig in case["records"]], names, lens, case["cutoff"], bool(case["g"]))
        assert [c[0] for c, rm in zip(case["contigs"], removed) if not rm] == case["kept"], case
    for case in GOLD["unique"]:
        keep = do.unique_names([c[0].encode() for c in case["contigs"]])
        assert [c[0] for c, k in zip(case["contigs"], keep) if k] == case["kept"], case
