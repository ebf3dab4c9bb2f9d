#!/usr/bin/env python
"""tests/golden/make_dedup_golden.py -- golden vectors for the dedup rules of TERefiner_1 (run HERE, where /root/reference
exists; the vectors are committed, the GPU box never needs the reference).

The reference's dedup stage (MergeContigs.py:15-70) is `TERefiner_1 -U` (unique names, TERefiner/refiner.cpp:1045-1140)
followed by `TERefiner_1 -P -b self.bam -r contigs.fa -o out.fa -c cutoff [-g]` (refiner.cpp:660-801 with
Alignment.cpp:397-437): rules over (query name, reference name, CIGAR) of every record of a BWA self-alignment.  BWA is
not vendored, but the RULES are in-tree and the prebuilt /root/reference/TERefiner/TERefiner_1 runs here, so this script
drives it with BAM files written below (BGZF + BAM records by hand; bamtools reads them without an index) and records which
contigs it removes: dedup_rules.json = [{contigs: [[name, len]...], records: [[q, r, [[op, len]...]]...], cutoff, g,
kept: [names in output order]}].  tests/test_dedup_rules.py pins gp_dedup_decide / gp_dedup_unique_names to them."""
import json
import os
import random
import struct
import subprocess
import tempfile
import zlib

TER = "/root/reference/TERefiner/TERefiner_1"
OPS = "MIDNSHP=X"
EOF_BLOCK = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


def bgzf_block(data):
    c = zlib.compressobj(6, zlib.DEFLATED, -15)
    comp = c.compress(data) + c.flush()
    hdr = struct.pack("<BBBBIBBHBBHH", 31, 139, 8, 4, 0, 0, 255, 6, 66, 67, 2, len(comp) + 25)
    return hdr + comp + struct.pack("<II", zlib.crc32(data) & 0xFFFFFFFF, len(data))


def write_bam(path, refs, recs):
    """refs: [(name, len)]; recs: [(qname, ref index, pos0, [(op, len)])]"""
    text = "@HD\tVN:1.0\tSO:unsorted\n" + "".join("@SQ\tSN:%s\tLN:%d\n" % r for r in refs)
    out = b"BAM\x01" + struct.pack("<i", len(text)) + text.encode() + struct.pack("<i", len(refs))
    for n, l in refs:
        out += struct.pack("<i", len(n) + 1) + n.encode() + b"\0" + struct.pack("<i", l)
    for q, rid, pos, cig in recs:
        qlen = sum(l for o, l in cig if o in "MIS=X")
        name = q.encode() + b"\0"
        body = struct.pack("<iiBBHHHiiii", rid, pos, len(name), 60, 4680, len(cig), 0, qlen, -1, -1, 0)
        body += name + b"".join(struct.pack("<I", (l << 4) | OPS.index(o)) for o, l in cig)
        body += bytes([0x11] * ((qlen + 1) // 2)) + b"\xff" * qlen
        out += struct.pack("<i", len(body)) + body
    with open(path, "wb") as f:
        for i in range(0, len(out), 60000):
            f.write(bgzf_block(out[i:i + 60000]))
        f.write(EOF_BLOCK)


def write_fasta(path, contigs):
    fai = []
    with open(path, "w") as f:
        off = 0
        for name, l in contigs:
            head = ">%s\n" % name
            f.write(head + "A" * l + "\n")
            fai.append("%s\t%d\t%d\t%d\t%d" % (name, l, off + len(head), l, l + 1))
            off += len(head) + l + 1
    open(path + ".fai", "w").write("\n".join(fai) + "\n")


def kept_names(path):
    return [ln[1:].strip() for ln in open(path) if ln.startswith(">")]


def random_cigar(rng, qlen):
    kind = rng.randrange(8)
    if kind == 0:
        return [("M", qlen)]
    if kind == 1:
        return [("M", max(1, qlen - rng.randrange(1, 4)))]            # single M shorter than the contig
    if kind == 2:
        return [("M", qlen + rng.randrange(1, 3))]                     # single M longer than the contig (never from BWA; the rule has a <=)
    m = max(1, int(qlen * rng.choice([0.5, 0.7, 0.84, 0.85, 0.86, 0.9, 0.95, 0.99])))
    rest = max(0, qlen - m)
    if kind == 3:
        return [("M", m), ("S", rest)] if rest else [("M", m)]
    if kind == 4:
        a = rest // 2
        return [x for x in [("S", a), ("M", m), ("S", rest - a)] if x[1] > 0]
    if kind == 5:
        a = m // 2
        return [x for x in [("M", a), ("I", rest), ("M", m - a)] if x[1] > 0]
    if kind == 6:
        a = m // 2
        return [x for x in [("H", rest), ("M", a), ("D", 3), ("M", m - a)] if x[1] > 0]
    return [x for x in [("M", m // 2), ("I", rest // 2), ("M", m - m // 2), ("S", rest - rest // 2)] if x[1] > 0]


def main():
    here = os.path.dirname(os.path.abspath(__file__))
    rng = random.Random(20261018)
    rule_cases, unique_cases = [], []
    with tempfile.TemporaryDirectory() as td:
        global TER
        import shutil
        ter = os.path.join(td, "TERefiner_1")                     # the reference tree is read-only and the binary has no x bit
        shutil.copy(TER, ter)
        os.chmod(ter, 0o755)
        TER = ter
        for case in range(240):
            n = rng.randrange(2, 9)
            names = rng.sample(["NODE_%d_length_%d_cov_%d" % (rng.randrange(1, 30), rng.randrange(40, 3000), rng.randrange(1, 9)) for _ in range(24)]
                               + ["NEW_CONTIG_MERGE_%d" % k for k in range(1, 12)] + ["a", "b", "B", "a_R", "10", "9"], n)
            if len(set(names)) != n:
                continue
            contigs = []
            for nm in names:
                l = rng.choice([40, 100, 100, 250, 1000, 1000, 1150, 1180, 2000])      # equal and similar lengths on purpose
                contigs.append([nm, l])
            recs = []
            for q in range(n):
                recs.append([q, q, [["M", contigs[q][1]]]])                              # bwa mem -a always reports the self hit
                for _ in range(rng.randrange(0, 4)):
                    r = rng.randrange(n)
                    recs.append([q, r, [list(x) for x in random_cigar(rng, contigs[q][1])]])
            rng.shuffle(recs)
            cutoff = rng.choice([0.85, 0.9, 0.95, 0.99, 0.6])
            g = rng.randrange(2)
            fa = os.path.join(td, "c%d.fa" % case)
            bam = os.path.join(td, "c%d.bam" % case)
            out = os.path.join(td, "c%d.out.fa" % case)
            write_fasta(fa, contigs)
            write_bam(bam, [tuple(c) for c in contigs], [(contigs[q][0], r, 0, [tuple(x) for x in cig]) for q, r, cig in recs])
            cmd = [TER, "-P", "-b", bam, "-r", fa, "-o", out, "-c", str(cutoff)] + (["-g"] if g else [])
            subprocess.run(cmd, check=True, capture_output=True)
            rule_cases.append({"contigs": contigs, "records": recs, "cutoff": cutoff, "g": g, "kept": kept_names(out)})
        for case in range(40):
            n = rng.randrange(1, 12)
            pool = ["n%d" % rng.randrange(1, 6) for _ in range(n)]
            contigs = [[nm, rng.randrange(5, 80)] for nm in pool]
            fa = os.path.join(td, "u%d.fa" % case)
            out = os.path.join(td, "u%d.out.fa" % case)
            write_fasta(fa, contigs)
            subprocess.run([TER, "-U", "-r", fa, "-o", out], check=True, capture_output=True)
            kept = kept_names(out) if os.path.exists(out) else [c[0] for c in contigs]   # nothing to remove + same file name: no output
            unique_cases.append({"contigs": contigs, "kept": kept, "kept_lens": None})
            # which of the equally named records survive: lengths tell them apart
            if os.path.exists(out):
                lens, cur = [], 0
                for ln in open(out):
                    if ln.startswith(">"):
                        if cur: lens.append(cur)
                        cur = 0
                    else:
                        cur += len(ln.strip())
                lens.append(cur)
                unique_cases[-1]["kept_lens"] = lens
    json.dump({"rules": rule_cases, "unique": unique_cases}, open(os.path.join(here, "dedup", "dedup_rules.json"), "w"))
    print(len(rule_cases), "rule cases,", len(unique_cases), "unique-name cases")


if __name__ == "__main__":
    main()
