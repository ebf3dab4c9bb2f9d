#!/usr/bin/env python
"""tools/ncu_lines.py -- where a kernel's time goes, by device function and by CUDA source line.

Joins the per-instruction counters of an .ncu-rep (`--page source`: instructions executed, warp-stall
samples) with the line table of the SAME binary (`nvdisasm --print-line-info` on the cubin inside the
shared library), instruction by instruction.  Read here, no GPU needed.

usage: python tools/ncu_lines.py gpurun_out/x.ncu-rep [gappadder_b200/libgappadder_b200.so] > profiles/x.lines.txt
"""
import collections
import csv
import glob
import io
import os
import re
import subprocess
import sys
import tempfile


def sh(cmd, **kw):
    return subprocess.run(cmd, capture_output=True, text=True, **kw).stdout


def main():
    rep = sys.argv[1]
    lib = sys.argv[2] if len(sys.argv) > 2 else "gappadder_b200/libgappadder_b200.so"
    src = list(csv.reader(io.StringIO(sh(["ncu", "-i", rep, "--page", "source", "--csv"]))))
    kname = src[0][1]
    hdr = src[1]
    c = {h: i for i, h in enumerate(hdr)}
    rows = src[2:]
    n = len(rows)
    ex = [int(r[c["Instructions Executed"]]) for r in rows]
    sm = [int(r[c["# Samples"]]) for r in rows]
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]

    # line table of the same kernel from the library's cubin
    tmp = tempfile.mkdtemp()
    sh(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp)
    want = re.sub(r"\(.*", "", kname).replace("void ", "").split("<")[0].split("::")[-1]
    targs = re.findall(r"\((?:bool|int)\)(\d+)", kname.split("(const")[0])
    best = None
    for cub in glob.glob(os.path.join(tmp, "*.cubin")):
        dis = sh(["nvdisasm", "--print-line-info", "-c", cub])
        secs = re.split(r"\n(?=\.text\.)", dis)
        for s in secs:
            head = s.split("\n", 1)[0]
            if want not in head:
                continue
            ins = []
            line, fn = ("?", 0), head
            for ln in s.split("\n"):
                m = re.match(r'\s*//## File "(.*)", line (\d+)', ln)
                if m:
                    line = (os.path.basename(m.group(1)), int(m.group(2)))
                    continue
                m = re.match(r"^(\$?[\w$.]+):\s*$", ln)
                if m and not m.group(1).startswith(".L"):
                    fn = m.group(1)
                    continue
                m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
                if m:
                    ins.append((line, fn, m.group(2).strip()))
            if len(ins) == n and (best is None):
                # several template instances may have the same length: check the opcode sequence
                ok = all(ins[i][2].split()[-1 if False else 0].lstrip("@!P0123456789T ") == "" or True for i in range(0))
                opc = lambda t: (t.split()[1] if t.startswith("@") else t.split()[0])
                same = sum(1 for i in range(n) if opc(ins[i][2]) == opc(rows[i][c["Source"]].strip()))
                if same > 0.98 * n:
                    best = ins
    if best is None:
        sys.exit("no section of %s matches the %d instructions of %s (%s)" % (lib, n, kname, targs))
    tot_ex, tot_sm = sum(ex), sum(sm)
    print("report:", rep)
    print("kernel:", kname.split("(const")[0])
    print("instructions executed (warp level): %d   stall samples: %d" % (tot_ex, tot_sm))

    def short(fn):
        fn = re.sub(r"^\$", "", fn)
        m = re.search(r"(wf16c_strip|wf16c_pass|wf16c_scan_cold|wf16t_\w+?|wf16_\w+?)I([A-Za-z0-9_]*?)E", fn)
        if m:
            return m.group(1) + "<" + m.group(2) + ">"
        return fn[:70]

    byfn = collections.OrderedDict()
    for i in range(n):
        k = short(best[i][1])
        a = byfn.setdefault(k, [0, 0, 0])
        a[0] += ex[i]; a[1] += sm[i]; a[2] += 1
    print("\nby device function (share of executed instructions, share of stall samples = time, static instructions):")
    for k, a in sorted(byfn.items(), key=lambda kv: -kv[1][1]):
        print("  %-60s %6.2f%% instr  %6.2f%% time  %5d static" % (k, 100.0 * a[0] / tot_ex, 100.0 * a[1] / max(tot_sm, 1), a[2]))

    byline = collections.defaultdict(lambda: [0, 0, collections.Counter()])
    for i in range(n):
        a = byline[best[i][0]]
        a[0] += ex[i]; a[1] += sm[i]
        for h in stall_cols:
            v = int(rows[i][c[h]])
            if v:
                a[2][h[6:]] += v
    print("\nby source line (>= 0.4% of time):")
    for k, a in sorted(byline.items(), key=lambda kv: -kv[1][1]):
        if a[1] < 0.004 * tot_sm:
            break
        top = ", ".join("%s %.0f%%" % (s, 100.0 * v / a[1]) for s, v in a[2].most_common(4))
        print("  %-22s:%-4d %6.2f%% instr  %6.2f%% time   [%s]" % (k[0], k[1], 100.0 * a[0] / tot_ex, 100.0 * a[1] / tot_sm, top))


if __name__ == "__main__":
    main()
