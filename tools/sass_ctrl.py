#!/usr/bin/env python
"""tools/sass_ctrl.py -- SASS of one kernel with the scheduling control fields decoded (stall count, yield,
write/read scoreboard, wait mask) and the source line of every instruction.  Runs here (nvdisasm on the cubin
inside the shared library); it is how scoreboard aliasing in the steady loops is found without a GPU.

usage: python tools/sass_ctrl.py KERNEL_SUBSTRING [first last] [lib.so]      (instruction index range)
       python tools/sass_ctrl.py KERNEL_SUBSTRING loops                      (densest VIADDMNMX regions)"""
import glob, os, re, subprocess, sys, tempfile


def load(lib, want):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
    for cub in glob.glob(os.path.join(tmp, "*.cubin")):
        txt = subprocess.run(["nvdisasm", "-hex", "--print-line-info", "-c", cub], capture_output=True, text=True).stdout
        for sec in re.split(r"\n(?=\.text\.)", txt):
            if not sec.startswith(".text.") or want not in sec.split("\n", 1)[0]:
                continue
            lines = sec.split("\n")
            ins, line, i = [], ("?", 0), 0
            while i < len(lines):
                m = re.match(r'\s*//## File "(.*)", line (\d+)', lines[i])
                if m:
                    line = (os.path.basename(m.group(1)), int(m.group(2)))
                m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);\s*/\* (0x[0-9a-f]+) \*/", lines[i])
                if m:
                    hi = int(re.match(r"\s*/\* (0x[0-9a-f]+) \*/", lines[i + 1]).group(1), 16)
                    ins.append((m.group(2).strip(), hi, line))
                    i += 1
                i += 1
            return ins
    sys.exit("kernel not found")


def fmt(k, t, hi, line):
    stall, y, wb, rb, wm = (hi >> 41) & 0xf, (hi >> 45) & 1, (hi >> 46) & 7, (hi >> 49) & 7, (hi >> 52) & 0x3f
    return "%5d %-60s st=%2d %s wr=%s rd=%s wait=%-4s %s:%d" % (
        k, t[:60], stall, "Y" if y else " ", wb if wb != 7 else "-", rb if rb != 7 else "-",
        "".join(str(b) for b in range(6) if wm >> b & 1) or "-", line[0], line[1])


def main():
    want = sys.argv[1]
    lib = sys.argv[-1] if sys.argv[-1].endswith(".so") else "gappadder_b200/libgappadder_b200.so"
    ins = load(lib, want)
    print(len(ins), "instructions")
    if len(sys.argv) > 2 and sys.argv[2] == "loops":
        idx = [i for i, x in enumerate(ins) if "VIADDMNMX" in x[0]]
        seen = []
        for a in range(len(idx) - 31):
            if idx[a + 31] - idx[a] < 90 and all(abs(idx[a] - s) > 100 for s in seen):
                seen.append(idx[a])
        for s in seen:
            print("---- region at", s)
            for k in range(max(s - 16, 0), min(s + 84, len(ins))):
                print(fmt(k, *ins[k]))
        return
    a, b = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (0, len(ins) - 1)
    for k in range(a, b + 1):
        print(fmt(k, *ins[k]))


if __name__ == "__main__":
    main()
