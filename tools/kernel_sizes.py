#!/usr/bin/env python
"""tools/kernel_sizes.py -- SASS bytes per kernel of libgappadder_b200.so (the instruction cache cliff: DESIGN.md section 4)."""
import re
import subprocess
import sys
so = sys.argv[1] if len(sys.argv) > 1 else "gappadder_b200/libgappadder_b200.so"
out = subprocess.run(["cuobjdump", "-elf", so], capture_output=True, text=True).stdout
rows = []
for ln in out.splitlines():
    m = re.match(r"\s*\d+\s+[0-9a-f]+\s+([0-9a-f]+)\s+.*PROGBITS.*\.text\.(\S+)", ln)
    if m:
        rows.append((int(m.group(1), 16), m.group(2)))
for sz, name in sorted(rows, reverse=True):
    d = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    print("%8d  %s" % (sz, d[:110]))
