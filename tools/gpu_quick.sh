#!/bin/bash
# Quick GPU check between kernel changes: parity tests of the kernels, one bench line, optionally one ncu capture.
# Usage (here): gpurun --timeout 900 -- 'bash tools/gpu_quick.sh <tag> [ncu]'
tag=${1:-q}
out=gpurun_out
mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> $out/${tag}_pytest.log
tail -3 $out/${tag}_pytest.log
timeout 600 python bench.py --no-cpu --no-dropin > $out/${tag}_bench.json 2> $out/${tag}_bench.err
echo "bench exit $?"; python - <<PY
import json
d=json.loads(open("$out/${tag}_bench.json").read().strip().splitlines()[-1])
print("GCUPS", d["value"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], "checksum", d["result_checksum"], d["kernel_split"])
PY
if [ "$2" = "ncu" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:${NCU_KERNEL:-overlap_wf16c} -s 1 -c 1 \
      -o $out/${tag}_wf16 -f python bench.py --steps 1 --warmup 1 --no-cpu --no-dropin > $out/${tag}_ncu_full.log 2>&1
  tail -2 $out/${tag}_ncu_full.log
fi
