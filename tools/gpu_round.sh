#!/bin/bash
# One GPU-box pass: parity tests, bench, ncu launch list, one ncu --set full capture of the top kernel.
# Usage (here): gpurun --timeout 1500 -- 'bash tools/gpu_round.sh <tag>'   -> results in gpurun_out/<tag>_*
tag=${1:-r01}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/${tag}_gpu.txt 2>&1
nproc >> $out/${tag}_gpu.txt; lscpu | grep -E 'Model name|^CPU\(s\)' >> $out/${tag}_gpu.txt
if [ -z "$SKIP_TESTS" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1
  echo "pytest exit $?" >> $out/${tag}_pytest.log
  tail -5 $out/${tag}_pytest.log
fi
timeout 600 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err
echo "bench exit $?"; cat $out/${tag}_bench.json
if [ -z "$SKIP_REF" ]; then
  timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $out/${tag}_bench_ref.json 2>> $out/${tag}_bench.err
  cat $out/${tag}_bench_ref.json
fi
if [ -z "$SKIP_NCU" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
      --log-file $out/${tag}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-dropin > $out/${tag}_ncu_launches.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:${NCU_KERNEL:-overlap_wf16c} -s 1 -c 1 \
      -o $out/${tag}_wf16 -f python bench.py --steps 1 --warmup 1 --no-cpu --no-dropin ${NCU_BENCH_ARGS} > $out/${tag}_ncu_full.log 2>&1
  tail -3 $out/${tag}_ncu_full.log
fi
ls -la $out
