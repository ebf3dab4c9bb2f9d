#!/bin/bash
# tools/gpu_final.sh TAG -- the round's closing GPU session: every GPU test, the affine bench line, smoke(), ncu of the affine kernels.
TAG=${1:-final}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_gpu.txt 2>&1
timeout 170 python -m pytest tests -m gpu -q > $O/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $O/${TAG}_pytest.log; tail -4 $O/${TAG}_pytest.log
timeout 60 python bench.py --config affine > $O/${TAG}_bench_affine.json 2> $O/${TAG}_bench_affine.err; echo "bench exit $?"; tail -c 700 $O/${TAG}_bench_affine.json
timeout 30 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; tail -2 $O/${TAG}_smoke.log
timeout 45 ncu --set full --clock-control none --import-source on -k regex:affine -c 2 -o $O/${TAG}_affine -f python tools/affine_bench.py --gaps 20 --reps 1 --check 0 > $O/${TAG}_ncu_affine.log 2>&1; tail -2 $O/${TAG}_ncu_affine.log
