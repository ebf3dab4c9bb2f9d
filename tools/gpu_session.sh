#!/bin/bash
# tools/gpu_session.sh TAG [what...] -- one gpurun call's worth of work on the GPU box; everything lands in gpurun_out/TAG_*.
# what: tests bench ref cfg3 cfg5 ncu_pair ncu_relax launches dropin qc affine ncu_affine ...   (default: tests bench)
TAG=$1; shift
WHAT="${*:-tests bench}"
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_gpu.txt 2>&1
for w in $WHAT; do
  case $w in
    tests)   timeout 1500 python -m pytest tests -m gpu -q > $O/${TAG}_pytest.log 2>&1; tail -5 $O/${TAG}_pytest.log ;;
    bench)   timeout 900 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; tail -c 600 $O/${TAG}_bench.json ;;
    benchq)  timeout 600 python bench.py --no-dropin --no-cpu > $O/${TAG}_benchq.json 2> $O/${TAG}_benchq.err; tail -c 400 $O/${TAG}_benchq.json ;;
    ref)     timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_bench_ref.json 2>&1 ;;
    cfg3)    timeout 600 python bench.py --config cfg3 --no-dropin > $O/${TAG}_bench_cfg3.json 2> $O/${TAG}_bench_cfg3.err ;;
    cfg5)    timeout 900 python bench.py --config cfg5 --steps 3 --no-dropin > $O/${TAG}_bench_cfg5.json 2> $O/${TAG}_bench_cfg5.err ;;
    cfg2)    timeout 600 python bench.py --config cfg2 --no-dropin > $O/${TAG}_bench_cfg2.json 2> $O/${TAG}_bench_cfg2.err; tail -c 400 $O/${TAG}_bench_cfg2.json ;;
    dropin)  timeout 600 python tools/dropin_bench.py --gaps 200 --ref-gaps 1 --repeat 3 > $O/${TAG}_dropin.json 2> $O/${TAG}_dropin.err; cat $O/${TAG}_dropin.json ;;
    launches) timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-dropin > $O/${TAG}_ncu_launches.log 2>&1 ;;
    ncu_pair) timeout 1200 ncu --set full --clock-control none --import-source on -k regex:overlap_wf16c -s 1 -c 1 -o $O/${TAG}_wf16c -f python bench.py --steps 1 --warmup 1 --no-cpu --no-dropin > $O/${TAG}_ncu_full.log 2>&1 ;;
    ncu_relax) mkdir -p /tmp/rl && python - <<'PY'
import sys, os
sys.path.insert(0, 'tools')
import synth_gaps
with open('/tmp/rl/list.tsv', 'w') as f:
    for g in range(200):
        fa = '/tmp/rl/g%d.fa' % g
        synth_gaps.write_fasta(fa, synth_gaps.make_gap(1 + g, synth_gaps.CONFIGS['cfg1']))
        f.write('%s\t/tmp/rl/g%d.out\t/tmp/rl/g%d.info\n' % (fa, g, g))
PY
             timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"relax_chain" -c 1 -o $O/${TAG}_relax -f build/ContigsMerger_b200 -s 0.4 -i1 -2.0 -i2 -2.0 -x 12 -y 50 -k 10 -m 1 -t 5 --batch /tmp/rl/list.tsv --no-gml > $O/${TAG}_ncu_relax.log 2>&1 ;;
    ncu_flank) timeout 900 ncu --set full --clock-control none --import-source on -k regex:flank_place -s 1 -c 1 -o $O/${TAG}_flank -f python bench.py --config cfg2 --steps 1 --warmup 1 --no-cpu > $O/${TAG}_ncu_flank.log 2>&1 ;;
    ncu_qc)  timeout 900 ncu --set full --clock-control none --import-source on -k regex:quick_check -s 1 -c 1 -o $O/${TAG}_qc -f python tools/quickcheck_bench.py --reps 2 > $O/${TAG}_ncu_qc.log 2>&1 ;;
    streams) for sN in 1 2 3; do timeout 300 python tools/dropin_bench.py --gaps 200 --ref-gaps 0 --repeat 3 --streams $sN > $O/${TAG}_dropin_s$sN.json 2>/dev/null; python -c "import json; d=json.loads(open('$O/${TAG}_dropin_s$sN.json').read().strip().splitlines()[-1]); print('streams $sN', d['merge_ms'], d['gaps_per_s'], d.get('detail_ms'))"; done ;;
    pipe)    for cg in 200 100 67 50; do for sN in 2 3; do timeout 300 python tools/dropin_bench.py --gaps 200 --ref-gaps 1 --repeat 3 --streams $sN --chunk-gaps $cg > $O/${TAG}_pipe_c${cg}_s$sN.json 2>/dev/null; python -c "import json; d=json.loads(open('$O/${TAG}_pipe_c${cg}_s$sN.json').read().strip().splitlines()[-1]); print('cfg1 x200 chunk $cg streams $sN', d['merge_ms'], round(d['gaps_per_s']), d['reference']['outputs_identical'], d.get('detail_ms'))"; done; done
             for cg in 400 200 100 64; do for sN in 2 3; do timeout 300 python tools/dropin_bench.py --config cfg3 --seed 5000 --gaps 1600 --ref-gaps 0 --repeat 2 --streams $sN --chunk-gaps $cg > $O/${TAG}_pipe3_c${cg}_s$sN.json 2>/dev/null; python -c "import json; d=json.loads(open('$O/${TAG}_pipe3_c${cg}_s$sN.json').read().strip().splitlines()[-1]); print('cfg3 x1600 chunk $cg streams $sN', d['merge_ms'], round(d['gaps_per_s']), d.get('detail_ms'))"; done; done ;;
    dedup)   timeout 900 python -m pytest tests/test_gpu_dedup.py -q > $O/${TAG}_pytest_dedup.log 2>&1; tail -5 $O/${TAG}_pytest_dedup.log
             timeout 600 python tools/dedup_bench.py > $O/${TAG}_dedup.json 2> $O/${TAG}_dedup.err; cat $O/${TAG}_dedup.json ;;
    cfg4)    timeout 1500 python tools/cfg4_bench.py --gpus ${CFG4_GPUS:-8,4,2,1} --gaps ${CFG4_GAPS:-50000} > $O/${TAG}_cfg4.json 2> $O/${TAG}_cfg4.err; cut -c1-3000 $O/${TAG}_cfg4.json; tail -3 $O/${TAG}_cfg4.err ;;
    refpool) timeout 900 python tools/dropin_bench.py --gaps 200 --ref-gaps 0 --ref-procs 16 --repeat 2 > $O/${TAG}_dropin_refpool.json 2> $O/${TAG}_dropin_refpool.err; cut -c1-2500 $O/${TAG}_dropin_refpool.json ;;
    rtrace)  timeout 600 python tools/relax_trace.py > $O/${TAG}_relax_trace.json 2> $O/${TAG}_relax_trace.err; cut -c1-600 $O/${TAG}_relax_trace.json ;;
    lone)    timeout 600 python tools/lone_pair_bench.py > $O/${TAG}_lone_pair.json 2> $O/${TAG}_lone_pair.err; cat $O/${TAG}_lone_pair.err ;;
    lonetrace) GAPPADDER_B200_LIB=build/libgappadder_b200_trace.so timeout 600 python tools/lone_pair_bench.py --trace > /dev/null 2> $O/${TAG}_lone_trace.txt; head -150 $O/${TAG}_lone_trace.txt ;;
    ppg)     timeout 600 python tools/process_per_gap_bench.py > $O/${TAG}_process_per_gap.json 2> $O/${TAG}_process_per_gap.err; cat $O/${TAG}_process_per_gap.json ;;
    qc)      timeout 600 python tools/quickcheck_bench.py > $O/${TAG}_quickcheck.json 2> $O/${TAG}_quickcheck.err; cat $O/${TAG}_quickcheck.json ;;
    affine)  timeout 300 python -m pytest tests/test_gpu_affine.py tests/test_gpu_terefiner.py -m gpu -q > $O/${TAG}_pytest_affine.log 2>&1; tail -3 $O/${TAG}_pytest_affine.log
             timeout 120 python tools/affine_bench.py --gaps 20 > $O/${TAG}_affine_kernels.json 2> $O/${TAG}_affine_kernels.err; cut -c1-400 $O/${TAG}_affine_kernels.json
             timeout 300 python bench.py --config affine > $O/${TAG}_bench_affine.json 2> $O/${TAG}_bench_affine.err; tail -c 500 $O/${TAG}_bench_affine.json ;;
    ncu_affine) timeout 300 ncu --set full --clock-control none --import-source on -k regex:affine_forward -c 1 -o $O/${TAG}_affine_forward -f python tools/affine_bench.py --gaps 20 --reps 1 --check 0 > $O/${TAG}_ncu_affine_forward.log 2>&1
             timeout 600 ncu --set full --clock-control none --import-source on -k regex:affine_epilogue -c 1 -o $O/${TAG}_affine_epilogue -f python tools/affine_bench.py --gaps 6 --reps 1 --check 0 > $O/${TAG}_ncu_affine_epilogue.log 2>&1 ;;
    *) echo "unknown: $w" ;;
  esac
done
