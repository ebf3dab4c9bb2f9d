#!/bin/bash
# ncu capture of the affine start-recovery kernel on a reduced cfg1 batch (sections only: a full set replays too often for a kernel
# that rewrites a few hundred MB of scratch per pass)
O=gpurun_out
mkdir -p $O
timeout 75 ncu --section SpeedOfLight --section SchedulerStats --section WarpStateStats --section InstructionStats --section LaunchStats --section Occupancy --section MemoryWorkloadAnalysis \
  --clock-control none -k regex:affine_epilogue -c 1 -o $O/r05g_affine_epilogue -f python tools/affine_bench.py --gaps 6 --reps 1 --check 0 > $O/r05g_ncu_epilogue.log 2>&1
tail -3 $O/r05g_ncu_epilogue.log
