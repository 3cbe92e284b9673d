#!/bin/bash
# Profiling recipe (run under gpurun, ONE GPU): launch list + one full ncu capture per hot kernel.
# Outputs land in gpurun_out/; summaries worth keeping are copied to profiles/ by hand.
set -u
mkdir -p gpurun_out
TAG=${1:-r01}
BENCH="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
# (1) every launch with its device time
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv $BENCH > gpurun_out/launches_${TAG}.log 2>&1
# (2) full sets of the hot kernels (second occurrence = the timed step)
for K in wb_shade_bwd_tc_kernel wb_shade_fwd_tc_kernel wb_march_count_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -o gpurun_out/prof_${K}_${TAG} -f $BENCH > gpurun_out/prof_${K}_${TAG}.log 2>&1
done
ls -la gpurun_out | tail -12
