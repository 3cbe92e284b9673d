#!/bin/bash
# Profiling recipe (run under gpurun, ONE GPU): launch list + one full ncu capture per hot kernel.
# Outputs land in gpurun_out/; `python tools/ncu_summary.py TAG` turns them into profiles/TAG_*.txt.
set -u
mkdir -p gpurun_out
TAG=${1:-r01}
shift
KERNELS=${@:-wb_mlp_bwd_tc_kernel wb_table_scatter_kernel wb_shade_fwd_tc_kernel wb_march_count_kernel}
BENCH="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv $BENCH > gpurun_out/launches_${TAG}.log 2>&1
for K in $KERNELS; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -o gpurun_out/prof_${K}_${TAG} -f $BENCH > gpurun_out/prof_${K}_${TAG}.log 2>&1
done
ls -la gpurun_out | tail -12
