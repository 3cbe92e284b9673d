#!/bin/bash
# Profiling recipe (run under gpurun, ONE GPU): launch list + one full ncu capture per hot kernel.
#   bash tools/profile_gpu.sh TAG [kernel regex ...]
# Outputs land in gpurun_out/ as CSV raw pages (the .ncu-rep files are deleted on the box: gpurun_out/ is capped at 64 MiB);
# `python tools/ncu_summary.py TAG` turns them into profiles/TAG_*.txt.
set -u
mkdir -p gpurun_out
TAG=${1:-r02}
shift
KERNELS=${@:-wb_mlp_bwd3_tc_kernel wb_shade_fwd_tc_kernel wb_march_count_kernel wb_composite_bwd_kernel}
BENCH="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv $BENCH > gpurun_out/launches_${TAG}.log 2>&1
for K in $KERNELS; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -o /tmp/prof_${K}_${TAG} -f $BENCH > gpurun_out/prof_${K}_${TAG}.log 2>&1
  ncu -i /tmp/prof_${K}_${TAG}.ncu-rep --page raw --csv > gpurun_out/prof_${K}_${TAG}.csv 2>/dev/null
  rm -f /tmp/prof_${K}_${TAG}.ncu-rep
done
# config 3: the persistent sphere tracer
B3="python bench.py --config 3 --steps 1 --warmup 1 --no-cpu-baseline"
ncu --set full --clock-control none --import-source on -k regex:wb_sdf_trace_kernel -s 1 -c 1 -o /tmp/prof_sdf_${TAG} -f $B3 > gpurun_out/prof_wb_sdf_trace_kernel_${TAG}.log 2>&1
ncu -i /tmp/prof_sdf_${TAG}.ncu-rep --page raw --csv > gpurun_out/prof_wb_sdf_trace_kernel_${TAG}.csv 2>/dev/null
rm -f /tmp/prof_sdf_${TAG}.ncu-rep
# hidden_dim = 128: the one-group decoder backward (same kernel template, NG = 1) and its forward
BH="python bench.py --hidden-dim 128 --steps 1 --warmup 1 --no-cpu-baseline"
for K in wb_mlp_bwd3_tc_kernel wb_shade_fwd_tc_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -o /tmp/prof_${K}_h128_${TAG} -f $BH > gpurun_out/prof_${K}_h128_${TAG}.log 2>&1
  ncu -i /tmp/prof_${K}_h128_${TAG}.ncu-rep --page raw --csv > gpurun_out/prof_${K}_h128_${TAG}.csv 2>/dev/null
  rm -f /tmp/prof_${K}_h128_${TAG}.ncu-rep
done
# config 4: the triplanar gather (channel-last planes) inside the tensor-core forward, and its plane scatter
B4="python bench.py --config 4 --steps 1 --warmup 1 --no-cpu-baseline"
for K in wb_shade_fwd_tc_kernel wb_featx_scatter_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -o /tmp/prof_${K}_cfg4_${TAG} -f $B4 > gpurun_out/prof_${K}_cfg4_${TAG}.log 2>&1
  ncu -i /tmp/prof_${K}_cfg4_${TAG}.ncu-rep --page raw --csv > gpurun_out/prof_${K}_cfg4_${TAG}.csv 2>/dev/null
  rm -f /tmp/prof_${K}_cfg4_${TAG}.ncu-rep
done
rm -f gpurun_out/prof_*_${TAG}.log
du -sh gpurun_out; ls -la gpurun_out | tail -14
