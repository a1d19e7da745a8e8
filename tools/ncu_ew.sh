#!/bin/bash
# ncu captures of the elementwise kernels (one launch each) -> gpurun_out/ncu_ew_*.ncu-rep
cd "$(dirname "$0")/.."
for k in in_apply_kernel in_bwd_reduce_kernel in_bwd_apply_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -f -o gpurun_out/ncu_ew_$k python tools/ew_bench.py --iters 2 > /dev/null 2>&1
done
ls -la gpurun_out/ncu_ew_*.ncu-rep
