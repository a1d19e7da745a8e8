#!/bin/bash
# Final evidence captures (1 GPU): ncu --set full of the hot kernels + a per-launch duration list.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python tools/kernel_bench.py > gpurun_out/kernel_bench_final.log 2>&1
python tools/ew_bench.py > gpurun_out/ew_bench_final.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_gemm2_kernel -s 3 -c 1 -f -o gpurun_out/ncu_conv2_fwd python tools/kernel_bench.py --only fwd --iters 2 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_gemm2_kernel -s 3 -c 1 -f -o gpurun_out/ncu_conv2_dgrad python tools/kernel_bench.py --only dgrad --iters 2 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:wgrad_gemm2_kernel -s 3 -c 1 -f -o gpurun_out/ncu_wgrad2 python tools/kernel_bench.py --only wgrad --iters 2 > /dev/null 2>&1
for k in in_apply_kernel in_bwd_reduce_kernel in_bwd_apply_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -f -o gpurun_out/ncu_ew_$k python tools/ew_bench.py --iters 2 > /dev/null 2>&1
done
# every launch of one training step with its device time (no CUDA graph so ncu sees each kernel)
B2_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum --clock-control none -s 1300 -c 650 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no_e2e > gpurun_out/launches_bench.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/launches.csv
cat gpurun_out/kernel_bench_final.log gpurun_out/ew_bench_final.log
