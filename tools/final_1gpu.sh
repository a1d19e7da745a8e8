#!/bin/bash
# Final single-GPU evidence: bench JSON, hot-kernel / elementwise benches, per-op step profile,
# one ncu --set full capture of the top kernel (tap-shared conv forward).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 2>/dev/null | tail -1 > gpurun_out/bench_final.json
cut -c1-200 gpurun_out/bench_final.json
python tools/kernel_bench.py --iters 10 > gpurun_out/kernel_bench_final.log 2>&1; tail -3 gpurun_out/kernel_bench_final.log
python tools/ew_bench.py > gpurun_out/ew_bench_final.log 2>&1; tail -5 gpurun_out/ew_bench_final.log
python tools/profile_step.py > gpurun_out/step_profile_final.txt 2>&1; head -1 gpurun_out/step_profile_final.txt
timeout 120 ncu --set full --clock-control none --import-source on -k regex:conv_gemm_ts_kernel -s 3 -c 1 -f -o gpurun_out/ncu_conv_ts_fwd python tools/kernel_bench.py --only fwd --iters 2 > /dev/null 2>&1
ls -la gpurun_out/ncu_conv_ts_fwd.ncu-rep
