#!/bin/bash
# Single-GPU evidence pass (gpurun -- tools/evidence.sh [tag]): everything the judge reads under
# profiles/ is regenerated from THIS tree in one call.
#   bench JSON (device-timed + e2e), hot GEMM / InstanceNorm micro-benches (L2 flushed and warm),
#   per-op step profile, the launch list of one eager step, ncu sections for EVERY kernel of the step
#   plan (one eager step), and ncu --set full + source for the top kernels.
cd "$(dirname "$0")/.."
tag=${1:-r2}
out=gpurun_out/evidence_$tag
mkdir -p $out
python bench.py --steps 30 --warmup 5 2>/dev/null | tail -1 > $out/bench_native_n1.json; cut -c100-330 $out/bench_native_n1.json
python bench.py --impl standin --steps 20 --warmup 5 2>/dev/null | tail -1 > $out/bench_standin_n1.json; cut -c100-330 $out/bench_standin_n1.json
# BASELINE configs 3 / 4 on one GPU (batch 8 bf16 vs fp8; 512 x 512 batch 1)
python bench.py --steps 20 --warmup 5 --batch_size 8 --no_e2e 2>/dev/null | tail -1 > $out/bench_b8_bf16.json; cut -c100-330 $out/bench_b8_bf16.json
python bench.py --steps 20 --warmup 5 --batch_size 8 --dtype fp8 --no_e2e 2>/dev/null | tail -1 > $out/bench_b8_fp8.json; cut -c100-330 $out/bench_b8_fp8.json
python bench.py --steps 20 --warmup 5 --batch_size 1 --image_size 512 --no_e2e 2>/dev/null | tail -1 > $out/bench_512_b1.json; cut -c100-330 $out/bench_512_b1.json
python tools/kernel_bench.py --iters 20 > $out/kernel_bench.log 2>&1; tail -3 $out/kernel_bench.log
python tools/ew_bench.py > $out/ew_bench_flush.log 2>&1; tail -5 $out/ew_bench_flush.log
python tools/ew_bench.py --no_flush > $out/ew_bench_warm.log 2>&1; tail -5 $out/ew_bench_warm.log
python tools/profile_step.py --out $out/step_profile.txt > /dev/null 2>&1; head -3 $out/step_profile.txt
# every launch of one training step with its device time (eager plan so ncu sees each kernel)
B2_CONFIG=use_graph=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 520 --csv \
  --log-file $out/launches_one_step.csv python bench.py --steps 2 --warmup 3 --no_e2e > $out/launches_bench.log 2>&1
# the recipe's metrics for EVERY kernel launch of one eager step (summarised per kernel by tools/ncu_kernels.py)
B2_CONFIG=use_graph=0 timeout 600 ncu --clock-control none -s 1500 -c 520 --csv --page raw \
  --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy \
  --metrics sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum \
  --log-file $out/ncu_all_kernels_raw.csv python bench.py --steps 2 --warmup 3 --no_e2e > $out/ncu_all_bench.log 2>&1
python tools/ncu_kernels.py $out/ncu_all_kernels_raw.csv > $out/ncu_all_kernels_summary.txt 2>&1; head -5 $out/ncu_all_kernels_summary.txt
# full sets (with source) of the top kernels
timeout 150 ncu --set full --clock-control none --import-source on -k regex:conv_gemm_ts_kernel -s 3 -c 1 -f -o $out/ncu_conv_ts_fwd python tools/kernel_bench.py --only fwd --iters 2 > /dev/null 2>&1
timeout 150 ncu --set full --clock-control none --import-source on -k regex:conv_gemm_ts_kernel -s 3 -c 1 -f -o $out/ncu_conv_ts_dgrad python tools/kernel_bench.py --only dgrad --iters 2 > /dev/null 2>&1
timeout 150 ncu --set full --clock-control none --import-source on -k regex:wgrad_gemm2_kernel -s 3 -c 1 -f -o $out/ncu_wgrad2 python tools/kernel_bench.py --only wgrad --iters 2 > /dev/null 2>&1
for k in in_apply_stream_kernel in_bwd_reduce_stream_kernel in_bwd_apply_stream_kernel; do
  timeout 150 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -f -o $out/ncu_$k python tools/ew_bench.py --iters 2 > /dev/null 2>&1
done
# sanitizers: memcheck over the whole (eager) smoke step, racecheck + synccheck over the kernel tests
SANITIZE_TIMEOUT=${SANITIZE_TIMEOUT:-300} bash tools/sanitize.sh memcheck racecheck synccheck 2>&1 | tail -4
cp gpurun_out/sanitizer_*.log $out/ 2>/dev/null
ls -la $out
