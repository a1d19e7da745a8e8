#!/bin/bash
# Single-GPU evidence for the MXFP8 path (BASELINE config 3) and the re-tuned conv kernels:
#   gpurun -- tools/evidence_mx.sh
# kernel micro-bench, stall attribution + time line of the MXFP8 pair kernel, ncu --set full of it,
# batch-8 benches bf16 / fp8 / mxfp8 on the SAME box, the default bench, the GPU test suite.
cd "$(dirname "$0")/.."
out=gpurun_out/evidence_mx
mkdir -p $out
timeout 600 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1; tail -2 $out/pytest_gpu.log
python tools/kernel_bench.py --iters 20 > $out/kernel_bench.log 2>&1; tail -6 $out/kernel_bench.log
python tools/kernel_bench.py --only mxfp8_ts --mx_profile > $out/mx_timeline.txt 2>&1
bash tools/mx_probe.sh > /dev/null 2>&1; cp gpurun_out/mx_probe.log $out/
python bench.py --steps 30 --warmup 5 2>/dev/null | tail -1 > $out/bench_native_n1.json; cut -c100-330 $out/bench_native_n1.json
for d in bf16 fp8 mxfp8; do
  python bench.py --steps 20 --warmup 5 --batch_size 8 --dtype $d 2>/dev/null | tail -1 > $out/bench_b8_$d.json; cut -c100-330 $out/bench_b8_$d.json
done
timeout 150 ncu --set full --clock-control none --import-source on -k regex:conv_gemm_mx_ts_kernel -s 3 -c 1 -f \
  -o $out/ncu_conv_mx_ts python tools/kernel_bench.py --only mxfp8_ts --iters 2 > /dev/null 2>&1
timeout 150 ncu --set full --clock-control none --import-source on -k regex:in_apply_stream_kernel -s 2 -c 1 -f \
  -o $out/ncu_in_apply_mx python -m pytest tests/test_native_kernels.py -q -k "mxfp8_conv and True" > /dev/null 2>&1
ls -la $out
