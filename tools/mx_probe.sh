#!/bin/bash
# Times the MXFP8 conv kernels with the timing switches (0 full, 1 no epilogue, 2 no MMAs, 5 no scale
# stores to TMEM) to show which pipe bounds them.
mkdir -p gpurun_out; : > gpurun_out/mx_probe.log
for d in 0 1 2 5; do
  for k in mxfp8 mxfp8_ts; do
    echo "== kernel_debug=$d $k" >> gpurun_out/mx_probe.log
    B2_CONFIG=kernel_debug=$d timeout 120 python tools/kernel_bench.py --only $k 2>&1 | tail -1 >> gpurun_out/mx_probe.log
  done
done
cat gpurun_out/mx_probe.log
