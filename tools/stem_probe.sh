#!/bin/bash
# Times the 7x7 stem forward conv with the kernel's timing switches (0 full, 1 no epilogue,
# 2 no MMAs, 3 no TMA loads, 4 no global stores) to show which pipe bounds the N = 64 kernel.
mkdir -p gpurun_out
for d in 0 1 2 3 4; do
  echo "== kernel_debug=$d" >> gpurun_out/stem_probe.log
  B2_CONFIG=kernel_debug=$d timeout 120 python tools/kernel_bench.py --layer stem --images 8 >> gpurun_out/stem_probe.log 2>&1
done
tail -40 gpurun_out/stem_probe.log
