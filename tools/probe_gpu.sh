#!/bin/bash
# Run every GPU test in its own process (a trapped kernel poisons the CUDA context),
# collect PASS/FAIL per test into gpurun_out/probe.log.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
LOG=gpurun_out/probe.log
: > $LOG
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv >> $LOG 2>&1
python -c "import torch; print(torch.__version__, torch.cuda.get_device_name(0))" >> $LOG 2>&1
FILES="${@:-tests/test_native_kernels.py}"
IDS=$(python -m pytest $FILES --collect-only -q -m gpu 2>/dev/null | grep "::")
for t in $IDS; do
  echo "=== $t" >> $LOG
  timeout 300 python -m pytest -x -q "$t" -m gpu > gpurun_out/_t.log 2>&1
  rc=$?
  if [ $rc -eq 0 ]; then echo "PASS $t" >> $LOG; else echo "FAIL($rc) $t" >> $LOG; grep -E "Error|error|assert|watchdog|rel_err|^E " gpurun_out/_t.log | head -30 >> $LOG; fi
done
grep -E "^(PASS|FAIL)" $LOG
