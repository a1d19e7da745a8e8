#!/bin/bash
# 1/2/4/8-GPU scaling + fused all-reduce validation + sweep, in one multi-GPU gpurun call.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NG=$(python -c "import torch; print(torch.cuda.device_count())")
echo "gpus=$NG"
timeout 600 python -m pytest tests/test_fused_allreduce.py tests/test_native_dp.py -m gpu -x -q -s > gpurun_out/pytest_ar_w$NG.log 2>&1; tail -4 gpurun_out/pytest_ar_w$NG.log
for n in ${SCALE_NS:-1 2 4 8}; do
  [ $n -le $NG ] || continue
  if [ $n -eq 1 ]; then
    timeout 300 python bench.py --gpus 1 --steps 15 --warmup 4 > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --steps 15 --warmup 4 > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  fi
  grep '"metric"' gpurun_out/scale_n$n.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('N=%d %.1f img/s %.2f ms e2e %.1f' % (d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value']))" || tail -3 gpurun_out/scale_n$n.err
done
# NCCL baseline path for comparison at full width
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29650 bench.py --gpus $NG --steps 15 --warmup 4 --fused_allreduce off --no_e2e > gpurun_out/scale_n${NG}_nccl.json 2> gpurun_out/scale_nccl.err
grep '"metric"' gpurun_out/scale_n${NG}_nccl.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('NCCL path N=%d %.1f img/s %.2f ms' % (d['n_gpus'], d['value'], d['ms_per_step']))" || tail -3 gpurun_out/scale_nccl.err
for w in ${SWEEP_WS:-2 8}; do
  [ $w -le $NG ] || continue
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $w --master-addr 127.0.0.1 --master-port $((29700+w)) tools/allreduce_bench.py > gpurun_out/sweep_w$w.log 2>&1
  python - <<PY
import json
try:
  d=json.load(open("gpurun_out/allreduce_sweep_w$w.json"))
  print("sweep world", d["world"], "multicast", d["multicast"])
  for r in d["rows"]:
    print("  %12d B fused %8.3f ms (%6.1f GB/s bus, %4.1f%% of 900)  nccl %8.3f ms (%6.1f)  nccl+adam %8.3f ms" % (r["bytes"], r["fused_ms"], r["fused_bus_GBs"], 100*r["fused_frac_of_900"], r["nccl_allreduce_ms"], r["nccl_bus_GBs"], r["nccl_plus_adam_ms"]))
except Exception as e:
  print("sweep failed", e); print(open("gpurun_out/sweep_w$w.log").read()[-1500:])
PY
done
