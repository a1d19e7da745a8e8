#!/bin/bash
# One multi-GPU validation pass on an N-GPU box (gpurun --gpus N -- tools/scale_run.sh N):
# fused all-reduce tests (both forms + stress), native DP test, quick all-reduce sweep, train-step
# bench in push and pull form, the stand-in baseline, and a 1-GPU bench on the same box.
N=${1:-8}
out=gpurun_out/scale_r2_n$N
mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 400 python -m pytest tests/test_fused_allreduce.py -m gpu -x -q > $out/pytest_fused_allreduce.log 2>&1; tail -4 $out/pytest_fused_allreduce.log
timeout 500 python -m pytest tests/test_native_dp.py -m gpu -x -q > $out/pytest_native_dp.log 2>&1; tail -4 $out/pytest_native_dp.log
timeout 300 $TR --master-port 29511 tools/allreduce_bench.py ${2:---quick} > $out/allreduce_sweep.log 2>&1; grep -v "^{" $out/allreduce_sweep.log | tail -20
cp gpurun_out/allreduce_sweep_w$N.json $out/ 2>/dev/null
for mode in push pull; do
  timeout 200 $TR --master-port 29512 bench.py --gpus $N --steps 30 --warmup 5 --ar_mode $mode > $out/bench_native_$mode.json 2> $out/bench_native_$mode.err
  tail -1 $out/bench_native_$mode.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$mode', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['collective'], d['clocks'])"
done
timeout 200 $TR --master-port 29513 bench.py --gpus $N --steps 20 --warmup 5 --impl standin > $out/bench_standin.json 2> $out/bench_standin.err
tail -1 $out/bench_standin.json | cut -c1-300
python bench.py --steps 30 --warmup 5 > $out/bench_native_n1_same_box.json 2>/dev/null; cut -c100-330 $out/bench_native_n1_same_box.json
