#!/bin/bash
# One multi-GPU validation pass on an N-GPU box:  gpurun --gpus N -- tools/scale_run.sh N [full]
#   fused all-reduce test (both forms + 200-replay stress), all-reduce sweep (quick / final: the sizes of
#   the train step; full: 1 KB - 1 GB), train-step bench in push and pull form (final: pull only), the
#   stand-in baseline (not in final).
# Everything lands in gpurun_out/scale_r2_n$N/. Keep it short: an N-GPU box is charged N x.
N=${1:-8}
mode=${2:-quick}
out=gpurun_out/scale_r2_n$N; [ "$mode" = final ] && out=gpurun_out/scale_r2_n${N}_final
mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
modes="push pull"; [ "$mode" = final ] && modes="pull"
for m in $modes; do
  timeout 150 $TR --master-port 29512 bench.py --gpus $N --steps 30 --warmup 5 --ar_mode $m > $out/bench_native_$m.json 2> $out/bench_native_$m.err
  tail -1 $out/bench_native_$m.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$m', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), round(d['e2e']['ms_per_step'],3), d['collective'], d['clocks'])"
done
if [ "$mode" = full ]; then sweep=""; else sweep="--quick"; fi
timeout 300 $TR --master-port 29511 tools/allreduce_bench.py $sweep > $out/allreduce_sweep.log 2>&1; grep -v "^{" $out/allreduce_sweep.log | tail -30
cp gpurun_out/allreduce_sweep_w$N.json $out/ 2>/dev/null
timeout 300 python -m pytest tests/test_fused_allreduce.py -m gpu -x -q -k auto > $out/pytest_fused_allreduce.log 2>&1; tail -3 $out/pytest_fused_allreduce.log
if [ "$mode" != final ]; then
  timeout 200 $TR --master-port 29513 bench.py --gpus $N --steps 20 --warmup 5 --impl standin > $out/bench_standin.json 2> $out/bench_standin.err
  tail -1 $out/bench_standin.json | cut -c100-330
fi
