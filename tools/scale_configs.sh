#!/bin/bash
# BASELINE configs 3 and 4 at N GPUs of one box:  gpurun --gpus N -- tools/scale_configs.sh N
#   batch 8 / GPU in bf16 and --dtype mxfp8 (config 3), 512 x 512 batch 1 / GPU (config 4).
# Lands in gpurun_out/scale_r2_n$N_configs/. Keep it short: an N-GPU box is charged N x.
N=${1:-8}
out=gpurun_out/scale_r2_n${N}_configs
mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
show() { tail -1 $1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$2', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['collective'], d['clocks'])"; }
timeout 120 $TR --master-port 29521 bench.py --gpus $N --steps 20 --warmup 5 --batch_size 8 > $out/bench_b8_bf16.json 2> $out/bench_b8_bf16.err; show $out/bench_b8_bf16.json b8_bf16
timeout 120 $TR --master-port 29522 bench.py --gpus $N --steps 20 --warmup 5 --batch_size 8 --dtype mxfp8 > $out/bench_b8_mxfp8.json 2> $out/bench_b8_mxfp8.err; show $out/bench_b8_mxfp8.json b8_mxfp8
timeout 120 $TR --master-port 29523 bench.py --gpus $N --steps 20 --warmup 5 --batch_size 1 --image_size 512 > $out/bench_512_b1.json 2> $out/bench_512_b1.err; show $out/bench_512_b1.json 512_b1
