#!/bin/bash
# 8-GPU bench lines of the final build (one box, 8 GPUs): gpurun --gpus 8 -- bash tools/bench_n8.sh
O=gpurun_out/n8; mkdir -p $O
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 8 "${@:2}" 2>/dev/null | tail -1; }
run 29511 --steps 20 --warmup 3 > $O/bench_n8.json
run 29512 --config 3 --steps 12 --warmup 3 > $O/bench_n8_config3.json
run 29513 --scaling weak --steps 10 --warmup 3 --no-e2e > $O/bench_n8_weak.json
ls -la $O
