#!/bin/bash
# 8-GPU config-2 lines only (strong + weak): gpurun --gpus 8 -- bash tools/bench_n8_c2.sh
O=gpurun_out/n8; mkdir -p $O
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 8 "${@:2}" 2>/dev/null | tail -1; }
run 29531 --steps 20 --warmup 3 > $O/bench_n8.json
run 29533 --scaling weak --steps 10 --warmup 3 --no-e2e > $O/bench_n8_weak.json
