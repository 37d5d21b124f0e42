#!/bin/bash
# 2-GPU bench lines of the final build: gpurun --gpus 2 -- bash tools/bench_n2.sh
O=gpurun_out/n2; mkdir -p $O
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 "${@:2}" 2>/dev/null | tail -1; }
run 29521 --steps 20 --warmup 3 > $O/bench_n2.json
run 29522 --config 3 --steps 12 --warmup 3 > $O/bench_n2_config3.json
ls -la $O
