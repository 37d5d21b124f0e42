#!/usr/bin/env python
"""Platform ceiling for the end-to-end number at N GPUs: every rank copies 1 GiB pinned host buffers to and from ITS GPU at the same
time as all the others (bare cudaMemcpyAsync on two streams), after a barrier.  Prints one JSON line: aggregate GB/s.
usage: python -m torch.distributed.run --nproc-per-node N tools/pcie_probe_multi.py"""
import json, os, time
import torch, torch.distributed as dist
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 1 << 30
h_in, h_out = torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory()
d_a, d_b = torch.empty(n, dtype=torch.uint8, device="cuda"), torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

def run(h2d, d2h, reps=4):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1): d_a.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h_out.copy_(d_b, non_blocking=True)
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    return world * reps * n * (h2d + d2h) / float(dt.item()) / 1e9
run(1, 1, 1)
res = {"n_gpus": world, "h2d_only_gbs": run(1, 0), "d2h_only_gbs": run(0, 1), "both_gbs": run(1, 1), "note": "aggregate over all ranks, 1 GiB pinned buffers, 4 repetitions, slowest rank's wall time"}
if rank == 0:
    print(json.dumps(res))
if world > 1:
    dist.destroy_process_group()
