#!/usr/bin/env python
"""Runs one configuration a few times on the device (for ncu captures / quick timing).
usage: python tools/run_config.py channels in_rate out_rate seconds streams [reps]"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import clownresampler_b200 as crb  # noqa: E402

ch, i, o, secs, streams = [int(x) for x in sys.argv[1:6]]
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 3
L = crb.lib()
assert L.ClownResamplerB200_Init(0) == 0
torch.cuda.set_device(0)
pre = crb.Precompute()
st = crb.LowLevel_Init(ch, i, o, o)
R = st.lowest_level.integer_stretched_kernel_radius
T = i * secs
n = crb.CountOutputFrames(st, T)
plan = crb.Plan(pre, st)
sptr = C.c_void_p(torch.cuda.current_stream().cuda_stream)
d_in = torch.zeros((streams, T + 2 * R, ch), dtype=torch.int16, device="cuda")
d_out = torch.empty((streams, n, ch), dtype=torch.int16, device="cuda")
for s in range(streams):
    L.ClownResamplerB200_FillNoiseDevice(C.c_void_p(d_in[s, R].data_ptr()), 7, s, 0, T, ch, sptr)
jobs = crb.Plan._jobs([crb.make_job(d_in[s].data_ptr(), d_out[s].data_ptr(), T, 0, 0, 0, n) for s in range(streams)])
ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
ev[0].record()
for r in range(reps):
    assert L.ClownResamplerB200_ResampleDevice(plan.handle, jobs, streams, crb.OUT_S16_CLAMPED, sptr) == 0, crb.last_error()
    ev[r + 1].record()
torch.cuda.synchronize()
ms = [ev[r].elapsed_time(ev[r + 1]) for r in range(reps)]
print("ms per launch:", [round(x, 3) for x in ms], "frames", streams * n, "Gsamples/s", streams * n * ch / min(ms) / 1e6)
