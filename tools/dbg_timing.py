#!/usr/bin/env python
"""Debug-build only (tools/build_variant.sh dbg -DCRB_DEBUG_TIMING): runs the bench workload once and prints where the warps of
the tiled kernel spent their cycles.  usage: CRB200_LIB=variants/dbg/libclownresampler_b200.so python tools/dbg_timing.py [streams seconds [channels in_rate out_rate]]"""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import clownresampler_b200 as crb
S = int(sys.argv[1]) if len(sys.argv) > 1 else 64
SEC = int(sys.argv[2]) if len(sys.argv) > 2 else 600
CH = int(sys.argv[3]) if len(sys.argv) > 3 else 2
IN, OUT = (int(sys.argv[4]), int(sys.argv[5])) if len(sys.argv) > 5 else (44100, 48000)
L = crb.lib(); L.ClownResamplerB200_Init(0)
pre = crb.Precompute(); st = crb.LowLevel_Init(CH, IN, OUT, OUT)
T = IN * SEC; n_out = crb.CountOutputFrames(st, T); R = st.lowest_level.integer_stretched_kernel_radius
plan = crb.Plan(pre, st)
d_in = torch.zeros((S, T + 2 * R, CH), dtype=torch.int16, device="cuda"); d_out = torch.empty((S, n_out, CH), dtype=torch.int16, device="cuda")
for s in range(S): L.ClownResamplerB200_FillNoiseDevice(C.c_void_p(d_in[s, R].data_ptr()), 1, s, 0, T, CH, None)
jobs = crb.Plan._jobs([crb.make_job(d_in[s].data_ptr(), d_out[s].data_ptr(), T, 0, 0, 0, n_out) for s in range(S)])
out = (C.c_ulonglong * 72)()
for rep in range(3):
    L.ClownResamplerB200_ResampleDevice(plan.handle, jobs, S, 1, None)
    L.ClownResamplerB200_DebugTiming(out, 1)
v = list(out)
print("consumer warps: wait for tile %.0f clk/tile, inside tile %.0f clk/tile (%d warp-tiles); producer: wait for free stage %.0f clk/tile, stage-free -> copies issued %.0f clk (%d tiles)"
      % (v[0] / max(v[1], 1), v[4] / max(v[1], 1), v[1], v[2] / max(v[3], 1), v[5] / max(v[3], 1), v[3]))
print("buckets (warp index mod 4, hardware warp slot mod 4): tiles, wait clk/tile, work clk/tile")
for b in range(16):
    if v[40 + b]:
        print("  idx%%4=%d sched=%d  tiles %8d  wait %6.0f  work %6.0f" % (b // 4, b % 4, v[40 + b], v[8 + b] / v[40 + b], v[24 + b] / v[40 + b]))
