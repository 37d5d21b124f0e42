#!/usr/bin/env python
"""Kernel-level timing of the BASELINE.json configurations other than the headline one (those are parity-test
cases, not bench lines; this script produces the numbers quoted in DESIGN.md).  Device-resident inputs, CUDA
events on the launching stream, inputs far larger than L2 (or an explicit L2 flush for the small ones).
usage: python tools/bench_configs.py [--quick]   -> one JSON line per configuration"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import clownresampler_b200 as crb  # noqa: E402

HBM = 6551.0
try:
    HBM = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def time_launch(fn, reps=10, flush=None):
    stream = torch.cuda.current_stream()
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(reps):
        if flush is not None:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        fn()
        b.record(stream)
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    return float(np.median(ms)), float(np.min(ms))


def run(name, channels, in_rate, out_rate, lpf, seconds, streams, pre, L, segments=1, flush=None):
    st = crb.LowLevel_Init(channels, in_rate, out_rate, lpf)
    R = st.lowest_level.integer_stretched_kernel_radius
    T = in_rate * seconds
    n_out = crb.CountOutputFrames(st, T)
    plan = crb.Plan(pre, st)
    sptr = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    d_in = torch.zeros((streams, T + 2 * R, channels), dtype=torch.int16, device="cuda")
    d_out = torch.empty((streams, n_out, channels), dtype=torch.int16, device="cuda")
    for s in range(streams):
        assert L.ClownResamplerB200_FillNoiseDevice(C.c_void_p(d_in[s, R].data_ptr()), 7, s, 0, T, channels, sptr) == 0
    jobs = []
    for s in range(streams):
        for g in range(segments):
            n0, n1 = n_out * g // segments, n_out * (g + 1) // segments
            jobs.append(crb.make_job(d_in[s].data_ptr(), d_out[s, n0].data_ptr() if n1 > n0 else d_out[s].data_ptr(), T, 0, 0, n0, n1 - n0))
    jarr = crb.Plan._jobs(jobs)

    def launch():
        rc = L.ClownResamplerB200_ResampleDevice(plan.handle, jarr, len(jobs), crb.OUT_S16_CLAMPED, sptr)
        assert rc == 0, crb.last_error()

    med, best = time_launch(launch, flush=flush)
    frames = streams * n_out
    bytes_algo = streams * (T + 2 * R) * channels * 2 + frames * channels * 2
    macs = frames * channels * plan.info.mean_taps
    print(json.dumps({"config": name, "channels": channels, "rates": [in_rate, out_rate, lpf], "streams": streams, "seconds": seconds, "jobs": len(jobs),
                      "output_frames": frames, "ms": med, "ms_best": best, "msamples_per_s": frames * channels / med / 1e3,
                      "gb_per_s_algorithmic": bytes_algo / med / 1e6, "hbm_frac": bytes_algo / med / 1e6 / HBM, "tmac_per_s": macs / med / 1e9,
                      "bytes_per_frame": bytes_algo / frames, "macs_per_frame": channels * plan.info.mean_taps, "mean_taps": plan.info.mean_taps,
                      "columns": plan.info.columns, "tile_out": plan.info.tile_output_frames, "smem_bytes": plan.info.smem_bytes, "kernel_kind": plan.info.kernel_kind}))
    sys.stdout.flush()
    del d_in, d_out
    plan.destroy()
    torch.cuda.empty_cache()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    L = crb.lib()
    assert L.ClownResamplerB200_Init(0) == 0, crb.last_error()
    torch.cuda.set_device(0)
    pre = crb.Precompute()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")    # > 126 MB L2, written between timed launches of small configs
    q = args.quick
    run("config2: 64 x stereo 10 min 44.1->48k", 2, 44100, 48000, 48000, 60 if q else 600, 64, pre, L)
    run("config3: 8ch 1 h 192k->44.1k, 1 job", 8, 192000, 44100, 44100, 360 if q else 3600, 1, pre, L)
    run("config3: 8ch 1 h 192k->44.1k, 8 time segments", 8, 192000, 44100, 44100, 360 if q else 3600, 1, pre, L, segments=8)
    run("config4 (bulk form): 1024 mono voices 10 s 22.05->48k", 1, 22050, 48000, 48000, 10, 1024, pre, L, flush=flush)
    for ch in (1, 2):
        for (i, o) in [(8000, 16000), (8000, 44100), (8000, 48000), (8000, 96000), (8000, 192000), (8000, 384000),
                       (384000, 192000), (384000, 96000), (384000, 48000), (384000, 44100), (384000, 16000), (384000, 8000)]:
            secs = 30
            streams = max(1, int(4e8 // (max(i, o) * secs * ch)))     # ~0.8 GB of the larger side so that L2 cannot hold it
            run(f"config5 sweep: {ch}ch {i}->{o}", ch, i, o, o, secs, min(streams, 256), pre, L)


if __name__ == "__main__":
    main()
