#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (BASELINE.json): output Msamples/s of the
Lanczos FIR resampler on B200, with the HBM roofline of the dominant kernel and the reference's
CPU path timed beside it.

  python bench.py --gpus N --steps K --warmup W            our arm (1 process per GPU; torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K ...  the reference's own CPU implementation (rank 0 only)

Workload (configs[1] of BASELINE.json, the one the metric is quoted on): a batch of 64 independent
stereo s16 streams, 10 minutes each, 44.1 kHz -> 48 kHz, clamped s16 output.  A step is one pass of
the hot path over the whole batch (one kernel launch).  Multi-GPU: streams are independent, every
rank resamples its own 64-stream batch (weak scaling, no data-path collective); time is the max over
ranks, value the samples all ranks produced per second.

PyTorch is plumbing here (device tensors, CUDA events on the launching stream, torch.distributed);
the product is libclownresampler_b200.so, called through its C ABI.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

STREAMS, CHANNELS, IN_RATE, OUT_RATE, SECONDS = 64, 2, 44100, 48000, 600
METRIC, UNIT = "output_msamples_per_s", "Msamples/s"
WORKLOAD = "64 x stereo s16 x 600 s, 44.1 kHz -> 48 kHz (BASELINE.json configs[1])"


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks and throttle reasons sampled DURING the timed region."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        clk = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(clk)) if clk else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(clk)}


def cpu_reference_run(n_streams, seconds, threads, o3=False):
    """Times the UNMODIFIED reference (oracle/_ref) on `n_streams` streams of `seconds` s of the same
    workload, spread over `threads` host threads (one independent resampler per stream, as the
    library is single-threaded by construction).  Only ClownResampler_LowLevel_Resample is timed per
    stream; the figure is samples / wall time of the slowest thread."""
    from oracle.cro import Oracle, Reference
    ref, orc = Reference(o3=o3), Oracle()
    T = IN_RATE * seconds
    R = 3
    base = np.zeros((T + 2 * R, CHANNELS), dtype=np.int16)
    base[R:R + T] = orc.noise(1, 0, 0, T, CHANNELS)
    per_thread = [n_streams // threads + (1 if t < n_streams % threads else 0) for t in range(threads)]
    frames = [0] * threads
    secs = [0.0] * threads

    def work(t):
        out = np.empty((T * OUT_RATE // IN_RATE + 16, CHANNELS), dtype=np.int16)
        for _ in range(per_thread[t]):
            s, f = ref.time_lowlevel(CHANNELS, IN_RATE, OUT_RATE, OUT_RATE, base, T, out)
            secs[t] += s
            frames[t] += f

    th = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
    t0 = time.perf_counter()
    [x.start() for x in th]
    [x.join() for x in th]
    wall = time.perf_counter() - t0
    busy = max(secs)
    return sum(frames) * CHANNELS / busy / 1e6, sum(frames), busy, wall


def run_reference(args, rank):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_streams, seconds = 2 * cores, SECONDS    # two full 10-minute streams per host thread per step (a bounded sample of the 64)
    vals = []
    for step in range(args.warmup + args.steps):
        v, frames, busy, wall = cpu_reference_run(n_streams, seconds, cores)
        if step >= args.warmup:
            vals.append((v, busy))
    value = float(np.mean([v for v, _ in vals]))
    ms = float(np.mean([b for _, b in vals])) * 1e3
    sample = f"{n_streams} of the 64 streams x {seconds} s per step, {cores} host threads (one independent resampler per stream), unmodified reference (oracle/_ref, gcc -O2)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64 (16.16 fixed point)", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--streams", type=int, default=STREAMS, help="streams per GPU (default: the BASELINE batch)")
    ap.add_argument("--seconds", type=int, default=SECONDS)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    import clownresampler_b200 as crb

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    L = crb.lib()
    if L.ClownResamplerB200_Init(local_rank) != 0:
        raise SystemExit("ClownResamplerB200_Init: " + crb.last_error())

    pre = crb.Precompute()
    st = crb.LowLevel_Init(CHANNELS, IN_RATE, OUT_RATE, OUT_RATE)
    R = st.lowest_level.integer_stretched_kernel_radius
    T = IN_RATE * args.seconds
    n_out = crb.CountOutputFrames(st, T)
    S = args.streams
    plan = crb.Plan(pre, st)
    stream = torch.cuda.current_stream()
    sptr = C.c_void_p(stream.cuda_stream)

    # ---- inputs resident in HBM: deterministic noise per (rank, stream), zero padding as tests/test-low-level.c:145-152
    d_in = torch.zeros((S, T + 2 * R, CHANNELS), dtype=torch.int16, device="cuda")
    d_out = torch.empty((S, n_out, CHANNELS), dtype=torch.int16, device="cuda")
    for s in range(S):
        rc = L.ClownResamplerB200_FillNoiseDevice(C.c_void_p(d_in[s, R].data_ptr()), 20261017, rank * S + s, 0, T, CHANNELS, sptr)
        assert rc == 0, crb.last_error()
    jobs = [crb.make_job(d_in[s].data_ptr(), d_out[s].data_ptr(), T, 0, 0, 0, n_out) for s in range(S)]
    jarr = crb.Plan._jobs(jobs)

    def step():
        rc = L.ClownResamplerB200_ResampleDevice(plan.handle, jarr, S, crb.OUT_S16_CLAMPED, sptr)
        if rc != 0:
            raise RuntimeError(crb.last_error())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    ev[0].record(stream)
    for k in range(args.steps):
        step()
        ev[k + 1].record(stream)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    per_launch_ms = [ev[k].elapsed_time(ev[k + 1]) for k in range(args.steps)]
    total_ms = ev[0].elapsed_time(ev[-1])
    t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    samples_per_step = S * n_out * CHANNELS * world
    value = samples_per_step / (ms_per_step * 1e-3) / 1e6

    # ---- roofline of the dominant (only) kernel: algorithmic bytes per launch / measured launch time
    bytes_in = S * (T + 2 * R) * CHANNELS * 2
    bytes_out = S * n_out * CHANNELS * 2
    launch_ms = float(np.mean(per_launch_ms))
    peak, peak_src = peaks()
    achieved = (bytes_in + bytes_out) / (launch_ms * 1e-3) / 1e9
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))["dram_bytes_per_launch_full_workload"]
    except Exception:
        pass
    macs_per_launch = S * n_out * CHANNELS * plan.info.mean_taps
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": peak_src, "kernel": "crb_tiled_kernel<2,1,1>", "algorithmic_bytes_per_launch": bytes_in + bytes_out,
                "launch_ms": launch_ms, "tmac_per_s": macs_per_launch / (launch_ms * 1e-3) / 1e12,
                "bytes_per_output_frame": (bytes_in + bytes_out) / (S * n_out), "macs_per_output_frame": CHANNELS * plan.info.mean_taps}

    # ---- end to end through the C ABI with HOST buffers (pinned): H2D + kernel + D2H inside the timed region
    e2e = None
    if not args.no_e2e:
        e2e_streams = S
        h_in = torch.empty((e2e_streams, T + 2 * R, CHANNELS), dtype=torch.int16, pin_memory=True)
        h_out = torch.empty((e2e_streams, n_out, CHANNELS), dtype=torch.int16, pin_memory=True)
        h_in.copy_(d_in[:e2e_streams])
        torch.cuda.synchronize()
        hjobs = crb.Plan._jobs([crb.make_job(h_in[s].data_ptr(), h_out[s].data_ptr(), T, 0, 0, 0, n_out) for s in range(e2e_streams)])
        e2e_steps = max(1, min(args.steps, 3))

        def e2e_step():
            rc = L.ClownResamplerB200_ResampleHost(plan.handle, hjobs, e2e_streams, crb.OUT_S16_CLAMPED)
            if rc != 0:
                raise RuntimeError(crb.last_error())
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        barrier()
        dt = torch.tensor([(time.perf_counter() - t0) / e2e_steps], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        # the device-resident run and the host run must agree (same kernel, same data)
        same = bool(torch.equal(h_out[0, :4096], d_out[0, :4096].cpu())) and bool(torch.equal(h_out[-1, -4096:], d_out[e2e_streams - 1, -4096:].cpu()))
        e2e = {"value": e2e_streams * n_out * CHANNELS * world / float(dt.item()) / 1e6, "unit": UNIT,
               "h2d_bytes_per_step": e2e_streams * (T + 2 * R) * CHANNELS * 2, "d2h_bytes_per_step": e2e_streams * n_out * CHANNELS * 2,
               "steps": e2e_steps, "s_per_step": float(dt.item()), "api": "ClownResamplerB200_ResampleHost (pinned host buffers)",
               "matches_device_run": same}
        del h_in, h_out

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        v, frames, busy, wall = cpu_reference_run(8, 600 if args.seconds >= 600 else args.seconds, 1)
        cpu = {"value": v, "unit": UNIT, "cores": 1, "kind": "reference",
               "sample": f"8 of the 64 streams ({frames} output frames), unmodified reference (oracle/_ref, gcc -O2), one thread, {busy:.1f} s inside ClownResampler_LowLevel_Resample"}
        try:    # BASELINE.md section 3 asks for the -O3 build beside the -O2 one
            v3, _, busy3, _ = cpu_reference_run(4, 600 if args.seconds >= 600 else args.seconds, 1, o3=True)
            cpu["value_gcc_O3_x86_64_v3"] = v3
        except Exception as e:  # pragma: no cover
            cpu["value_gcc_O3_x86_64_v3"] = f"unavailable: {e}"

    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32 (16.16 fixed point, exact per-tap truncation)", "data": "synthetic",
            "config": {"workload": WORKLOAD, "streams_per_gpu": S, "channels": CHANNELS, "input_frames_per_stream": T, "output_frames_per_stream": n_out,
                       "output_format": "s16 clamped", "l2": "inputs (%.2f GB per GPU) are far larger than the 126 MB L2; no explicit flush" % (bytes_in / 1e9),
                       "parallelism": f"streams x{world} ranks, no collective"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": args.steps, "clocks": clocks,
        }))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
