#!/usr/bin/env python
"""bench.py -- benchmark of the hot path (BASELINE.json): output Msamples/s of the Lanczos FIR resampler on B200, with the
roofline of the dominant kernel and the reference's CPU path timed beside it.

  python bench.py --gpus N --steps K --warmup W [--config C] [--scaling strong|weak]    our arm (1 process per GPU; torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K --warmup W [--config C]            the reference's own CPU implementation (rank 0)

--config selects the BASELINE.json shape (default 2, the one the metric is quoted on):
  1  tests/test.flac workload (stereo, 4 s) at the CTest rates 8000 -> 44100 and 44100 -> 8000 (correctness-sized; launch bound)
  2  64 x stereo x 600 s, 44.1 -> 48 kHz.  N GPUs: the 64 streams are dealt 64/N per GPU (--scaling strong, the default: BASELINE
     says "sharded by stream across 1/2/4/8 GPUs") or every GPU takes its own 64 (--scaling weak)
  3  one 8-channel stream, 1 hour, 192 -> 44.1 kHz with low-pass.  N GPUs: N contiguous output-time segments, each rank reads only
     its slice of the input plus the 14-frame halo; the final gather of the outputs to rank 0 (NCCL) is timed separately
  4  1024 mono voices 22.05 -> 48 kHz through the HighLevel-style streaming front ends, one 1024-frame tick at a time (the batched
     ClownResamplerB200_VoiceBatch and the unmodified callback API), plus the same audio as one bulk launch for the roofline
  5  ratio sweep 8 kHz <-> 384 kHz, mono and stereo, 30 s each: one line, per-ratio table in config.cases

A step is one pass of the hot path over the workload.  Device-timed `value`: inputs resident in HBM, CUDA events on the launching
stream, max over ranks.  `e2e`: the same through the C ABI with HOST buffers (pinned), H2D + kernel + D2H inside the timed region.
PyTorch is plumbing here (device tensors, CUDA events, torch.distributed); the product is libclownresampler_b200.so.
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

METRIC, UNIT = "output_msamples_per_s", "Msamples/s"
# measured issue rate of IMAD.HI / IMAD.WIDE (any multiply with a 64-bit product) on B200: 0.917 warp instructions per clock per SM
# (profiles/r01_pipe_microbench.jsonl, tools/microbench/overlap.cu) -> 148 SMs x 32 lanes x 0.917 x 1.965 GHz
EXACT_MAC_ROOF_TMACS = 148 * 32 * 0.917 * 1.965e9 / 1e12

SHAPES = {
    1: dict(name="tests/test.flac workload (BASELINE.json configs[0])", channels=2),
    2: dict(name="64 x stereo s16 x 600 s, 44.1 kHz -> 48 kHz (BASELINE.json configs[1])", streams=64, channels=2, in_rate=44100, out_rate=48000, lpf=48000, seconds=600),
    3: dict(name="1 x 8-channel s16 x 3600 s, 192 kHz -> 44.1 kHz with low-pass (BASELINE.json configs[2])", streams=1, channels=8, in_rate=192000, out_rate=44100, lpf=44100, seconds=3600),
    4: dict(name="1024 mono voices x 10 s, 22.05 kHz -> 48 kHz, 1024-frame ticks (BASELINE.json configs[3])", streams=1024, channels=1, in_rate=22050, out_rate=48000, lpf=48000, seconds=10),
    5: dict(name="ratio sweep 8 kHz <-> 384 kHz, mono and stereo, 30 s each (BASELINE.json configs[4])", seconds=30),
}
SWEEP = [(8000, o) for o in (16000, 44100, 48000, 96000, 192000, 384000)] + [(384000, o) for o in (192000, 96000, 48000, 44100, 16000, 8000)]


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def traffic_record(config):
    """DRAM bytes per launch from the committed ncu capture of the same kernel and workload (NOT measured in this run)."""
    for name in ("r02_traffic.json", "r01_traffic.json"):
        try:
            t = json.load(open(os.path.join(ROOT, "profiles", name)))
            if config == 2 and "dram_bytes_per_launch_full_workload" in t:
                return t["dram_bytes_per_launch_full_workload"], "profiles/" + name
        except Exception:
            pass
    return None, None


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


# ---------------------------------------------------------------------------------------------------------------------
# the reference's own CPU implementation (oracle/_ref: the unmodified header compiled in place)
# ---------------------------------------------------------------------------------------------------------------------
def cpu_reference_run(channels, in_rate, out_rate, lpf, n_streams, seconds, threads, o3=False):
    """Times the UNMODIFIED reference on `n_streams` streams of `seconds` s, spread over `threads` host threads (one independent
    resampler per stream: the library is single-threaded by construction).  Only ClownResampler_LowLevel_Resample is timed, with the
    examples' clamp-and-store-s16 callback; the figure is samples / busy time of the slowest thread."""
    from oracle.cro import Oracle, Reference
    ref, orc = Reference(o3=o3), Oracle()
    T = in_rate * seconds
    R = ref.configure(in_rate, out_rate, lpf)[1]
    base = np.zeros((T + 2 * R, channels), dtype=np.int16)
    base[R:R + T] = orc.noise(1, 0, 0, T, channels)
    per_thread = [n_streams // threads + (1 if t < n_streams % threads else 0) for t in range(threads)]
    frames, secs = [0] * threads, [0.0] * threads

    inc = ref.ratio(in_rate, out_rate)      # the reference steps by the TRUNCATED 16.16 ratio: a few frames more than T * out / in

    def work(t):
        out = np.empty(((T * 65536 + inc - 1) // inc + 16, channels), dtype=np.int16)
        for _ in range(per_thread[t]):
            s, f = ref.time_lowlevel(channels, in_rate, out_rate, lpf, base, T, out)
            secs[t] += s
            frames[t] += f

    th = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
    [x.start() for x in th]
    [x.join() for x in th]
    busy = max(secs)
    return sum(frames) * channels / busy / 1e6, sum(frames), busy


def reference_sample(config, cores):
    """(channels, in, out, lpf, streams, seconds, description) of the bounded CPU sample of a config: about 10-30 s of CPU work per step."""
    if config == 2:
        return 2, 44100, 48000, 48000, 2 * cores, 600, f"{2 * cores} of the 64 streams x 600 s per step"
    if config == 3:
        return 8, 192000, 44100, 44100, cores, 60, f"{cores} x 60 s slices of the 3600 s stream per step (cost per frame is constant)"
    if config == 4:
        return 1, 22050, 48000, 48000, 16 * cores, 10, f"{16 * cores} of the 1024 voices x 10 s per step, through ClownResampler_LowLevel_Resample"
    if config == 5:
        return 2, 384000, 44100, 44100, cores, 30, f"the stereo 384 -> 44.1 kHz point of the sweep (median tap count), {cores} x 30 s per step"
    return 2, 8000, 44100, 44100, cores, 24, f"{cores} x 24 s of stereo 8000 -> 44100 (the CTest rates) per step"


def run_reference(args, rank):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    ch, i, o, lpf, n_streams, seconds, what = reference_sample(args.config, cores)
    vals = []
    for step in range(args.warmup + args.steps):
        v, frames, busy = cpu_reference_run(ch, i, o, lpf, n_streams, seconds, cores)
        if step >= args.warmup:
            vals.append((v, busy))
    value = float(np.mean([v for v, _ in vals]))
    ms = float(np.mean([b for _, b in vals])) * 1e3
    sample = f"{what}, {cores} host threads (one independent resampler per stream), unmodified reference (oracle/_ref, gcc -O2)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "int64 (16.16 fixed point)", "data": "synthetic",
        "config": {"workload": SHAPES[args.config]["name"], "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ---------------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------------
class Ctx:
    def __init__(self, args):
        import torch
        import torch.distributed as dist
        import clownresampler_b200 as crb
        self.torch, self.dist, self.crb, self.args = torch, dist, crb, args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
        torch.cuda.set_device(self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
        self.L = crb.lib()
        if self.L.ClownResamplerB200_Init(self.local_rank) != 0:
            raise SystemExit("ClownResamplerB200_Init: " + crb.last_error())
        self.pre = crb.Precompute()
        self.stream = torch.cuda.current_stream()
        self.sptr = C.c_void_p(self.stream.cuda_stream)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, x):
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def time_steps(self, step, steps, warmup):
        """W warm-up steps, then K timed steps with CUDA events on the launching stream; returns (ms per step max over ranks, per-launch ms, clocks)."""
        torch = self.torch
        for _ in range(warmup):
            step()
        self.barrier()
        sampler = ClockSampler(self.local_rank)
        if self.rank == 0:
            sampler.start()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        ev[0].record(self.stream)
        for k in range(steps):
            step()
            ev[k + 1].record(self.stream)
        self.barrier()
        clocks = sampler.stop() if self.rank == 0 else None
        per_launch = [ev[k].elapsed_time(ev[k + 1]) for k in range(steps)]
        total_ms = self.max_over_ranks(ev[0].elapsed_time(ev[-1]))
        return total_ms / steps, per_launch, clocks

    def time_host(self, step, steps):
        """wall-clock seconds per step of a host-synchronous call, max over ranks (one untimed call first)"""
        step()
        self.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            step()
        self.barrier()
        return self.max_over_ranks((time.perf_counter() - t0) / steps)


class BulkJobs:
    """A set of independent device-resident jobs of one plan: noise inputs, one launch per step, and the same through host buffers."""

    def __init__(self, cx, channels, in_rate, out_rate, lpf, items, seed_base=0):
        """items: [(stream id, T, first padded frame, padded frames, pos_int, pos_frac, n_out)] -- a whole stream, or one output-time segment of it."""
        crb, torch, L = cx.crb, cx.torch, cx.L
        self.cx, self.channels = cx, channels
        self.st = crb.LowLevel_Init(channels, in_rate, out_rate, lpf)
        self.R = self.st.lowest_level.integer_stretched_kernel_radius
        self.plan = crb.Plan(cx.pre, self.st)
        self.items = items
        self.d_in, self.d_out, jobs = [], [], []
        for (sid, T, padded_first, padded_frames, pos_int, pos_frac, n_out) in items:
            d_in = torch.zeros((padded_frames, channels), dtype=torch.int16, device="cuda")
            d_out = torch.empty((max(n_out, 1), channels), dtype=torch.int16, device="cuda")
            # padded-buffer frame p holds real frame p - R; zeros outside [0, T) as tests/test-low-level.c:145-152
            lo, hi = max(padded_first - self.R, 0), min(padded_first + padded_frames - self.R, T)
            if hi > lo:
                rc = L.ClownResamplerB200_FillNoiseDevice(C.c_void_p(d_in[lo + self.R - padded_first].data_ptr()), 20261017, seed_base + sid, lo, hi - lo, channels, cx.sptr)
                assert rc == 0, crb.last_error()
            self.d_in.append(d_in)
            self.d_out.append(d_out)
            jobs.append(crb.make_job(d_in.data_ptr(), d_out.data_ptr(), padded_frames - 2 * self.R, pos_int, pos_frac, 0, n_out))
        self.jarr = crb.Plan._jobs(jobs)
        self.n_jobs = len(jobs)
        self.out_frames = sum(it[6] for it in items)
        self.bytes_in = sum(it[3] for it in items) * channels * 2
        self.bytes_out = self.out_frames * channels * 2

    def step(self):
        cx = self.cx
        rc = cx.L.ClownResamplerB200_ResampleDevice(self.plan.handle, self.jarr, self.n_jobs, cx.crb.OUT_S16_CLAMPED, cx.sptr)
        if rc != 0:
            raise RuntimeError(cx.crb.last_error())

    def host_setup(self):
        torch, crb = self.cx.torch, self.cx.crb
        self.h_in = [torch.empty(d.shape, dtype=torch.int16, pin_memory=True) for d in self.d_in]
        self.h_out = [torch.empty(d.shape, dtype=torch.int16, pin_memory=True) for d in self.d_out]
        for h, d in zip(self.h_in, self.d_in):
            h.copy_(d)
        torch.cuda.synchronize()
        jobs = [crb.make_job(self.h_in[k].data_ptr(), self.h_out[k].data_ptr(), it[3] - 2 * self.R, it[4], it[5], 0, it[6]) for k, it in enumerate(self.items)]
        self.hjarr = crb.Plan._jobs(jobs)

    def host_step(self):
        cx = self.cx
        rc = cx.L.ClownResamplerB200_ResampleHost(self.plan.handle, self.hjarr, self.n_jobs, cx.crb.OUT_S16_CLAMPED)
        if rc != 0:
            raise RuntimeError(cx.crb.last_error())

    def host_matches_device(self):
        torch = self.cx.torch
        ok = True
        for h, d, it in ((self.h_out[0], self.d_out[0], self.items[0]), (self.h_out[-1], self.d_out[-1], self.items[-1])):
            n = min(4096, it[6])
            ok = ok and bool(torch.equal(h[:n], d[:n].cpu())) and bool(torch.equal(h[it[6] - n:it[6]], d[it[6] - n:it[6]].cpu()))
        return ok

    def host_free(self):
        del self.h_in, self.h_out, self.hjarr


def whole_stream_items(cx, st, channels, T, ids):
    R = st.lowest_level.integer_stretched_kernel_radius
    n_out = cx.crb.CountOutputFrames(st, T)
    return [(sid, T, 0, T + 2 * R, 0, 0, n_out) for sid in ids]


def roofline(bytes_algorithmic, launch_ms, macs, kernel, traffic=None, traffic_source=None):
    peak, peak_src = peaks()
    achieved = bytes_algorithmic / (launch_ms * 1e-3) / 1e9
    tmacs = macs / (launch_ms * 1e-3) / 1e12
    return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic, "traffic_source": traffic_source if traffic is not None else None,
            "peak_source": peak_src, "kernel": kernel, "algorithmic_bytes_per_launch": bytes_algorithmic, "launch_ms": launch_ms,
            "exact_mac_issue": {"achieved": tmacs, "peak": EXACT_MAC_ROOF_TMACS, "unit": "TMAC/s", "frac": tmacs / EXACT_MAC_ROOF_TMACS,
                                "note": "peak = one IMAD.HI per exact MAC at its measured issue rate (0.917 warp instructions per clock per SM): the roof of "
                                        "the IMAD.HI kernels (general kernel for 1, 2, odd and 9-16 channels, slightly stretched kernels). The unstretched "
                                        "mono / stereo kernels and the 4/6/8-channel general kernel use the two-instruction MAC (PRMT + IMAD) instead; "
                                        "their binding unit is the ALU pipe, 69-74 % busy in profiles/r02_ncu_*_summary.txt"}}


def emit(cx, args, shape_name, ms_per_step, samples_all_ranks, cfg_extra, roof, cpu, e2e, launches, clocks, extra=None):
    if cx.rank != 0:
        return
    line = {
        "metric": METRIC, "value": samples_all_ranks / (ms_per_step * 1e-3) / 1e6, "unit": UNIT, "n_gpus": cx.world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "int32 (16.16 fixed point, exact per-tap truncation)", "data": "synthetic",
        "config": dict({"workload": shape_name, "output_format": "s16 clamped"}, **cfg_extra),
        "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
    }
    if extra:
        line.update(extra)
    print(json.dumps(line))


def cpu_baseline(cx, args, config):
    """the reference on ONE host core, rank 0 at N = 1 only, on a bounded sample of the workload"""
    if cx.rank != 0 or cx.world != 1 or args.no_cpu:
        return None
    ch, i, o, lpf, _, seconds, _ = reference_sample(config, 1)
    n = {2: 8, 3: 4, 4: 1024, 5: 20}.get(config, 8)       # about 10 s of one core each
    secs = {3: 300}.get(config, seconds)
    v, frames, busy = cpu_reference_run(ch, i, o, lpf, n, secs, 1)
    cpu = {"value": v, "unit": UNIT, "cores": 1, "kind": "reference",
           "sample": f"{n} x {secs} s of {ch}-channel {i} -> {o} Hz ({frames} output frames), unmodified reference (oracle/_ref, gcc -O2), one thread, {busy:.1f} s inside ClownResampler_LowLevel_Resample"}
    if config == 2:
        try:    # BASELINE.md section 3 asks for the -O3 build beside the -O2 one
            cpu["value_gcc_O3_x86_64_v3"] = cpu_reference_run(ch, i, o, lpf, 4, secs, 1, o3=True)[0]
        except Exception as e:  # pragma: no cover
            cpu["value_gcc_O3_x86_64_v3"] = f"unavailable: {e}"
    return cpu


def run_bulk_config(cx, args, config):
    """configs 2 and 3: one launch over the rank's share of the workload"""
    crb, sh = cx.crb, SHAPES[config]
    ch, i, o, lpf = sh["channels"], sh["in_rate"], sh["out_rate"], sh["lpf"]
    seconds = args.seconds or sh["seconds"]
    T = i * seconds
    st = crb.LowLevel_Init(ch, i, o, lpf)
    R = st.lowest_level.integer_stretched_kernel_radius
    gather = None
    if config == 2:
        S = args.streams or sh["streams"]
        if args.scaling == "strong":
            from clownresampler_b200.sharding import stream_shard
            ids = list(stream_shard(S, cx.rank, cx.world))
        else:
            ids = [cx.rank * S + k for k in range(S)]
        items = whole_stream_items(cx, st, ch, T, ids)
        parallelism = f"{S} streams dealt over {cx.world} rank(s), no collective" if args.scaling == "strong" else f"{S} streams per rank x {cx.world} rank(s), no collective"
    else:
        from clownresampler_b200.sharding import segment_for_rank
        seg = segment_for_rank(st, T, cx.rank, cx.world)      # strong by construction: one stream, N output-time segments
        items = [(0, T, seg.first_padded_input_frame, seg.padded_input_frames, seg.position_integer, seg.position_fractional, seg.output_frames)]
        parallelism = f"1 stream cut into {cx.world} contiguous output-time segment(s), each with its own input slice + {R}-frame halo; no data-path collective"
    work = BulkJobs(cx, ch, i, o, lpf, items)
    ms_per_step, per_launch, clocks = cx.time_steps(work.step, args.steps, args.warmup)
    samples = cx.sum_over_ranks(work.out_frames * ch)
    launch_ms = float(np.mean(per_launch))
    if config == 3 and cx.world > 1:
        # the optional final gather of the segments onto rank 0 (NCCL over NVLink), timed on its own
        torch, dist = cx.torch, cx.dist
        n_max = int(cx.max_over_ranks(work.out_frames))
        mine16 = torch.zeros((n_max, ch), dtype=torch.int16, device="cuda")
        mine16[:work.out_frames] = work.d_out[0][:work.out_frames]
        mine = mine16.view(torch.uint8)          # NCCL has no 16-bit integer type: the frames travel as bytes
        parts = [torch.empty_like(mine) for _ in range(cx.world)] if cx.rank == 0 else None
        dist.gather(mine, parts, dst=0)
        cx.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(cx.stream)
        dist.gather(mine, parts, dst=0)
        e1.record(cx.stream)
        cx.barrier()
        g_ms = cx.max_over_ranks(e0.elapsed_time(e1))
        gather = {"ms": g_ms, "bytes_to_rank0": int(samples * 2 * (cx.world - 1) / cx.world), "api": "torch.distributed.gather (NCCL)",
                  "ms_per_step_with_gather": ms_per_step + g_ms}
        del parts, mine, mine16
    traffic, tsrc = traffic_record(config) if (cx.world == 1 and not args.streams and not args.seconds) else (None, None)
    kernel = "crb_tiled_kernel<%d,1,%d>" % (ch, 1 if work.plan.info.kernel_kind == 0 and config == 2 else 0)
    roof = roofline(work.bytes_in + work.bytes_out, launch_ms, work.out_frames * ch * work.plan.info.mean_taps, kernel, traffic, tsrc)
    roof["bytes_per_output_frame"] = (work.bytes_in + work.bytes_out) / max(work.out_frames, 1)
    roof["macs_per_output_frame"] = ch * work.plan.info.mean_taps
    e2e = None
    if not args.no_e2e:
        work.host_setup()
        e2e_steps = max(1, min(args.steps, 3))
        dt = cx.time_host(work.host_step, e2e_steps)
        e2e = {"value": samples / dt / 1e6, "unit": UNIT, "h2d_bytes_per_step": work.bytes_in, "d2h_bytes_per_step": work.bytes_out,
               "steps": e2e_steps, "s_per_step": dt, "api": "ClownResamplerB200_ResampleHost (pinned host buffers), per rank",
               "matches_device_run": work.host_matches_device()}
        work.host_free()
    cpu = cpu_baseline(cx, args, config)
    cfg = {"channels": ch, "jobs_per_gpu": work.n_jobs, "input_frames_per_stream": T, "output_frames_this_rank": work.out_frames,
           "l2": "inputs (%.2f GB on this rank) are far larger than the 126 MB L2; no explicit flush" % (work.bytes_in / 1e9), "parallelism": parallelism}
    emit(cx, args, sh["name"], ms_per_step, samples, cfg, roof, cpu, e2e, args.steps, clocks, {"gather": gather} if gather else None)


def run_config1(cx, args):
    """the reference's own test workload: both CTest rate pairs over tests/test.flac (committed decoded fixture), bit-compared with the tripwires"""
    import gzip
    import hashlib
    crb, torch = cx.crb, cx.torch
    raw = gzip.open(os.path.join(ROOT, "tests", "golden", "test_flac_s16le.bin.gz"), "rb").read()
    pcm = np.frombuffer(raw, dtype="<i2").reshape(-1, 2).copy()
    trip = json.load(open(os.path.join(ROOT, "tests", "golden", "tripwires.json")))
    cases, total_frames, total_bytes = [], 0, 0
    for (i, o) in ((8000, 44100), (44100, 8000)):
        st = crb.LowLevel_Init(2, i, o, o)
        R = st.lowest_level.integer_stretched_kernel_radius
        padded = np.zeros((pcm.shape[0] + 2 * R, 2), dtype=np.int16)
        padded[R:R + pcm.shape[0]] = pcm
        n_out = crb.CountOutputFrames(st, pcm.shape[0])
        d_in = torch.from_numpy(padded).cuda()
        d_out = torch.empty((n_out, 2), dtype=torch.int32, device="cuda")
        plan = crb.Plan(cx.pre, st)
        jarr = crb.Plan._jobs([crb.make_job(d_in.data_ptr(), d_out.data_ptr(), pcm.shape[0], 0, 0, 0, n_out)])
        cases.append((plan, jarr, d_in, d_out, i, o))
        total_frames += n_out
        total_bytes += padded.nbytes + n_out * 8

    def step():
        for plan, jarr, *_ in cases:
            rc = cx.L.ClownResamplerB200_ResampleDevice(plan.handle, jarr, 1, crb.OUT_S32, cx.sptr)
            if rc != 0:
                raise RuntimeError(crb.last_error())
    ms_per_step, per_launch, clocks = cx.time_steps(step, args.steps, args.warmup)
    sha = {f"{i}->{o}": hashlib.sha256(d_out.cpu().numpy().astype("<i4").tobytes()).hexdigest() for _, _, _, d_out, i, o in cases}
    known = {v["sha256"] for v in trip.get("ctest_outputs", {}).values()}
    roof = roofline(total_bytes, float(np.mean(per_launch)), 0, "two launches (unstretched + general kernel)")
    emit(cx, args, SHAPES[1]["name"], ms_per_step, total_frames * 2 * cx.world, {"output_format": "s32 unclamped (what tests/test-low-level.c writes)", "launches_per_step": 2,
         "sha256": sha, "sha256_match_reference_tripwires": all(v in known for v in sha.values()), "parallelism": "replicas only (correctness-sized workload)"},
         roof, cpu_baseline(cx, args, 1), None, 2 * args.steps, clocks)


def run_config4(cx, args):
    """1024 voices: the streaming front ends through the C harness tools/bench_highlevel.c (host buffers by construction), and the
    same audio as one bulk launch for the kernel roofline"""
    crb, sh = cx.crb, SHAPES[4]
    ch, i, o, lpf = sh["channels"], sh["in_rate"], sh["out_rate"], sh["lpf"]
    voices, seconds = args.streams or sh["streams"], args.seconds or sh["seconds"]
    st = crb.LowLevel_Init(ch, i, o, lpf)
    items = whole_stream_items(cx, st, ch, i * seconds, [cx.rank * voices + k for k in range(voices)])
    work = BulkJobs(cx, ch, i, o, lpf, items)
    ms_bulk, per_launch, clocks = cx.time_steps(work.step, args.steps, args.warmup)
    roof = roofline(work.bytes_in + work.bytes_out, float(np.mean(per_launch)), work.out_frames * ch * work.plan.info.mean_taps, "crb_tiled_kernel<1,1,1>")
    harness = os.path.join(ROOT, "clownresampler_b200", "lib", "bench-highlevel")
    runs = {}
    env = dict(os.environ, CUDA_VISIBLE_DEVICES=str(cx.local_rank))
    if os.path.exists(harness):
        for name, extra in (("voice_batch", ["batch"]), ("dropin_highlevel", [])):
            best = None
            for _ in range(max(1, min(args.steps, 3)) + 1):      # first run warms the context up
                out = subprocess.run([harness, str(voices), str(seconds)] + extra, capture_output=True, text=True, env=env).stdout.strip().splitlines()
                rec = json.loads(out[-1]) if out else None
                if rec and (best is None or rec["wall_s"] < best["wall_s"]):
                    best = rec
            runs[name] = best
    vb = runs.get("voice_batch")
    samples_job = cx.sum_over_ranks(vb["output_frames"] if vb else 0)
    wall = cx.max_over_ranks(vb["wall_s"] if vb else 1.0)
    e2e = {"value": samples_job / wall / 1e6, "unit": UNIT, "h2d_bytes_per_step": voices * i * seconds * 2, "d2h_bytes_per_step": int(vb["output_frames"] * 2) if vb else 0,
           "s_per_step": wall, "api": "ClownResamplerB200_VoiceBatch{Push,Tick}: 1024-frame ticks, one upload + launch + download per tick, host buffers",
           "voice_ticks_per_s": vb["voice_ticks_per_s"] if vb else None, "dropin_highlevel": runs.get("dropin_highlevel")} if vb else None
    cfg = {"channels": ch, "voices_per_gpu": voices, "tick_frames": 1024, "value_is": "the same audio as ONE bulk launch, inputs resident (kernel roofline); e2e is the tick-by-tick streaming front end",
           "parallelism": f"{voices} voices per rank x {cx.world} rank(s), no collective", "l2": "bulk form: inputs %.0f MB per rank" % (work.bytes_in / 1e6)}
    emit(cx, args, sh["name"], ms_bulk, cx.sum_over_ranks(work.out_frames * ch), cfg, roof, cpu_baseline(cx, args, 4), e2e, args.steps, clocks)


def run_config5(cx, args):
    """ratio sweep: every (channels, in, out) point is its own plan and launch; one line with the per-point table"""
    crb = cx.crb
    seconds = args.seconds or SHAPES[5]["seconds"]
    cases = []
    tot_frames_ch = tot_ms = tot_bytes = tot_macs = 0.0
    clocks = None
    for ch in (1, 2):
        for (i, o) in SWEEP:
            st = crb.LowLevel_Init(ch, i, o, o)
            n_streams = 16           # 16 equal streams per point keep the persistent grid busy at the short end of the sweep
            work = BulkJobs(cx, ch, i, o, o, whole_stream_items(cx, st, ch, i * seconds, list(range(n_streams))), seed_base=1000 * ch)
            ms, per_launch, clocks = cx.time_steps(work.step, max(3, args.steps // 4), args.warmup)
            macs = work.out_frames * ch * work.plan.info.mean_taps
            peak, _ = peaks()
            cases.append({"channels": ch, "in": i, "out": o, "mean_taps": round(work.plan.info.mean_taps, 2), "ms": ms, "msamples_per_s": work.out_frames * ch / (ms * 1e-3) / 1e6,
                          "hbm_frac": (work.bytes_in + work.bytes_out) / (ms * 1e-3) / 1e9 / peak, "tmacs": macs / (ms * 1e-3) / 1e12})
            tot_frames_ch += work.out_frames * ch
            tot_ms += ms
            tot_bytes += work.bytes_in + work.bytes_out
            tot_macs += macs
            del work
            cx.torch.cuda.empty_cache()
    roof = roofline(tot_bytes, tot_ms, tot_macs, "all kernels of the sweep (sum of launch times)")
    emit(cx, args, SHAPES[5]["name"], tot_ms, cx.sum_over_ranks(tot_frames_ch), {"streams_per_point": 16, "seconds_per_stream": seconds, "cases": cases,
         "parallelism": "replicas only (every point is one launch)", "l2": "no explicit flush; every point's input exceeds L2 except the shortest"},
         roof, cpu_baseline(cx, args, 5), None, len(cases) * max(3, args.steps // 4), clocks)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[1, 2, 3, 4, 5])
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"], help="config 2 on N > 1 GPUs: deal the 64 streams (strong) or 64 per GPU (weak)")
    ap.add_argument("--streams", type=int, default=0, help="override the number of streams / voices (smaller profiling runs)")
    ap.add_argument("--seconds", type=int, default=0, help="override the stream length")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.config in (3,):
        args.scaling = "strong"
    if args.config in (1, 4, 5):
        args.scaling = "weak"
    if args.impl == "reference":
        run_reference(args, int(os.environ.get("RANK", "0")))
        return
    cx = Ctx(args)
    if args.config in (2, 3):
        run_bulk_config(cx, args, args.config)
    elif args.config == 1:
        run_config1(cx, args)
    elif args.config == 4:
        run_config4(cx, args)
    else:
        run_config5(cx, args)
    if cx.world > 1:
        cx.dist.destroy_process_group()


if __name__ == "__main__":
    main()
