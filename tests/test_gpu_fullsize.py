"""Parity at BASELINE.json's FULL sizes ("parity is checked on every sample for configs 1, 4, 5 and on >= 2 whole
streams + 1000 random 4096-frame windows for configs 2, 3"), without the CPU doing the whole job:

* every sample of whole streams: the tiled kernel's output against the DIRECT kernel's -- an independent device implementation of
  the reference's formulas (64-bit arithmetic, the original strided table, H:993-1033) -- by order-dependent checksum;
* 1000 random 4096-frame windows per configuration against the ORACLE, bit for bit;
* config 4: all 1024 voices through both streaming front ends (the unmodified callback API and the VoiceBatch), every sample,
  against the oracle's HighLevel stream by the harness's checksum;
* sine and adversarial full-scale inputs (the signals SURVEY.md 8d lists beside the noise) against the oracle, every sample.
"""
import ctypes as C
import json
import os
import subprocess

import numpy as np
import pytest

import clownresampler_b200 as crb
from conftest import ROOT, pad

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pre():
    assert crb.lib().ClownResamplerB200_Init(0) == 0, crb.last_error()
    return crb.Precompute()


def _noise_stream_on_device(L, ch, T, R, stream_id):
    d_in = crb.DeviceBuffer((T + 2 * R) * ch * 2)
    zeros = np.zeros(R * ch, dtype=np.int16)
    L.ClownResamplerB200_CopyToDevice(d_in.ptr, zeros.ctypes.data, zeros.nbytes)
    L.ClownResamplerB200_CopyToDevice(d_in.ptr + (R + T) * ch * 2, zeros.ctypes.data, zeros.nbytes)
    assert L.ClownResamplerB200_FillNoiseDevice(d_in.ptr + R * ch * 2, 20261017, stream_id, 0, T, ch, None) == 0
    return d_in


def _checksum(L, buf, words):
    v = C.c_ulong(0)
    assert L.ClownResamplerB200_ChecksumDevice(buf.ptr, words, 2, C.byref(v), None) == 0
    return v.value


def _check_windows(oracle, st, ch, rates, T, R, n, d_in, d_out, firsts):
    inc = st.increment
    for first in firsts:
        p0 = first * inc
        f0 = p0 >> 16                                        # first padded frame the window needs
        span = ((first + 4095) * inc >> 16) - f0 + 2 * R + 1
        span = min(span, T + 2 * R - f0)
        win = d_in.to_numpy(np.int16, span * ch, f0 * ch * 2).reshape(-1, ch)
        want = oracle.lowlevel(ch, *rates, win, span - 2 * R, 0, p0 & 0xFFFF, max_frames=4096)[0]
        got = d_out.to_numpy(np.int16, 4096 * ch, first * ch * 2).reshape(-1, ch)
        m = min(len(want), 4096)
        assert m > 0 and np.array_equal(got[:m], np.clip(want[:m], -0x7FFF, 0x7FFF).astype(np.int16)), first


def test_config2_four_whole_streams_and_1000_windows(pre, oracle):
    """Four DIFFERENT 10-minute stereo streams of config 2 (one lockstep job of four on the tiled kernel): every sample of all four
    against the direct kernel, 1000 random windows (250 per stream) against the oracle."""
    L = crb.lib()
    ch, i, o, T = 2, 44100, 48000, 44100 * 600
    st = crb.LowLevel_Init(ch, i, o, o)
    R = st.lowest_level.integer_stretched_kernel_radius
    n = crb.CountOutputFrames(st, T)
    assert n == 28_800_096
    ins = [_noise_stream_on_device(L, ch, T, R, 100 + k) for k in range(4)]
    outs = [crb.DeviceBuffer(n * ch * 2) for _ in range(4)]
    plan = crb.Plan(pre, st)
    plan.resample_device([crb.make_job(ins[k].ptr, outs[k].ptr, T, 0, 0, 0, n) for k in range(4)], fmt=crb.OUT_S16_CLAMPED)
    sums = [_checksum(L, outs[k], n * ch) for k in range(4)]
    assert len(set(sums)) == 4                               # four different streams
    os.environ["CRB200_FORCE_DIRECT"] = "1"
    try:
        dplan = crb.Plan(pre, st)
    finally:
        del os.environ["CRB200_FORCE_DIRECT"]
    chk = crb.DeviceBuffer(n * ch * 2)
    for k in range(4):
        dplan.resample_device([crb.make_job(ins[k].ptr, chk.ptr, T, 0, 0, 0, n)], fmt=crb.OUT_S16_CLAMPED)
        assert _checksum(L, chk, n * ch) == sums[k], k
    rng = np.random.default_rng(2)
    for k in range(4):
        firsts = [0, n - 4096] + [int(x) for x in rng.integers(0, n - 4096, size=248)]
        _check_windows(oracle, st, ch, (i, o, o), T, R, n, ins[k], outs[k], firsts)


def test_config3_the_whole_hour_and_1000_windows(pre, oracle):
    """Config 3 at its full size: one 8-channel stream, 3600 s, 192 -> 44.1 kHz (11 GB in): every sample against the direct kernel,
    1000 random windows against the oracle; and the same hour cut into 8 output-time segments gives the same checksum."""
    L = crb.lib()
    ch, i, o, T = 8, 192000, 44100, 192000 * 3600
    st = crb.LowLevel_Init(ch, i, o, o)
    R = st.lowest_level.integer_stretched_kernel_radius
    n = crb.CountOutputFrames(st, T)
    assert n == 158_760_447 and R == 14
    d_in = _noise_stream_on_device(L, ch, T, R, 7)
    d_out = crb.DeviceBuffer(n * ch * 2)
    plan = crb.Plan(pre, st)
    plan.resample_device([crb.make_job(d_in.ptr, d_out.ptr, T, 0, 0, 0, n)], fmt=crb.OUT_S16_CLAMPED)
    total = _checksum(L, d_out, n * ch)
    os.environ["CRB200_FORCE_DIRECT"] = "1"
    try:
        dplan = crb.Plan(pre, st)
    finally:
        del os.environ["CRB200_FORCE_DIRECT"]
    chk = crb.DeviceBuffer(n * ch * 2)
    dplan.resample_device([crb.make_job(d_in.ptr, chk.ptr, T, 0, 0, 0, n)], fmt=crb.OUT_S16_CLAMPED)
    assert _checksum(L, chk, n * ch) == total
    # the same hour as 8 contiguous output-time segments, each from its own slice of the input (as 8 GPUs would run it)
    from clownresampler_b200.sharding import segment_for_rank
    jobs = []
    for r in range(8):
        seg = segment_for_rank(st, T, r, 8)
        jobs.append(crb.make_job(d_in.ptr + seg.first_padded_input_frame * ch * 2, chk.ptr + seg.first_output_frame * ch * 2,
                                 seg.total_input_frames(R), seg.position_integer, seg.position_fractional, 0, seg.output_frames))
    L.ClownResamplerB200_CopyToDevice(chk.ptr, np.zeros(4096, dtype=np.int16).ctypes.data, 8192)
    plan.resample_device(jobs, fmt=crb.OUT_S16_CLAMPED)
    assert _checksum(L, chk, n * ch) == total
    rng = np.random.default_rng(3)
    firsts = [0, n - 4096] + [int(x) for x in rng.integers(0, n - 4096, size=998)]
    _check_windows(oracle, st, ch, (i, o, o), T, R, n, d_in, d_out, firsts)


def _harness_checksum(stream_s16, voices):
    """bench_highlevel.c: checksum = checksum * 31 + (unsigned short)sample over every voice's output in turn (mod 2^64)."""
    s = stream_s16.astype(np.uint16).astype(np.uint64)
    n = len(s)
    with np.errstate(over="ignore"):
        pw = np.ones(n + 1, dtype=np.uint64)
        for k in range(1, n + 1):                            # 31^k mod 2^64
            pw[k] = pw[k - 1] * np.uint64(31)
        h = np.uint64((s * pw[n - 1::-1][:n]).sum())         # hash of one voice's stream
        c = np.uint64(0)
        for _ in range(voices):
            c = c * pw[n] + h
    return int(c)


def test_config4_all_1024_voices_through_both_front_ends(pre, oracle):
    """BASELINE configs[3] at full size through tools/bench_highlevel.c: 1024 mono voices x 10 s, 22.05 -> 48 kHz, one 1024-frame tick
    at a time, (a) every voice through the unmodified ClownResampler_HighLevel_* callback API and (b) through the VoiceBatch.  Both
    runs' checksum over every output sample of every voice must equal the one computed from the oracle's HighLevel stream."""
    harness = os.path.join(ROOT, "clownresampler_b200", "lib", "bench-highlevel")
    if not os.path.exists(harness):
        pytest.skip("clownresampler_b200/lib/bench-highlevel not built (make voices-bench)")
    voices, seconds = 1024, 10
    T = 22050 * seconds
    x, data = 12345, np.empty(T, dtype=np.int16)             # the harness's LCG input (the same for every voice)
    for k in range(T):
        x = (x * 1664525 + 1013904223) & 0xFFFFFFFF
        data[k] = np.int16(np.uint16(x >> 16))
    want = np.clip(oracle.highlevel(1, 22050, 48000, 48000, data.reshape(-1, 1)), -0x7FFF, 0x7FFF).astype(np.int16).ravel()
    expect = _harness_checksum(want, voices)
    for extra in (["batch"], []):
        out = subprocess.run([harness, str(voices), str(seconds)] + extra, capture_output=True, text=True, timeout=600)
        assert out.returncode == 0, out.stderr[-400:]
        rec = json.loads(out.stdout.strip().splitlines()[-1])
        assert rec["output_frames"] == voices * len(want), rec
        assert rec["checksum"] == expect, rec["impl"]


SIGNALS = ["sine_1k", "sine_045_nyquist", "all_min", "all_max", "alternating", "period3_full_scale", "random_full_scale", "lobe_aligned"]


def _signal(kind, T, ch, in_rate, out_rate, rng):
    t = np.arange(T, dtype=np.float64)
    if kind == "sine_1k":
        col = np.rint(0.5 * 32767 * np.sin(2 * np.pi * 1000.0 * t / in_rate))
    elif kind == "sine_045_nyquist":
        col = np.rint(0.5 * 32767 * np.sin(2 * np.pi * 0.45 * min(in_rate, out_rate) / 2 * t / in_rate))
    elif kind == "all_min":
        col = np.full(T, -32768.0)
    elif kind == "all_max":
        col = np.full(T, 32767.0)
    elif kind == "alternating":
        col = np.where(t % 2 == 0, 32767.0, -32768.0)
    elif kind == "period3_full_scale":
        col = np.where(t % 3 == 0, -32768.0, 32767.0)
    elif kind == "random_full_scale":
        col = np.where(rng.integers(0, 2, size=T) == 0, 32767.0, -32768.0)
    else:
        # the sign pattern of the kernel's lobes around the centre, repeated: -32768 where the weight is positive and +32767 where it
        # is negative drives the accumulators to their extremes whenever a window lines up with it (SURVEY.md 8d iii, 7.3 item 2)
        width = max(1, int(round(in_rate / min(in_rate, out_rate))))
        lobe = (np.arange(T) // width) % 2
        col = np.where(lobe == 0, -32768.0, 32767.0)
    data = np.repeat(col.astype(np.int16)[:, None], ch, axis=1)
    for c in range(1, ch):                                   # the other channels: the same signal shifted and (odd channels) inverted
        data[:, c] = np.roll(data[:, 0], 7 * c)
        if c % 2:
            data[:, c] = np.where(data[:, c] == -32768, 32767, -data[:, c].astype(np.int32)).astype(np.int16)
    return data


@pytest.mark.parametrize("shape", [(2, 44100, 48000), (1, 22050, 48000), (8, 192000, 44100), (2, 48000, 44100), (2, 384000, 8000), (3, 8000, 44100)])
@pytest.mark.parametrize("kind", SIGNALS)
def test_sine_and_adversarial_full_scale_inputs(pre, oracle, shape, kind):
    """Sines (1 kHz and 0.45 x Nyquist, half scale, host lrint) and adversarial full-scale patterns -- constant extremes, alternating
    signs, lobe-aligned sign patterns -- through every kernel kind, every sample and both output formats against the oracle: the 32-bit
    ranges the plan proves (accumulator chains, the 16-bit chains of the unstretched kernel, the normaliser) hold at the extremes."""
    ch, i, o = shape
    T = 40000 if i <= 48000 else 160000
    data = _signal(kind, T, ch, i, o, np.random.default_rng(hash(kind) % 1000))
    st = crb.LowLevel_Init(ch, i, o, o)
    padded = pad(data, st.lowest_level.integer_stretched_kernel_radius)
    want = oracle.lowlevel(ch, i, o, o, padded, T)[0]
    got = crb.resample_array(pre, st, padded, T, fmt=crb.OUT_S32)
    assert np.array_equal(got, want)
    got16 = crb.resample_array(pre, st, padded, T, fmt=crb.OUT_S16_CLAMPED)
    assert np.array_equal(got16, np.clip(want, -0x7FFF, 0x7FFF).astype(np.int16))
