"""GPU tests of the library's behaviour AROUND the hot path (threads, re-entrant callbacks, devices, the plan cache and the
failure contract) -- the parts of the drop-in boundary SURVEY.md 8b lists beside the arithmetic: the reference is re-entrant
per state (examples/low-level.c:87-102 call it from the audio thread) and its states are plain memcpy-able structs."""
import ctypes as C
import threading

import numpy as np
import pytest

import clownresampler_b200 as crb
from conftest import pad

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pre():
    assert crb.lib().ClownResamplerB200_Init(0) == 0, crb.last_error()
    return crb.Precompute()


def _stream(oracle, seed, ch, i, o, T):
    st = crb.LowLevel_Init(ch, i, o, min(i, o))
    R = st.lowest_level.integer_stretched_kernel_radius
    padded = pad(oracle.noise(seed, 0, 0, T, ch), R)
    want = oracle.lowlevel(ch, i, o, min(i, o), padded, T)[0]
    return st, padded, want


def test_concurrent_states_from_four_threads(pre, oracle):
    """Four threads, four different states, at once (H:677-681: one Precomputed shared by any number of resamplers)."""
    cases = [(1, 22050, 48000, 60000), (2, 44100, 48000, 50000), (2, 48000, 44100, 40000), (8, 192000, 44100, 30000)]
    work = [_stream(oracle, 10 + k, *c) for k, c in enumerate(cases)]
    got, errors = [None] * 4, []

    def run(k):
        try:
            for _ in range(3):
                st = crb.LowLevel_Init(cases[k][0], cases[k][1], cases[k][2], min(cases[k][1], cases[k][2]))
                got[k] = crb.LowLevel_Resample(st, pre, work[k][1], cases[k][3])[0]
        except Exception as e:  # pragma: no cover
            errors.append(e)

    threads = [threading.Thread(target=run, args=(k,)) for k in range(4)]
    [t.start() for t in threads]
    [t.join() for t in threads]
    assert not errors
    for k in range(4):
        assert np.array_equal(got[k], work[k][2]), cases[k]


def test_output_callback_may_drive_another_resampler(pre, oracle):
    """A mixer's output callback that itself pulls from a second resampler (a chained resampler): no lock is held across callbacks."""
    st_o, padded_o, want_o = _stream(oracle, 21, 2, 44100, 48000, 20000)
    st_i, padded_i, want_i = _stream(oracle, 22, 1, 48000, 44100, 15000)
    L = crb.lib()
    outer, inner = [], []

    def on_inner(_u, frame, n):
        inner.append([frame[k] for k in range(n)])
        return 1
    inner_cb = crb.ClownResampler_OutputCallback(on_inner)

    def on_outer(_u, frame, n):
        outer.append([frame[k] for k in range(n)])
        if len(outer) == 1000:      # in the middle of the outer stream, run the whole inner one
            total = C.c_size_t(15000)
            L.ClownResampler_LowLevel_Resample(C.byref(st_i), C.byref(pre), padded_i.ctypes.data, C.byref(total), inner_cb, None)
        return 1
    outer_cb = crb.ClownResampler_OutputCallback(on_outer)
    total = C.c_size_t(20000)
    assert L.ClownResampler_LowLevel_Resample(C.byref(st_o), C.byref(pre), padded_o.ctypes.data, C.byref(total), outer_cb, None) == 1
    assert np.array_equal(np.array(outer, dtype=np.int64), want_o)
    assert np.array_equal(np.array(inner, dtype=np.int64), want_i)


def test_continuous_pitch_bend_builds_one_plan(pre, reference):
    """LowLevel_Adjust to a new ratio every 400 frames (H:1052-1056): identical to the live reference call by call, and the cache
    serves every ratio from the plan of the kernel geometry (all up-sampling ratios share one) instead of building one per ratio."""
    L = crb.lib()
    ch, T, R = 2, 30000, 3
    data = np.random.default_rng(5).integers(-32768, 32768, size=(T, ch), dtype=np.int16)
    padded = pad(data, R)
    rates = [(44100, 48000 + 37 * k) for k in range(40)]
    st = crb.LowLevel_Init(ch, *rates[0], rates[0][1])
    crb.LowLevel_Resample(crb.LowLevel_Init(ch, *rates[0], rates[0][1]), pre, padded, 64)     # the geometry's plan exists from here on
    built = L.ClownResamplerB200_PlansBuilt()
    consumed, calls = 0, 0
    for k in range(200):
        i, o = rates[k % len(rates)]
        assert L.ClownResampler_LowLevel_Adjust(C.byref(st), i, o, o)
        remaining = T - consumed
        view = padded[consumed:]
        ref_out, ref_ret, ref_left, ref_pi, ref_pf = reference.lowlevel(ch, i, o, o, view, remaining, st.position_integer, st.position_fractional, 400)
        got, ret, left = crb.LowLevel_Resample(st, pre, view, remaining, 400)
        assert np.array_equal(got, ref_out.astype(np.int64)), (k, i, o)
        assert (ret, left, st.position_integer, st.position_fractional) == (ref_ret, ref_left, ref_pi, ref_pf), (k, i, o)
        consumed += remaining - left
        calls += 1
        if ret == 1:
            break
    assert calls > 50
    assert L.ClownResamplerB200_PlansBuilt() == built, "a ratio change within one kernel geometry must not build a plan"


@pytest.mark.parametrize("n_jobs", [5, 1])
def test_resample_host_multi(pre, oracle, n_jobs):
    """ClownResamplerB200_ResampleHostMulti: jobs dealt over the listed devices, or -- with fewer jobs than devices -- every job
    cut into output-time segments.  (On a one-GPU box the same device is listed twice: same code path, two workers.)"""
    L = crb.lib()
    n_dev = max(1, L.ClownResamplerB200_DeviceCount())
    devices = (C.c_int * 3)(*[k % n_dev for k in range(3)])
    ch, i, o, T = 2, 44100, 48000, 70001
    st = crb.LowLevel_Init(ch, i, o, o)
    R = st.lowest_level.integer_stretched_kernel_radius
    n_out = crb.CountOutputFrames(st, T)
    ins = [pad(oracle.noise(30 + k, 0, 0, T, ch), R) for k in range(n_jobs)]
    outs = [np.zeros((n_out, ch), dtype=np.int16) for _ in range(n_jobs)]
    jobs = crb.Plan._jobs([crb.make_job(ins[k].ctypes.data, outs[k].ctypes.data, T, 0, 0, 0, n_out) for k in range(n_jobs)])
    rc = L.ClownResamplerB200_ResampleHostMulti(C.byref(pre), C.byref(st), devices, 3, jobs, n_jobs, crb.OUT_S16_CLAMPED)
    assert rc == 0, crb.last_error()
    for k in range(n_jobs):
        want = np.clip(oracle.lowlevel(ch, i, o, o, ins[k], T)[0], -0x7FFF, 0x7FFF).astype(np.int16)
        assert np.array_equal(outs[k], want), k


def test_kept_frames_follow_the_whole_table(pre, oracle):
    """Frames computed ahead of a stopping callback are reused only while the caller's table is unchanged -- ALL of it."""
    ch, i, o, T = 1, 22050, 48000, 20000
    st, padded, want = _stream(oracle, 41, ch, i, o, T)
    mine = crb.ClownResampler_Precomputed()
    C.memmove(C.byref(mine), C.byref(pre), C.sizeof(mine))
    got1, ret, remaining = crb.LowLevel_Resample(st, mine, padded, T, max_frames=256)
    assert ret == 0 and np.array_equal(got1, want[:256])
    for k in range(2049, 4096, 2):                # every second entry of the main lobe, the first and the last entries untouched
        mine.lanczos_kernel_table[k] -= 300
    got2, _, _ = crb.LowLevel_Resample(st, mine, padded[T - remaining:], remaining, max_frames=256)
    want2 = oracle.lowlevel(ch, i, o, o, padded, T, table=crb.table_of(mine))[0][256:512]
    assert not np.array_equal(want2, want[256:512])          # the edit is audible ...
    assert np.array_equal(got2, want2)                       # ... and the library used the edited table, not the kept frames


def test_bad_state_stops_without_consuming_input(pre, capfd):
    """A state the reference itself cannot run (17 channels: H:1071 has 16 accumulator slots) fails loudly and consumes nothing."""
    st = crb.LowLevel_Init(2, 44100, 48000, 48000)
    st.channels = 17
    padded = np.zeros((1000 + 6, 17), dtype=np.int16)
    out, ret, remaining = crb.LowLevel_Resample(st, pre, padded, 1000)
    assert out.shape[0] == 0 and ret == 0 and remaining == 1000
    assert "channels" in crb.last_error()
    assert "clownresampler_b200" in capfd.readouterr().err


@pytest.mark.parametrize("case", [(2, 44100, 48000), (6, 22050, 48000), (3, 48000, 44100), (8, 192000, 44100)])
def test_planar_device_io(pre, oracle, case):
    """a stream held as one plane per channel (de-interleaved on the device from the interleaved s16 a decoder
    produces) resampled plane by plane -- lockstep mono streams -- and re-interleaved, against the oracle on the interleaved stream."""
    L = crb.lib()
    ch, i, o = case
    T = 30011
    st = crb.LowLevel_Init(ch, i, o, min(i, o))
    mono = crb.LowLevel_Init(1, i, o, min(i, o))
    R = st.lowest_level.integer_stretched_kernel_radius
    padded = pad(oracle.noise(50, 0, 0, T, ch), R)
    want = oracle.lowlevel(ch, i, o, min(i, o), padded, T)[0]
    n = want.shape[0]
    d_inter = crb.DeviceBuffer.from_numpy(padded)
    in_planes = [crb.DeviceBuffer((T + 2 * R) * 2) for _ in range(ch)]
    out_planes = [crb.DeviceBuffer(n * 4) for _ in range(ch)]
    ip = (C.c_void_p * ch)(*[b.ptr for b in in_planes])
    op = (C.c_void_p * ch)(*[b.ptr for b in out_planes])
    assert L.ClownResamplerB200_DeinterleaveDevice(d_inter.ptr, ip, T + 2 * R, ch, 2, None) == 0, crb.last_error()
    for c in range(ch):                                   # the planes hold the channels
        assert np.array_equal(in_planes[c].to_numpy(np.int16, T + 2 * R), padded[:, c])
    plan = crb.Plan(pre, mono)
    job = crb.ClownResamplerB200_PlanarJob(ip, op, ch, T, 0, 0, 0, n)
    assert L.ClownResamplerB200_ResamplePlanarDevice(plan.handle, C.byref(job), 1, crb.OUT_S32, None) == 0, crb.last_error()
    d_out = crb.DeviceBuffer(n * ch * 4)
    assert L.ClownResamplerB200_InterleaveDevice(op, d_out.ptr, n, ch, 4, None) == 0, crb.last_error()
    assert L.ClownResamplerB200_Synchronize(None) == 0
    got = d_out.to_numpy(np.int32, n * ch).reshape(n, ch)
    assert np.array_equal(got, want)
