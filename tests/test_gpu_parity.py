"""Parity tests proper: the CUDA path, called through the C ABI of libclownresampler_b200.so,
against the oracle (oracle/cr_oracle.c), the committed reference vectors and the reference's
tripwire hashes.  Integer arithmetic throughout -> every comparison is bit-exact."""
import ctypes as C
import gzip
import hashlib
import os
import subprocess

import numpy as np
import pytest

import clownresampler_b200 as crb
from conftest import CTEST_CASES, GOLD, ROOT, pad

pytestmark = pytest.mark.gpu


def sha32(a):
    return hashlib.sha256(np.ascontiguousarray(a).astype("<i4").tobytes()).hexdigest()


@pytest.fixture(scope="module")
def pre():
    assert crb.lib().ClownResamplerB200_Init(0) == 0, crb.last_error()
    return crb.Precompute()


def state_for(ch, i, o, l, pos_int=0, pos_frac=0):
    st = crb.LowLevel_Init(ch, i, o, l)
    st.position_integer, st.position_fractional = pos_int, pos_frac
    return st


@pytest.mark.parametrize("rates", CTEST_CASES)
def test_ctest_workload_bulk_device(pre, flac_pcm, tripwires, rates):
    """tests/test-low-level.c's workload through the bulk device entry point."""
    i, o, l = rates
    st = state_for(2, i, o, l)
    R = st.lowest_level.integer_stretched_kernel_radius
    out = crb.resample_array(pre, st, pad(flac_pcm, R), flac_pcm.shape[0])
    want = tripwires["ctest_outputs"][f"low:{i}:{o}:{l}"]
    assert out.size * 4 == want["bytes"]
    assert sha32(out) == want["sha256"]


@pytest.mark.parametrize("rates", [CTEST_CASES[0], CTEST_CASES[2]])
def test_ctest_workload_dropin_lowlevel_callbacks(pre, flac_pcm, tripwires, oracle, rates):
    """The drop-in ClownResampler_LowLevel_Resample: per-frame callbacks, return value, state, leftovers."""
    i, o, l = rates
    st = state_for(2, i, o, l)
    R = st.lowest_level.integer_stretched_kernel_radius
    frames = 60000   # keep the Python per-frame callback affordable
    out, ret, remaining = crb.LowLevel_Resample(st, pre, pad(flac_pcm[:frames], R), frames)
    want = oracle.lowlevel(2, i, o, l, pad(flac_pcm[:frames], R), frames)
    assert np.array_equal(out, want[0].astype(np.int64))
    assert (ret, remaining, st.position_integer, st.position_fractional) == want[1:]


@pytest.mark.parametrize("case", [(1, 22050, 48000, 1024), (2, 44100, 8000, 37), (3, 48000, 44100, 700), (1, 22050, 48000, 20000)])
def test_dropin_tick_style_calls_reuse_frames_computed_ahead(pre, oracle, case):
    """A mixer takes one tick's worth of frames per call (the callback says stop, H:746-748) and comes back with the
    advanced buffer.  The frames the GPU computed ahead are kept and served to the following calls -- after a
    byte-for-byte comparison of the input they came from -- so the stream must equal the one-shot oracle, with far
    fewer kernel launches than calls; and a caller that CHANGES its input between calls must get the new result."""
    ch, i, o, tick = case
    rng = np.random.default_rng(tick)
    R = oracle.configure(i, o, o)[1]
    T = 30000 if tick < 4096 else 300000      # ticks beyond the first speculative chunk (16384 frames for mono): the kept frames start at a later chunk
    data = rng.integers(-32768, 32768, size=(T, ch), dtype=np.int16)
    padded = pad(data, R)
    want = oracle.lowlevel(ch, i, o, o, padded, T)[0]
    st = crb.LowLevel_Init(ch, i, o, o)
    launches0, kept0 = crb.counters()
    offset, remaining, got, calls = 0, T, [], 0
    while True:
        out, ret, left = crb.LowLevel_Resample(st, pre, padded[offset:], remaining, tick)
        got.append(out)
        offset += remaining - left
        remaining = left
        calls += 1
        if ret:
            break
    assert np.array_equal(np.concatenate(got), want.astype(np.int64))
    launches, kept = crb.counters()
    assert calls > 20 and kept - kept0 >= calls // 2, (calls, launches - launches0, kept - kept0)   # big ticks: several chunks per call, most calls still start from kept frames
    if tick < 4096:
        assert kept - kept0 >= calls * 2 // 3 and launches - launches0 <= calls // 2, (calls, launches - launches0, kept - kept0)
    # changed input: the kept frames must not be used
    st = crb.LowLevel_Init(ch, i, o, o)
    out1, ret, left = crb.LowLevel_Resample(st, pre, padded, T, tick)
    consumed = T - left
    changed = padded.copy()
    changed[consumed + R + 5:] = (changed[consumed + R + 5:].astype(np.int32) // 2).astype(np.int16)
    out2, ret, left2 = crb.LowLevel_Resample(st, pre, changed[consumed:], left, tick)
    ref = oracle.lowlevel(ch, i, o, o, changed, T)[0].astype(np.int64)
    assert np.array_equal(out1, want[:tick].astype(np.int64)) and np.array_equal(out2, ref[tick:2 * tick])
    assert not np.array_equal(out2, want[tick:2 * tick].astype(np.int64))


def test_legacy_normaliser_kat_against_shipped_golden(pre, flac_pcm, tripwires, oracle):
    """tests/test3 pins everything before the normaliser: take the GPU's un-normalised accumulators
    (diagnostic format) through the legacy normaliser on the host and compare with the golden."""
    raw = gzip.open(os.path.join(GOLD, "ref_test3_s32le.bin.gz"), "rb").read()
    gold = np.frombuffer(raw, dtype="<i4").reshape(-1, 2)
    st = state_for(2, 44100, 8000, 44100)
    R = st.lowest_level.integer_stretched_kernel_radius
    acc = crb.resample_array(pre, st, pad(flac_pcm, R), flac_pcm.shape[0], fmt=crb.OUT_S32_RAW)
    scale = oracle.ratio(8000, 44100)
    legacy = np.trunc(acc[:, :2].astype(np.int64) * scale / 65536.0)   # placeholder, exact integer form below
    a = acc[:, :2].astype(np.int64) * scale
    legacy = np.where(a >= 0, a // 65536, -((-a) // 65536))
    assert np.array_equal(legacy, gold.astype(np.int64))
    # and the reciprocal column is the reference's 0x80000000 / tap_sum
    norm = np.where(a >= 0, 0, 0)  # noqa: F841
    cur = acc[:, :2].astype(np.int64) * acc[:, 2:3].astype(np.int64)
    cur = np.where(cur >= 0, cur // 32768, -((-cur) // 32768))
    assert sha32(cur) == tripwires["ctest_outputs"]["low:44100:8000:44100"]["sha256"]


def test_reference_vectors_bulk(pre, ref_vectors):
    meta, data = ref_vectors
    n = 0
    for m in meta:
        if m["kind"] != "lowlevel":
            continue
        k = m["id"]
        st = state_for(m["channels"], m["in"], m["out"], m["lpf"], m["pos_int"], m["pos_frac"])
        frames = m["frames"]
        out = crb.resample_array(pre, st, data[f"ll{k}_in"], m["T"], output_frames=frames)
        assert np.array_equal(out, data[f"ll{k}_out"]), m
        n += 1
    assert n > 60


def test_reference_vectors_dropin_state(pre, ref_vectors):
    """Return value, leftover input and end state of the drop-in call, incl. early stop (H:1084-1088)."""
    meta, data = ref_vectors
    for m in meta:
        if m["kind"] != "lowlevel":
            continue
        k = m["id"]
        st = state_for(m["channels"], m["in"], m["out"], m["lpf"], m["pos_int"], m["pos_frac"])
        out, ret, remaining = crb.LowLevel_Resample(st, pre, data[f"ll{k}_in"], m["T"], m["limit"])
        assert np.array_equal(out, data[f"ll{k}_out"].astype(np.int64)), m
        assert (ret, remaining, st.position_integer, st.position_fractional) == (m["ret"], m["remaining"], m["end_pos_int"], m["end_pos_frac"]), m


def test_reference_vectors_highlevel(pre, ref_vectors):
    meta, data = ref_vectors
    n = 0
    for m in meta:
        if m["kind"] != "highlevel":
            continue
        k = m["id"]
        out = crb.HighLevel_Stream(pre, m["channels"], m["in"], m["out"], m["lpf"], data[f"hl{k}_in"], m["chunk"])
        assert np.array_equal(out, data[f"hl{k}_out"].astype(np.int64)), m
        n += 1
    assert n > 10


def test_random_sweep_vs_oracle(pre, oracle):
    """Every channel count 1..16, up/down ratios incl. the BASELINE shapes, random start state; s32 and clamped s16."""
    rng = np.random.default_rng(2026)
    rates = [8000, 11025, 16000, 22050, 44100, 48000, 88200, 96000, 192000, 384000]
    done = 0
    for ch in list(range(1, 17)) * 2:
        i, o = int(rng.choice(rates)), int(rng.choice(rates))
        l = int(rng.choice([i, o, 44100]))
        cfg = oracle.configure(i, o, l)
        R = cfg[1]
        T = int(rng.integers(2000, 30000))
        T = max(16, min(T, 60000 * oracle.ratio(i, o) // 65536))
        data = rng.integers(-32768, 32768, size=(T, ch), dtype=np.int16)
        data[: T // 8] = np.where(rng.random((T // 8, ch)) < 0.5, -32768, 32767)
        pi, pf = int(rng.integers(0, 3)), int(rng.integers(0, 65536))
        want = oracle.lowlevel(ch, i, o, l, pad(data, R), T, pi, pf)[0]
        got = crb.resample_array(pre, state_for(ch, i, o, l, pi, pf), pad(data, R), T)
        assert np.array_equal(got, want), (ch, i, o, l, T, pi, pf)
        got16 = crb.resample_array(pre, state_for(ch, i, o, l, pi, pf), pad(data, R), T, fmt=crb.OUT_S16_CLAMPED)
        assert np.array_equal(got16, np.clip(want, -0x7FFF, 0x7FFF).astype(np.int16)), (ch, i, o, l)
        done += 1
    assert done == 32


@pytest.mark.parametrize("case", [(4, 96000, 44100, 44100), (6, 192000, 44100, 44100), (8, 384000, 44100, 44100), (4, 44100, 16000, 16000),
                                  (8, 192000, 44100, 30000), (6, 88200, 32000, 11025)])
def test_chain_form_general_kernel_vs_oracle(pre, oracle, case):
    """4, 6, 8 channels, down-sampling without column rotation: the general kernel in its chain form (16-bit chains + single
    columns, crb_kernels.cuh frame_chains).  Full-scale runs of both signs and alternating signs stress the chains' range."""
    ch, i, o, l = case
    geo, _ = crb.debug_plan_host(pre, state_for(ch, i, o, l))
    assert geo["chain_mode"] == 1 and geo["kernel_kind"] == 0, geo
    rng = np.random.default_rng(ch * 31 + o)
    R = oracle.configure(i, o, l)[1]
    T = 3000 * max(1, i // o) + int(rng.integers(0, 777))
    data = rng.integers(-32768, 32768, size=(T, ch), dtype=np.int16)
    q = T // 5
    data[:q] = 32767
    data[q: 2 * q] = -32768
    data[2 * q: 3 * q] = np.where((np.arange(q)[:, None] + np.arange(ch)[None, :]) % 2 == 0, 32767, -32768)
    for pi, pf in [(0, 0), (2, int(rng.integers(1, 65536)))]:
        want = oracle.lowlevel(ch, i, o, l, pad(data, R), T, pi, pf)[0]
        got = crb.resample_array(pre, state_for(ch, i, o, l, pi, pf), pad(data, R), T)
        assert np.array_equal(got, want), (case, pi, pf)
        got16 = crb.resample_array(pre, state_for(ch, i, o, l, pi, pf), pad(data, R), T, fmt=crb.OUT_S16_CLAMPED)
        assert np.array_equal(got16, np.clip(want, -0x7FFF, 0x7FFF).astype(np.int16)), (case, pi, pf)


@pytest.mark.parametrize("case", [(2, 384000, 48000), (1, 384000, 8000), (2, 384000, 8000), (8, 192000, 48000), (4, 192000, 48000),
                                  (1, 384000, 48000), (2, 96000, 48000), (2, 384000, 96000), (6, 48000, 8000)])
def test_integer_ratio_downsampling_vs_oracle(pre, oracle, case):
    """Integer ratios: every lane of a warp sits on the same phase row and a multiple of the bank count apart, so the
    plan rotates the column order per lane (crb_plan.c step 7b); several tiles, ragged last tile, random start state."""
    ch, i, o = case
    rng = np.random.default_rng(ch * 7 + i // o)
    R = oracle.configure(i, o, o)[1]
    T = (i // o) * 5000 + int(rng.integers(0, 999))
    data = rng.integers(-32768, 32768, size=(T, ch), dtype=np.int16)
    data[: T // 8] = np.where(rng.random((T // 8, ch)) < 0.5, -32768, 32767)
    for pi, pf in [(0, 0), (1, int(rng.integers(1, 65536)))]:
        want = oracle.lowlevel(ch, i, o, o, pad(data, R), T, pi, pf)[0]
        got = crb.resample_array(pre, state_for(ch, i, o, o, pi, pf), pad(data, R), T)
        assert np.array_equal(got, want), (case, pi, pf)


@pytest.mark.parametrize("ch", [1, 2, 3, 6, 8])
@pytest.mark.parametrize("rates", [(48000, 44100), (48000, 32000), (96000, 48000), (44100, 32000), (44100, 22050), (48000, 47999), (88200, 48000)])
def test_slightly_stretched_kernels_vs_oracle(pre, oracle, ch, rates, monkeypatch):
    """Down-sampling by less than about 2 with up to eight channels runs the kernel that is unrolled over 6..12 signed
    taps (crb_device.cu frame_sk): s32, clamped s16 and the diagnostic format against the oracle, several tiles with
    a ragged last one, and the same plan forced onto the general kernel as a cross-check."""
    i, o = rates
    st = state_for(ch, i, o, o)
    geo, _ = crb.debug_plan_host(pre, st)
    assert geo["small_taps"] in ((6, 8, 10, 12) if ch <= 2 else (0, 6, 8, 10, 12))   # strides that pile onto few banks stay on the general kernel
    rng = np.random.default_rng(i + o + ch)
    R = oracle.configure(i, o, o)[1]
    T = 40000 + int(rng.integers(0, 999))
    data = rng.integers(-32768, 32768, size=(T, ch), dtype=np.int16)
    data[: T // 8] = np.where(rng.random((T // 8, ch)) < 0.5, -32768, 32767)
    data[T // 8: T // 4] = rng.integers(-2, 3, size=(T // 4 - T // 8, ch))
    pi, pf = 1, int(rng.integers(0, 65536))
    want = oracle.lowlevel(ch, i, o, o, pad(data, R), T, pi, pf)[0]
    got = crb.resample_array(pre, state_for(ch, i, o, o, pi, pf), pad(data, R), T)
    assert np.array_equal(got, want)
    got16 = crb.resample_array(pre, state_for(ch, i, o, o, pi, pf), pad(data, R), T, fmt=crb.OUT_S16_CLAMPED)
    assert np.array_equal(got16, np.clip(want, -0x7FFF, 0x7FFF).astype(np.int16))
    raw = crb.resample_array(pre, state_for(ch, i, o, o, pi, pf), pad(data, R), T, fmt=crb.OUT_S32_RAW)
    unnorm = oracle.lowlevel(ch, i, o, o, pad(data, R), T, pi, pf, norm=2)[0]
    assert np.array_equal(raw[:, :ch], unnorm)
    monkeypatch.setenv("CRB200_NO_SMALL", "1")
    assert crb.debug_plan_host(pre, st)[0]["small_taps"] == 0
    general = crb.resample_array(pre, state_for(ch, i, o, o, pi, pf), pad(data, R), T)
    assert np.array_equal(general, want)


SWEEP = [(8000, o) for o in (16000, 44100, 48000, 96000, 192000, 384000)] + [(384000, o) for o in (192000, 96000, 48000, 44100, 16000, 8000)]


@pytest.mark.parametrize("ch", [1, 2])
@pytest.mark.parametrize("rates", SWEEP)
def test_config5_ratio_sweep_every_sample(pre, oracle, ch, rates):
    """BASELINE config 5: every ratio of the sweep, mono and stereo, low-pass at the output rate, every sample against the
    oracle; sizes give several tiles of every kernel kind involved (5 to 288 taps)."""
    i, o = rates
    R = oracle.configure(i, o, o)[1]
    T = int(min(400000, max(3000, 20000 * i // o)))
    data = oracle.noise(55, ch, 0, T, ch)
    want = oracle.lowlevel(ch, i, o, o, pad(data, R), T)[0]
    got = crb.resample_array(pre, state_for(ch, i, o, o), pad(data, R), T, fmt=crb.OUT_S16_CLAMPED)
    assert got.shape[0] > 10000 or i > o
    assert np.array_equal(got, np.clip(want, -0x7FFF, 0x7FFF).astype(np.int16))


def test_six_channel_input_alignment(pre, oracle):
    """6-channel frames (12 bytes) are read with 32-bit loads: a device pointer that is 4- but not 16-byte aligned
    works (the tile's lead is then not a whole number of frames), a 2-byte aligned one is refused loudly."""
    ch, i, o = 6, 44100, 48000
    st = state_for(ch, i, o, o)
    R = st.lowest_level.integer_stretched_kernel_radius
    T = 30000
    data = np.random.default_rng(6).integers(-32768, 32768, size=(T, ch), dtype=np.int16)
    padded = pad(data, R)
    want = oracle.lowlevel(ch, i, o, o, padded, T, 0, 0)[0]
    n = want.shape[0]
    d_in = crb.DeviceBuffer(padded.nbytes + 64)
    d_out = crb.DeviceBuffer(n * ch * 4)
    plan = crb.Plan(pre, st)
    for shift in (4, 8, 12):
        crb.lib().ClownResamplerB200_CopyToDevice(C.c_void_p(d_in.ptr + shift), padded.ctypes.data, padded.nbytes)
        plan.resample_device([crb.make_job(d_in.ptr + shift, d_out.ptr, T, 0, 0, 0, n)], fmt=crb.OUT_S32)
        assert np.array_equal(d_out.to_numpy(np.int32, n * ch).reshape(n, ch), want), shift
    with pytest.raises(crb.Error, match="aligned"):
        plan.resample_device([crb.make_job(d_in.ptr + 2, d_out.ptr, T, 0, 0, 0, n)], fmt=crb.OUT_S32)
    plan.destroy()


def test_mono_s16_output_at_odd_sample_offsets(pre, oracle):
    """Mono s16 output: full tiles store adjacent frame pairs as one 32-bit word when the pairs are 4-byte aligned; an output
    pointer on an odd sample takes the frame-by-frame path.  Both must match, for one stream and for lockstep groups."""
    ch, i, o = 1, 22050, 48000
    st = state_for(ch, i, o, o)
    R = st.lowest_level.integer_stretched_kernel_radius
    T = 40000                                # several full tiles per stream
    rng = np.random.default_rng(61)
    streams = [rng.integers(-32768, 32768, size=(T, ch), dtype=np.int16) for _ in range(4)]
    streams[1][:2000] = 32767                # full scale: the clamp at work
    streams[2][:2000] = -32768
    padded = [pad(d, R) for d in streams]
    want = [np.clip(oracle.lowlevel(ch, i, o, o, p, T, 0, 0)[0], -0x7FFF, 0x7FFF).astype(np.int16) for p in padded]
    n = want[0].shape[0]
    d_in = [crb.DeviceBuffer(p.nbytes) for p in padded]
    for b, p in zip(d_in, padded):
        crb.lib().ClownResamplerB200_CopyToDevice(C.c_void_p(b.ptr), p.ctypes.data, p.nbytes)
    stride = (n + 8) * 2
    d_out = crb.DeviceBuffer(4 * stride + 64)
    plan = crb.Plan(pre, st)
    for shift in (0, 2):                     # bytes: 2 = odd sample
        for count in (1, 4):                 # four equal streams form a lockstep group
            jobs = [crb.make_job(d_in[s].ptr, d_out.ptr + shift + s * stride, T, 0, 0, 0, n) for s in range(count)]
            plan.resample_device(jobs, fmt=crb.OUT_S16_CLAMPED)
            got = d_out.to_numpy(np.int16, (4 * stride + 64) // 2)
            for s in range(count):
                first = (shift + s * stride) // 2
                assert np.array_equal(got[first: first + n], want[s][:, 0]), (shift, count, s)
    plan.destroy()


def test_device_noise_matches_oracle_generator(pre, oracle):
    n, ch = 5000, 3
    buf = crb.DeviceBuffer(n * ch * 2)
    assert crb.lib().ClownResamplerB200_FillNoiseDevice(buf.ptr, 77, 5, 123456789012, n, ch, None) == 0
    assert crb.lib().ClownResamplerB200_Synchronize(None) == 0
    got = buf.to_numpy(np.int16).reshape(n, ch)
    assert np.array_equal(got, oracle.noise(77, 5, 123456789012, n, ch))


@pytest.mark.parametrize("prog", ["dropin-test-low-level", "dropin-test-high-level"])
@pytest.mark.parametrize("rates", [CTEST_CASES[0], CTEST_CASES[3]])
def test_reference_test_programs_unmodified_against_the_library(tmp_path, tripwires, prog, rates):
    """The reference's tests/test-low-level.c and tests/test-high-level.c, compiled UNMODIFIED against
    include/clownresampler.h and linked with the CUDA library (Makefile target `dropin`), must write
    byte-identical files to what the reference build writes (tests/CMakeLists.txt:25-47)."""
    exe = os.path.join(ROOT, "oracle", "_ref", prog)
    flac = os.path.join(ROOT, "oracle", "_ref", "test.flac")
    if not (os.path.exists(exe) and os.path.exists(flac)):
        pytest.skip("drop-in binaries were not prebuilt (they need /root/reference at build time)")
    i, o, l = rates
    out = tmp_path / "out.bin"
    subprocess.check_call([exe, flac, str(out), str(i), str(o), str(l)], stderr=subprocess.DEVNULL)
    kind = "low" if "low" in prog else "high"
    want = tripwires["ctest_outputs"][f"{kind}:{i}:{o}:{l}"]
    data = out.read_bytes()
    assert len(data) == want["bytes"]
    assert hashlib.sha256(data).hexdigest() == want["sha256"]


def test_segment_sharding_on_device(pre, oracle):
    """SURVEY.md 8e on the GPU: 8 output-time segments of an 8-channel 192 -> 44.1 kHz stream, each run as its
    own job on a private copy of its slice (what 8 GPUs would each do), concatenate to the one-shot result."""
    from clownresampler_b200.sharding import segment_for_rank
    ch, i, o, l, world = 8, 192000, 44100, 44100, 8
    st = state_for(ch, i, o, l)
    R = st.lowest_level.integer_stretched_kernel_radius
    T = 300000
    data = oracle.noise(3, 0, 0, T, ch)
    padded = pad(data, R)
    whole = oracle.lowlevel(ch, i, o, l, padded, T)[0]
    parts = []
    for rank in range(world):
        seg = segment_for_rank(st, T, rank, world)
        private = padded[seg.first_padded_input_frame: seg.first_padded_input_frame + seg.padded_input_frames].copy()
        s2 = state_for(ch, i, o, l, seg.position_integer, seg.position_fractional)
        parts.append(crb.resample_array(pre, s2, private, seg.total_input_frames(R), output_frames=seg.output_frames))
    assert np.array_equal(np.concatenate(parts), whole)
    # the same segments as jobs of ONE launch on the shared buffer (first_output_frame addressing)
    plan = crb.Plan(pre, st)
    d_in = crb.DeviceBuffer.from_numpy(padded)
    d_out = crb.DeviceBuffer(whole.size * 4)
    jobs = []
    for rank in range(world):
        seg = segment_for_rank(st, T, rank, world)
        jobs.append(crb.make_job(d_in.ptr, d_out.ptr + seg.first_output_frame * ch * 4, T, 0, 0, seg.first_output_frame, seg.output_frames))
    plan.resample_device(jobs)
    assert np.array_equal(d_out.to_numpy(np.int32).reshape(-1, ch), whole)


def test_many_jobs_one_launch(pre, oracle):
    """A batch of independent mono voices (more than fit the kernel-parameter job table) in one launch,
    with different lengths and start states: the batched form config 4 needs."""
    ch, i, o = 1, 22050, 48000
    st = state_for(ch, i, o, o)
    R = st.lowest_level.integer_stretched_kernel_radius
    plan = crb.Plan(pre, st)
    rng = np.random.default_rng(8)
    jobs, wants, bufs = [], [], []
    for v in range(37):
        T = int(rng.integers(1, 9000))
        pi, pf = int(rng.integers(0, 2)), int(rng.integers(0, 65536))
        data = oracle.noise(9, v, 0, T, ch)
        want = oracle.lowlevel(ch, i, o, o, pad(data, R), T, pi, pf)[0]
        d_in = crb.DeviceBuffer.from_numpy(pad(data, R))
        d_out = crb.DeviceBuffer(max(want.size, 1) * 2)
        jobs.append(crb.make_job(d_in.ptr, d_out.ptr, T, pi, pf, 0, want.shape[0]))
        wants.append(want)
        bufs.append((d_in, d_out))
    plan.resample_device(jobs, fmt=crb.OUT_S16_CLAMPED)
    for (d_in, d_out), want in zip(bufs, wants):
        got = d_out.to_numpy(np.int16, want.size).reshape(-1, ch)
        assert np.array_equal(got, np.clip(want, -0x7FFF, 0x7FFF).astype(np.int16))


def test_host_bulk_entry_point(pre, oracle):
    """ClownResamplerB200_ResampleHost: pageable host buffers in, host buffers out, chunked copies overlapped."""
    for ch, i, o in ((2, 44100, 48000), (8, 192000, 44100)):
        st = state_for(ch, i, o, o, 1, 4242)
        R = st.lowest_level.integer_stretched_kernel_radius
        T = 3_000_000 if ch == 2 else 1_500_000     # several 2M-frame chunks for the stereo case
        data = oracle.noise(4, 1, 0, T, ch)
        want = oracle.lowlevel(ch, i, o, o, pad(data, R), T, 1, 4242)[0]
        got = crb.resample_array(pre, st, pad(data, R), T, fmt=crb.OUT_S16_CLAMPED, via="host")
        assert np.array_equal(got, np.clip(want, -0x7FFF, 0x7FFF).astype(np.int16))


def test_direct_kernel_matches_tiled_kernel(pre, oracle, monkeypatch):
    """The global-memory kernel (used when a tile's window cannot fit shared memory) evaluates the reference's
    formulas independently of the per-phase plan rows: both kernels must agree with the oracle."""
    cases = [(2, 44100, 48000, 48000), (8, 192000, 44100, 44100), (1, 384000, 8000, 8000), (5, 8000, 44100, 8000)]
    for ch, i, o, l in cases:
        R = oracle.configure(i, o, l)[1]
        T = max(64, min(40000, 30000 * oracle.ratio(i, o) // 65536))
        data = oracle.noise(6, ch, 0, T, ch)
        want = oracle.lowlevel(ch, i, o, l, pad(data, R), T, 0, 777)[0]
        monkeypatch.setenv("CRB200_FORCE_DIRECT", "1")
        st = state_for(ch, i, o, l, 0, 777)
        plan = crb.Plan(pre, st)
        assert plan.info.kernel_kind == 1
        plan.destroy()
        direct = crb.resample_array(pre, st, pad(data, R), T)
        monkeypatch.delenv("CRB200_FORCE_DIRECT")
        tiled = crb.resample_array(pre, st, pad(data, R), T)
        assert np.array_equal(direct, want) and np.array_equal(tiled, want), (ch, i, o, l)


def test_lowest_level_single_frame(pre, oracle):
    """ClownResampler_LowestLevel_Resample (H:688): one frame, accumulating into the caller's zeroed frame."""
    L = crb.lib()
    ch, i, o = 3, 48000, 32000
    st = state_for(ch, i, o, o)
    R = st.lowest_level.integer_stretched_kernel_radius
    data = oracle.noise(2, 0, 0, 400, ch)
    padded = pad(data, R)
    for pos_int, frac in ((0, 0), (7, 12345), (200, 65535)):
        frame = (crb.cc_s32f * ch)(*([0] * ch))
        L.ClownResampler_LowestLevel_Resample(C.byref(st.lowest_level), C.byref(pre), frame, ch, padded.ctypes.data, pos_int, frac)
        want = oracle.lowlevel(ch, i, o, o, padded, pos_int + 1, pos_int, frac, max_frames=1)[0][0]
        assert list(frame) == list(want)


@pytest.mark.parametrize("case", [(1, 22050, 48000, 48000), (2, 48000, 44100, 44100), (3, 44100, 8000, 8000)])
def test_voice_batch_matches_highlevel_streams(pre, oracle, case):
    """The batched streaming front end: every voice must emit exactly what the
    reference's HighLevel_Resample + HighLevel_ResampleEnd emit for its input, whatever the push/tick pattern."""
    ch, i, o, l = case
    rng = np.random.default_rng(ch)
    voices = 23
    lengths = [int(x) for x in rng.integers(0, 9000, size=voices)]
    lengths[0], lengths[1] = 0, 2            # shorter than the kernel radius, and empty
    data = [oracle.noise(21, v, 0, n, ch) if n else np.zeros((0, ch), dtype=np.int16) for v, n in enumerate(lengths)]
    want = [oracle.highlevel(ch, i, o, l, d) for d in data]
    vb = crb.VoiceBatch(pre, voices, ch, i, o, l)
    pos = [0] * voices
    got = [[] for _ in range(voices)]
    tick_frames = 700
    for _ in range(400):
        for v in range(voices):
            if pos[v] < lengths[v]:
                n = int(rng.integers(0, 600))
                vb.push(v, data[v][pos[v]:pos[v] + n])
                pos[v] = min(pos[v] + n, lengths[v])
            if pos[v] >= lengths[v]:
                vb.end(v)
        out, produced = vb.tick(tick_frames)
        for v in range(voices):
            got[v].append(out[v, :produced[v]].copy())
        if all(p >= n for p, n in zip(pos, lengths)) and produced.sum() == 0:
            break
    for v in range(voices):
        g = np.concatenate(got[v]) if got[v] else np.zeros((0, ch), dtype=np.int32)
        assert g.shape == want[v].shape, (v, g.shape, want[v].shape)
        assert np.array_equal(g, want[v]), v
    vb.destroy()


def test_voice_batch_pinned_output_takes_the_download_directly(pre, oracle):
    """A pinned caller buffer receives the tick's download without a staging copy (voice v at its stride): same frames
    as the staged path, for voices that run dry at different times."""
    ch, i, o = 2, 22050, 48000
    voices, tick_frames = 9, 512
    lengths = [0, 3, 900, 2500, 4000, 4001, 7000, 7000, 123]
    data = [oracle.noise(5, v, 0, n, ch) if n else np.zeros((0, ch), dtype=np.int16) for v, n in enumerate(lengths)]
    want = [oracle.highlevel(ch, i, o, o, d) for d in data]
    L = crb.lib()
    L.ClownResamplerB200_PinnedAlloc.restype = C.c_void_p
    nbytes = voices * tick_frames * ch * 4
    ptr = L.ClownResamplerB200_PinnedAlloc(C.c_size_t(nbytes))
    assert ptr
    pinned = np.ctypeslib.as_array((C.c_int32 * (nbytes // 4)).from_address(ptr)).reshape(voices, tick_frames, ch)
    vb = crb.VoiceBatch(pre, voices, ch, i, o, o)
    for v in range(voices):
        vb.push(v, data[v])
        vb.end(v)
    got = [[] for _ in range(voices)]
    for _ in range(64):
        pinned[:] = -1
        out, produced = vb.tick(tick_frames, out=pinned)
        assert out is pinned
        for v in range(voices):
            got[v].append(pinned[v, :produced[v]].copy())
        if produced.sum() == 0:
            break
    for v in range(voices):
        g = np.concatenate(got[v])
        assert np.array_equal(g, want[v]), v
    vb.destroy()
    L.ClownResamplerB200_PinnedFree(C.c_void_p(ptr))


def test_midstream_adjust_against_live_reference(pre, reference):
    """ClownResampler_LowLevel_Adjust between calls (pitch-bend).  The drop-in keeps the
    caller's position, switches plans by configuration, and must match the unmodified reference run through the
    same call sequence (oracle/_ref; skipped where it was not prebuilt)."""
    L = crb.lib()
    rng = np.random.default_rng(77)
    ch, T = 2, 60000
    data = rng.integers(-32768, 32768, size=(T, ch), dtype=np.int16)
    segments = [(44100, 48000, 48000, 5000), (44100, 32000, 32000, 3000), (48000, 44100, 20000, 4000), (22050, 48000, 48000, 7000), (44100, 44100, 44100, 0)]
    R = 8     # >= every segment's radius: the same padded buffer serves all of them
    padded = pad(data, R)
    for seg in segments:
        assert reference.configure(*seg[:3])[1] <= R
    # the reference reads `integer_stretched_kernel_radius` frames before the pointer it is given, so every
    # segment is handed the pointer (current position - its own radius), as the low-level contract asks
    st = crb.LowLevel_Init(ch, *segments[0][:3])
    out_frames, consumed = [], 0
    for seg in segments:
        assert L.ClownResampler_LowLevel_Adjust(C.byref(st), *seg[:3])
        r = st.lowest_level.integer_stretched_kernel_radius
        remaining = T - consumed
        if remaining == 0:
            break
        view = padded[R + consumed - r:]
        # reference for this segment, from the same start state
        ref_out, ref_ret, ref_left, ref_pi, ref_pf = reference.lowlevel(ch, *seg[:3], view, remaining, st.position_integer, st.position_fractional, seg[3])
        got, ret, left = crb.LowLevel_Resample(st, pre, view, remaining, seg[3])
        assert np.array_equal(got, ref_out.astype(np.int64)), seg
        assert (ret, left, st.position_integer, st.position_fractional) == (ref_ret, ref_left, ref_pi, ref_pf), seg
        consumed += remaining - left
        out_frames.append(got)
    assert sum(len(x) for x in out_frames) > 19000


def test_voice_batch_pitch_bend_against_live_reference(pre, reference):
    """SURVEY.md 8f rank 2 in the batched front end: voices whose ratio changes mid-stream (VoiceBatchAdjust) must emit
    what the unmodified reference's HighLevel_Resample / HighLevel_Adjust / HighLevel_ResampleEnd sequence emits when
    the ratio is switched at the same output frames.  All ratios here are up-sampling (one kernel geometry)."""
    ch = 1
    rng = np.random.default_rng(31)
    voices = 9
    tick = 512
    plans = []
    for v in range(voices):
        n_seg = int(rng.integers(1, 5))
        segs = [(22050, int(rng.choice([24000, 32000, 44100, 48000, 96000])), 384000) for _ in range(n_seg)]
        switch = sorted(set(int(x) * tick for x in rng.integers(1, 12, size=n_seg - 1)))     # switches happen between ticks
        segs = segs[: len(switch) + 1]
        data = rng.integers(-32768, 32768, size=(int(rng.integers(500, 6000)), ch), dtype=np.int16)
        want = reference.highlevel_adjust(ch, segs, switch, data, 200000)
        plans.append((segs, switch, data, want))
    vb = crb.VoiceBatch(pre, voices, ch, *plans[0][0][0])
    for v, (segs, switch, data, want) in enumerate(plans):
        vb.adjust(v, *segs[0])
        vb.push(v, data)
        vb.end(v)
    emitted = [0] * voices
    seg_index = [0] * voices
    got = [[] for _ in range(voices)]
    for _ in range(400):
        for v, (segs, switch, data, want) in enumerate(plans):
            if seg_index[v] < len(switch) and emitted[v] == switch[seg_index[v]]:
                seg_index[v] += 1
                vb.adjust(v, *segs[seg_index[v]])
        out, produced = vb.tick(tick)
        for v in range(voices):
            got[v].append(out[v, :produced[v]].copy())
            emitted[v] += int(produced[v])
        if produced.sum() == 0:
            break
    for v, (segs, switch, data, want) in enumerate(plans):
        g = np.concatenate(got[v])
        assert g.shape == want.shape, (v, segs, switch, g.shape, want.shape)
        assert np.array_equal(g, want), (v, segs, switch)
    # a ratio whose kernel is wider than the one the voices were created with is refused, as H:1195 refuses it
    with pytest.raises(crb.Error, match="kernel radius beyond"):
        vb.adjust(0, 48000, 22050, 22050)
    vb.destroy()


def test_voice_batch_adjust_across_kernel_geometries(pre, reference):
    """SURVEY.md 8f rank 2, the remainder: voices of ONE batch running different kernel geometries after HighLevel_Adjust -- down-sampling
    by different amounts, up-sampling, back again -- grouped by plan tick by tick, against the unmodified reference's
    HighLevel_Resample / HighLevel_Adjust / HighLevel_ResampleEnd sequence.  Created with a wide kernel (R = 7), so that every
    later radius is allowed (H:1195)."""
    ch = 2
    rng = np.random.default_rng(47)
    voices, tick = 7, 256
    create = (48000, 22050, 22050)                       # R = 7
    choices = [(48000, 22050, 22050), (48000, 32000, 32000), (44100, 48000, 48000), (48000, 24000, 24000), (22050, 48000, 48000), (48000, 44100, 30000)]
    plans = []
    for v in range(voices):
        n_seg = int(rng.integers(2, 6))
        segs = [create] + [choices[int(rng.integers(0, len(choices)))] for _ in range(n_seg - 1)]
        switch = sorted(set(int(x) * tick for x in rng.integers(1, 10, size=n_seg - 1)))
        segs = segs[: len(switch) + 1]
        data = rng.integers(-32768, 32768, size=(int(rng.integers(3000, 9000)), ch), dtype=np.int16)
        want = reference.highlevel_adjust(ch, segs, switch, data, 200000)
        plans.append((segs, switch, data, want))
    vb = crb.VoiceBatch(pre, voices, ch, *create)
    for v, (segs, switch, data, want) in enumerate(plans):
        vb.push(v, data)
        vb.end(v)
    emitted, seg_index, got = [0] * voices, [0] * voices, [[] for _ in range(voices)]
    for _ in range(400):
        for v, (segs, switch, data, want) in enumerate(plans):
            if seg_index[v] < len(switch) and emitted[v] == switch[seg_index[v]]:
                seg_index[v] += 1
                vb.adjust(v, *segs[seg_index[v]])
        out, produced = vb.tick(tick)
        for v in range(voices):
            got[v].append(out[v, :produced[v]].copy())
            emitted[v] += int(produced[v])
        if produced.sum() == 0:
            break
    for v, (segs, switch, data, want) in enumerate(plans):
        g = np.concatenate(got[v])
        assert g.shape == want.shape, (v, segs, switch, g.shape, want.shape)
        assert np.array_equal(g, want), (v, segs, switch)
    vb.destroy()


def test_voice_batch_lockstep_voices_and_split_tick(pre, oracle):
    """Many voices started together (the 1024-voice shape of BASELINE configs[3] in small): they walk through the same phases, so
    the batch merges them four at a time into lockstep jobs; the tick is driven through TickBegin / TickEnd with the next tick's
    input pushed in between.  Every voice against the oracle's HighLevel stream."""
    ch, i, o = 1, 22050, 48000
    voices, tick, T = 23, 1024, 9000                     # 23: lockstep jobs of 4, 2 and 1
    data = [oracle.noise(60 + v, 0, 0, T, ch) for v in range(voices)]
    want = [oracle.highlevel(ch, i, o, o, d) for d in data]
    vb = crb.VoiceBatch(pre, voices, ch, i, o, o)
    L = crb.lib()
    per_tick = tick * i // o + 2
    pos = 0
    out = np.zeros((voices, tick, ch), dtype=np.int32)
    produced = (C.c_size_t * voices)()
    got = [[] for _ in range(voices)]

    def push(upto):
        nonlocal pos
        n = min(upto, T) - pos
        if n > 0:
            for v in range(voices):
                vb.push(v, data[v][pos:pos + n])
            pos += n
            if pos == T:
                for v in range(voices):
                    vb.end(v)
    push(per_tick)
    for _ in range(200):
        assert L.ClownResamplerB200_VoiceBatchTickBegin(vb.handle, tick, crb.OUT_S32, out.ctypes.data, out.strides[0], produced) == 0, crb.last_error()
        push(pos + per_tick)                              # the next tick's input arrives while the GPU works
        assert L.ClownResamplerB200_VoiceBatchTickEnd(vb.handle) == 0, crb.last_error()
        n = np.ctypeslib.as_array(produced).astype(np.int64)
        for v in range(voices):
            got[v].append(out[v, :n[v]].copy())
        if n.sum() == 0 and pos == T:
            break
    for v in range(voices):
        assert np.array_equal(np.concatenate(got[v]), want[v]), v
    vb.destroy()
