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
