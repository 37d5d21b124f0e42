"""CPU-only checks of the host layer of libclownresampler_b200.so: exported symbols, ABI, the
configuration/ratio code, closed-form state updates, plan construction, and (through
tests/device_model.py) the integer identities the CUDA kernel relies on, all against the oracle.
No device call is made here."""
import ctypes as C
import json
import os
import subprocess

import numpy as np
import pytest

import clownresampler_b200 as crb
from conftest import GOLD, ROOT, pad
from device_model import resample as model_resample


@pytest.fixture(scope="module")
def pre():
    return crb.Precompute()


def test_library_exports_every_declared_symbol():
    L = crb.lib()
    for name in crb.DROPIN_SYMBOLS + crb.EXTENSION_SYMBOLS:
        assert hasattr(L, name), name
    # and the headers declare exactly these
    for header, names in (("clownresampler.h", crb.DROPIN_SYMBOLS), ("clownresampler_b200.h", crb.EXTENSION_SYMBOLS)):
        text = open(os.path.join(ROOT, "include", header)).read()
        for name in names:
            assert name + "(" in text, (header, name)


def test_headers_compile_as_c89_and_match_reference_abi(tmp_path, tripwires):
    src = tmp_path / "abi.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#define CLOWNRESAMPLER_IMPLEMENTATION\n#define CLOWNRESAMPLER_STATIC\n'
                   '#include "clownresampler_b200.h"\nint main(void){printf("%lu %lu %lu %lu %lu %lu %lu %lu %lu %lu %lu %lu %lu %lu %lu %lu\\n",'
                   '(unsigned long)sizeof(cc_s16l),(unsigned long)sizeof(cc_s32l),(unsigned long)sizeof(cc_s32f),(unsigned long)sizeof(cc_u32f),'
                   '(unsigned long)sizeof(cc_u8f),(unsigned long)sizeof(cc_bool),(unsigned long)sizeof(size_t),'
                   '(unsigned long)sizeof(ClownResampler_Precomputed),(unsigned long)sizeof(ClownResampler_LowestLevel_Configuration),'
                   '(unsigned long)sizeof(ClownResampler_LowLevel_State),(unsigned long)sizeof(ClownResampler_HighLevel_State),'
                   '(unsigned long)offsetof(ClownResampler_LowLevel_State,channels),(unsigned long)offsetof(ClownResampler_LowLevel_State,position_integer),'
                   '(unsigned long)offsetof(ClownResampler_LowLevel_State,position_fractional),(unsigned long)offsetof(ClownResampler_LowLevel_State,increment),'
                   '(unsigned long)offsetof(ClownResampler_HighLevel_State,input_buffer));return 0;}\n')
    exe = tmp_path / "abi"
    subprocess.check_call(["gcc", "-std=c89", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert got == tripwires["abi"]       # sizes/offsets printed by the reference build (oracle/ref_shim.c: ref_abi)


def test_precompute_is_bit_identical(pre, oracle):
    assert np.array_equal(crb.table_of(pre), oracle.table.astype(np.int64))


def test_ratio_and_configure_grid(pre):
    g = json.load(open(os.path.join(GOLD, "ref_ratio_config.json")))
    L = crb.lib()
    for a, b, l, want in g["configure"]:
        cfg = crb.ClownResampler_LowestLevel_Configuration()
        ok = L.ClownResampler_LowestLevel_Configure(C.byref(cfg), a, b, l)
        got = [cfg.stretched_kernel_radius, cfg.integer_stretched_kernel_radius, cfg.stretched_kernel_radius_delta, cfg.kernel_step_size] if ok else []
        assert got == want, (a, b, l)
    for a, b, want in g["ratio"]:
        st = crb.ClownResampler_LowLevel_State()
        L.ClownResampler_LowLevel_Adjust(C.byref(st), a, b, b)
        assert st.increment == want, (a, b)


def test_closed_form_state_matches_reference_vectors(ref_vectors):
    """CountOutputFrames / AdvanceState against what the reference's loop left behind."""
    meta, _ = ref_vectors
    L = crb.lib()
    for m in meta:
        if m["kind"] != "lowlevel":
            continue
        st = crb.LowLevel_Init(m["channels"], m["in"], m["out"], m["lpf"])
        st.position_integer, st.position_fractional = m["pos_int"], m["pos_frac"]
        n = crb.CountOutputFrames(st, m["T"])
        stopped = bool(m["limit"]) and m["limit"] <= n
        emitted = m["limit"] if stopped else n
        assert emitted == m["frames"], m
        total = C.c_size_t(m["T"])
        L.ClownResamplerB200_AdvanceState(C.byref(st), C.byref(total), emitted, int(stopped))
        assert (total.value, st.position_integer, st.position_fractional) == (m["remaining"], m["end_pos_int"], m["end_pos_frac"]), m
        assert (0 if stopped else 1) == m["ret"]


PLAN_CASES = [
    (2, 44100, 48000, 48000), (1, 22050, 48000, 48000), (8, 192000, 44100, 44100), (2, 8000, 44100, 44100),
    (2, 44100, 8000, 44100), (1, 384000, 8000, 8000), (1, 8000, 384000, 384000), (3, 48000, 44100, 44100),
    (16, 96000, 48000, 48000), (2, 48000, 48000, 48000), (1, 3, 2, 2), (2, 44100, 48000, 10000), (5, 7, 1000, 1000),
    (2, 384000, 48000, 48000), (8, 192000, 48000, 48000),   # integer ratios: rotated column layout
    (2, 48000, 44100, 44100), (1, 48000, 32000, 32000), (4, 48000, 16000, 16000), (11, 48000, 44100, 44100),   # stretched kernels: taps whose sign depends on the phase
    (4, 96000, 44100, 44100), (6, 192000, 44100, 44100), (8, 384000, 44100, 44100), (4, 44100, 16000, 16000),   # chain form of the general kernel
]


def check_chain_layout(geo, rows):
    """Chain form of the general kernel: plain weights; a chain group's |k| sum to at most 65535 in every phase row; single
    columns stay at or below 65536 (65535 in the signed class); signed columns only where the sign really changes."""
    assert geo["n_runs"] == 0 and 1 <= geo["n_groups"] <= 12
    seen = 0
    for gi in range(geo["n_groups"]):
        first, count, rotates = geo["groups"][gi]
        cls, single = geo["group_kinds"][gi] >> 1, geo["group_kinds"][gi] & 1
        assert count >= 2 and count % 2 == 0 and cls <= 2
        cols = rows[:, first: first + count].astype(np.int64)
        if cls != 2:
            assert (cols >= 0).all()
        else:
            real = [c for c in range(count) if cols[:, c].any()]
            assert all((cols[:, c] > 0).any() and (cols[:, c] < 0).any() for c in real)
        if single:
            assert np.abs(cols).max() <= (65535 if cls == 2 else 65536)
        else:
            assert np.abs(cols).sum(axis=1).max() <= 65535
            assert sum(1 for c in range(count) if cols[:, c].any()) >= 2      # a chain of one is a single column
        seen += count
    assert seen <= geo["n_cols"]


@pytest.mark.parametrize("case", PLAN_CASES)
def test_plan_and_device_arithmetic_model_match_oracle(pre, oracle, case):
    ch, i, o, l = case
    st = crb.LowLevel_Init(ch, i, o, l)
    geo, rows = crb.debug_plan_host(pre, st)
    cfg = oracle.configure(i, o, l)
    assert (geo["radius_fx"], geo["radius_int"], geo["delta"], geo["step"]) == cfg
    assert geo["increment"] == oracle.ratio(i, o)
    if geo["chain_mode"]:
        check_chain_layout(geo, rows)
    elif not geo["unstretched5"]:
        signed_cols = [c for (col, length, off, neg, big) in geo["runs"] if neg == 2 for c in range(col, col + length)]
        unsigned_cols = [c for (col, length, off, neg, big) in geo["runs"] if neg != 2 for c in range(col, col + length)]
        assert (rows[:, unsigned_cols] >= 0).all()          # |k| columns, sign carried by the run
        for c in signed_cols if not geo["small_taps"] else []:   # general kernel: signed columns only where the sign really changes between rows
            assert (rows[:, c] > 0).any() and (rows[:, c] < 0).any()
    rng = np.random.default_rng(ch * 1000 + i % 997)
    R = cfg[1]
    T = max(4, min(700, 1500 * geo["increment"] // 65536))
    data = rng.integers(-32768, 32768, size=(T, ch), dtype=np.int16)
    data[: T // 4] = np.where(rng.random((T // 4, ch)) < 0.5, -32768, 32767)   # full-scale stretch
    data[T // 4: T // 2] = rng.integers(-2, 3, size=(T // 2 - T // 4, ch))     # tiny values: truncation toward zero
    padded = pad(data, R)
    p_int, p_frac = int(rng.integers(0, 3)), int(rng.integers(0, 65536))
    want = oracle.lowlevel(ch, i, o, l, padded, T, p_int, p_frac)[0]
    q0 = (p_int << 16) + p_frac + geo["delta"]
    got = model_resample(geo, rows, padded, q0, 0, want.shape[0])
    assert np.array_equal(got, want.astype(np.int64))
    raw = model_resample(geo, rows, padded, q0, 0, min(64, want.shape[0]), fmt=2)
    unnorm = oracle.lowlevel(ch, i, o, l, padded, T, p_int, p_frac, norm=2)[0][: raw.shape[0]]
    assert np.array_equal(raw[:, :ch], unnorm.astype(np.int64))


def test_plan_shapes_for_the_baseline_configs(pre):
    """tap counts and phase counts of the named configurations."""
    geo, rows = crb.debug_plan_host(pre, crb.LowLevel_Init(2, 44100, 48000, 48000))
    assert geo["unstretched5"] == 1 and geo["n_rows"] == 1024 and geo["n_cols"] == 5 and geo["n_breaks"] == 0 and geo["kernel_kind"] == 0
    assert geo["row_words"] == 4 and rows.shape == (1024, 4)            # 16-byte packed rows: one LDS.128 per frame
    assert [(r[3], r[4]) for r in geo["runs"]] == [(0, 0), (1, 0), (0, 1), (1, 0)]   # (negative, big) per run
    geo, rows = crb.debug_plan_host(pre, crb.LowLevel_Init(8, 192000, 44100, 44100))
    assert geo["radius_int"] == 14 and geo["delta"] == 61526 and geo["step"] == 235 and geo["taps_max"] == 26
    assert geo["chain_mode"] == 1 and sum(np.count_nonzero(np.abs(rows[:, f: f + n]).max(axis=0)) for f, n, _ in geo["groups"][: geo["n_groups"]]) == 26   # one column per tap: mixed-sign taps are signed columns, not two
    geo, rows = crb.debug_plan_host(pre, crb.LowLevel_Init(2, 48000, 44100, 44100))
    assert geo["taps_max"] == 6 and geo["small_taps"] == 6 and geo["row_words"] == 8 and geo["runs"] == [(0, 6, 0, 2, 1)]   # slightly stretched kernel
    geo, rows = crb.debug_plan_host(pre, crb.LowLevel_Init(1, 48000, 32000, 32000))
    assert geo["taps_max"] == 9 and geo["small_taps"] == 10 and geo["row_words"] == 12 and (rows[:, 9] == 0).all()
    geo, rows = crb.debug_plan_host(pre, crb.LowLevel_Init(12, 48000, 44100, 44100))
    assert geo["small_taps"] == 0 and geo["chain_mode"] == 0 and any(r[3] == 2 for r in geo["runs"]) and sum(r[1] for r in geo["runs"]) == 6       # general kernel (IMAD.HI form: 12 channels), signed columns
    assert geo["kernel_kind"] == 0 and geo["norm_mode"] >= 1
    geo, rows = crb.debug_plan_host(pre, crb.LowLevel_Init(1, 384000, 8000, 8000))
    assert geo["radius_int"] == 144 and geo["step"] == 21 and geo["taps_max"] == 288


@pytest.mark.parametrize("case", [(2, 384000, 48000), (1, 384000, 8000), (2, 384000, 8000), (8, 192000, 48000), (4, 192000, 48000), (1, 384000, 48000)])
def test_column_rotation_layout(pre, case):
    """Integer down-sampling ratios put every lane's frame a multiple of the bank count apart; the plan then rotates
    the column order per lane.  A rotating group must be followed by a faithful copy of its first 2 * rot_mask columns
    (weights of every row and frame offsets), and every group must stay inside the row."""
    ch, i, o = case
    geo, rows = crb.debug_plan_host(pre, crb.LowLevel_Init(ch, i, o, o))
    assert geo["rot"] >= 1 and geo["rot"] % 2 == 1 and geo["rot_mask"] in (1, 3, 7, 15, 31), geo
    offs = np.array(geo["col_offsets"])
    assert offs.shape[0] == geo["n_cols"] and geo["row_words"] > geo["n_cols"]
    end = 0
    for first, count, rotates in geo["groups"]:
        assert first >= end and count % 2 == 0
        extra = 2 * geo["rot_mask"] if rotates else 0
        if rotates:
            assert count // 2 > geo["rot_mask"]
            assert np.array_equal(rows[:, first + count: first + count + extra], rows[:, first: first + extra])
            assert np.array_equal(offs[first + count: first + count + extra], offs[first: first + extra])
        end = first + count + extra
    assert end == geo["n_cols"]
    assert any(r for _, _, r in geo["groups"])
    # the largest start pair of any lane stays inside the copy
    assert max(((geo["rot"] * l) >> geo["rot_shift"]) & geo["rot_mask"] for l in range(32)) <= geo["rot_mask"]


def test_plan_rejects_what_the_reference_cannot_run(pre):
    st = crb.LowLevel_Init(1, 3000000, 1000, 1000)     # kernel_step_size == 0 -> the reference divides by zero
    assert st is not None and st.lowest_level.kernel_step_size == 0
    with pytest.raises(crb.Error, match="kernel_step_size is 0"):
        crb.debug_plan_host(pre, st)
    st = crb.LowLevel_Init(17, 44100, 48000, 48000)    # LowLevel_Init itself does not check (H:1044-1050) ...
    with pytest.raises(crb.Error, match="channels"):   # ... the plan does, instead of overrunning 16 accumulators (H:1071)
        crb.debug_plan_host(pre, st)
    assert crb.LowLevel_Init(1, 1 << 29, 1, 1) is None  # scale >= 0x1000: Configure returns cc_false (H:974)


def _has_cuda_device():
    return crb.lib().ClownResamplerB200_DeviceCount() > 0


@pytest.mark.skipif(_has_cuda_device(), reason="only meaningful on a machine without a CUDA device")
def test_no_device_means_loud_failure_not_a_cpu_fallback(pre, capfd):
    """Without a usable sm_100 device nothing is computed anywhere else: plans cannot be created, the bulk entry
    points return an error, and the drop-in call emits no frames, says so on stderr and leaves an error message."""
    st = crb.LowLevel_Init(2, 44100, 48000, 48000)
    with pytest.raises(crb.Error, match="no usable CUDA device|no CPU fallback"):
        crb.Plan(pre, st)
    assert crb.lib().ClownResamplerB200_Init(0) != 0
    padded = np.zeros((1000 + 6, 2), dtype=np.int16)
    out, ret, remaining = crb.LowLevel_Resample(st, pre, padded, 1000)
    # no frames, cc_false ("stopped"), and the input that produced no output is NOT consumed: the failure is visible
    assert out.shape[0] == 0 and ret == 0 and remaining == 1000
    assert st.position_integer == 0 and st.position_fractional == 0
    assert "no usable CUDA device" in crb.last_error()
    assert "clownresampler_b200" in capfd.readouterr().err
