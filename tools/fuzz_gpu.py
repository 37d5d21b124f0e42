#!/usr/bin/env python
"""Randomised parity run on the device: arbitrary rates (not only the audio standards), low-pass rates, channel counts
and start states, bulk path against the oracle, bit-exact.  Not part of the test suite (run time); prints one line per
kernel kind and fails loudly on the first mismatch.   usage: python tools/fuzz_gpu.py [cases] [seed]"""
import collections
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import clownresampler_b200 as crb  # noqa: E402
from cro import Oracle  # noqa: E402

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 200
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(seed)
assert crb.lib().ClownResamplerB200_Init(0) == 0, crb.last_error()
pre = crb.Precompute()
oracle = Oracle()
std = [8000, 11025, 12000, 16000, 22050, 24000, 32000, 44100, 48000, 88200, 96000, 176400, 192000, 384000]
kinds = collections.Counter()
done = 0
while done < cases:
    if rng.random() < 0.5:
        i, o = int(rng.choice(std)), int(rng.choice(std))
    else:
        i, o = int(rng.integers(1000, 400000)), int(rng.integers(1000, 400000))
    ch = int(rng.choice([1, 1, 2, 2, 2, 3, 4, 5, 6, 8, 8, 16]))
    l = int(rng.choice([min(i, o), o, i, max(1000, min(i, o) // 2), int(rng.integers(1000, 400000))]))
    st = crb.LowLevel_Init(ch, i, o, l)
    try:
        geo, _ = crb.debug_plan_host(pre, st)
    except crb.Error:
        continue                                  # configurations the reference cannot run either
    R = oracle.configure(i, o, l)[1]
    inc = oracle.ratio(i, o)
    n_out = int(rng.integers(1, 3)) * geo["tile_out"] + int(rng.integers(1, 999))     # a few tiles and a ragged one
    T = max(4, min(n_out * inc // 65536, 2_000_000 // ch))
    data = rng.integers(-32768, 32768, size=(T, ch), dtype=np.int16)
    data[: T // 8] = np.where(rng.random((T // 8, ch)) < 0.5, -32768, 32767)
    padded = np.concatenate([np.zeros((R, ch), np.int16), data, np.zeros((R, ch), np.int16)])
    pi, pf = int(rng.integers(0, 3)), int(rng.integers(0, 65536))
    want = oracle.lowlevel(ch, i, o, l, padded, T, pi, pf)[0]
    st.position_integer, st.position_fractional = pi, pf
    fmt = crb.OUT_S16_CLAMPED if rng.random() < 0.5 else crb.OUT_S32
    got = crb.resample_array(pre, st, padded, T, fmt=fmt)
    ref = np.clip(want, -0x7FFF, 0x7FFF).astype(np.int16) if fmt == crb.OUT_S16_CLAMPED else want
    kind = "direct" if geo["kernel_kind"] else "unstretched" if geo["unstretched5"] else f"small{geo['small_taps']}" if geo["small_taps"] \
        else "general+rot" if geo["rot"] else "general"
    if not np.array_equal(got, ref):
        bad = np.argwhere(got != ref)[0]
        print(f"MISMATCH ch={ch} in={i} out={o} lpf={l} T={T} pos=({pi},{pf}) fmt={fmt} kind={kind} first at frame {bad[0]} ch {bad[1]}: {got[tuple(bad)]} != {ref[tuple(bad)]}")
        print({k: geo[k] for k in geo if k not in ("runs", "col_offsets")})
        sys.exit(1)
    kinds[kind] += 1
    done += 1
print(f"{done} cases bit-exact against the oracle:", dict(kinds))

# drop-in callback path, tick style: random tick sizes, the caller comes back with the advanced buffer; now and then it
# changes its pending input between calls (the kept frames must then not be used)
ticks_done = 0
for case in range(max(4, cases // 20)):
    i, o = int(rng.choice(std)), int(rng.choice(std))
    ch = int(rng.choice([1, 2, 3, 6]))
    st = crb.LowLevel_Init(ch, i, o, min(i, o))
    R = oracle.configure(i, o, min(i, o))[1]
    inc = oracle.ratio(i, o)
    T = int(max(64, min(40000, 60000 * inc // 65536)))
    data = rng.integers(-32768, 32768, size=(T, ch), dtype=np.int16)
    padded = np.concatenate([np.zeros((R, ch), np.int16), data, np.zeros((R, ch), np.int16)])
    offset, remaining, got = 0, T, []
    while True:
        tick = int(rng.choice([1, 7, 100, 512, 1024, 3000, 6000]))
        if rng.random() < 0.15 and remaining > 2 * R + 8:
            padded = padded.copy()
            padded[offset + 2 * R + 4:] = padded[offset + 2 * R + 4:] ^ 0x55      # the caller rewrites input it has not consumed yet
        exp = oracle.lowlevel(ch, i, o, min(i, o), padded[offset:], remaining, st.position_integer, st.position_fractional, max_frames=tick)[0]
        out, ret, left = crb.LowLevel_Resample(st, pre, padded[offset:], remaining, tick)
        if not np.array_equal(out, exp.astype(np.int64)):
            print(f"DROP-IN MISMATCH ch={ch} in={i} out={o} tick={tick} offset={offset}")
            sys.exit(1)
        offset += remaining - left
        remaining = left
        ticks_done += 1
        if ret:
            break
print(f"{ticks_done} tick-style drop-in calls bit-exact against the oracle; counters (launches, calls served from kept frames): {crb.counters()}")
