#!/usr/bin/env python
"""Small mixed workload for compute-sanitizer (memcheck / racecheck): a few configurations through the tiled and
direct kernels, bulk and callback paths.  usage: compute-sanitizer --tool memcheck python tools/sanitize_run.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import clownresampler_b200 as crb
from oracle.cro import Oracle

o = Oracle()
assert crb.lib().ClownResamplerB200_Init(0) == 0
pre = crb.Precompute()
for ch, i, r in [(2, 44100, 48000), (8, 192000, 44100), (1, 22050, 48000), (3, 48000, 32000), (1, 384000, 8000),
                 (2, 48000, 44100), (6, 48000, 44100), (2, 384000, 48000), (8, 192000, 48000), (5, 8000, 48000), (12, 96000, 48000),
                 (4, 96000, 44100), (6, 192000, 44100)]:
    st = crb.LowLevel_Init(ch, i, r, r)
    R = st.lowest_level.integer_stretched_kernel_radius
    T = 20011
    data = o.noise(1, 0, 0, T, ch)
    padded = np.zeros((T + 2 * R, ch), dtype=np.int16); padded[R:R + T] = data
    want = o.lowlevel(ch, i, r, r, padded, T)[0]
    got = crb.resample_array(pre, st, padded, T)
    assert np.array_equal(got, want), (ch, i, r)
    got16 = crb.resample_array(pre, st, padded, T, fmt=crb.OUT_S16_CLAMPED)     # packed stores of the s16 paths
    assert np.array_equal(got16, np.clip(want, -0x7FFF, 0x7FFF).astype(np.int16)), (ch, i, r)
    st2 = crb.LowLevel_Init(ch, i, r, r)
    out, ret, left = crb.LowLevel_Resample(st2, pre, padded[:3000 + 2 * R], 3000, 500)
    ref = o.lowlevel(ch, i, r, r, padded[:3000 + 2 * R], 3000, max_frames=500)[0]
    assert np.array_equal(out, ref.astype(np.int64))
print("sanitize workload ok")
