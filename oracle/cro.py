"""ctypes front-end for the parity checker.  TEST INFRASTRUCTURE ONLY.

`Oracle` wraps oracle/_build/libcr_oracle.so (the C restatement, built on demand with gcc);
`Reference` wraps oracle/_ref/libclownref.so (the unmodified reference compiled in place from
/root/reference by oracle/Makefile; prebuilt copy on the GPU box).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
TABLE_SIZE = 6144
NORM_CURRENT, NORM_LEGACY, NORM_NONE = 0, 1, 2

_u64p = C.POINTER(C.c_uint64)


def _ptr(a, ty):
    return a.ctypes.data_as(C.POINTER(ty)) if a is not None else None


def build_oracle() -> str:
    path = os.path.join(HERE, "_build", "libcr_oracle.so")
    src = [os.path.join(HERE, "cr_oracle.c"), os.path.join(HERE, "cr_oracle.h")]
    if not os.path.exists(path) or any(os.path.getmtime(s) > os.path.getmtime(path) for s in src):
        subprocess.check_call(["make", "-C", HERE, "oracle"], stdout=subprocess.DEVNULL)
    return path


def build_ref(reference_dir: str = "/root/reference") -> str | None:
    """Builds oracle/_ref from the reference sources where they lie (build container only)."""
    path = os.path.join(HERE, "_ref", "libclownref.so")
    if os.path.isdir(reference_dir):
        subprocess.check_call(["make", "-C", HERE, "ref", f"REFERENCE_DIR={reference_dir}"], stdout=subprocess.DEVNULL)
    return path if os.path.exists(path) else None


class CroConfig(C.Structure):
    _fields_ = [("radius_fx", C.c_uint64), ("radius_int", C.c_uint64), ("radius_delta", C.c_uint64), ("step", C.c_uint64)]


class CroState(C.Structure):
    _fields_ = [("cfg", CroConfig), ("channels", C.c_uint32), ("pos_int", C.c_uint64), ("pos_frac", C.c_uint64), ("increment", C.c_uint64)]


class Oracle:
    def __init__(self):
        L = self.lib = C.CDLL(build_oracle())
        L.cro_ratio.restype = C.c_uint64
        L.cro_ratio.argtypes = [C.c_uint64, C.c_uint64]
        L.cro_configure.argtypes = [C.POINTER(CroConfig), C.c_uint64, C.c_uint64, C.c_uint64]
        L.cro_init.argtypes = [C.POINTER(CroState), C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint64]
        L.cro_lowlevel_resample.argtypes = [C.POINTER(CroState), C.POINTER(C.c_int32), C.POINTER(C.c_int16), _u64p,
                                            C.POINTER(C.c_int32), C.c_uint64, _u64p, C.c_int, C.c_uint64]
        L.cro_highlevel_stream.restype = C.c_uint64
        L.cro_highlevel_stream.argtypes = [C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint64, C.POINTER(C.c_int32),
                                           C.POINTER(C.c_int16), C.c_uint64, C.c_uint64, C.POINTER(C.c_int32), C.c_uint64]
        L.cro_count_output_frames.restype = C.c_uint64
        L.cro_count_output_frames.argtypes = [C.c_uint64] * 4
        L.cro_fill_noise.argtypes = [C.POINTER(C.c_int16), C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint32]
        self.table = np.zeros(TABLE_SIZE, dtype=np.int32)
        L.cro_precompute(_ptr(self.table, C.c_int32))

    def ratio(self, a, b):
        return int(self.lib.cro_ratio(a, b))

    def configure(self, in_rate, out_rate, lpf):
        cfg = CroConfig()
        ok = self.lib.cro_configure(C.byref(cfg), in_rate, out_rate, lpf)
        return (cfg.radius_fx, cfg.radius_int, cfg.radius_delta, cfg.step) if ok else None

    def state(self, channels, in_rate, out_rate, lpf, pos_int=0, pos_frac=0):
        st = CroState()
        if not self.lib.cro_init(C.byref(st), channels, in_rate, out_rate, lpf):
            return None
        st.pos_int, st.pos_frac = pos_int, pos_frac
        return st

    def count(self, pos_int, pos_frac, increment, total):
        return int(self.lib.cro_count_output_frames(pos_int, pos_frac, increment, total))

    def lowlevel(self, channels, in_rate, out_rate, lpf, padded, total_frames, pos_int=0, pos_frac=0,
                 max_frames=0, norm=NORM_CURRENT, legacy_scale=0, table=None):
        """Returns (out[int32 frames x channels], ret, remaining_input_frames, pos_int, pos_frac)."""
        st = self.state(channels, in_rate, out_rate, lpf, pos_int, pos_frac)
        if st is None:
            raise ValueError("configuration rejected")
        padded = np.ascontiguousarray(padded, dtype=np.int16)
        n_max = self.count(pos_int, pos_frac, st.increment, total_frames)
        if max_frames:
            n_max = min(n_max, max_frames)
        out = np.zeros((max(n_max, 1), channels), dtype=np.int32)
        total = C.c_uint64(total_frames)
        wrote = C.c_uint64(0)
        tab = self.table if table is None else np.ascontiguousarray(table, dtype=np.int32)
        ret = self.lib.cro_lowlevel_resample(C.byref(st), _ptr(tab, C.c_int32), _ptr(padded, C.c_int16), C.byref(total),
                                             _ptr(out, C.c_int32), max_frames, C.byref(wrote), norm, legacy_scale)
        return out[: wrote.value], int(ret), int(total.value), int(st.pos_int), int(st.pos_frac)

    def highlevel(self, channels, in_rate, out_rate, lpf, data, chunk=0, capacity=None):
        data = np.ascontiguousarray(data, dtype=np.int16).reshape(-1, channels)
        n = data.shape[0]
        if capacity is None:
            inc = self.ratio(in_rate, out_rate)
            capacity = ((n + 2 * 4096) * 65536) // max(inc, 1) + 16
        out = np.zeros((capacity, channels), dtype=np.int32)
        wrote = self.lib.cro_highlevel_stream(channels, in_rate, out_rate, lpf, _ptr(self.table, C.c_int32),
                                              _ptr(data, C.c_int16), n, chunk, _ptr(out, C.c_int32), capacity)
        if wrote == 2**64 - 1:
            raise ValueError("configuration rejected")
        return out[:wrote]

    def noise(self, seed, stream, first_frame, n_frames, channels):
        a = np.zeros((n_frames, channels), dtype=np.int16)
        self.lib.cro_fill_noise(_ptr(a, C.c_int16), seed, stream, first_frame, n_frames, channels)
        return a


class Reference:
    """The unmodified reference, through oracle/ref_shim.c."""

    def __init__(self, o3: bool = False):
        path = os.path.join(HERE, "_ref", "libclownref_o3.so" if o3 else "libclownref.so")
        if not os.path.exists(path):
            build_ref()
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        L = self.lib = C.CDLL(path)
        L.ref_ratio.restype = C.c_uint64
        L.ref_ratio.argtypes = [C.c_uint64, C.c_uint64]
        L.ref_configure.argtypes = [_u64p, C.c_uint64, C.c_uint64, C.c_uint64]
        L.ref_lowlevel_bulk.argtypes = [C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint64, C.POINTER(C.c_int16), _u64p, _u64p,
                                        C.POINTER(C.c_int32), C.c_uint64, _u64p]
        L.ref_highlevel_stream.restype = C.c_uint64
        L.ref_highlevel_stream.argtypes = [C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint64, C.POINTER(C.c_int16), C.c_uint64,
                                           C.c_uint64, C.POINTER(C.c_int32), C.c_uint64]
        L.ref_lowlevel_adjust_sequence.restype = C.c_uint64
        L.ref_lowlevel_adjust_sequence.argtypes = [C.c_uint32, _u64p, _u64p, C.c_uint32, C.POINTER(C.c_int16), C.c_uint64, C.POINTER(C.c_int32), _u64p]
        L.ref_highlevel_adjust_stream.restype = C.c_uint64
        L.ref_highlevel_adjust_stream.argtypes = [C.c_uint32, _u64p, _u64p, C.c_uint32, C.POINTER(C.c_int16), C.c_uint64, C.POINTER(C.c_int32), C.c_uint64]
        L.ref_time_lowlevel.restype = C.c_double
        L.ref_time_lowlevel.argtypes = [C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint64, C.POINTER(C.c_int16), C.c_uint64,
                                        C.POINTER(C.c_int16), _u64p]
        self.table = np.zeros(TABLE_SIZE, dtype=np.int32)
        L.ref_table(_ptr(self.table, C.c_int32))

    def abi(self):
        a = (C.c_uint64 * 16)()
        self.lib.ref_abi(a)
        return list(a)

    def ratio(self, a, b):
        return int(self.lib.ref_ratio(a, b))

    def configure(self, in_rate, out_rate, lpf):
        a = (C.c_uint64 * 4)()
        return tuple(a) if self.lib.ref_configure(a, in_rate, out_rate, lpf) else None

    def lowlevel(self, channels, in_rate, out_rate, lpf, padded, total_frames, pos_int=0, pos_frac=0, max_frames=0, capacity=None):
        padded = np.ascontiguousarray(padded, dtype=np.int16)
        if capacity is None:
            inc = self.ratio(in_rate, out_rate)
            capacity = (total_frames * 65536) // max(inc, 1) + 2
            if max_frames:
                capacity = min(capacity, max_frames)
        out = np.zeros((max(capacity, 1), channels), dtype=np.int32)
        total = C.c_uint64(total_frames)
        st = (C.c_uint64 * 2)(pos_int, pos_frac)
        wrote = C.c_uint64(0)
        ret = self.lib.ref_lowlevel_bulk(channels, in_rate, out_rate, lpf, _ptr(padded, C.c_int16), C.byref(total), st,
                                         _ptr(out, C.c_int32), max_frames, C.byref(wrote))
        if ret < 0:
            raise ValueError("configuration rejected")
        return out[: wrote.value], int(ret), int(total.value), int(st[0]), int(st[1])

    def highlevel(self, channels, in_rate, out_rate, lpf, data, chunk=0, capacity=None):
        data = np.ascontiguousarray(data, dtype=np.int16).reshape(-1, channels)
        n = data.shape[0]
        if capacity is None:
            inc = self.ratio(in_rate, out_rate)
            capacity = ((n + 2 * 4096) * 65536) // max(inc, 1) + 16
        out = np.zeros((capacity, channels), dtype=np.int32)
        wrote = self.lib.ref_highlevel_stream(channels, in_rate, out_rate, lpf, _ptr(data, C.c_int16), n, chunk,
                                              _ptr(out, C.c_int32), capacity)
        if wrote == 2**64 - 1:
            raise ValueError("configuration rejected")
        return out[:wrote]

    def adjust_sequence(self, channels, segments, padded, total_frames, capacity):
        """segments: [(in_rate, out_rate, lpf, frame_limit)]; returns (frames, (pos_int, pos_frac, remaining))."""
        padded = np.ascontiguousarray(padded, dtype=np.int16)
        rates = (C.c_uint64 * (3 * len(segments)))(*[x for seg in segments for x in seg[:3]])
        limits = (C.c_uint64 * len(segments))(*[seg[3] for seg in segments])
        out = np.zeros((capacity, channels), dtype=np.int32)
        st = (C.c_uint64 * 3)()
        n = self.lib.ref_lowlevel_adjust_sequence(channels, rates, limits, len(segments), _ptr(padded, C.c_int16), total_frames, _ptr(out, C.c_int32), st)
        if n == 2**64 - 1:
            raise ValueError("configuration rejected")
        return out[:n], tuple(int(x) for x in st)

    def highlevel_adjust(self, channels, segments, switch_at, data, capacity):
        """segments: [(in_rate, out_rate, lpf)], switch_at: output frame counts where the next segment starts."""
        data = np.ascontiguousarray(data, dtype=np.int16).reshape(-1, channels)
        rates = (C.c_uint64 * (3 * len(segments)))(*[x for seg in segments for x in seg])
        sw = (C.c_uint64 * max(len(switch_at), 1))(*switch_at)
        out = np.zeros((capacity, channels), dtype=np.int32)
        n = self.lib.ref_highlevel_adjust_stream(channels, rates, sw, len(segments), _ptr(data, C.c_int16), data.shape[0], _ptr(out, C.c_int32), capacity)
        if n >= 2**64 - 2:
            raise ValueError("configuration rejected")
        return out[:n]

    def time_lowlevel(self, channels, in_rate, out_rate, lpf, padded, total_frames, out_s16):
        frames = C.c_uint64(0)
        secs = self.lib.ref_time_lowlevel(channels, in_rate, out_rate, lpf, _ptr(padded, C.c_int16), total_frames,
                                          _ptr(out_s16, C.c_int16), C.byref(frames))
        return secs, int(frames.value)
