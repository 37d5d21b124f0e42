"""clownresampler_b200 -- Python (ctypes) mirror of the C API of libclownresampler_b200.so.

The product is the shared library: hand-written sm_100a CUDA kernels behind the reference's C89
API (include/clownresampler.h) plus bulk extensions (include/clownresampler_b200.h).  This
module only binds it, with the same names and argument meaning as the reference header
(/root/reference/clownresampler.h), so that tests and benchmarks read like the reference's own
harnesses.  There is no Python or CPU implementation of the hot path here: if the library is not
built, importing fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# CRB200_LIB selects another build of the same library (kernel-variant A/B runs); never a different implementation
LIB_PATH = os.environ.get("CRB200_LIB") or os.path.join(_HERE, "lib", "libclownresampler_b200.so")

KERNEL_RADIUS = 3
KERNEL_RESOLUTION = 0x400
MAXIMUM_CHANNELS = 16
TABLE_SIZE = KERNEL_RADIUS * 2 * KERNEL_RESOLUTION

OUT_S32, OUT_S16_CLAMPED, OUT_S32_RAW = 0, 1, 2

# ---- the reference's types in its default (C89) integer mode, LP64 --------------------------
cc_s16l = C.c_short
cc_s32l = C.c_long
cc_s32f = C.c_long
cc_u8f = C.c_uint
cc_u32f = C.c_ulong
cc_bool = C.c_ubyte


class ClownResampler_Precomputed(C.Structure):
    _fields_ = [("lanczos_kernel_table", cc_s32l * TABLE_SIZE)]


class ClownResampler_LowestLevel_Configuration(C.Structure):
    _fields_ = [("stretched_kernel_radius", C.c_size_t), ("integer_stretched_kernel_radius", C.c_size_t),
                ("stretched_kernel_radius_delta", C.c_size_t), ("kernel_step_size", C.c_size_t)]


class ClownResampler_LowLevel_State(C.Structure):
    _fields_ = [("lowest_level", ClownResampler_LowestLevel_Configuration), ("channels", cc_u8f),
                ("position_integer", C.c_size_t), ("position_fractional", cc_u32f), ("increment", cc_u32f)]


class ClownResampler_HighLevel_State(C.Structure):
    _fields_ = [("low_level", ClownResampler_LowLevel_State), ("input_buffer", cc_s16l * 0x1000),
                ("input_buffer_start", C.POINTER(cc_s16l)), ("input_buffer_end", C.POINTER(cc_s16l)),
                ("maximum_integer_stretched_kernel_radius", C.c_size_t),
                ("leading_padding_frames_needed", C.c_size_t), ("trailing_padding_frames_remaining", C.c_size_t)]


class ClownResamplerB200_Job(C.Structure):
    _fields_ = [("input", C.c_void_p), ("output", C.c_void_p), ("total_input_frames", C.c_size_t),
                ("position_integer", C.c_size_t), ("position_fractional", cc_u32f),
                ("first_output_frame", C.c_size_t), ("output_frames", C.c_size_t)]


class ClownResamplerB200_PlanarJob(C.Structure):
    _fields_ = [("input_planes", C.POINTER(C.c_void_p)), ("output_planes", C.POINTER(C.c_void_p)), ("channels", C.c_size_t),
                ("total_input_frames", C.c_size_t), ("position_integer", C.c_size_t), ("position_fractional", cc_u32f),
                ("first_output_frame", C.c_size_t), ("output_frames", C.c_size_t)]


class ClownResamplerB200_PlanInfo(C.Structure):
    _fields_ = [("channels", C.c_uint), ("increment", C.c_ulong), ("phases", C.c_uint), ("taps_max", C.c_uint),
                ("columns", C.c_uint), ("runs", C.c_uint), ("tile_output_frames", C.c_uint), ("tile_input_frames", C.c_uint),
                ("smem_bytes", C.c_uint), ("kernel_kind", C.c_uint), ("mean_taps", C.c_double)]


ClownResampler_InputCallback = C.CFUNCTYPE(C.c_size_t, C.c_void_p, C.POINTER(cc_s16l), C.c_size_t)
ClownResampler_OutputCallback = C.CFUNCTYPE(cc_bool, C.c_void_p, C.POINTER(cc_s32f), cc_u8f)

assert C.sizeof(ClownResampler_Precomputed) == 49152 and C.sizeof(ClownResampler_LowestLevel_Configuration) == 32
assert C.sizeof(ClownResampler_LowLevel_State) == 64 and C.sizeof(ClownResampler_HighLevel_State) == 8296

# every symbol include/clownresampler.h and include/clownresampler_b200.h declare
DROPIN_SYMBOLS = [
    "ClownResampler_Precompute", "ClownResampler_LowestLevel_Configure", "ClownResampler_LowestLevel_Resample",
    "ClownResampler_LowLevel_Init", "ClownResampler_LowLevel_Adjust", "ClownResampler_LowLevel_Resample",
    "ClownResampler_HighLevel_Init", "ClownResampler_HighLevel_Resample", "ClownResampler_HighLevel_Adjust",
    "ClownResampler_HighLevel_ResampleEnd",
]
EXTENSION_SYMBOLS = [
    "ClownResamplerB200_GetCounters", "ClownResamplerB200_Init", "ClownResamplerB200_Shutdown", "ClownResamplerB200_GetLastError", "ClownResamplerB200_DeviceCount",
    "ClownResamplerB200_CountOutputFrames", "ClownResamplerB200_AdvanceState", "ClownResamplerB200_PlanCreate",
    "ClownResamplerB200_PlanDestroy", "ClownResamplerB200_PlanGetInfo", "ClownResamplerB200_ResampleDevice",
    "ClownResamplerB200_ResampleHost", "ClownResamplerB200_SegmentStream", "ClownResamplerB200_DeviceAlloc",
    "ClownResamplerB200_DeviceFree", "ClownResamplerB200_PinnedAlloc", "ClownResamplerB200_PinnedFree",
    "ClownResamplerB200_CopyToDevice", "ClownResamplerB200_CopyToHost", "ClownResamplerB200_Synchronize",
    "ClownResamplerB200_FillNoiseDevice", "ClownResamplerB200_ChecksumDevice", "ClownResamplerB200_DebugBuildPlanHost",
    "ClownResamplerB200_VoiceBatchCreate", "ClownResamplerB200_VoiceBatchDestroy", "ClownResamplerB200_VoiceBatchPush",
    "ClownResamplerB200_VoiceBatchEnd", "ClownResamplerB200_VoiceBatchTick", "ClownResamplerB200_VoiceBatchAdjust",
    "ClownResamplerB200_PlanCreateOnDevice", "ClownResamplerB200_DeviceAllocOn", "ClownResamplerB200_SynchronizeOn",
    "ClownResamplerB200_ResampleHostMulti", "ClownResamplerB200_PlansBuilt", "ClownResamplerB200_VoiceBatchTickBegin", "ClownResamplerB200_VoiceBatchTickEnd",
    "ClownResamplerB200_ResamplePlanarDevice", "ClownResamplerB200_DeinterleaveDevice", "ClownResamplerB200_InterleaveDevice",
]


class Error(RuntimeError):
    pass


_lib = None


def lib() -> C.CDLL:
    """Loads the shared library (never a fallback: a missing build is an error)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is not built; run `make` at the repository root (or __graft_entry__.build()). "
                          "There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    P = C.POINTER
    L.ClownResampler_Precompute.argtypes = [P(ClownResampler_Precomputed)]
    L.ClownResampler_Precompute.restype = None
    L.ClownResampler_LowestLevel_Configure.argtypes = [P(ClownResampler_LowestLevel_Configuration), cc_u32f, cc_u32f, cc_u32f]
    L.ClownResampler_LowestLevel_Configure.restype = cc_bool
    L.ClownResampler_LowestLevel_Resample.argtypes = [P(ClownResampler_LowestLevel_Configuration), P(ClownResampler_Precomputed),
                                                     P(cc_s32f), cc_u8f, C.c_void_p, C.c_size_t, cc_u32f]
    L.ClownResampler_LowestLevel_Resample.restype = None
    for name in ("ClownResampler_LowLevel_Init",):
        getattr(L, name).argtypes = [P(ClownResampler_LowLevel_State), cc_u8f, cc_u32f, cc_u32f, cc_u32f]
        getattr(L, name).restype = cc_bool
    L.ClownResampler_LowLevel_Adjust.argtypes = [P(ClownResampler_LowLevel_State), cc_u32f, cc_u32f, cc_u32f]
    L.ClownResampler_LowLevel_Adjust.restype = cc_bool
    L.ClownResampler_LowLevel_Resample.argtypes = [P(ClownResampler_LowLevel_State), P(ClownResampler_Precomputed), C.c_void_p,
                                                  P(C.c_size_t), ClownResampler_OutputCallback, C.c_void_p]
    L.ClownResampler_LowLevel_Resample.restype = cc_bool
    L.ClownResampler_HighLevel_Init.argtypes = [P(ClownResampler_HighLevel_State), cc_u8f, cc_u32f, cc_u32f, cc_u32f]
    L.ClownResampler_HighLevel_Init.restype = cc_bool
    L.ClownResampler_HighLevel_Resample.argtypes = [P(ClownResampler_HighLevel_State), P(ClownResampler_Precomputed),
                                                   ClownResampler_InputCallback, ClownResampler_OutputCallback, C.c_void_p]
    L.ClownResampler_HighLevel_Resample.restype = cc_bool
    L.ClownResampler_HighLevel_Adjust.argtypes = [P(ClownResampler_HighLevel_State), cc_u32f, cc_u32f, cc_u32f]
    L.ClownResampler_HighLevel_Adjust.restype = cc_bool
    L.ClownResampler_HighLevel_ResampleEnd.argtypes = [P(ClownResampler_HighLevel_State), P(ClownResampler_Precomputed),
                                                      ClownResampler_OutputCallback, C.c_void_p]
    L.ClownResampler_HighLevel_ResampleEnd.restype = cc_bool

    L.ClownResamplerB200_Init.argtypes = [C.c_int]
    L.ClownResamplerB200_GetLastError.restype = C.c_char_p
    L.ClownResamplerB200_CountOutputFrames.argtypes = [P(ClownResampler_LowLevel_State), C.c_size_t]
    L.ClownResamplerB200_CountOutputFrames.restype = C.c_size_t
    L.ClownResamplerB200_AdvanceState.argtypes = [P(ClownResampler_LowLevel_State), P(C.c_size_t), C.c_size_t, C.c_int]
    L.ClownResamplerB200_AdvanceState.restype = None
    L.ClownResamplerB200_PlanCreate.argtypes = [P(ClownResampler_Precomputed), P(ClownResampler_LowLevel_State)]
    L.ClownResamplerB200_PlanCreate.restype = C.c_void_p
    L.ClownResamplerB200_PlanCreateOnDevice.argtypes = [P(ClownResampler_Precomputed), P(ClownResampler_LowLevel_State), C.c_int]
    L.ClownResamplerB200_PlanCreateOnDevice.restype = C.c_void_p
    L.ClownResamplerB200_DeviceAllocOn.argtypes = [C.c_int, C.c_size_t]
    L.ClownResamplerB200_DeviceAllocOn.restype = C.c_void_p
    L.ClownResamplerB200_SynchronizeOn.argtypes = [C.c_int, C.c_void_p]
    L.ClownResamplerB200_ResampleHostMulti.argtypes = [P(ClownResampler_Precomputed), P(ClownResampler_LowLevel_State), P(C.c_int), C.c_size_t,
                                                      P(ClownResamplerB200_Job), C.c_size_t, C.c_int]
    L.ClownResamplerB200_PlansBuilt.restype = C.c_ulong
    L.ClownResamplerB200_ResamplePlanarDevice.argtypes = [C.c_void_p, P(ClownResamplerB200_PlanarJob), C.c_size_t, C.c_int, C.c_void_p]
    L.ClownResamplerB200_DeinterleaveDevice.argtypes = [C.c_void_p, P(C.c_void_p), C.c_size_t, C.c_uint, C.c_int, C.c_void_p]
    L.ClownResamplerB200_InterleaveDevice.argtypes = [P(C.c_void_p), C.c_void_p, C.c_size_t, C.c_uint, C.c_int, C.c_void_p]
    L.ClownResamplerB200_PlanDestroy.argtypes = [C.c_void_p]
    L.ClownResamplerB200_PlanDestroy.restype = None
    L.ClownResamplerB200_PlanGetInfo.argtypes = [C.c_void_p, P(ClownResamplerB200_PlanInfo)]
    L.ClownResamplerB200_ResampleDevice.argtypes = [C.c_void_p, P(ClownResamplerB200_Job), C.c_size_t, C.c_int, C.c_void_p]
    L.ClownResamplerB200_ResampleHost.argtypes = [C.c_void_p, P(ClownResamplerB200_Job), C.c_size_t, C.c_int]
    L.ClownResamplerB200_SegmentStream.argtypes = [P(ClownResampler_LowLevel_State), C.c_size_t, C.c_size_t, C.c_size_t] + [P(C.c_size_t)] * 5 + [P(cc_u32f)]
    L.ClownResamplerB200_DeviceAlloc.argtypes = [C.c_size_t]
    L.ClownResamplerB200_DeviceAlloc.restype = C.c_void_p
    L.ClownResamplerB200_DeviceFree.argtypes = [C.c_void_p]
    L.ClownResamplerB200_DeviceFree.restype = None
    L.ClownResamplerB200_PinnedAlloc.argtypes = [C.c_size_t]
    L.ClownResamplerB200_PinnedAlloc.restype = C.c_void_p
    L.ClownResamplerB200_PinnedFree.argtypes = [C.c_void_p]
    L.ClownResamplerB200_PinnedFree.restype = None
    L.ClownResamplerB200_CopyToDevice.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    L.ClownResamplerB200_CopyToHost.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    L.ClownResamplerB200_Synchronize.argtypes = [C.c_void_p]
    L.ClownResamplerB200_FillNoiseDevice.argtypes = [C.c_void_p, C.c_uint, C.c_uint, C.c_size_t, C.c_size_t, C.c_uint, C.c_void_p]
    L.ClownResamplerB200_ChecksumDevice.argtypes = [C.c_void_p, C.c_size_t, C.c_int, P(C.c_ulong), C.c_void_p]
    L.ClownResamplerB200_DebugBuildPlanHost.argtypes = [P(ClownResampler_Precomputed), P(ClownResampler_LowLevel_State), C.c_uint,
                                                       P(C.c_uint), C.c_size_t, P(C.c_int), C.c_size_t]
    L.ClownResamplerB200_VoiceBatchCreate.argtypes = [P(ClownResampler_Precomputed), C.c_size_t, cc_u8f, cc_u32f, cc_u32f, cc_u32f]
    L.ClownResamplerB200_VoiceBatchCreate.restype = C.c_void_p
    L.ClownResamplerB200_VoiceBatchDestroy.argtypes = [C.c_void_p]
    L.ClownResamplerB200_VoiceBatchDestroy.restype = None
    L.ClownResamplerB200_VoiceBatchPush.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
    L.ClownResamplerB200_VoiceBatchEnd.argtypes = [C.c_void_p, C.c_size_t]
    L.ClownResamplerB200_VoiceBatchAdjust.argtypes = [C.c_void_p, C.c_size_t, cc_u32f, cc_u32f, cc_u32f]
    L.ClownResamplerB200_VoiceBatchTick.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_size_t, P(C.c_size_t)]
    L.ClownResamplerB200_VoiceBatchTickBegin.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_size_t, P(C.c_size_t)]
    L.ClownResamplerB200_VoiceBatchTickEnd.argtypes = [C.c_void_p]
    _lib = L
    return L


def last_error() -> str:
    return (lib().ClownResamplerB200_GetLastError() or b"").decode()


def _check(rc: int, what: str):
    if rc != 0:
        raise Error(f"{what} failed ({rc}): {last_error()}")


# ---- thin object layer over the C API --------------------------------------------------------
def Precompute() -> ClownResampler_Precomputed:
    """ClownResampler_Precompute (H:682)."""
    pre = ClownResampler_Precomputed()
    lib().ClownResampler_Precompute(C.byref(pre))
    return pre


def table_of(pre: ClownResampler_Precomputed) -> np.ndarray:
    return np.ctypeslib.as_array(pre.lanczos_kernel_table).astype(np.int64)


def LowLevel_Init(channels, input_sample_rate, output_sample_rate, low_pass_filter_sample_rate):
    """ClownResampler_LowLevel_Init (H:711): returns the state, or None when the reference would return cc_false."""
    st = ClownResampler_LowLevel_State()
    ok = lib().ClownResampler_LowLevel_Init(C.byref(st), channels, input_sample_rate, output_sample_rate, low_pass_filter_sample_rate)
    return st if ok else None


def CountOutputFrames(state, total_input_frames) -> int:
    return int(lib().ClownResamplerB200_CountOutputFrames(C.byref(state), total_input_frames))


def LowLevel_Resample(state, pre, padded_input: np.ndarray, total_input_frames: int, max_frames: int = 0):
    """ClownResampler_LowLevel_Resample (H:749) with an output callback that stores each frame and
    returns 0 on frame number `max_frames` (0 = never), as examples/low-level.c:84 does.
    Returns (frames[int64, n x channels], return value, remaining input frames)."""
    L = lib()
    padded_input = np.ascontiguousarray(padded_input, dtype=np.int16)
    ch = state.channels
    frames = []

    def on_frame(_user, frame, n):
        frames.append([frame[i] for i in range(n)])
        return 0 if (max_frames and len(frames) == max_frames) else 1

    cb = ClownResampler_OutputCallback(on_frame)
    total = C.c_size_t(total_input_frames)
    ret = L.ClownResampler_LowLevel_Resample(C.byref(state), C.byref(pre), padded_input.ctypes.data, C.byref(total), cb, None)
    out = np.array(frames, dtype=np.int64).reshape(-1, ch)
    return out, int(ret), int(total.value)


def HighLevel_Stream(pre, channels, in_rate, out_rate, lpf, data: np.ndarray, chunk: int = 0, max_frames: int = 0):
    """HighLevel_Init + HighLevel_Resample + HighLevel_ResampleEnd (H:770, H:825, H:847) over an in-memory
    stream, the way tests/test-high-level.c:116-127 drives them; `chunk` caps frames per input callback."""
    L = lib()
    st = ClownResampler_HighLevel_State()
    if not L.ClownResampler_HighLevel_Init(C.byref(st), channels, in_rate, out_rate, lpf):
        return None
    data = np.ascontiguousarray(data, dtype=np.int16).reshape(-1, channels)
    pos = [0]
    frames = []

    def on_input(_user, buffer, total_frames):
        n = min(total_frames, data.shape[0] - pos[0])
        if chunk:
            n = min(n, chunk)
        if n:
            C.memmove(buffer, data[pos[0]:pos[0] + n].ctypes.data, n * channels * 2)
        pos[0] += n
        return n

    def on_frame(_user, frame, n):
        frames.append([frame[i] for i in range(n)])
        return 0 if (max_frames and len(frames) == max_frames) else 1

    icb, ocb = ClownResampler_InputCallback(on_input), ClownResampler_OutputCallback(on_frame)
    if L.ClownResampler_HighLevel_Resample(C.byref(st), C.byref(pre), icb, ocb, None):
        L.ClownResampler_HighLevel_ResampleEnd(C.byref(st), C.byref(pre), ocb, None)
    return np.array(frames, dtype=np.int64).reshape(-1, channels)


class DeviceBuffer:
    """A cudaMalloc'd buffer owned through the library's helpers (no torch needed)."""

    def __init__(self, nbytes: int):
        self.nbytes = int(nbytes)
        self.ptr = lib().ClownResamplerB200_DeviceAlloc(max(self.nbytes, 16))
        if not self.ptr:
            raise Error(f"device allocation of {nbytes} bytes failed: {last_error()}")

    @classmethod
    def from_numpy(cls, a: np.ndarray) -> "DeviceBuffer":
        a = np.ascontiguousarray(a)
        b = cls(a.nbytes)
        if a.nbytes:
            _check(lib().ClownResamplerB200_CopyToDevice(b.ptr, a.ctypes.data, a.nbytes), "CopyToDevice")
        return b

    def to_numpy(self, dtype, count=None, offset_bytes=0) -> np.ndarray:
        dtype = np.dtype(dtype)
        count = (self.nbytes - offset_bytes) // dtype.itemsize if count is None else count
        out = np.empty(count, dtype=dtype)
        if out.nbytes:
            _check(lib().ClownResamplerB200_CopyToHost(out.ctypes.data, self.ptr + offset_bytes, out.nbytes), "CopyToHost")
        return out

    def free(self):
        if self.ptr:
            lib().ClownResamplerB200_DeviceFree(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Plan:
    """ClownResamplerB200_Plan: the device-resident per-phase tap table of one configuration."""

    def __init__(self, pre, state, device=None):
        if device is None:
            self.handle = lib().ClownResamplerB200_PlanCreate(C.byref(pre), C.byref(state))
        else:
            self.handle = lib().ClownResamplerB200_PlanCreateOnDevice(C.byref(pre), C.byref(state), device)
        if not self.handle:
            raise Error(f"PlanCreate failed: {last_error()}")
        self.channels = state.channels
        self.info = ClownResamplerB200_PlanInfo()
        _check(lib().ClownResamplerB200_PlanGetInfo(self.handle, C.byref(self.info)), "PlanGetInfo")

    def frame_bytes(self, fmt):
        return {OUT_S32: 4 * self.channels, OUT_S16_CLAMPED: 2 * self.channels, OUT_S32_RAW: 4 * (self.channels + 1)}[fmt]

    @staticmethod
    def _jobs(jobs):
        arr = (ClownResamplerB200_Job * len(jobs))()
        for i, j in enumerate(jobs):
            arr[i] = j
        return arr

    def resample_device(self, jobs, fmt=OUT_S32, stream=None, sync=True):
        _check(lib().ClownResamplerB200_ResampleDevice(self.handle, self._jobs(jobs), len(jobs), fmt, stream), "ResampleDevice")
        if sync:
            _check(lib().ClownResamplerB200_Synchronize(stream), "Synchronize")

    def resample_host(self, jobs, fmt=OUT_S32):
        _check(lib().ClownResamplerB200_ResampleHost(self.handle, self._jobs(jobs), len(jobs), fmt), "ResampleHost")

    def destroy(self):
        if self.handle:
            lib().ClownResamplerB200_PlanDestroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


def make_job(input_ptr, output_ptr, total_input_frames, position_integer=0, position_fractional=0,
             first_output_frame=0, output_frames=0) -> ClownResamplerB200_Job:
    return ClownResamplerB200_Job(input_ptr, output_ptr, total_input_frames, position_integer, position_fractional,
                                  first_output_frame, output_frames)


def resample_array(pre, state, padded_input: np.ndarray, total_input_frames: int, fmt=OUT_S32,
                   first_output_frame=0, output_frames=None, via="device") -> np.ndarray:
    """Bulk resample of one padded host array through the device (or host-staged) path; returns frames."""
    padded_input = np.ascontiguousarray(padded_input, dtype=np.int16)
    plan = Plan(pre, state)
    try:
        n_all = CountOutputFrames(state, total_input_frames)
        n = n_all - first_output_frame if output_frames is None else output_frames
        fb = plan.frame_bytes(fmt)
        dtype = np.int16 if fmt == OUT_S16_CLAMPED else np.int32
        width = fb // np.dtype(dtype).itemsize
        if via == "host":
            out = np.zeros(max(n, 1) * width, dtype=dtype)
            job = make_job(padded_input.ctypes.data, out.ctypes.data, total_input_frames, state.position_integer,
                           state.position_fractional, first_output_frame, n)
            plan.resample_host([job], fmt)
            return out[: n * width].reshape(n, width)
        d_in = DeviceBuffer.from_numpy(padded_input)
        d_out = DeviceBuffer(max(n, 1) * fb)
        job = make_job(d_in.ptr, d_out.ptr, total_input_frames, state.position_integer, state.position_fractional,
                       first_output_frame, n)
        plan.resample_device([job], fmt)
        out = d_out.to_numpy(dtype, n * width).reshape(n, width)
        d_in.free()
        d_out.free()
        return out
    finally:
        plan.destroy()


def counters():
    """(kernel launches of the drop-in calls, drop-in calls served from frames kept from an earlier call)."""
    a, b = C.c_ulong(0), C.c_ulong(0)
    lib().ClownResamplerB200_GetCounters(C.byref(a), C.byref(b))
    return a.value, b.value


def debug_plan_host(pre, state, smem_budget=227 * 1024):
    """Host-only plan (no device): returns (geometry dict, rows[int32 n_rows x row_words])."""
    words = (C.c_uint * 256)()
    rows = (C.c_int * (1 << 16))()
    n = lib().ClownResamplerB200_DebugBuildPlanHost(C.byref(pre), C.byref(state), smem_budget, words, 256, rows, 1 << 16)
    if n < 0:
        raise Error(f"plan rejected ({n}): {last_error()}")
    w = list(words[:n])
    names = ["channels", "increment", "step", "delta", "radius_int", "radius_fx", "ks0", "n_breaks"]
    geo = dict(zip(names, w[:8]))
    geo["breaks"] = w[8:12][: geo["n_breaks"]]
    rest = ["n_rows", "n_cols", "row_words", "taps_max", "n_runs", "tile_out", "tile_in_frames", "stage_bytes", "unstretched5",
            "norm_mode", "kernel_kind", "smem_bytes", "lane_stride"]
    geo.update(dict(zip(rest, w[12:25])))
    geo["rot"], geo["rot_shift"], geo["rot_mask"] = w[25], w[26], w[27]
    quads = [tuple(w[28 + 4 * i: 32 + 4 * i]) for i in range(12)]
    geo["groups"] = [q[:3] for q in quads]             # (first column, columns, rotates) per group
    geo["group_kinds"] = [q[3] for q in quads]         # chain form: class * 2 + single (class 0 / 1 / 2 = positive / negative / signed)
    geo["small_taps"], geo["chain_mode"], geo["n_groups"] = w[76], w[77], w[78]
    runs = w[79:]
    geo["runs"] = [(int(C.c_int(runs[4 * i]).value), int(C.c_int(runs[4 * i + 1]).value), int(C.c_int(runs[4 * i + 2]).value),
                    runs[4 * i + 3] & 3, (runs[4 * i + 3] >> 2) & 1) for i in range(geo["n_runs"])]   # (col, len, off, negative: 0 / 1 / 2 = signed, big)
    flat = np.ctypeslib.as_array(rows)
    n = geo["n_rows"] * geo["row_words"]
    r = flat[:n].reshape(geo["n_rows"], geo["row_words"]).copy()
    geo["col_offsets"] = [] if geo["unstretched5"] else flat[n: n + geo["n_cols"]].tolist()   # bytes after the window start, per column
    return geo, r


class VoiceBatch:
    """ClownResamplerB200_VoiceBatch: many HighLevel-style voices advanced together, one launch per tick."""

    def __init__(self, pre, voices, channels, in_rate, out_rate, lpf):
        self.handle = lib().ClownResamplerB200_VoiceBatchCreate(C.byref(pre), voices, channels, in_rate, out_rate, lpf)
        if not self.handle:
            raise Error(f"VoiceBatchCreate failed: {last_error()}")
        self.voices, self.channels = voices, channels

    def push(self, voice, frames: np.ndarray):
        frames = np.ascontiguousarray(frames, dtype=np.int16).reshape(-1, self.channels)
        _check(lib().ClownResamplerB200_VoiceBatchPush(self.handle, voice, frames.ctypes.data, frames.shape[0]), "VoiceBatchPush")

    def end(self, voice):
        _check(lib().ClownResamplerB200_VoiceBatchEnd(self.handle, voice), "VoiceBatchEnd")

    def adjust(self, voice, in_rate, out_rate, lpf):
        _check(lib().ClownResamplerB200_VoiceBatchAdjust(self.handle, voice, in_rate, out_rate, lpf), "VoiceBatchAdjust")

    def tick(self, max_frames, fmt=OUT_S32, out=None):
        """Returns (out[voices, max_frames, channels], produced[voices])."""
        dtype = np.int16 if fmt == OUT_S16_CLAMPED else np.int32
        if out is None:
            out = np.zeros((self.voices, max_frames, self.channels), dtype=dtype)
        produced = (C.c_size_t * self.voices)()
        _check(lib().ClownResamplerB200_VoiceBatchTick(self.handle, max_frames, fmt, out.ctypes.data, out.strides[0], produced), "VoiceBatchTick")
        return out, np.ctypeslib.as_array(produced).astype(np.int64)

    def destroy(self):
        if self.handle:
            lib().ClownResamplerB200_VoiceBatchDestroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass
