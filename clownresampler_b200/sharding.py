"""Multi-GPU partitioning of the hot path: no data-path collective is needed.

* independent streams / voices are dealt to ranks in contiguous blocks (`stream_shard`);
* one long stream is cut into contiguous OUTPUT-time segments, each needing only its own slice of the
  padded input plus a kernel-radius halo (`segment_for_rank`, a thin wrapper over the C entry point
  ClownResamplerB200_SegmentStream, which works from the closed-form position generator);
* an optional final gather of the per-rank outputs (`gather_frames`) is the only collective and is off
  the critical path.  It uses torch.distributed (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import ClownResampler_LowLevel_State, _check, cc_u32f, lib


def stream_shard(n_streams: int, rank: int, world: int) -> range:
    """Contiguous block of stream indices owned by `rank` (sizes differ by at most one)."""
    base, extra = divmod(n_streams, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


@dataclass
class Segment:
    first_output_frame: int      # index in the whole stream's output
    output_frames: int
    first_padded_input_frame: int  # slice of the padded input buffer this segment reads (halo included)
    padded_input_frames: int
    position_integer: int        # start position relative to the slice
    position_fractional: int

    def total_input_frames(self, radius: int) -> int:
        """`total_input_frames` to pass with a private copy of the slice (its padding is the halo)."""
        return max(self.padded_input_frames - 2 * radius, 0)


def segment_for_rank(state: ClownResampler_LowLevel_State, total_input_frames: int, rank: int, world: int) -> Segment:
    out = [C.c_size_t() for _ in range(5)]
    frac = cc_u32f()
    _check(lib().ClownResamplerB200_SegmentStream(C.byref(state), total_input_frames, world, rank,
                                                 *[C.byref(x) for x in out], C.byref(frac)), "SegmentStream")
    return Segment(out[0].value, out[1].value, out[2].value, out[3].value, out[4].value, frac.value)


def gather_frames(local: np.ndarray, group=None) -> np.ndarray | None:
    """Concatenates per-rank output frames on rank 0 (returns None elsewhere)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    device = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    n = torch.tensor([local.shape[0]], dtype=torch.int64, device=device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c.item()) for c in counts]
    width = local.shape[1]
    mine = torch.zeros((max(counts), width), dtype=torch.int32, device=device)
    mine[: local.shape[0]] = torch.from_numpy(np.ascontiguousarray(local, dtype=np.int32)).to(device)
    parts = [torch.zeros_like(mine) for _ in range(world)] if rank == 0 else None
    dist.gather(mine, parts, dst=0, group=group)
    if rank != 0:
        return None
    return np.concatenate([p[:c].cpu().numpy() for p, c in zip(parts, counts)], axis=0)
