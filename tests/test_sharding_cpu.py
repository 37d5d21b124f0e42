"""The N > 1 path on CPU: stream sharding, output-time segmentation with halo and the final
gather, exercised with world_size 2 over gloo.  The per-segment arithmetic is done by the oracle here (this
file checks the partitioning and plumbing, not the kernels -- tests/test_gpu_parity.py does that on the GPU)."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

import clownresampler_b200 as crb
from clownresampler_b200.sharding import gather_frames, segment_for_rank, stream_shard
from conftest import pad


def test_stream_shard_covers_everything_once():
    for n in (0, 1, 7, 64, 1024):
        for world in (1, 2, 3, 8):
            got = [i for r in range(world) for i in stream_shard(n, r, world)]
            assert got == list(range(n))
            sizes = [len(stream_shard(n, r, world)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.parametrize("case", [(8, 192000, 44100, 44100, 8), (2, 44100, 48000, 48000, 7), (1, 384000, 8000, 8000, 4), (3, 8000, 384000, 384000, 8)])
def test_segments_reproduce_the_one_shot_stream(oracle, case):
    """Each segment, run on a PRIVATE COPY of only its slice with the relative start state, concatenates to
    the one-shot output (the invariant time-segment sharding relies on; same cases as SURVEY.md 8e)."""
    ch, i, o, l, world = case
    st = crb.LowLevel_Init(ch, i, o, l)
    R = st.lowest_level.integer_stretched_kernel_radius
    T = 20000 if i > o else 1500
    data = np.random.default_rng(5).integers(-32768, 32768, size=(T, ch), dtype=np.int16)
    padded = pad(data, R)
    whole = oracle.lowlevel(ch, i, o, l, padded, T)[0]
    parts, covered = [], 0
    for rank in range(world):
        seg = segment_for_rank(st, T, rank, world)
        assert seg.first_output_frame == covered
        covered += seg.output_frames
        if seg.output_frames == 0:
            continue
        private = padded[seg.first_padded_input_frame: seg.first_padded_input_frame + seg.padded_input_frames].copy()
        out = oracle.lowlevel(ch, i, o, l, private, seg.total_input_frames(R), seg.position_integer, seg.position_fractional,
                              max_frames=seg.output_frames)[0]
        parts.append(out)
    assert covered == whole.shape[0]
    assert np.array_equal(np.concatenate(parts), whole)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.cro import Oracle
    oracle = Oracle()
    ch, i, o, l, T = 2, 48000, 44100, 44100, 30000
    st = crb.LowLevel_Init(ch, i, o, l)
    R = st.lowest_level.integer_stretched_kernel_radius
    data = np.random.default_rng(11).integers(-32768, 32768, size=(T, ch), dtype=np.int16)   # same on every rank
    padded = pad(data, R)
    seg = segment_for_rank(st, T, rank, world)
    private = padded[seg.first_padded_input_frame: seg.first_padded_input_frame + seg.padded_input_frames].copy()
    local = oracle.lowlevel(ch, i, o, l, private, seg.total_input_frames(R), seg.position_integer, seg.position_fractional,
                            max_frames=seg.output_frames)[0]
    # streams: every rank owns a block of 5 streams; nothing to exchange, only a count check
    mine = list(stream_shard(5, rank, world))
    total = gather_frames(local)
    if rank == 0:
        whole = oracle.lowlevel(ch, i, o, l, padded, T)[0]
        q.put((bool(np.array_equal(total, whole)), mine))
    else:
        q.put((total is None, mine))
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_gloo_segments_and_gather():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    results = [q.get(timeout=120) for _ in procs]
    [p.join(timeout=60) for p in procs]
    assert all(ok for ok, _ in results)
    assert sorted(i for _, mine in results for i in mine) == [0, 1, 2, 3, 4]
