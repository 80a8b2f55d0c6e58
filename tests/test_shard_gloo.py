"""The multi-process (one rank per GPU) sharding arithmetic bench.py uses, exercised with
world_size 2 over gloo on the CPU: every rank derives its frame range and sample span
from (rank, world), no data-path collective; the ranges must tile the recording and each
span must carry the (N - hop) halo."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from glfer_b200 import shard


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, hop, nframes, sub_mean, depth, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, count = shard.frame_range(nframes, world, rank)
    lo, hi = shard.sample_span(n, hop, first, count, sub_mean=sub_mean, avg_depth=depth)
    mine = torch.tensor([first, count, lo, hi], dtype=torch.int64)
    got = [torch.zeros(4, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(got, mine)
    # the timing reduction bench.py does: max over ranks
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.barrier()
    if rank == 0:
        q.put(([g.tolist() for g in got], t.item()))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_ranks_tile_the_recording(world):
    n, hop, nframes = 4096, 2048, 84375
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, hop, nframes, True, 4, q)) for r in range(world)]
    for p in procs:
        p.start()
    got, tmax = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert tmax == float(world)
    pos = 0
    for first, count, lo, hi in got:
        assert first == pos
        pos += count
        assert hi == (first + count) * hop
        halo_frames = min(first, 3)
        assert lo == max(0, ((first - halo_frames) * hop - (n - hop)))
    assert pos == nframes


def test_shard_matches_library_partition(api):
    for nframes in (1, 10, 84375):
        for world in (1, 2, 4, 8):
            for r in range(world):
                assert shard.frame_range(nframes, world, r) == api.shard_range(nframes, world, r)


def test_span_block_alignment_for_odd_hops():
    # hop 409 (overlap 0.9 at N=4096): with sub_mean the span starts on a block boundary
    lo, hi = shard.sample_span(4096, 409, 100, 10, sub_mean=True, avg_depth=0)
    assert lo % 409 == 0 and lo <= 100 * 409 - (4096 - 409) and hi == 110 * 409
    lo2, _ = shard.sample_span(4096, 409, 100, 10, sub_mean=False, avg_depth=0)
    assert lo2 == 100 * 409 - (4096 - 409)
