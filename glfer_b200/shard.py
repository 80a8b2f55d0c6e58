"""Time-sharding arithmetic for one-process-per-GPU runs (bench.py under torchrun).

Same rule as glfer_gram_shard_range / glfer_gram_required_span in the C host layer
(glfer_b200/host/gram.c): frames are split into contiguous ranges, the first
(nframes % world) ranks take one extra frame; rank g needs stream samples
[F_g*hop - (N-hop) - halo, F_{g+1}*hop): the (N - hop)-sample overlap history
(fft.c:98-113), (depth-1) extra frames when averaging so its first row sees a full
window (avg.c:116-127), rounded down to a hop-block boundary when block means are
removed (fft.c:86-96).  No exchange between ranks."""
from __future__ import annotations


def frame_range(nframes: int, world: int, rank: int) -> tuple[int, int]:
    base, extra = divmod(nframes, world)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def sample_span(n: int, hop: int, first_frame: int, nframes: int, sub_mean: bool = False,
                avg_depth: int = 0) -> tuple[int, int]:
    halo = min(first_frame, max(avg_depth - 1, 0))
    lo = (first_frame - halo) * hop - (n - hop)
    if sub_mean and lo > 0:
        lo = (lo // hop) * hop
    return max(lo, 0), (first_frame + nframes) * hop
