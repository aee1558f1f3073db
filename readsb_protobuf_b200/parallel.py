"""Multi-GPU plumbing: independent receiver streams, one per rank; one collective for the stats.

The path shards by stream: every rank demodulates its own file with its own ICAO
filter and counters, so there is no data-path exchange.  The only collective mirrors add_stats()
(stats.c:195-288): counters are summed, peak_signal_power is the maximum.
"""
from __future__ import annotations

import numpy as np

from .results import STATS_DTYPE

SUM_INT_FIELDS = ("demod_preambles", "demod_rejected_bad", "demod_rejected_unknown_icao", "demod_accepted",
                  "demod_preamblePhase", "demod_bestPhase", "strong_signal_count", "messages_total",
                  "samples_processed", "noise_power_count", "signal_power_count")
SUM_FLOAT_FIELDS = ("noise_power_sum", "signal_power_sum")
MAX_FLOAT_FIELDS = ("peak_signal_power",)


def pack_stats(stats: np.ndarray):
    ints = np.concatenate([np.atleast_1d(stats[f]).astype(np.int64) for f in SUM_INT_FIELDS])
    sums = np.array([float(stats[f]) for f in SUM_FLOAT_FIELDS], dtype=np.float64)
    maxs = np.array([float(stats[f]) for f in MAX_FLOAT_FIELDS], dtype=np.float64)
    return ints, sums, maxs


def unpack_stats(ints, sums, maxs) -> np.ndarray:
    out = np.zeros(1, dtype=STATS_DTYPE)[0]
    o = 0
    for f in SUM_INT_FIELDS:
        n = int(np.prod(STATS_DTYPE[f].shape)) if STATS_DTYPE[f].shape else 1
        v = np.asarray(ints[o:o + n])
        out[f] = v.reshape(STATS_DTYPE[f].shape) if STATS_DTYPE[f].shape else int(v[0])
        o += n
    for f, v in zip(SUM_FLOAT_FIELDS, sums):
        out[f] = float(v)
    for f, v in zip(MAX_FLOAT_FIELDS, maxs):
        out[f] = float(v)
    return out


def reduce_stats(stats: np.ndarray, device=None, group=None) -> np.ndarray:
    """All-reduce one rank's demodulator stats over the process group (NCCL on GPUs, gloo on CPU)."""
    import torch
    import torch.distributed as dist

    ints, sums, maxs = pack_stats(stats)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return unpack_stats(ints, sums, maxs)
    dev = device if device is not None else ("cuda" if dist.get_backend(group) == "nccl" else "cpu")
    ti = torch.from_numpy(ints).to(dev)
    ts = torch.from_numpy(sums).to(dev)
    tm = torch.from_numpy(maxs).to(dev)
    dist.all_reduce(ti, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(ts, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(tm, op=dist.ReduceOp.MAX, group=group)
    return unpack_stats(ti.cpu().numpy(), ts.cpu().numpy(), tm.cpu().numpy())


def stream_seed(base_seed: int, rank: int) -> int:
    """Seed of rank's receiver file (BASELINE.json configs[4]: seeds 10..17, one per GPU)."""
    return base_seed + rank
