"""CPU, world_size 2 over gloo: the stats collective of the multi-GPU path (one stream per rank)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import port
from readsb_protobuf_b200 import parallel, results, synth


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port_no, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port_no))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # each rank demodulates its own receiver stream (here with the CPU oracle: the collective is
        # what is under test, not the kernels)
        cfg = synth.SynthConfig(seed=parallel.stream_seed(40, rank), nsamples=300_000, frames_per_s=3000, frac_biterror=0.2)
        iq, _ = synth.generate(cfg)
        mine = port.run(iq, "uc8").stats
        merged = parallel.reduce_stats(mine)
        out[rank] = (mine.tobytes(), merged.tobytes())
    finally:
        dist.destroy_process_group()


def test_stats_reduce_world_size_2():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    per_rank = [np.frombuffer(out[r][0], dtype=results.STATS_DTYPE)[0] for r in range(world)]
    merged = [np.frombuffer(out[r][1], dtype=results.STATS_DTYPE)[0] for r in range(world)]
    assert merged[0].tobytes() == merged[1].tobytes()
    m = merged[0]
    for f in parallel.SUM_INT_FIELDS:
        assert np.array_equal(m[f], per_rank[0][f] + per_rank[1][f]), f
    for f in parallel.SUM_FLOAT_FIELDS:
        assert float(m[f]) == pytest.approx(float(per_rank[0][f]) + float(per_rank[1][f]), rel=1e-12)
    assert float(m["peak_signal_power"]) == max(float(per_rank[0]["peak_signal_power"]), float(per_rank[1]["peak_signal_power"]))
    assert int(m["messages_total"]) > 0


def test_pack_roundtrip_single_process():
    cfg = synth.SynthConfig(seed=41, nsamples=200_000, frames_per_s=3000)
    iq, _ = synth.generate(cfg)
    st = port.run(iq, "uc8").stats
    back = parallel.reduce_stats(st)
    for f in parallel.SUM_INT_FIELDS + parallel.SUM_FLOAT_FIELDS + parallel.MAX_FLOAT_FIELDS:
        assert np.array_equal(back[f], st[f]), f
