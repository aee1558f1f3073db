"""Prints the host-side timeline of the chunk pipeline for a few steps of a BASELINE config (B200_SPAN_TRACE).

    python scripts/span_trace.py [config index = 1] [format = uc8] [steps = 4] [host]

Each line is one span: the time (ms since the span began) at which a chunk was issued, the host started waiting
for a chunk, the chunk's results were there, and its resolve was done.
"""
import os
import sys

os.environ["B200_SPAN_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

import bench  # noqa: E402
from readsb_protobuf_b200 import synth  # noqa: E402


def main():
    idx = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    fmt = sys.argv[2] if len(sys.argv) > 2 else "uc8"
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
    host = len(sys.argv) > 4 and sys.argv[4] == "host"
    bench.bind_rank_cpus(0, 0, 1)
    cfg = synth.baseline_config(idx, seed=2)
    sb = bench.StreamBench(torch, cfg, fmt, 0)
    for _ in range(steps):
        (sb.step_host if host else sb.step_device)()
        t = sb.demod.timing()
        print("  timing:", {k: round(v, 3) if isinstance(v, float) else v for k, v in t.items()}, file=sys.stderr)
    sb.close()


if __name__ == "__main__":
    main()
