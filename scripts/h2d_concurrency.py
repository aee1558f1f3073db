"""Measures the box's host-to-device ceiling: N concurrent pinned H2D copies of the bench's 288 MB stream, one per GPU.

    python scripts/h2d_concurrency.py [out.json]        # on a box with 8 GPUs: N = 1, 2, 4, 8

The end-to-end number of bench.py at N GPUs is bounded by this (every step moves 288 MB per GPU over PCIe from
page-locked host memory); what it shows is how far the aggregate falls short of N x the single-GPU rate when all
GPUs pull from the same host memory system at once.  One process, one thread: the copies are asynchronous
(cudaMemcpyAsync on a stream per device) and are timed from the first launch to the last completion.
"""
import json
import sys
import time

import torch

NBYTES = 288_000_000
REPS = 10


def measure(n):
    hosts = [torch.empty(NBYTES, dtype=torch.uint8, pin_memory=True) for _ in range(n)]
    devs, streams = [], []
    for i in range(n):
        with torch.cuda.device(i):
            devs.append(torch.empty(NBYTES, dtype=torch.uint8, device=f"cuda:{i}"))
            streams.append(torch.cuda.Stream(device=i))
    for h in hosts:
        h.fill_(7)  # touch the pages

    def one_round():
        for i in range(n):
            with torch.cuda.device(i), torch.cuda.stream(streams[i]):
                devs[i].copy_(hosts[i], non_blocking=True)
        for i in range(n):
            streams[i].synchronize()

    for _ in range(3):
        one_round()
    t0 = time.perf_counter()
    for _ in range(REPS):
        one_round()
    dt = (time.perf_counter() - t0) / REPS
    return {"gpus": n, "ms_per_round": dt * 1e3, "aggregate_GBps": n * NBYTES / dt / 1e9, "per_gpu_GBps": NBYTES / dt / 1e9}


def main():
    ngpu = torch.cuda.device_count()
    rows = [measure(n) for n in (1, 2, 4, 8) if n <= ngpu]
    out = {"what": "concurrent pinned H2D copies of 288 MB, one per GPU (cudaMemcpyAsync, a stream per device)",
           "gpus_visible": ngpu, "bytes_per_copy": NBYTES, "rows": rows}
    text = json.dumps(out, indent=1)
    print(text)
    if len(sys.argv) > 1:
        open(sys.argv[1], "w").write(text + "\n")


if __name__ == "__main__":
    main()
