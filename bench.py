#!/usr/bin/env python
"""bench.py -- IQ Msamples/s of the Mode S demodulation path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (config.workload): BASELINE.json configs[1] -- 60 s of synthetic 2.4 MS/s uc8 IQ
(144 000 000 samples, 288 MB, ~200 frames/s, seed 2 + rank), one independent receiver stream per
GPU.  A step is one pass of the hot path over that stream.

  value     whole-job Msamples/s with the stream already resident in HBM (kernels + survivor
            download + host resolve), CUDA events on the launching stream, max over ranks
  e2e       the same through b200_demod_process() on a pinned HOST buffer: H2D of the 288 MB
            inside the timed region, decoded messages back on the host
  roofline  the scan kernel (K1a: IQ -> magnitude + preamble scan, + candidate list and the u16
            magnitudes K1b slices from): algorithmic bytes (2 B/sample uc8) / mean K1a duration
            (CUDA events inside the library, on the launching stream) against MEASURED_PEAKS.json
            hbm_gbs; scan_only = the same kernel without the candidate list and magnitude store;
            other_kernels = K1b (slice + CRC class of every candidate phase) and K2 (classify)
  cpu_baseline  the unmodified reference (oracle/_ref/ref_demod) on one host core, bounded sample

--impl reference times the reference's own CPU path (oracle/_ref/ref_demod, else the C port)
with one process per host core, each on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

SAMPLE_RATE = 2_400_000
WORKLOAD_SECONDS = 60.0
WORKLOAD = "configs[1]: 60 s synthetic uc8 2.4 MS/s, ~200 frames/s, --preamble-threshold 58 --fix"
METRIC = "IQ Msamples/s"
UNIT = "Msamples/s"


def measured_peak_gbs():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic_ratio():
    """DRAM bytes (read + write) per algorithmic byte of a K1a launch, from the committed ncu capture."""
    p = ROOT / "profiles" / "k1_traffic.json"
    if p.exists():
        try:
            j = json.loads(p.read_text())
            return float(j["dram_bytes_per_launch"]) / float(j["algorithmic_bytes_per_launch"])
        except Exception:
            return None
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.QUERY}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                for line in out.strip().splitlines():
                    self.rows.append([x.strip() for x in line.split(",")])
            except Exception:
                pass
            self._stop.wait(0.5)  # nvidia-smi takes driver locks: sample sparsely

    def start(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(timeout=6)
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                smax.append(float(r[2]))
                for name, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------
# reference arm and cpu baseline (the ONLY places bench.py runs anything under oracle/)
# ------------------------------------------------------------------------------------------

def _reference_runner():
    """(callable(path, repeat) -> (samples, cpu_seconds)), kind"""
    from oracle import ref as oracle_ref
    if oracle_ref.available():
        def run(path, repeat):
            r = oracle_ref.run_file(path, "uc8", repeat=repeat)
            return r.n_samples, float(r.stats["convert_cpu_s"] + r.stats["demod_cpu_s"])
        return run, "reference"
    from oracle import port as oracle_port

    def run(path, repeat):
        iq = np.fromfile(path, dtype=np.uint8)
        total, cpu = 0, 0.0
        for _ in range(repeat):
            r = oracle_port.run(iq, "uc8")
            total += r.n_samples
            cpu += float(r.stats["convert_cpu_s"] + r.stats["demod_cpu_s"])
        return total, cpu
    return run, "port"


def cpu_baseline(iq_sample: np.ndarray, target_cpu_s: float = 12.0):
    """The reference on ONE host core over a bounded sample of the workload."""
    run, kind = _reference_runner()
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "sample.bin")
        iq_sample.tofile(path)
        n1, c1 = run(path, 1)
        repeat = max(1, min(64, int(target_cpu_s / max(c1, 1e-3))))
        n, c = run(path, repeat)
    secs = iq_sample.size // 2 / SAMPLE_RATE
    return {"value": n / c / 1e6, "unit": UNIT, "cores": 1, "kind": kind,
            "sample": f"first {secs:.0f} s of the workload stream replayed {repeat}x as one stream "
                      f"({n} samples, {c:.1f} s thread-CPU: converter + demodulate2400, readsb.c:828-837 clock)"}


def run_reference_arm(args):
    """--impl reference: the reference's CPU path on all host cores, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from readsb_protobuf_b200 import synth
    ncores = os.cpu_count() or 1
    sample_seconds = 10.0
    cfg = synth.baseline_config(1, seconds=sample_seconds)
    iq, _ = synth.generate(cfg)
    nsamples = cfg.nsamples
    _, kind = _reference_runner()
    from readsb_protobuf_b200 import build
    exe = build.ensure_ref()
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "sample.bin")
        iq.tofile(path)

        def one_step():
            t0 = time.perf_counter()
            if exe is not None:
                procs = [subprocess.Popen([str(exe), "--in", path, "--out", os.path.join(td, f"o{i}.res")],
                                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) for i in range(ncores)]
                for p in procs:
                    p.wait()
            else:  # C port, one thread per core through ctypes (releases the GIL)
                from oracle import port as oracle_port
                ths = [threading.Thread(target=oracle_port.run, args=(iq, "uc8")) for _ in range(ncores)]
                for t in ths:
                    t.start()
                for t in ths:
                    t.join()
            return time.perf_counter() - t0

        for _ in range(args.warmup):
            one_step()
        t = sum(one_step() for _ in range(args.steps))
    value = ncores * nsamples * args.steps / t / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "streams": ncores, "samples_per_stream_step": nsamples},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": ncores, "kind": kind,
                         "sample": f"{ncores} independent processes, each the first {sample_seconds:.0f} s of the workload stream per step "
                                   "(whole process wall clock: read + convert + demodulate2400)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------

def bind_to_gpu_numa_node(gpu_index: int):
    """Keep this rank's threads (host resolver, CUDA workers) and its first-touch allocations on the NUMA
    node its GPU hangs off: the resolver reads memory the GPU has just written over PCIe.  Returns a
    short description, or None when the topology is not visible (single node, container without sysfs)."""
    try:
        out = subprocess.run(["nvidia-smi", f"--id={gpu_index}", "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=10).stdout.strip()
        bdf = out.lower()
        if bdf.startswith("00000000:"):
            bdf = bdf[4:]
        node = int(Path(f"/sys/bus/pci/devices/{bdf}/numa_node").read_text())
        if node < 0:
            return None
        cpus = set()
        for part in Path(f"/sys/devices/system/node/node{node}/cpulist").read_text().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if len(cpus) < 2:
            return None
        os.sched_setaffinity(0, cpus)
        return f"numa node {node}, {len(cpus)} cpus"
    except Exception:
        return None


def run_ours(args):
    import torch
    from readsb_protobuf_b200 import api, synth

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the demodulator has no CPU fallback")
    binding = bind_to_gpu_numa_node(local_rank)
    print(f"rank {rank}: gpu {local_rank}, cpu binding: {binding}", file=sys.stderr)
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    seconds = args.seconds
    cfg = synth.baseline_config(1, seed=2 + rank, seconds=seconds)
    nsamples = cfg.nsamples
    nbytes = nsamples * 2

    # the stream, rendered straight into pinned host memory
    host = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    frames = synth.plan(cfg)
    synth.render(cfg, frames, out=host.numpy())
    dev = host.to("cuda", non_blocking=False)
    stream = torch.cuda.current_stream()
    sptr = stream.cuda_stream

    demod = api.Demodulator(fmt="uc8", nfix=1, threshold=58, device=local_rank, max_span_samples=nsamples + (1 << 20))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    stats_vec = torch.zeros(16, dtype=torch.int64, device="cuda")

    def reduce_stats():
        # the only collective of the path: merged demodulator counters (add_stats, stats.c:195-288)
        if dist is None:
            return
        st = demod.stats()
        vals = [int(st["demod_preambles"]), int(st["demod_rejected_bad"]), int(st["demod_rejected_unknown_icao"]),
                *[int(x) for x in st["demod_accepted"]], int(st["messages_total"]), int(st["samples_processed"])]
        stats_vec[: len(vals)] = torch.tensor(vals, dtype=torch.int64)
        dist.all_reduce(stats_vec)

    # A step is one pass of the path over one stream per GPU; the ranks' streams are independent, so the
    # steps run unsynchronised and the merged-statistics all-reduce (the path's only collective, "final
    # message-count/stats reduction") happens once, after the last step, inside the timed region.
    def step_device():
        demod.reset()
        return demod.process_device(dev.data_ptr(), nsamples, final=True, stream=sptr)

    def step_host():
        demod.reset()
        return demod.process_ptr(host.data_ptr(), nsamples, final=True)

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        reduce_stats()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k1, k1b, k2, nmsg, d2h, launches, chunks = [], [], [], 0, 0, 0, 1
        e0.record(stream)
        for _ in range(steps):
            r = fn()
            k1.append(r.timing["scan_ms"])
            k1b.append(r.timing["slice_ms"])
            k2.append(r.timing["classify_ms"])
            nmsg = len(r.msgs)
            d2h = r.timing["d2h_bytes"]
            launches += r.timing["scan_launches"]
            chunks = r.timing["chunks"]
        reduce_stats()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, k1, k1b, k2, nmsg, d2h, launches, chunks

    sampler = ClockSampler(local_rank) if rank == 0 else None  # one sampler per job: nvidia-smi is not free
    if sampler:
        sampler.start()
    ms_dev, k1_ms, k1b_ms, k2_ms, nmsg, _, n_launches, n_chunks = timed(step_device, args.steps, args.warmup)
    clocks = sampler.stop() if sampler else None
    ms_host, _, _, _, _, d2h_bytes, _, _ = timed(step_host, args.steps, max(1, args.warmup // 2))

    # scan kernel alone, both modes (device-resident, same stream)
    scan_only, scan_full, scan_slice = [], [], []
    for i in range(3 + args.steps):
        a, _ = demod.scan_device(dev.data_ptr(), nsamples, mode=0, stream=sptr)
        b, _ = demod.scan_device(dev.data_ptr(), nsamples, mode=1, stream=sptr)
        c, _ = demod.scan_device(dev.data_ptr(), nsamples, mode=2, stream=sptr)
        if i >= 3:
            scan_only.append(a)
            scan_full.append(b)
            scan_slice.append(c)

    total_samples = nsamples * world
    value = total_samples * args.steps / (ms_dev * 1e-3) / 1e6
    e2e_value = total_samples * args.steps / (ms_host * 1e-3) / 1e6

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        k1_mean = float(np.mean(k1_ms))
        ratio = ncu_traffic_ratio()
        traffic = None if ratio is None else ratio * nbytes / max(int(n_chunks), 1)
        achieved = nsamples * 2 / (k1_mean * 1e-3) / 1e9
        so_mean = float(np.mean(scan_only))
        sf_mean = float(np.mean(scan_full))
        k1b_mean = float(np.mean(k1b_ms))
        k2_mean = float(np.mean(k2_ms))
        k1b_alone = max(float(np.mean(scan_slice)) - sf_mean, 1e-6)
        try:
            cpu = cpu_baseline(host.numpy()[: int(10 * SAMPLE_RATE) * 2])
        except Exception as exc:  # the checker failing must not hide the GPU numbers
            cpu = {"value": None, "unit": UNIT, "cores": 1, "kind": "unavailable", "sample": repr(exc)}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "samples_per_gpu": nsamples, "bytes_per_gpu": nbytes,
                       "l2": "input 288 MB per GPU > 126 MB L2, no flush needed", "streams": world, "cpu_binding": binding,
                       "decoded_msgs_per_stream": nmsg, "msgs_per_s": nmsg * world * args.steps / (ms_dev * 1e-3)},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": int(d2h_bytes),
                    "ms_per_step": ms_host / args.steps},
            "gpu_launches": int(n_launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "scan_kernel<uc8> (K1a: IQ -> magnitude + preamble scan + candidates)",
                         "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": "profiles/k1_traffic.json: DRAM bytes per algorithmic byte of the ncu-captured launch x this run's bytes per launch", "peak_source": peak_src,
                         "launches_per_step": int(n_chunks), "algorithmic_bytes_per_launch": nbytes // max(int(n_chunks), 1),
                         "kernel_ms_per_launch": k1_mean / max(int(n_chunks), 1), "kernel_ms_per_step": k1_mean,
                         "scan_only": {"kernel": "scan_kernel<uc8, scan only> (magnitude + preamble scan)",
                                       "kernel_ms": so_mean, "achieved": nbytes / (so_mean * 1e-3) / 1e9,
                                       "frac": nbytes / (so_mean * 1e-3) / 1e9 / peak},
                         "k1a_whole_span_ms": sf_mean,
                         "k1a_whole_span_frac": nbytes / (sf_mean * 1e-3) / 1e9 / peak,
                         "other_kernels": {
                             "K1b slice_kernel (PPM slice + CRC class of every candidate phase; reads the u16 magnitudes, 2 B/sample)":
                                 {"ms_per_step": k1b_mean, "whole_span_ms": k1b_alone,
                                  "achieved": nbytes / (k1b_mean * 1e-3) / 1e9, "frac": nbytes / (k1b_mean * 1e-3) / 1e9 / peak},
                             "K2 classify_warp_kernel (address-set test, dead/live lists, survivors re-sliced; touches candidates only)":
                                 {"ms_per_step": k2_mean}}},
            "cpu_baseline": cpu,
        }
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--seconds", type=float, default=WORKLOAD_SECONDS, help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: relaunch under torchrun
        port = 29500 + (os.getpid() % 1000)
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(port), str(Path(__file__).resolve()),
               "--gpus", str(args.gpus), "--steps", str(args.steps), "--warmup", str(args.warmup), "--seconds", str(args.seconds)]
        return subprocess.call(cmd)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
