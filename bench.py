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

  other_configs  the remaining BASELINE configs, each with value / e2e / msgs_per_s / per-stage ms:
            N=1: configs[2] (60 s sc16 and sc16q11, seed 3), configs[3] (600 s dense uc8, seed 4) and eight
            --dcfilter receiver streams side by side on the one GPU;
            N>1: configs[4] (one 600 s dense uc8 file per GPU, seeds 10..)
  sustained the configs[1] device-resident step repeated for >= 3 s with nvidia-smi clock sampling
            at 5 Hz (the K-step timed region itself lasts tens of milliseconds)

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
            self._stop.wait(0.15)  # one rank samples, ~5 Hz (each query costs ~50 ms and takes driver locks)

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

def _cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        part = part.strip()
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_rank_cpus(gpu_index: int, local_rank: int, local_world: int):
    """Keep this rank's threads (host resolver, CUDA workers) and its first-touch allocations near its GPU and
    away from the other ranks' threads.  Order of preference: the GPU's NUMA node from sysfs; the CPU affinity
    `nvidia-smi topo -m` reports for the GPU; in both cases (and when neither is visible) the candidate CPUs
    are then cut into distinct, equal slices per local rank, so that eight resolvers never share cores."""
    avail = sorted(os.sched_getaffinity(0))
    pool, how = None, "all cpus"
    try:
        out = subprocess.run(["nvidia-smi", f"--id={gpu_index}", "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=10).stdout.strip()
        bdf = out.lower()
        if bdf.startswith("00000000:"):
            bdf = bdf[4:]
        node = int(Path(f"/sys/bus/pci/devices/{bdf}/numa_node").read_text())
        if node >= 0:
            cpus = _cpulist(Path(f"/sys/devices/system/node/node{node}/cpulist").read_text()) & set(avail)
            if len(cpus) >= 2:
                pool, how = sorted(cpus), f"numa node {node}"
    except Exception:
        pass
    if pool is None:
        try:
            topo = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=10).stdout
            for line in topo.splitlines():
                cols = line.split()
                if cols and cols[0] == f"GPU{gpu_index}":
                    for c in cols[1:]:
                        if c and c[0].isdigit() and all(ch.isdigit() or ch in ",-" for ch in c):
                            cpus = _cpulist(c) & set(avail)
                            if len(cpus) >= 2:
                                pool, how = sorted(cpus), "nvidia-smi topo affinity"
                            break
        except Exception:
            pass
    if pool is None:
        pool = avail
    # ranks whose pools coincide (single NUMA node, or no topology at all) take distinct slices
    per = max(2, len(pool) // max(local_world, 1))
    lo = (local_rank * per) % max(len(pool), 1)
    mine = pool[lo:lo + per] if lo + per <= len(pool) else pool[-per:]
    try:
        os.sched_setaffinity(0, set(mine))
        return f"{how}: cpus {mine[0]}-{mine[-1]} ({len(mine)})"
    except Exception:
        return None


class StreamBench:
    """One receiver stream of a BASELINE config on this rank's GPU: pinned host copy, device copy, demodulator."""

    def __init__(self, torch, cfg, fmt, local_rank):
        from readsb_protobuf_b200 import api, synth
        self.torch = torch
        self.cfg = cfg
        self.nsamples = cfg.nsamples
        self.nbytes = cfg.nsamples * synth.BYTES_PER_SAMPLE[fmt]
        # the stream, rendered straight into pinned host memory
        self.host = torch.empty(self.nbytes, dtype=torch.uint8, pin_memory=True)
        synth.render(cfg, synth.plan(cfg), out=self.host.numpy())
        self.dev = self.host.to("cuda", non_blocking=False)
        self.stream = torch.cuda.current_stream()
        self.sptr = self.stream.cuda_stream
        self.demod = api.Demodulator(fmt=fmt, nfix=1, threshold=58, device=local_rank, max_span_samples=self.nsamples + (1 << 20))

    def step_device(self):
        self.demod.reset()
        # copy=False: the decoded messages are in host memory the library owns (a view, not a second copy into numpy)
        return self.demod.process_device(self.dev.data_ptr(), self.nsamples, final=True, stream=self.sptr, copy=False)

    def step_host(self):
        self.demod.reset()
        return self.demod.process_ptr(self.host.data_ptr(), self.nsamples, final=True, copy=False)

    def close(self):
        self.demod.close()
        del self.dev, self.host
        self.torch.cuda.empty_cache()


def run_ours(args):
    import torch
    from readsb_protobuf_b200 import parallel, synth

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the demodulator has no CPU fallback")
    binding = bind_rank_cpus(local_rank, local_rank, local_world)
    print(f"rank {rank}: gpu {local_rank}, cpu binding: {binding}", file=sys.stderr)
    torch.cuda.set_device(local_rank)

    seconds = args.seconds
    cfg = synth.baseline_config(1, seed=2 + rank, seconds=seconds)
    sb = StreamBench(torch, cfg, "uc8", local_rank)
    nsamples, nbytes = sb.nsamples, sb.nbytes
    stream = sb.stream

    # the reference on one host core: BEFORE the process group exists, so that the other ranks wait for rank 0
    # asleep in the rendezvous instead of spinning seven GPUs in an NCCL barrier
    cpu = None
    if rank == 0:
        try:
            cpu = cpu_baseline(sb.host.numpy()[: int(10 * SAMPLE_RATE) * 2])
        except Exception as exc:  # the checker failing must not hide the GPU numbers
            cpu = {"value": None, "unit": UNIT, "cores": 1, "kind": "unavailable", "sample": repr(exc)}

    dist = None
    if world > 1:
        import datetime
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), timeout=datetime.timedelta(minutes=20))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    merged_holder = {}

    def reduce_stats(bench):
        # the only collective of the path: the merged demodulator statistics (add_stats, stats.c:195-288):
        # SUM of the int64 counters, SUM of the double power sums, MAX of peak_signal_power
        merged_holder["stats"] = parallel.reduce_stats(bench.demod.stats(), device="cuda")

    # A step is one pass of the path over one stream per GPU; the ranks' streams are independent, so the
    # steps run unsynchronised and the merged-statistics all-reduce (the path's only collective, "final
    # message-count/stats reduction") happens once, after the last step, inside the timed region.
    def timed(bench, fn, steps, warmup):
        for _ in range(warmup):
            fn()
        reduce_stats(bench)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        acc = {"scan_ms": [], "slice_ms": [], "classify_ms": [], "resolve_ms": [], "h2d_ms": [], "total_ms": []}
        nmsg, d2h, launches, chunks = 0, 0, 0, 1
        e0.record(stream)
        for _ in range(steps):
            r = fn()
            for k in acc:
                acc[k].append(r.timing[k])
            nmsg = len(r.msgs)
            d2h = r.timing["d2h_bytes"]
            launches += r.timing["scan_launches"]
            chunks = r.timing["chunks"]
        reduce_stats(bench)
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, {k: float(np.mean(v)) for k, v in acc.items()}, nmsg, d2h, launches, chunks

    def job_total(x):
        if dist is None:
            return float(x)
        t = torch.tensor([float(x)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t)
        return float(t.item())

    sampler = ClockSampler(local_rank) if rank == 0 else None  # one sampler per job: nvidia-smi is not free
    if sampler:
        sampler.start()
    ms_dev, st_dev, nmsg, _, n_launches, n_chunks = timed(sb, sb.step_device, args.steps, args.warmup)
    # merged == sum (max for the peak) of the ranks' own statistics: checked on the numbers NCCL returned
    stats_check = None
    if dist is not None:
        mine = parallel.pack_stats(sb.demod.stats())
        gi = [torch.zeros(len(mine[0]), dtype=torch.int64, device="cuda") for _ in range(world)]
        gs = [torch.zeros(len(mine[1]), dtype=torch.float64, device="cuda") for _ in range(world)]
        gm = [torch.zeros(len(mine[2]), dtype=torch.float64, device="cuda") for _ in range(world)]
        dist.all_gather(gi, torch.from_numpy(mine[0]).cuda())
        dist.all_gather(gs, torch.from_numpy(mine[1]).cuda())
        dist.all_gather(gm, torch.from_numpy(mine[2]).cuda())
        mi, ms_, mm = parallel.pack_stats(merged_holder["stats"])
        stats_check = bool(np.array_equal(mi, torch.stack(gi).sum(0).cpu().numpy())
                           and np.allclose(ms_, torch.stack(gs).sum(0).cpu().numpy(), rtol=1e-12, atol=0)
                           and np.array_equal(mm, torch.stack(gm).max(0).values.cpu().numpy()))
    merged_headline = merged_holder.get("stats")

    # sustained: the same device-resident step back to back for >= 3 s (--sustain), so that the clock record holds
    # more than a handful of samples under load
    barrier()
    sus_steps, sus_t0 = 0, time.perf_counter()
    es0, es1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    es0.record(stream)
    while time.perf_counter() - sus_t0 < args.sustain:
        sb.step_device()
        sus_steps += 1
    es1.record(stream)
    torch.cuda.synchronize()
    sus_ms = es0.elapsed_time(es1)
    clocks = sampler.stop() if sampler else None
    ms_host, st_host, _, d2h_bytes, _, _ = timed(sb, sb.step_host, args.steps, max(3, args.warmup // 2))

    # scan kernel alone, both modes (device-resident, same stream)
    scan_only, scan_full, scan_slice = [], [], []
    for i in range(3 + args.steps):
        a, _ = sb.demod.scan_device(sb.dev.data_ptr(), nsamples, mode=0, stream=sb.sptr)
        b, _ = sb.demod.scan_device(sb.dev.data_ptr(), nsamples, mode=1, stream=sb.sptr)
        c, _ = sb.demod.scan_device(sb.dev.data_ptr(), nsamples, mode=2, stream=sb.sptr)
        if i >= 3:
            scan_only.append(a)
            scan_full.append(b)
            scan_slice.append(c)
    sb.close()

    # ---- the other BASELINE configs ----
    def run_config(label, cfg_i, fmt, steps, warmup):
        c = synth.baseline_config(cfg_i, seed=None if cfg_i != 4 else 10 + rank)
        if fmt != c.fmt:
            import dataclasses
            c = dataclasses.replace(c, fmt=fmt)
        t0 = time.perf_counter()
        b = StreamBench(torch, c, fmt, local_rank)
        gen_s = time.perf_counter() - t0
        msd, sd, nm, _, _, ch = timed(b, b.step_device, steps, warmup)
        msh, sh, _, d2h, _, _ = timed(b, b.step_host, steps, 3)
        merged = merged_holder.get("stats")
        tot = job_total(b.nsamples)
        msgs_job = job_total(nm)
        out = {
            "workload": label, "samples_per_gpu": b.nsamples, "bytes_per_gpu": b.nbytes, "steps": steps, "warmup": warmup,
            "value": tot * steps / (msd * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": msd / steps,
            "e2e": {"value": tot * steps / (msh * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": msh / steps,
                    "h2d_bytes_per_step": b.nbytes, "d2h_bytes_per_step": int(d2h)},
            "decoded_msgs_per_step": int(msgs_job), "msgs_per_s": msgs_job * steps / (msd * 1e-3),
            "stage_ms_per_step": {"K1a_scan": sd["scan_ms"], "K1b_slice": sd["slice_ms"], "K2_classify_order": sd["classify_ms"],
                                  "host_resolve": sd["resolve_ms"], "call_wall": sd["total_ms"]},
            "e2e_stage_ms_per_step": {"h2d": sh["h2d_ms"], "K1a_scan": sh["scan_ms"], "K1b_slice": sh["slice_ms"],
                                      "K2_classify_order": sh["classify_ms"], "host_resolve": sh["resolve_ms"], "call_wall": sh["total_ms"]},
            "chunks": int(ch), "generate_s": gen_s,
            "k1a_frac_of_hbm_peak": b.nbytes / (sd["scan_ms"] * 1e-3) / 1e9 / measured_peak_gbs()[0],
        }
        if merged is not None and world > 1:
            out["merged_stats"] = {"demod_preambles": int(merged["demod_preambles"]), "messages_total": int(merged["messages_total"]),
                                   "demod_accepted": [int(x) for x in merged["demod_accepted"]],
                                   "samples_processed": int(merged["samples_processed"]),
                                   "peak_signal_power": float(merged["peak_signal_power"])}
        b.close()
        return out

    def run_dcfilter_streams(nstreams=8, seconds=10.0):
        """--dcfilter (convert_*_generic, convert.c:113-213): the DC block is one sequential chain per receiver stream
        (a single warp walks it), so one stream leaves the GPU idle; several receivers side by side -- one context and
        one host thread each, as a multi-receiver host would run them -- fill it."""
        from readsb_protobuf_b200 import api
        n = int(seconds * SAMPLE_RATE)
        ctxs = []
        for i in range(nstreams):
            c = synth.SynthConfig(seed=40 + i + 100 * rank, nsamples=n, frames_per_s=500.0)
            iq, _ = synth.generate(c)
            dev = torch.from_numpy(iq).cuda()
            ctxs.append((api.Demodulator(fmt="uc8", nfix=1, threshold=58, device=local_rank, max_span_samples=n + (1 << 20), dcfilter=True), dev))
        nmsg = [0] * nstreams

        def one(i):
            d, dev = ctxs[i]
            d.reset()
            nmsg[i] = len(d.process_device(dev.data_ptr(), n, final=True, copy=False).msgs)

        def all_streams(k):
            ths = [threading.Thread(target=one, args=(i,)) for i in range(k)]
            t0 = time.perf_counter()
            for t in ths:
                t.start()
            for t in ths:
                t.join()
            torch.cuda.synchronize()
            return time.perf_counter() - t0

        all_streams(nstreams)  # warm-up (allocations)
        t1 = min(all_streams(1) for _ in range(2))
        tn = min(all_streams(nstreams) for _ in range(2))
        for d, _ in ctxs:
            d.close()
        return {"workload": f"{nstreams} receiver streams of {seconds:.0f} s uc8 with --dcfilter on one GPU, one context and host thread each",
                "one_stream_value": n / t1 / 1e6, "value": nstreams * n / tn / 1e6, "unit": UNIT, "streams": nstreams,
                "seconds_wall": tn, "decoded_msgs": int(sum(nmsg)), "timing": "host wall clock around the process calls (device-resident input)"}

    others = {}
    want = args.other_configs
    if want == "auto":
        want = "2,3,dc" if world == 1 else "4"
    try:
        for tok in [t for t in want.split(",") if t and t != "none"]:
            if tok == "2":
                others["configs[2] sc16"] = run_config("configs[2]: 60 s sc16, ~200 frames/s, seed 3", 2, "sc16", 5, 3)
                others["configs[2] sc16q11"] = run_config("configs[2]: 60 s sc16q11, ~200 frames/s, seed 3", 2, "sc16q11", 5, 3)
            elif tok == "3":
                others["configs[3]"] = run_config("configs[3]: 600 s dense uc8, 5000 frames/s, 20 % one-bit errors, seed 4", 3, "uc8", 3, 3)
            elif tok == "dc":
                others["dcfilter x8 streams"] = run_dcfilter_streams()
            elif tok == "4":
                others["configs[4]"] = run_config(f"configs[4]: {world} independent 600 s dense uc8 files, one per GPU, seeds 10..{9 + world}",
                                                  4, "uc8", 3, 3)
    except Exception as exc:
        others["error"] = repr(exc)

    total_samples = nsamples * world
    value = total_samples * args.steps / (ms_dev * 1e-3) / 1e6
    e2e_value = total_samples * args.steps / (ms_host * 1e-3) / 1e6

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        k1_mean = st_dev["scan_ms"]
        ratio = ncu_traffic_ratio()
        traffic = None if ratio is None else ratio * nbytes / max(int(n_chunks), 1)
        achieved = nsamples * 2 / (k1_mean * 1e-3) / 1e9
        so_mean = float(np.mean(scan_only))
        sf_mean = float(np.mean(scan_full))
        k1b_mean = st_dev["slice_ms"]
        k2_mean = st_dev["classify_ms"]
        k1b_alone = max(float(np.mean(scan_slice)) - sf_mean, 1e-6)
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
            "stage_ms_per_step": {"K1a_scan": st_dev["scan_ms"], "K1b_slice": st_dev["slice_ms"], "K2_classify_order": st_dev["classify_ms"],
                                  "host_resolve": st_dev["resolve_ms"], "call_wall": st_dev["total_ms"]},
            "sustained": {"seconds": sus_ms * 1e-3, "steps": sus_steps, "value": nsamples * sus_steps / (sus_ms * 1e-3) / 1e6,
                          "unit": UNIT + " (this rank, device-resident, back to back)"},
            "stats_reduce": {"collectives": "SUM int64 + SUM f64 + MAX f64 (parallel.reduce_stats, add_stats stats.c:195-288)",
                             "backend": "nccl" if dist is not None else "single rank", "merged_equals_sum_of_ranks": stats_check,
                             "merged": None if merged_headline is None else {
                                 "demod_preambles": int(merged_headline["demod_preambles"]),
                                 "messages_total": int(merged_headline["messages_total"]),
                                 "samples_processed": int(merged_headline["samples_processed"]),
                                 "peak_signal_power": float(merged_headline["peak_signal_power"])}},
            "roofline": {"bound": "hbm", "kernel": "scan2_kernel (K1a, uc8: IQ -> magnitude + preamble scan + candidates + u16 magnitudes)",
                         "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": "profiles/k1_traffic.json: DRAM bytes per algorithmic byte of the ncu-captured launch x this run's bytes per launch", "peak_source": peak_src,
                         "launches_per_step": int(n_chunks), "algorithmic_bytes_per_launch": nbytes // max(int(n_chunks), 1),
                         "kernel_ms_per_launch": k1_mean / max(int(n_chunks), 1), "kernel_ms_per_step": k1_mean,
                         "scan_only": {"kernel": "scan2_kernel, scan only (magnitude + preamble scan; no candidate list, no magnitude store)",
                                       "kernel_ms": so_mean, "achieved": nbytes / (so_mean * 1e-3) / 1e9,
                                       "frac": nbytes / (so_mean * 1e-3) / 1e9 / peak},
                         "k1a_whole_span_ms": sf_mean,
                         "k1a_whole_span_frac": nbytes / (sf_mean * 1e-3) / 1e9 / peak,
                         "other_kernels": {
                             "K1b slice_kernel (PPM slice + CRC class of every candidate phase; reads the u16 magnitudes, 2 B/sample)":
                                 {"ms_per_step": k1b_mean, "whole_span_ms": k1b_alone,
                                  "achieved": nbytes / (k1b_mean * 1e-3) / 1e9, "frac": nbytes / (k1b_mean * 1e-3) / 1e9 / peak},
                             "K2 classify_warp_kernel + order_live (address-set test, dead/live lists, live records and hidden counts into pinned host memory; touches candidates only)":
                                 {"ms_per_step": k2_mean}}},
            "cpu_baseline": cpu,
            "other_configs": others,
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
    ap.add_argument("--other-configs", default="auto",
                    help="auto (N=1: configs[2], configs[3] and the --dcfilter multi-stream case; N>1: configs[4]), none, or a list such as 2,3,dc")
    ap.add_argument("--sustain", type=float, default=3.0, help="seconds of back-to-back steps for the clock record")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: relaunch under torchrun
        port = 29500 + (os.getpid() % 1000)
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(port), str(Path(__file__).resolve()),
               "--gpus", str(args.gpus), "--steps", str(args.steps), "--warmup", str(args.warmup), "--seconds", str(args.seconds),
               "--other-configs", args.other_configs, "--sustain", str(args.sustain)]
        return subprocess.call(cmd)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
