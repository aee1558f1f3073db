"""Turns an ncu report into what profiles/ keeps: a JSON summary per captured launch and the hot SASS of each kernel.

    python scripts/ncu_summarize.py gpurun_out/r02_all.ncu-rep profiles/r02_all [label,label,...]

Writes  <out>_ncu_summary.json   one entry per launch: duration, issue / warp / pipe utilisation, DRAM bytes,
                                 shared-memory wavefronts and conflicts, registers, the top stall reasons
        <out>_sass/<kernel>.txt  for the first launch of every kernel: the instructions of its hottest loops (the
                                 contiguous SASS regions that carry >= 5 % of the executed warp instructions), with
                                 executed counts and stall samples -- the listing DESIGN.md quotes from
Runs `ncu -i` (no GPU needed).  Labels (optional, comma separated, one per launch) name what each launch processed.
"""
import csv
import io
import json
import os
import re
import subprocess
import sys

KEYS = {
    "duration_us": ("gpu__time_duration.sum", 1e-3),  # ns in the raw page when no unit scaling is applied; fixed below from the unit row
    "issue_active_pct": ("smsp__issue_active.avg.pct_of_peak_sustained_active", 1),
    "warps_active_pct": ("sm__warps_active.avg.pct_of_peak_sustained_active", 1),
    "sm_throughput_pct": ("sm__throughput.avg.pct_of_peak_sustained_elapsed", 1),
    "dram_throughput_pct": ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1),
    "lts_throughput_pct": ("lts__throughput.avg.pct_of_peak_sustained_elapsed", 1),
    "l1tex_throughput_pct": ("l1tex__throughput.avg.pct_of_peak_sustained_active", 1),
    "lts_hit_rate_pct": ("lts__t_sector_hit_rate.pct", 1),
    "pipe_alu_pct": ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", 1),
    "pipe_fma_pct": ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", 1),
    "pipe_fmaheavy_pct": ("sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", 1),
    "pipe_lsu_pct": ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", 1),
    "warp_instructions": ("smsp__inst_executed.sum", 1),
    "dram_bytes_read": ("dram__bytes_read.sum", None),
    "dram_bytes_write": ("dram__bytes_write.sum", None),
    "smem_wavefronts": ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", 1),
    "smem_bank_conflicts": ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", 1),
    "registers_per_thread": ("launch__registers_per_thread", 1),
    "grid": ("launch__grid_size", 1),
    "block": ("launch__block_size", 1),
    "dyn_smem_bytes": ("launch__shared_mem_per_block_dynamic", None),
}
STALLS = ["long_scoreboard", "short_scoreboard", "wait", "not_selected", "math_pipe_throttle", "mio_throttle", "lg_throttle",
          "no_instruction", "dispatch_stall", "barrier", "branch_resolving", "membar", "sleeping", "tex_throttle", "imc_miss", "drain"]
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6, "usecond": 1, "msecond": 1e3, "nsecond": 1e-3}


def ncu(*args):
    return subprocess.run(["ncu", *args], capture_output=True, text=True, check=True).stdout


def num(x):
    try:
        return float(x.replace(",", ""))
    except Exception:
        return None


def main():
    rep, out = sys.argv[1], sys.argv[2]
    labels = sys.argv[3].split(",") if len(sys.argv) > 3 else []
    rows = list(csv.reader(io.StringIO(ncu("-i", rep, "--page", "raw", "--csv"))))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {n: i for i, n in enumerate(hdr)}
    summary = []
    for k, r in enumerate(data):
        e = {"launch": k, "kernel": r[col["Kernel Name"]], "label": labels[k] if k < len(labels) else None}
        for key, (metric, scale) in KEYS.items():
            if metric not in col:
                continue
            v = num(r[col[metric]])
            if v is None:
                continue
            u = units[col[metric]]
            if key == "duration_us" or scale is None:
                v *= UNIT.get(u, 1)
            e[key] = v
        st = {}
        for s in STALLS:
            m = f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio"
            if m in col and num(r[col[m]]) is not None:
                st[s] = round(num(r[col[m]]), 3)
        e["stalls_per_issue"] = dict(sorted(st.items(), key=lambda kv: -kv[1])[:6])
        summary.append(e)
    with open(out + "_ncu_summary.json", "w") as f:
        json.dump({"report": os.path.basename(rep), "launches": summary}, f, indent=1)
        f.write("\n")

    # hot SASS of the first launch of every kernel
    os.makedirs(out + "_sass", exist_ok=True)
    seen = set()
    for k, r in enumerate(data):
        name = r[col["Kernel Name"]]
        short = re.sub(r"[^A-Za-z0-9_<>,]+", "_", re.sub(r"\(.*", "", name)).strip("_")
        if short in seen:
            continue
        seen.add(short)
        txt = ncu("-i", rep, "--page", "source", "--csv", "--launch-skip", str(k), "--launch-count", "1")
        srows = [x for x in csv.reader(io.StringIO(txt))]
        h = next((x for x in srows if x and x[0] == "Address"), None)
        if h is None:
            continue
        ia, isrc, ie, ismp, ithr = h.index("Address"), h.index("Source"), h.index("Instructions Executed"), h.index("# Samples"), h.index("Avg. Threads Executed")
        ins, addrs = [], set()
        for x in srows[srows.index(h) + 1:]:
            if len(x) > ie and x[ie].isdigit() and x[ia] not in addrs:
                addrs.add(x[ia])
                ins.append((x[isrc].strip(), int(x[ie]), int(x[ismp] or 0), x[ithr]))
        total = sum(i[1] for i in ins) or 1
        # contiguous regions with a similar executed count
        regions, cur = [], None
        for idx, (s, e, smp, thr) in enumerate(ins):
            if cur and abs(e - cur["e"]) <= max(0.15 * cur["e"], 50):
                cur["end"] = idx
                cur["sum"] += e
            else:
                if cur:
                    regions.append(cur)
                cur = {"start": idx, "end": idx, "e": e, "sum": e}
        if cur:
            regions.append(cur)
        with open(os.path.join(out + "_sass", short + ".txt"), "w") as f:
            f.write(f"{name}\nlaunch {k} of {os.path.basename(rep)}; {len(ins)} SASS instructions, {total} warp instructions executed\n")
            f.write("regions carrying >= 5 % of the executed warp instructions (executed count, stall samples, avg threads, instruction):\n")
            for reg in regions:
                if reg["sum"] < 0.05 * total:
                    continue
                n = reg["end"] - reg["start"] + 1
                f.write(f"\n--- {n} instructions x {reg['e']} executions = {100 * reg['sum'] / total:.1f} % ---\n")
                for s, e, smp, thr in ins[reg["start"]:reg["end"] + 1]:
                    f.write(f"{e:>10} {smp:>6} {thr:>5}  {s}\n")
    print("wrote", out + "_ncu_summary.json", "and", len(seen), "listings under", out + "_sass/")


if __name__ == "__main__":
    main()
