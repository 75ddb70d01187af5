#!/usr/bin/env python
"""Reads `ncu --set full` reports (run here, no GPU needed) and writes
  * profiles/r02_attn_store_traffic.json : {key: {"dram_bytes": read+write per launch, "source": report, ...}} -- what
    bench.py's `roofline.traffic` cites;
  * a markdown summary (duration, DRAM bytes, issue utilisation, instruction count, top stall reasons) on stdout.
    python scripts/ncu_traffic.py key=report.ncu-rep [key=report.ncu-rep ...] [--json profiles/r02_attn_store_traffic.json]"""
import csv
import json
import os
import subprocess
import sys

WANT = {"gpu__time_duration.sum": "duration_us", "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
        "smsp__inst_executed.sum": "warp_instructions", "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
        "launch__registers_per_thread": "registers", "launch__grid_size": "grid", "launch__block_size": "block",
        "launch__shared_mem_per_block_dynamic": "smem_dynamic", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed": "smem_wavefront_pct"}
UNIT = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "Kbyte/block": 1e3, "byte/block": 1.0, "register/thread": 1.0, "byte": 1.0, "us": 1.0, "ms": 1e3, "ns": 1e-3, "usecond": 1.0, "msecond": 1e3, "nsecond": 1e-3}


def read(report):
    raw = subprocess.run(["ncu", "-i", report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    launches = []
    for v in rows[2:]:
        d = {"kernel": v[hdr.index("Kernel Name")] if "Kernel Name" in hdr else ""}
        stalls = {}
        for i, h in enumerate(hdr):
            if h in WANT and v[i] != "":
                d[WANT[h]] = float(v[i].replace(",", "")) * UNIT.get(units[i], 1.0)
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and v[i] != "":
                stalls[h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = float(v[i])
        d["top_stalls"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:5])
        launches.append(d)
    return launches


def main():
    out_json = sys.argv[sys.argv.index("--json") + 1] if "--json" in sys.argv else None
    table = {}
    for a in sys.argv[1:]:
        if "=" not in a:
            continue
        key, rep = a.split("=", 1)
        ls = read(rep)
        d = ls[-1]
        n = len(ls)
        table[key] = {"dram_bytes": int(sum(l["dram_read"] + l["dram_write"] for l in ls) / n), "source": os.path.relpath(rep),
                      "kernel": d["kernel"][:90], "duration_us_under_ncu": round(sum(l["duration_us"] for l in ls) / n, 2),
                      "launches_averaged": n}
        print(f"### {key}: `{d['kernel'][:100]}`\n")
        print(f"| launches | duration under ncu (us) | DRAM read (MB) | DRAM write (MB) | DRAM % of peak | warp instructions | issue active % | "
              f"warps active % | tensor pipe % | shared wavefronts % | regs | grid x block | dyn smem (KB) |")
        print("|---|---|---|---|---|---|---|---|---|---|---|---|---|")
        print(f"| {n} | {d['duration_us']:.1f} | {d['dram_read'] / 1e6:.2f} | {d['dram_write'] / 1e6:.2f} | {d.get('dram_pct_of_peak', 0):.1f} | "
              f"{d.get('warp_instructions', 0) / 1e6:.1f} M | {d.get('issue_active_pct', 0):.1f} | {d.get('warps_active_pct', 0):.1f} | "
              f"{d.get('tensor_pipe_pct', 0):.1f} | {d.get('smem_wavefront_pct', 0):.1f} | {int(d.get('registers', 0))} | "
              f"{int(d.get('grid', 0))} x {int(d.get('block', 0))} | {d.get('smem_dynamic', 0) / 1e3:.1f} |")
        print("\ntop stall reasons (warps per issue-active cycle): " + ", ".join(f"{k} {v:.2f}" for k, v in d["top_stalls"].items()) + "\n")
    if out_json:
        prev = json.load(open(out_json)) if os.path.exists(out_json) else {}
        prev.update(table)
        json.dump(prev, open(out_json, "w"), indent=1)


if __name__ == "__main__":
    main()
