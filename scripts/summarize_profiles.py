#!/usr/bin/env python
"""Turn the raw profiler outputs in gpurun_out/ into the committed summaries under profiles/.
  python scripts/summarize_profiles.py <launches.csv> <step_table.json> <ncu-rep or ''> <out.md> [title]"""
import collections, csv, json, re, subprocess, sys

launch_csv, table_json, rep, out = sys.argv[1:5]
title = sys.argv[5] if len(sys.argv) > 5 else "profile"
md = [f"# {title}\n"]
if launch_csv:
    lines = [l for l in open(launch_csv) if not l.startswith("==")]
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", "")); u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
        name = re.sub(r"\(.*", "", row["Kernel Name"]); name = re.sub(r"^void ", "", name)[:90]
        tot[name] += v; cnt[name] += 1
    T = sum(tot.values())
    ours = sum(v for k, v in tot.items() if k.startswith("skp::"))
    md.append("## ncu launch list of ONE eager optimizer step (cfg2, N=77, R=128, fp32 trunk on tcgen05)\n")
    md.append("`ncu --nvtx --nvtx-include \"skp_step\" --metrics gpu__time_duration.sum --clock-control none --csv python scripts/profile_step.py`  "
              "(cold-cache, serialised launches: compare SHARES).\n")
    md.append(f"total {T/1e3:.1f} ms over {sum(cnt.values())} launches; libskp_b200 kernels: {ours/1e3:.1f} ms = {100*ours/T:.1f}% of GPU time\n")
    md.append("| us | share | launches | kernel |\n|---:|---:|---:|---|")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:40]:
        md.append(f"| {v:.0f} | {100*v/T:.1f}% | {cnt[k]} | `{k}` |")
    md.append("")
if table_json:
    d = json.load(open(table_json))
    T = d["total_gpu_us"]
    md.append("## torch.profiler (CUPTI) GPU time of one eager step, warm caches\n")
    md.append(f"total {T/1e3:.1f} ms\n\n| us | share | calls | kernel |\n|---:|---:|---:|---|")
    for r in d["kernels"][:40]:
        md.append(f"| {r['us']:.0f} | {100*r['us']/T:.1f}% | {r['calls']} | `{r['kernel'][:90]}` |")
    md.append("")
if rep:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "smsp__inst_executed.sum", "sm__issue_active.avg.pct_of_peak_sustained_elapsed"]
    keys = [k for k in keys if k in hdr]
    md.append(f"## ncu --set full ({rep.split('/')[-1]})\n")
    md.append("| kernel | " + " | ".join(f"{k} [{units[hdr.index(k)]}]" for k in keys) + " |")
    md.append("|---|" + "---:|" * len(keys))
    seen = collections.Counter()
    for r in rows[2:]:
        name = re.sub(r"\(.*", "", r[hdr.index("Kernel Name")])[:60]
        grid = r[hdr.index("launch__grid_size")]
        seen[(name, grid)] += 1
        if seen[(name, grid)] > 1:
            continue
        md.append(f"| `{name}` | " + " | ".join(r[hdr.index(k)] for k in keys) + " |")
    md.append("")
open(out, "w").write("\n".join(md))
print("wrote", out)
