#!/usr/bin/env python
"""Turn the ncu artefacts a gpurun call brought back (gpurun_out/) into the small, tracked
summaries under profiles/:

  profiles/<tag>_launches.csv      the launch list (kernel, device time) of `bench.py`
  profiles/<tag>_summary.md        launch shares, key raw metrics, instruction mix, stall reasons
  profiles/ncu_summary.json        {"dram_bytes_per_launch": ...} read by bench.py for `traffic`

usage: python scripts/summarize_ncu.py <tag> [aux]    (expects prof_<tag>.ncu-rep, launches_<tag>.csv;
       `aux`: a secondary kernel, summary json goes to profiles/<tag>_ncu.json)
"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "sass__inst_executed_local_loads",
        "sass__inst_executed_local_stores", "smsp__cycles_active.avg",
        "sm__cycles_elapsed.max", "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum"]


def export(tag, page):
    rep = os.path.join(G, f"prof_{tag}.ncu-rep")
    out = os.path.join(G, f"prof_{tag}_{page}.csv")
    # the CSV pages may have been exported on the GPU box already (reports of several captures exceed gpurun's
    # 64 MiB return limit): then the .ncu-rep is not needed here
    if os.path.exists(out) and not os.path.exists(rep):
        return out
    if not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(rep):
        with open(out, "w") as fh:
            subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], stdout=fh,
                           stderr=subprocess.DEVNULL, check=True)
    return out


def main():
    tag = sys.argv[1]
    os.makedirs(P, exist_ok=True)
    md = [f"# ncu summary `{tag}`", ""]
    # ---- launch list ----
    lpath = os.path.join(G, f"launches_{tag}.csv")
    if os.path.exists(lpath):
        rows = [r for r in csv.reader(open(lpath)) if len(r) > 10]
        hdr = rows[0]
        ix = {h: i for i, h in enumerate(hdr)}
        agg = collections.OrderedDict()
        with open(os.path.join(P, f"{tag}_launches.csv"), "w") as fh:
            fh.write("id,kernel,grid,block,gpu_time_ns\n")
            for r in rows[1:]:
                k = r[ix["Kernel Name"]]
                v = float(r[ix["Metric Value"]].replace(",", ""))
                fh.write(f'{r[ix["ID"]]},"{k}","{r[ix["Grid Size"]]}","{r[ix["Block Size"]]}",{v:.0f}\n')
                a = agg.setdefault(k, [0, 0.0])
                a[0] += 1
                a[1] += v
        tot = sum(v[1] for v in agg.values())
        md += ["## Launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`, "
               "`bench.py --steps 3 --warmup 3 --no-cpu-baseline`)", "",
               "Cold-cache, serialised times: compare shares, not absolutes.  `fp64_peak_kernel` is the "
               "roofline-denominator microbenchmark that bench.py runs OUTSIDE the timed region; inside "
               "the timed step the evaluation kernel is the only launch (share of the step: 100 %).", "",
               "| launches | total us | share | kernel |", "|---|---|---|---|"]
        for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
            md.append(f"| {n} | {t/1e3:.1f} | {100*t/tot:.1f} % | `{k[:100]}` |")
        md.append("")
    # ---- raw page ----
    raw = list(csv.reader(open(export(tag, "raw"))))
    hdr, units = raw[0], raw[1]
    summary = {}
    for r in raw[2:]:
        name = r[hdr.index("Kernel Name")]
        md += [f"## `ncu --set full` — `{name[:110]}`", "", "| metric | unit | value |", "|---|---|---|"]
        vals = dict(zip(hdr, r))
        for k in KEYS:
            if k in vals:
                md.append(f"| {k} | {units[hdr.index(k)]} | {vals[k]} |")

        def byt(k):
            v = float(vals[k].replace(",", ""))
            u = units[hdr.index(k)].lower()
            return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]
        dr, dw = byt("dram__bytes_read.sum"), byt("dram__bytes_write.sum")
        summary = {"kernel": name, "dram_bytes_per_launch": dr + dw, "dram_read": dr, "dram_write": dw,
                   "gpu_time_us": vals["gpu__time_duration.sum"], "tag": tag,
                   "fp64_pipe_pct_active": vals.get("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
                   "regs": vals.get("launch__registers_per_thread")}
        md += ["", f"DRAM traffic per launch: {dr/1e6:.1f} MB read + {dw/1e6:.1f} MB written = "
               f"{(dr+dw)/1e6:.1f} MB.", ""]
    # ---- source page: instruction mix and stall reasons (one section per captured launch; the last
    # capture of every distinct kernel is summarised) ----
    src = list(csv.reader(open(export(tag, "source"))))
    sections, cur = collections.OrderedDict(), None
    for r in src:
        if r and r[0] == "Kernel Name":
            cur = r[1]
            sections[cur] = []
        elif cur is not None:
            sections[cur].append(r)
    for kname, rows in sections.items():
        hdr = rows[0]
        ix = {h: i for i, h in enumerate(hdr)}
        st = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        mix, stall = collections.Counter(), collections.Counter()
        tot_inst = 0
        for r in rows[1:]:
            if len(r) < len(hdr):
                continue
            toks = [t for t in r[ix["Source"]].split() if not t.startswith("@")]
            op = toks[0].split(".")[0] if toks else "?"
            n = int(r[ix["Instructions Executed"]] or 0)
            mix[op] += n
            tot_inst += n
            for h in st:
                stall[h] += int(r[ix[h]] or 0)
        md += [f"## Instruction mix — `{kname[:110]}` (warp-level instructions executed, source page)", "",
               "| opcode | executed | share |", "|---|---|---|"]
        for k, v in mix.most_common(12):
            md.append(f"| {k} | {v} | {100*v/max(tot_inst, 1):.1f} % |")
        fp = sum(v for k, v in mix.items() if k in ("DFMA", "DMUL", "DADD"))
        md += ["", f"FP64 arithmetic = {100*fp/max(tot_inst, 1):.1f} % of all issued warp instructions.", "",
               "## Warp stall sampling (all samples)", "", "| reason | samples | share |", "|---|---|---|"]
        sm_ = max(sum(stall.values()), 1)
        for k, v in stall.most_common(8):
            md.append(f"| {k} | {v} | {100*v/sm_:.1f} % |")
        md.append("")
        summary["inst_mix"] = dict(mix.most_common(8))
        summary["stalls"] = dict(stall.most_common(6))
    with open(os.path.join(P, f"{tag}_summary.md"), "w") as fh:
        fh.write("\n".join(md) + "\n")
    # a second argument marks a capture of a SECONDARY kernel: its json does not feed bench.py's `traffic`
    with open(os.path.join(P, "ncu_summary.json" if len(sys.argv) < 3 else f"{tag}_ncu.json"), "w") as fh:
        json.dump(summary, fh, indent=1)
    print("\n".join(md))


if __name__ == "__main__":
    main()
