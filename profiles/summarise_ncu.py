#!/usr/bin/env python
"""profiles/summarise_ncu.py REPORT.ncu-rep [sites_per_launch] -- key metrics, stall ratios, per-opcode executed
instructions and shared-memory wavefronts of the (single) kernel in an `ncu --set full --import-source on` report.
Reads the report with `ncu -i ... --page raw|source --csv`; prints JSON.  This is how the r01_* summaries were made."""
import csv
import json
import subprocess
import sys
from collections import defaultdict

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active",
        "smsp__warps_eligible.avg.per_cycle_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__cycles_active.avg"]


def page(rep, which):
    out = subprocess.run(["ncu", "-i", rep, "--page", which, "--csv"] + (["--print-source", "sass"] if which == "source" else []),
                         capture_output=True, text=True).stdout
    return list(csv.reader(out.splitlines()))


def main():
    rep = sys.argv[1]
    sites = float(sys.argv[2]) if len(sys.argv) > 2 else None
    raw = page(rep, "raw")
    hdr, units, vals = raw[0], raw[1], raw[2]
    res = {"report": rep.split("/")[-1], "kernel": vals[hdr.index("Kernel Name")]}
    for k in KEYS:
        if k in hdr:
            res[k] = "%s %s" % (vals[hdr.index(k)], units[hdr.index(k)])
    stalls = {h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]: float(vals[i])
              for i, h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")}
    res["stalled_warps_per_issue_active"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:10])
    src = page(rep, "source")
    h = src[1]
    ie, so = h.index("Instructions Executed"), h.index("Source")
    wf = h.index("L1 Wavefronts Shared") if "L1 Wavefronts Shared" in h else None   # absent when the kernel uses no shared memory
    ops, wfs, total = defaultdict(float), 0.0, 0.0
    for r in src[2:]:
        if len(r) <= max(ie, so, wf or 0):
            continue
        if r[ie] == "Instructions Executed":   # a second captured launch starts here: the summary is of the first
            break
        toks = r[so].split()
        op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")
        n = float(r[ie] or 0)
        ops[op.split(".")[0]] += n
        total += n
        wfs += float(r[wf] or 0) if wf is not None else 0.0
    res["warp_instructions_executed"] = total
    res["shared_wavefronts"] = wfs
    if sites:
        per = 32.0 / sites
        res["per_32_sites"] = {"instructions": round(total * per, 1), "shared_wavefronts": round(wfs * per, 1),
                               "by_opcode": {k: round(v * per, 1) for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:24]}}
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
