#!/usr/bin/env python
"""Condense an `ncu --page raw --csv` export into one small JSON per captured launch: duration, DRAM bytes, instruction count,
issue utilisation, occupancy, pipe utilisation and the stall-reason split.  Runs on the build container (no GPU needed).

    python tools/ncu_summary.py gpurun_out/r02_c2_raw.csv > profiles/r02_c2_kernels.json
"""
import csv
import json
import sys

KEYS = {
    "gpu__time_duration.sum": "duration_us",
    "dram__bytes_read.sum": "dram_bytes_read",
    "dram__bytes_write.sum": "dram_bytes_write",
    "smsp__inst_executed.sum": "warp_instructions",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct_of_64",
    "sm__cycles_active.avg": "sm_cycles_active_avg",
    "sm__cycles_elapsed.max": "sm_cycles_elapsed",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__grid_size": "grid_size",
    "launch__block_size": "block_size",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "pipe_fma_pct",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "pipe_alu_pct",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "pipe_lsu_pct",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "pipe_xu_pct",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active": "pipe_fp64_pct",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
}


def num(v, unit):
    try:
        x = float(v.replace(",", ""))
    except ValueError:
        return v
    scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ms": 1e3, "ns": 1e-3, "s": 1e6}.get(unit)
    return x * scale if scale else x


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    out = []
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        d = {"kernel": r[ix["Kernel Name"]].split("(")[0][:80]}
        for k, name in KEYS.items():
            if k in ix:
                d[name] = num(r[ix[k]], units[ix[k]])
        stalls = {}
        for h in hdr:
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                v = num(r[ix[h]], "")
                if isinstance(v, float) and v > 0.02:
                    stalls[h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = round(v, 3)
        d["warp_cycles_stalled_per_issue"] = stalls
        if "dram_bytes_read" in d and "dram_bytes_write" in d:
            d["dram_bytes"] = d["dram_bytes_read"] + d["dram_bytes_write"]
        out.append(d)
    json.dump(out, sys.stdout, indent=1)


if __name__ == "__main__":
    main()
