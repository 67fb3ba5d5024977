#!/usr/bin/env python
"""Instruction budget of one profiled kernel from an `ncu --set full --import-source on` report.

    python tools/sass_budget.py gpurun_out/r01d_force.ncu-rep [--top 25] [--dump A:B]

Reads the report's source page (`ncu -i REP --page source --csv`: one row per SASS instruction with "Instructions
Executed" and the warp-stall samples) and prints, for the first kernel in the report,
  * the opcode mix (share of executed warp instructions per SASS mnemonic),
  * the stall-reason split of the samples,
  * the straight-line blocks (runs of instructions with the same execution count) ranked by executed instructions —
    the loop bodies; divide a block's count by the loop's trip count to get instructions per unit of work,
  * with --dump A:B the SASS of instruction indices [A, B) with execution counts and samples.
This is how profiles/r01d_force_sass_budget.md was produced.  Runs on the build container (no GPU needed).
"""
import argparse
import collections
import csv
import io
import re
import subprocess


def load(report, kernel_re=None):
    out = subprocess.run(["ncu", "-i", report, "--page", "source", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    # a report may hold several launches: one section per launch, each starting with its "Kernel Name" line
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] or [0]
    pick = next((i for i in starts if re.search(kernel_re, rows[i][1])), starts[0]) if kernel_re else starts[0]
    end = next((i for i in starts if i > pick), len(rows))
    rows = rows[pick:end]
    h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    kernel = next((r[1] for r in rows[:h] if r and r[0] == "Kernel Name"), "?")
    hdr = rows[h]
    ix = {name: i for i, name in enumerate(hdr)}
    data = [r for r in rows[h + 1:] if len(r) == len(hdr)]
    return kernel, hdr, ix, data


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--top", type=int, default=25)
    ap.add_argument("--dump", default=None, help="A:B = print the SASS of instruction indices [A, B)")
    ap.add_argument("--kernel", default=None, help="regular expression: the first captured launch whose kernel name matches")
    a = ap.parse_args()
    kernel, hdr, ix, data = load(a.report, a.kernel)
    ex = [int(r[ix["Instructions Executed"]] or 0) for r in data]
    sm = [int(r[ix["# Samples"]] or 0) for r in data]
    tot, stot = sum(ex), max(sum(sm), 1)
    print(f"kernel: {kernel}\n{len(data)} SASS instructions, {tot} warp instructions executed, {sum(sm)} stall samples\n")

    ops = collections.Counter()
    for r, e in zip(data, ex):
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[ix["Source"]])
        ops[m.group(2) if m else "?"] += e
    print("opcode mix:")
    for op, c in ops.most_common(24):
        print(f"  {op:10s} {c:11d} {100.0 * c / tot:5.1f} %")

    stalls = {n: sum(int(r[ix[n]] or 0) for r in data) for n in hdr if n.startswith("stall_") and "Not Issued" not in n}
    st = max(sum(stalls.values()), 1)
    print("\nstall reasons (all samples):")
    for n, c in sorted(stalls.items(), key=lambda kv: -kv[1])[:10]:
        print(f"  {n:24s} {c:6d} {100.0 * c / st:5.1f} %")

    blocks, cur = [], [0]
    for k in range(1, len(data)):
        if abs(ex[k] - ex[k - 1]) > 0.02 * max(ex[k], ex[k - 1], 1):
            blocks.append(cur)
            cur = []
        cur.append(k)
    blocks.append(cur)
    ranked = sorted(blocks, key=lambda b: -sum(ex[k] for k in b))
    print("\nstraight-line blocks by executed instructions:")
    print("  first   len   executions   warp instr   share  samples")
    for b in ranked[:a.top]:
        e, s = sum(ex[k] for k in b), sum(sm[k] for k in b)
        print(f"  {b[0]:5d} {len(b):5d} {ex[b[0]]:12d} {e:12d} {100.0 * e / tot:6.1f} % {100.0 * s / stot:5.1f} %")

    if a.dump:
        lo, hi = (int(v) for v in a.dump.split(":"))
        print()
        for k in range(lo, min(hi, len(data))):
            print(f"{k:6d} {ex[k]:9d} {sm[k]:4d}  {data[k][ix['Source']].strip()}")


if __name__ == "__main__":
    main()
