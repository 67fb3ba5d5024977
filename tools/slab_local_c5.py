"""The C5 workload (compacted C2 state tiled 8 x 8, 1 043 200 beads) as 8 slab ranks on ONE GPU (host threads, in-process
transport): per-kernel durations of a slab rank's rebuild and step for an ncu launch list — what a rank launches at N = 8, without
the NVLink transfers.
    ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s <skip> -c <count> --csv --log-file out.csv python tools/slab_local_c5.py [world]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from opencafemol_b200 import synth  # noqa: E402
from opencafemol_b200.capi import product_library  # noqa: E402
from test_gpu_slab import SlabRun  # noqa: E402

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
lib = product_library()
spec = synth.c2_equilibrated(tile=(8, 8))
run = SlabRun(lib, spec, world, 1, nstlist=24)
run.each(lambda c: c.step(48))
n0 = [c.counter("kernel_launches") for c in run.ctx]
run.each(lambda c: c.step(48))
print("launches per rank in 48 steps (2 windows):", [c.counter("kernel_launches") - a for c, a in zip(run.ctx, n0)])
print("owned", [c.counter("owned_beads") for c in run.ctx])
