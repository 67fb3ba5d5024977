"""Where the end-to-end loop of bench.py spends its time: per bench step the wall time of upload / step / download, the device
time of the step, rollbacks, rebuilds and the rebuild interval.  python tools/e2e_probe.py [bench_steps]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from opencafemol_b200 import synth, toml_input  # noqa: E402
from opencafemol_b200.capi import product_library  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
ctx = toml_input.make_context(product_library(), synth.c2_equilibrated())
for _ in range(3):
    ctx.step(1000)
keys = ("rollbacks", "rebuilds", "nstlist", "kernel_launches")
st = ctx.get_state(positions=True, velocities=True)
pos, vel = st["positions"], st["velocities"]
for mode in ("resident", "e2e", "e2e_blocking"):
    for k in range(steps):
        c0 = {q: ctx.counter(q) for q in keys}
        t0 = time.perf_counter()
        if mode == "e2e":
            ctx.set_state_async(pos, vel)
        elif mode == "e2e_blocking":
            ctx.set_positions(pos); ctx.set_velocities(vel)
        t1 = time.perf_counter()
        ctx.step(1000)
        t2 = time.perf_counter()
        if mode == "e2e":
            st = ctx.get_state_end(ctx.get_state_begin(positions=True, velocities=True, energy=True))
        elif mode == "e2e_blocking":
            st = ctx.get_state(positions=True, velocities=True, energy=True)
        if mode != "resident":
            pos, vel = st["positions"], st["velocities"]
        t3 = time.perf_counter()
        c1 = {q: ctx.counter(q) for q in keys}
        print(f"{mode:13s} {k}: upload {1e3 * (t1 - t0):6.2f} step {1e3 * (t2 - t1):6.2f} (device {ctx.counter('last_step_ms'):6.2f}) download {1e3 * (t3 - t2):6.2f} ms  "
              f"rollbacks +{c1['rollbacks'] - c0['rollbacks']:.0f} rebuilds +{c1['rebuilds'] - c0['rebuilds']:.0f} nstlist {c1['nstlist']:.0f}", flush=True)
