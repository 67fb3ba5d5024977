"""C5 (the compacted C2 state tiled 8 x 8, 1 043 200 beads) for profilers: a few dozen MD steps across one list rebuild."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opencafemol_b200 import synth, toml_input as ti  # noqa: E402
from opencafemol_b200.capi import product_library  # noqa: E402

ctx = ti.make_context(product_library(), synth.c2_equilibrated(tile=(8, 8)))
ctx.step(30)
print("ms/step", ctx.counter("last_step_ms") / 30)
if len(sys.argv) > 1:   # python tools/c5_profile_run.py time: a timed run and the per-kernel figures of bench.py's c5 record
    ctx.step(200)
    ctx.step(400)
    print("ms/step", ctx.counter("last_step_ms") / 400, "nstlist", ctx.counter("nstlist"))
    for k in ("force", "integrate", "rebuild"):
        print(k, "ms", ctx.profile_kernel(k, 20))
