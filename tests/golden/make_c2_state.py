#!/usr/bin/env python
"""Generates opencafemol_b200/data/c2_state.npz: the synthetic C2 condensate (synth.hps_slab, 16 300 beads) after
`steps` Langevin steps from its serpentine start.

    python tests/golden/make_c2_state.py [steps=40000]

Why: the serpentine start has the bead density of the reference's sample file (1.12 nm^-3, 2.5 M in-cutoff pairs), but
under the HPS force field the slab compacts for tens of thousands of steps (3.1 M pairs after 3 000 steps, 4.5 M after
20 000, 4.76 M after 32 000, 4.79 M after 40 000), so a benchmark that starts there times a different amount of pair work depending on how long it warms up, and
a CPU arm that can only afford a few hundred steps would time a much lighter state than the GPU arm.  bench.py and
tools/profile_run.py therefore start both arms from this committed state.  It is an INPUT file (positions and velocities
as float32, nm and nm/ps); it is produced offline by the double-precision CPU restatement (oracle/) because that runs in
the build container — any integrator of the same system yields a statistically equivalent state.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402
from opencafemol_b200 import synth, toml_input  # noqa: E402
from opencafemol_b200.capi import Library  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
g.build_oracle()
lib = Library(g.ORACLE_SO, "oracle_")
spec = synth.hps_slab(n_chains=100)
ctx = toml_input.make_context(lib, spec)
t0 = time.time()
done = 0
while done < steps:
    n = min(1000, steps - done)
    ctx.step(n)
    done += n
    print(done, "steps,", len(ctx.neighbor_pairs(1)), "Ashbaugh-Hatch pairs inside the cutoff, %.0f s" % (time.time() - t0), flush=True)
st = ctx.get_state(positions=True, velocities=True)
out = os.path.join(ROOT, "opencafemol_b200", "data", "c2_state.npz")
os.makedirs(os.path.dirname(out), exist_ok=True)
np.savez_compressed(out, x=st["positions"].astype(np.float32), v=st["velocities"].astype(np.float32),
                    steps=np.int64(steps), pairs=np.int64(len(ctx.neighbor_pairs(1))))
print("wrote", out)
