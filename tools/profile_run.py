"""Short C2 run for profilers: equilibrate, then a few hundred MD steps.
usage: python tools/profile_run.py [equil_steps] [steps] [n_chains]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opencafemol_b200 import synth, toml_input  # noqa: E402
from opencafemol_b200.capi import product_library  # noqa: E402

equil = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
chains = int(sys.argv[3]) if len(sys.argv) > 3 else 100
lib = product_library()
spec = synth.c2_equilibrated() if chains == 100 and os.path.exists(synth.C2_STATE) else synth.hps_slab(n_chains=chains)
ctx = toml_input.make_context(lib, spec)
for k, v in [a.split("=") for a in sys.argv[4:]]:
    ctx.set_option(k, float(v))
ctx.step(equil)
r0 = ctx.counter("rollbacks")
ctx.step(steps)
print("ms/step", ctx.counter("last_step_ms") / steps, "nstlist", ctx.counter("nstlist"), "rollbacks", r0,
      ctx.counter("rollbacks"), "max_nbr", ctx.counter("max_neighbors"), "evals", ctx.counter("pair_evals_per_step"))
for k in ("force", "integrate", "rebuild"):
    print(k, "ms", ctx.profile_kernel(k, 50))
