"""C2 tuning sweep on one GPU: force-kernel timings for library variants x threads_per_bead x beads_per_block.
usage: python tools/sweep_c2.py [equil_steps] [lib paths...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opencafemol_b200 import synth, toml_input  # noqa: E402
from opencafemol_b200.capi import PRODUCT_LIB, Library  # noqa: E402

equil = int(sys.argv[1]) if len(sys.argv) > 1 else 6000
paths = sys.argv[2:] or [PRODUCT_LIB]
spec = synth.hps_slab(n_chains=100)
for path in paths:
    lib = Library(path, "ocm_")
    ctx = toml_input.make_context(lib, spec)
    ctx.step(equil)
    print(os.path.basename(path), "nstlist", ctx.counter("nstlist"), "max_nbr", ctx.counter("max_neighbors"), "evals",
          ctx.counter("pair_evals_per_step"))
    for staging in (0, 1):
        ctx.set_option("row_staging", staging)
        for tpb in (16, 32):
            for chunk in (0, 4, 8, 16):
                ctx.set_option("threads_per_bead", tpb)
                ctx.set_option("beads_per_block", chunk)
                print(f"  staging {staging} tpb {tpb:2d} chunk {chunk:3d} force_ms {ctx.profile_kernel('force', 100):.4f}", flush=True)
    ctx.set_option("row_staging", 0)
    ctx.set_option("threads_per_bead", 0)
    ctx.set_option("beads_per_block", 0)
    print("  integrate ms", ctx.profile_kernel("integrate", 100), "rebuild ms", ctx.profile_kernel("rebuild", 30))
    ctx.step(2000)
    print("  us/step", ctx.counter("last_step_ms") / 2000 * 1e3, "nstlist", ctx.counter("nstlist"))
    del ctx

# fixed overhead of the force launch: same bead count, a cutoff so short that rows are nearly empty
lib = Library(paths[0], "ocm_")
for cutoff in (1.2, 3.0):
    sp = synth.hps_slab(n_chains=100, cutoff=cutoff)
    ctx = toml_input.make_context(lib, sp)
    ctx.step(200)
    print(f"cutoff ratio {cutoff}: evals {ctx.counter('pair_evals_per_step'):.0f} force_ms {ctx.profile_kernel('force', 100):.4f} "
          f"integrate_ms {ctx.profile_kernel('integrate', 100):.4f} rebuild_ms {ctx.profile_kernel('rebuild', 30):.4f}")
    ctx.step(2000)
    print("   us/step", ctx.counter("last_step_ms") / 2000 * 1e3, "nstlist", ctx.counter("nstlist"))
    del ctx
