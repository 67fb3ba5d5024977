"""Force-kernel time of the tile path on the C2 bench state for a few launch shapes (one GPU)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opencafemol_b200 import synth, toml_input as ti  # noqa: E402
from opencafemol_b200.capi import product_library  # noqa: E402

lib = product_library()
spec = synth.c2_equilibrated()
for skin in (0.4, 0.2):
    for nb in (148 * 3, 148 * 4, 148 * 5, 148 * 10, 148 * 15, 148 * 20):
        ctx = ti.make_context(lib, spec)
        ctx.set_option("skin_nm", skin)
        ctx.set_option("tile_blocks", nb)
        ctx.get_state(energy=True)
        print(f"skin {skin} blocks {nb}: force {ctx.profile_kernel('force', 100) * 1e3:.2f} us, rebuild {ctx.profile_kernel('rebuild', 20) * 1e3:.1f} us, "
              f"lane evals {ctx.counter('pair_evals_per_step') / 1e6:.2f} M", flush=True)
