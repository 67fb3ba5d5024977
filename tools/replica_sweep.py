"""A/B of the replica kernel's launch options: python tools/replica_sweep.py [steps]
(replica_carveout: shared-memory carve-out in per cent, -1 = what the resident blocks need; replica_block: threads per block)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_golden_spec  # noqa: E402

from opencafemol_b200 import capi, toml_input  # noqa: E402

lib = capi.product_library()
sh3 = load_golden_spec("sh3")
nsteps = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
cases = [(None, {}), (None, {"replica_carveout": 100}), (None, {"replica_block": 512}),
         (64, {}), (64, {"replica_carveout": 100}),
         (512, {}), (512, {"replica_carveout": 100}), (512, {"replica_block": 512}), (512, {"replica_carveout": 70}),
         (2048, {}), (2048, {"replica_carveout": 100})]
for reps, opts in cases:
    n = nsteps if reps is None else max(2000, nsteps // 10)
    ctx = toml_input.make_context(lib, sh3, replicas=reps)
    for k, v in opts.items():
        ctx.set_option(k, v)
    ctx.step(500)
    best = 0.0
    for _ in range(3):
        ctx.step(n)
        best = max(best, n / (ctx.counter("last_step_ms") * 1e-3))
    r = reps or 1
    print("replicas %5d %-28s %9.0f MD steps/s  %7.2f M replica-steps/s" % (r, str(opts), best, r * best / 1e6), flush=True)
    ctx.close()
