"""Whole-step replica kernel (replica_kernels.cuh) against the general path: parity with the oracle and steps/s for one SH3
replica (C1) and the 512-replica ensemble (C4).  usage: python tools/replica_bench.py [steps]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_golden_spec  # noqa: E402

from opencafemol_b200 import capi, toml_input  # noqa: E402

lib = capi.product_library()
ora_lib = capi.Library(os.path.join(ROOT, "oracle", "libocm_oracle.so"), "oracle_")
sh3 = load_golden_spec("sh3")
nsteps = int(sys.argv[1]) if len(sys.argv) > 1 else 20000


def ferr(a, b):
    d = np.linalg.norm(a - b, axis=1)
    return (d / np.maximum(np.linalg.norm(b, axis=1), np.sqrt((b ** 2).sum(1).mean()))).max()


for reps in (None, 5):
    ctx = toml_input.make_context(lib, sh3, replicas=reps)
    ora = toml_input.make_context(ora_lib, sh3, replicas=reps)
    ctx.set_option("replica_kernel", 1)
    print("replicas", reps, "replica_kernel in use:", ctx.get_option("replica_kernel"))
    for k in (1, 19, 80):
        ctx.step(k); ora.step(k)
        a = ctx.get_state(positions=True, velocities=True, forces=True); b = ora.get_state(positions=True, velocities=True, forces=True)
        print("  after +%d steps: max |dx| %.3e nm, max |dv| %.3e, force err (own states) %.3e" % (
            k, np.abs(a["positions"] - b["positions"]).max(), np.abs(a["velocities"] - b["velocities"]).max(), ferr(a["forces"], b["forces"])))
    ctx.close(); ora.close()

for reps, n in ((None, nsteps), (512, max(2000, nsteps // 10)), (2048, max(2000, nsteps // 10))):
    for opt in (0, 1):
        ctx = toml_input.make_context(lib, sh3, replicas=reps)
        ctx.set_option("replica_kernel", opt)
        ctx.step(500)
        ctx.get_state(positions=True)
        t0 = time.perf_counter()
        ctx.step(n)
        ctx.get_state(positions=True)
        wall = time.perf_counter() - t0
        dev = ctx.counter("last_step_ms") * 1e-3
        r = reps or 1
        print("replicas %5d replica_kernel %d: %.0f MD steps/s, %.2f M replica-steps/s (device), wall %.0f steps/s" % (
            r, opt, n / dev, r * n / dev / 1e6, n / wall))
        ctx.close()
