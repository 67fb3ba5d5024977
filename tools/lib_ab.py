#!/usr/bin/env python
"""A/B of engine BUILDS (compile-time variants) on the committed C2 bench state, one GPU, one gpurun call:

    python tools/lib_ab.py opencafemol_b200/lib/libocm_b200.so scratch/libocm_mb7.so ...

Every library runs in its own process: 3000 MD steps to settle the rebuild interval, then 3000 device-timed steps; prints MD steps/s,
the per-kernel times of ocm_profile_kernel and, for the first library as the reference, the largest force difference of the others
at the start state (a wrong build must not look fast)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r"""
import sys, json
sys.path.insert(0, %r)
import numpy as np
from opencafemol_b200 import synth, toml_input
from opencafemol_b200.capi import Library
lib = Library(sys.argv[1], "ocm_")
spec = synth.c2_equilibrated()
ctx = toml_input.make_context(lib, spec)
f0 = ctx.get_state(forces=True)["forces"]
np.save(sys.argv[2], f0)
ctx.step(3000)
ctx.step(3000)
ms = ctx.counter("last_step_ms")
out = {"lib": sys.argv[1], "md_steps_per_s": 3000.0 / (ms * 1e-3), "nstlist": ctx.counter("nstlist")}
for k in ("force", "integrate", "rebuild"):
    out[k + "_ms"] = ctx.profile_kernel(k, 50)
print(json.dumps(out))
""" % ROOT


def main():
    import numpy as np

    ref = None
    for k, path in enumerate(sys.argv[1:]):
        fpath = os.path.join("/tmp", f"lib_ab_{k}.npy")
        r = subprocess.run([sys.executable, "-c", CHILD, os.path.abspath(path), fpath], capture_output=True, text=True, timeout=600)
        if r.returncode != 0:
            print(json.dumps({"lib": path, "error": r.stderr[-400:]}))
            continue
        out = json.loads(r.stdout.strip().splitlines()[-1])
        f = np.load(fpath)
        if ref is None:
            ref = f
        else:
            frms = np.sqrt((ref ** 2).sum(1).mean())
            out["max_force_diff_rel"] = float((np.linalg.norm(f - ref, axis=1) / np.maximum(np.linalg.norm(ref, axis=1), frms)).max())
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
