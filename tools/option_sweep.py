#!/usr/bin/env python
"""A/B engine tunables on one equilibrated state (one GPU, one gpurun call).

    python tools/option_sweep.py [--chains 100] [--tile 1x1] [--equil 6000] [--steps 4000] \
        skin_nm=0.3,0.4,0.5 beads_per_block=0,8,16

The workload (synth.hps_slab: C2 with the defaults, --tile 8x8 --chains 6400 = C5 on one GPU) is equilibrated once with
the default options; every combination of the listed option values then starts from that same state (positions and
velocities through the C ABI, same integrator seed), runs --steps MD steps twice (the first run settles the adaptive
rebuild interval) and prints one JSON line: MD steps/s from the device-timed second run, the per-kernel times of
ocm_profile_kernel, the rebuild interval and the list size.  Options are the context tunables of include/ocm_b200.h
("skin_nm", "nstlist", "threads_per_bead", "beads_per_block", "pair_kernel", "tile_blocks", "pdl", ...).
"""
import argparse
import itertools
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chains", type=int, default=100)
    ap.add_argument("--tile", default="1x1")
    ap.add_argument("--equil", type=int, default=6000)
    ap.add_argument("--steps", type=int, default=4000)
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("options", nargs="*", help="key=v1,v2,...")
    a = ap.parse_args()

    from opencafemol_b200 import synth, toml_input
    from opencafemol_b200.capi import product_library

    lib = product_library()  # raises without the CUDA extension: there is no CPU fallback
    tile = tuple(int(v) for v in a.tile.split("x"))
    spec = synth.hps_slab(n_chains=a.chains, tile=tile)
    ctx = toml_input.make_context(lib, spec, device=a.device)
    ctx.step(a.equil)
    st = ctx.get_state(positions=True, velocities=True)
    ctx.close()
    spec.positions, spec.velocities = st["positions"], st["velocities"]

    keys = [o.split("=")[0] for o in a.options]
    values = [[float(v) for v in o.split("=")[1].split(",")] for o in a.options]
    for combo in itertools.product(*values) if keys else [()]:
        ctx = toml_input.make_context(lib, spec, device=a.device)
        for k, v in zip(keys, combo):
            ctx.set_option(k, v)
        ctx.step(a.steps)
        ctx.step(a.steps)
        dev_s = ctx.counter("last_step_ms") * 1e-3
        line = {"options": dict(zip(keys, combo)), "beads": int(spec.mass.size), "md_steps_per_s": a.steps / dev_s,
                "us_per_step": 1e6 * dev_s / a.steps, "nstlist": ctx.counter("nstlist"),
                "list_entries": ctx.counter("pair_evals_per_step"), "rollbacks": ctx.counter("rollbacks")}
        for which, reps in (("force", 100), ("integrate", 100), ("rebuild", 20)):
            line[which + "_us"] = 1e3 * ctx.profile_kernel(which, reps)
        print(json.dumps(line), flush=True)
        ctx.close()


if __name__ == "__main__":
    main()
