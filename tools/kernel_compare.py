"""Rows vs tiles on one GPU: parity against the oracle (small case) and kernel timings on the C2 bench state.
usage: python tools/kernel_compare.py [tile=1] [n_chains=100] [key=value ...]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opencafemol_b200 import synth, toml_input as ti  # noqa: E402
from opencafemol_b200.capi import Library, product_library  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = product_library()
args = sys.argv[1:]
opts = [a.split("=") for a in args if "=" in a]


def ferr(f, fref):
    frms = np.sqrt((fref ** 2).sum(1).mean())
    return (np.linalg.norm(f - fref, axis=1) / np.maximum(np.linalg.norm(fref, axis=1), frms)).max()


ora_lib = Library(os.path.join(ROOT, "oracle", "libocm_oracle.so"), "oracle_")
for tile in (0, 1):
    spec = synth.hps_slab(n_chains=12)
    ctx = ti.make_context(lib, spec)
    ctx.set_option("pair_kernel", tile)
    ora = ti.make_context(ora_lib, spec)
    a, b = ctx.get_state(forces=True, energy=True), ora.get_state(forces=True, energy=True)
    same = all(np.array_equal(ctx.neighbor_pairs(t), ora.neighbor_pairs(t)) for t in range(ctx.num_pair_terms()))
    print(f"12 chains pair_kernel={tile}: force err {ferr(a['forces'], b['forces']):.2e} energies {a['energy_by_group'][:4]} vs {b['energy_by_group'][:4]} pair sets equal {same}", flush=True)

spec = synth.c2_equilibrated()
ora = ti.make_context(ora_lib, spec)
b = ora.get_state(forces=True, energy=True)
for tile in (0, 1):
    ctx = ti.make_context(lib, spec)
    ctx.set_option("pair_kernel", tile)
    for k, v in opts:
        ctx.set_option(k, float(v))
    a = ctx.get_state(forces=True, energy=True)
    print(f"C2 pair_kernel={tile}: force err {ferr(a['forces'], b['forces']):.2e} E {a['energy_by_group'][:4]} vs {b['energy_by_group'][:4]}", flush=True)
    print("   evals", ctx.counter("pair_evals_per_step"), "skin", ctx.get_option("skin_nm") if hasattr(ctx, "get_option") else "")
    for k in ("force", "integrate", "rebuild"):
        print("   ", k, "ms", ctx.profile_kernel(k, 50))
    t0 = time.time()
    ctx.step(2000)
    ctx.step(2000)
    print("    2000 steps:", ctx.counter("last_step_ms") / 2000, "ms/step, nstlist", ctx.counter("nstlist"), "rollbacks", ctx.counter("rollbacks"), flush=True)
    a = ctx.get_state(positions=True, energy=True)
    print("    E after", a["potential_energy"], a["kinetic_energy"])
