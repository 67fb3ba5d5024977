import sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
from test_oracle_kat import _da_system
from opencafemol_b200.capi import Context, System, Library, product_library
import __graft_entry__ as g
eng = product_library(); ora_lib = Library(g.ORACLE_SO, "oracle_")
n = 16
for seed in (1,):
    ctxs = []
    for lib in (eng, ora_lib):
        rng = np.random.default_rng(100 + seed)
        s = System(lib, np.full(n, 60.0), None)
        pos = _da_system(s, rng, "3spn2_base_pair", False, n)
        c = Context(s, 300.0, 0.2, 0.002, seed=3)
        c.set_positions(pos)
        ctxs.append(c)
    ctx, ora = ctxs
    for step in range(12):
        a = ctx.get_state(positions=True, velocities=True, forces=True, energy=True)
        b = ora.get_state(positions=True, velocities=True, forces=True, energy=True)
        # forces of the oracle at the ENGINE's positions
        chk = Context.__new__(Context)
        ora2_pos = ora.get_state(positions=True)["positions"]
        ora.set_positions(a["positions"])
        fb = ora.get_state(forces=True, energy=True)
        ora.set_positions(ora2_pos)
        print(f"step {step}: dpos {np.abs(a['positions']-b['positions']).max():.3e} dvel {np.abs(a['velocities']-b['velocities']).max():.3e} "
              f"|F|max {np.abs(b['forces']).max():.3e} dF(same x) {np.abs(a['forces']-fb['forces']).max():.3e} E {a['potential_energy']:.6f} {fb['potential_energy']:.6f} "
              f"pairs {len(ctx.da_pairs(0))} {len(ora.da_pairs(0))}")
        ctx.step(1); ora.step(1)
