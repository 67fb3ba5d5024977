"""Statistical parity (BASELINE.json north_star: "Long Langevin runs: equilibrium observables (Rg, native-contact fraction, slab
density profile) agree within statistical error"; "Short NVE runs: energy drift matched").

The engine (FP32 pair arithmetic, its own noise realisations) and the CPU oracle (double precision) are run from INDEPENDENT seeds,
four of each, and the observables are compared as ensemble means: |mean_engine - mean_oracle| <= 3 combined standard errors (plus a
small absolute floor where the seed-to-seed spread happens to be tiny).  Sizes are chosen so that the oracle part takes well under a
minute on the GPU box's host cores."""
import numpy as np
import pytest

from conftest import load_golden_spec
from opencafemol_b200 import synth, toml_input as ti

pytestmark = pytest.mark.gpu
SEEDS = 4


def agree(a, b, floor):
    """ensemble means of two sets of independent runs within 3 standard errors (or an absolute floor)"""
    a, b = np.asarray(a, float), np.asarray(b, float)
    se = np.sqrt(a.var(ddof=1) / a.size + b.var(ddof=1) / b.size)
    return abs(a.mean() - b.mean()) <= max(3.0 * se, floor), (a.mean(), b.mean(), se)


def radius_of_gyration(x, m):
    c = (x * m[:, None]).sum(0) / m.sum()
    return float(np.sqrt((m * ((x - c) ** 2).sum(1)).sum() / m.sum()))


def test_sh3_rg_and_native_contacts(engine, oracle):
    """sample/toml_input/sh3_AICG2+.toml at 300 K: radius of gyration and fraction of native contacts (Go contacts closer than
    1.2 r0), time-averaged over the second half of 40 000-step runs, 4 engine runs (one replica batch) against 4 oracle runs."""
    spec = load_golden_spec("sh3")
    n, m = spec.mass.size, spec.mass
    go = next(t for t in spec.terms if t.kind == "go_contact")
    ij, r0 = np.asarray(go.args["ij"]).reshape(-1, 2), np.asarray(go.args["r0"])

    def observables(x):
        d = np.linalg.norm(x[ij[:, 0]] - x[ij[:, 1]], axis=1)
        return radius_of_gyration(x, m), float((d < 1.2 * r0).mean())

    def sample(step_fn, state_fn, replicas):
        rg, q = [[] for _ in range(replicas)], [[] for _ in range(replicas)]
        step_fn(20000)
        for _ in range(50):
            step_fn(400)
            x = state_fn().reshape(replicas, n, 3)
            for r in range(replicas):
                a, b = observables(x[r])
                rg[r].append(a); q[r].append(b)
        return [np.mean(v) for v in rg], [np.mean(v) for v in q]

    eng = ti.make_context(engine, spec, seed=1000, replicas=SEEDS)   # replica r runs with seed 1000 + r
    eng.set_velocities_to_temperature(spec.temperature, 77)
    rg_e, q_e = sample(eng.step, lambda: eng.get_state(positions=True)["positions"], SEEDS)
    rg_o, q_o = [], []
    for s in range(SEEDS):
        ora = ti.make_context(oracle, spec, seed=5000 + s)
        ora.set_option("threads", 1)   # 56 beads: one thread is the fastest oracle (19 k steps/s)
        ora.set_velocities_to_temperature(spec.temperature, 900 + s)
        a, b = sample(ora.step, lambda: ora.get_state(positions=True)["positions"], 1)
        rg_o += a; q_o += b
    ok, info = agree(rg_e, rg_o, 0.004)
    assert ok, f"Rg (engine mean, oracle mean, SE) = {info}"
    ok, info = agree(q_e, q_o, 0.01)
    assert ok, f"native-contact fraction = {info}"
    assert 0.8 < np.mean(rg_e) < 1.3 and np.mean(q_e) > 0.7   # a folded SH3 domain (Rg ~ 1 nm), not a blown-up one


def test_hps_slab_density_profile(engine, oracle):
    """6-chain HPS slab (978 beads): bead density along z after 1 500 Langevin steps from the common start, 4 noise
    realisations each; every occupied bin of the 24-bin profile agrees within 3 SE, and so do the slab's second moment and
    the potential energy per bead."""
    spec = synth.hps_slab(n_chains=6)
    lz = spec.box[2]
    edges = np.linspace(0, lz, 25)

    def run(lib, seed):
        ctx = ti.make_context(lib, spec, seed=seed)
        ctx.set_velocities_to_temperature(spec.temperature, seed + 17)
        ctx.step(1500)
        st = ctx.get_state(positions=True, energy=True)
        z = st["positions"][:, 2]
        zc = z - lz * np.floor(z / lz)
        h = np.histogram(zc, edges)[0] / z.size
        return h, float(z.var()), st["potential_energy"] / z.size

    eng = [run(engine, 100 + s) for s in range(SEEDS)]
    ora = [run(oracle, 700 + s) for s in range(SEEDS)]
    he, ho = np.array([r[0] for r in eng]), np.array([r[0] for r in ora])
    for b in range(24):
        if he[:, b].mean() + ho[:, b].mean() < 1e-3:
            continue
        ok, info = agree(he[:, b], ho[:, b], 0.004)
        assert ok, f"density bin {b}: {info}"
    ok, info = agree([r[1] for r in eng], [r[1] for r in ora], 0.01)
    assert ok, f"second moment of z: {info}"
    ok, info = agree([r[2] for r in eng], [r[2] for r in ora], 0.01)
    assert ok, f"potential energy per bead: {info}"


def test_nve_drift_matches_the_oracle(engine, oracle):
    """gamma = 0 (leap-frog): total-energy drift of the engine against the oracle's on the SAME start over 1 600 steps.  Both
    integrate the same equations with the same dt, so the O(dt^2) fluctuation is common and the secular drift of either must
    be a small fraction of the kinetic energy; the two drifts are compared directly."""
    spec = synth.hps_slab(n_chains=6)
    spec.friction = 0.0
    tot = {}
    for name, lib in (("engine", engine), ("oracle", oracle)):
        ctx = ti.make_context(lib, spec)
        ctx.set_velocities_to_temperature(spec.temperature, 5)
        ctx.step(400)  # past the first relaxation of the serpentine start
        e = []
        for _ in range(16):
            ctx.step(100)
            st = ctx.get_state(energy=True)
            e.append(st["potential_energy"] + st["kinetic_energy"])
        tot[name] = (np.array(e), st["kinetic_energy"])
    (ee, ke), (eo, _) = tot["engine"], tot["oracle"]
    t = np.arange(ee.size)
    drift_e, drift_o = np.polyfit(t, ee, 1)[0], np.polyfit(t, eo, 1)[0]   # kJ/mol per 100 steps
    assert abs(drift_o) * t.size < 2e-3 * ke, "the oracle itself conserves energy"
    assert abs(drift_e - drift_o) * t.size < 2e-3 * ke, (drift_e, drift_o, ke)
    assert np.abs((ee - ee.mean()) - (eo - eo.mean())).max() < 5e-3 * ke   # the fluctuation is the common O(dt^2) term
