"""Edge cases of the hot path: empty and ragged inputs, beads without partners, systems without any force term, a single bead.
The same builders run on the CPU oracle (`-m "not gpu"`: the checker must accept them and behave sensibly) and on the CUDA engine
against the oracle (`-m gpu`)."""
import numpy as np
import pytest

from opencafemol_b200.capi import Context, PairTopology, System


def _free_particles(lib, n=9, seed=4):
    """no force term at all: a Langevin gas"""
    rng = np.random.default_rng(seed)
    s = System(lib, rng.uniform(10.0, 200.0, n), None)
    c = Context(s, 300.0, 0.5, 0.005, seed=21)
    c.set_positions(rng.normal(scale=2.0, size=(n, 3)))
    c.set_velocities_to_temperature(300.0, 9)
    return c


def _single_bead(lib):
    s = System(lib, np.array([55.0]), None)
    s.add_position_restraint(1, [0], [[0.3, -0.2, 0.1]], [40.0], [0.0])
    c = Context(s, 300.0, 0.5, 0.005, seed=2)
    c.set_positions(np.zeros((1, 3)))
    c.set_velocities_to_temperature(300.0, 3)
    return c


def _empty_terms(lib, periodic):
    """terms with zero entries beside real ones, and a pair term none of whose beads meets a partner"""
    n = 12
    rng = np.random.default_rng(8)
    box = np.array([9.0, 9.5, 10.0]) if periodic else None
    s = System(lib, np.full(n, 40.0), box)
    none2, none3, none4 = np.zeros((0, 2), np.int32), np.zeros((0, 3), np.int32), np.zeros((0, 4), np.int32)
    s.add_harmonic_bond(0, none2, [], [], periodic)
    s.add_harmonic_angle(0, none3, [], [], periodic)
    s.add_gaussian_dihedral(0, none4, [], [], [], periodic)
    s.add_pulling(0, [], np.zeros((0, 3)))
    ij = np.array([[i, i + 1] for i in range(0, n - 1, 2)], np.int32)   # six dimers
    s.add_harmonic_bond(1, ij, np.full(len(ij), 0.38), np.full(len(ij), 500.0), periodic)
    # excluded volume on every second bead only (ragged participation); the dimers sit 3 nm apart, beyond cutoff + skin:
    # every Verlet row is empty
    rad = np.full(n, np.nan)
    rad[::2] = 0.25
    part = list(range(0, n, 2))   # (OpenMM: beads without parameters need an interaction group that leaves them out)
    s.add_excluded_volume(2, 0.6, 1.0, rad, PairTopology(igroups=[(part, part)], periodic=periodic))
    pos = np.zeros((n, 3))
    for d in range(n // 2):
        base = np.array([1.0 + 3.0 * (d % 3), 1.0 + 3.0 * (d // 3), 1.5])
        dirn = rng.normal(size=3)
        dirn /= np.linalg.norm(dirn)
        pos[2 * d], pos[2 * d + 1] = base, base + 0.41 * dirn
    c = Context(s, 300.0, 0.5, 0.002, seed=6)
    c.set_positions(pos)
    c.set_velocities_to_temperature(300.0, 4)
    return c


def _ragged_chains(lib, periodic):
    """chains of 1, 2, 5 and 17 beads (a lone bead has no bonded term at all), excluded volume with exclusions, some beads of the pair
    term far from everything else"""
    lens = [1, 2, 5, 17, 1, 3]
    n = sum(lens)
    rng = np.random.default_rng(12)
    box = np.array([7.0, 7.5, 8.0]) if periodic else None
    s = System(lib, rng.uniform(30.0, 120.0, n), box)
    pos, ij, start = np.zeros((n, 3)), [], 0
    for k, L in enumerate(lens):
        p = np.array([1.0 + 1.3 * k, 1.0 + 0.9 * k, 2.0])
        for m in range(L):
            pos[start + m] = p
            step = rng.normal(size=3)
            p = p + 0.38 * step / np.linalg.norm(step)
            if m + 1 < L:
                ij.append([start + m, start + m + 1])
        start += L
    ij = np.array(ij, np.int32)
    s.add_harmonic_bond(0, ij, np.full(len(ij), 0.38), np.full(len(ij), 800.0), periodic)
    s.add_excluded_volume(1, 0.6, 0.9, np.full(n, 0.22), PairTopology(excl=ij, periodic=periodic))
    c = Context(s, 300.0, 0.5, 0.002, seed=13)
    c.set_positions(pos)
    c.set_velocities_to_temperature(300.0, 14)
    return c


BUILDERS = {
    "free_particles": _free_particles,
    "single_bead": _single_bead,
    "empty_terms": lambda lib: _empty_terms(lib, False),
    "empty_terms_periodic": lambda lib: _empty_terms(lib, True),
    "ragged_chains": lambda lib: _ragged_chains(lib, False),
    "ragged_chains_periodic": lambda lib: _ragged_chains(lib, True),
}


@pytest.mark.parametrize("case", list(BUILDERS))
def test_oracle_accepts_edge_cases(oracle, case):
    c = BUILDERS[case](oracle)
    st = c.get_state(forces=True, energy=True)
    assert np.isfinite(st["forces"]).all() and np.isfinite(st["potential_energy"]) and st["kinetic_energy"] > 0
    if case == "free_particles":
        assert np.abs(st["forces"]).max() == 0.0 and st["potential_energy"] == 0.0
    if case.startswith("empty_terms"):
        assert c.num_pair_terms() == 1 and len(c.neighbor_pairs(0)) == 0
    c.step(30)
    st = c.get_state(positions=True, energy=True)
    assert np.isfinite(st["positions"]).all() and np.isfinite(st["potential_energy"])


def test_zero_steps_and_repeated_queries_are_idempotent(oracle):
    c = _ragged_chains(oracle, True)
    a = c.get_state(positions=True, velocities=True, forces=True, energy=True)
    c.step(0)
    b = c.get_state(positions=True, velocities=True, forces=True, energy=True)
    for k in ("positions", "velocities"):
        assert np.array_equal(a[k], b[k])
    # (the oracle's force pass after a step call sums per-thread slabs of its Verlet lists: same value, other rounding order)
    assert np.allclose(a["forces"], b["forces"], rtol=1e-12, atol=1e-11)
    assert a["potential_energy"] == pytest.approx(b["potential_energy"], rel=1e-13) and a["kinetic_energy"] == b["kinetic_energy"]


@pytest.mark.gpu
@pytest.mark.parametrize("case", list(BUILDERS))
def test_engine_edge_cases_match_the_oracle(engine, oracle, case):
    from test_gpu_parity import force_error

    ctx, ora = BUILDERS[case](engine), BUILDERS[case](oracle)
    a, b = ctx.get_state(forces=True, energy=True), ora.get_state(forces=True, energy=True)
    if np.abs(b["forces"]).max() > 0:
        assert force_error(a["forces"], b["forces"]).max() < 1e-5
    else:
        assert np.abs(a["forces"]).max() == 0.0
    assert np.allclose(a["energy_by_group"], b["energy_by_group"], rtol=1e-5, atol=1e-7)
    assert a["kinetic_energy"] == pytest.approx(b["kinetic_energy"], rel=1e-5)
    for t in range(ctx.num_pair_terms()):
        for ax in "xyz":
            ora.set_option("virtual_box_" + ax, ctx.counter("virtual_box_" + ax))
        assert np.array_equal(ctx.neighbor_pairs(t), ora.neighbor_pairs(t))
    ctx.step(0)
    ctx.step(30); ora.step(30)
    pa, pb = ctx.get_state(positions=True)["positions"], ora.get_state(positions=True)["positions"]
    assert np.abs(pa - pb).max() < 5e-5, f"{case}: trajectories deviate by {np.abs(pa - pb).max():.3e} nm after 30 steps"
    ea, eb = ctx.get_state(energy=True), ora.get_state(energy=True)
    assert ea["potential_energy"] == pytest.approx(eb["potential_energy"], rel=2e-3, abs=1e-3)
