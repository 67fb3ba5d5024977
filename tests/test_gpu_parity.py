"""Parity of the CUDA engine (through the C ABI) against the CPU oracle on identical inputs.

Bars (BASELINE.json north_star): neighbour pairs and cell assignment bit-exact as integer sets;
per-particle forces and per-term energies within 1e-5 relative (FP32 pair arithmetic against the
double-precision oracle; bonded terms are evaluated in double on the device)."""
import numpy as np
import pytest

from conftest import load_golden_spec, load_golden_values
from opencafemol_b200 import synth, toml_input as ti
from opencafemol_b200.capi import Context, PairTopology, System

pytestmark = pytest.mark.gpu
FORCE_TOL = 1e-5
ENERGY_TOL = 1e-5


def force_error(f, fref):
    """norm-relative per-particle error with the rms force as floor."""
    frms = np.sqrt((fref ** 2).sum(1).mean())
    return np.linalg.norm(f - fref, axis=1) / np.maximum(np.linalg.norm(fref, axis=1), frms)


def check_state(ctx, ora, spec, ftol=FORCE_TOL, etol=ENERGY_TOL):
    a = ctx.get_state(forces=True, energy=True)
    b = ora.get_state(forces=True, energy=True)
    err = force_error(a["forces"], b["forces"])
    assert err.max() < ftol, f"force error {err.max():.3e}"
    escale = max(np.abs(b["energy_by_group"]).max(), 1.0)
    for name, g in spec.group_ids.items():
        ea, eb = a["energy_by_group"][g], b["energy_by_group"][g]
        assert abs(ea - eb) <= etol * max(abs(eb), 1e-3 * escale), (name, ea, eb)
    assert a["kinetic_energy"] == pytest.approx(b["kinetic_energy"], rel=1e-5)
    return a, b


def sync_canon(ctx, ora):
    """Unlimited boundary: the engine quantises pair-kernel positions in a virtual periodic box (fixed point, uniform
    resolution); the oracle's canonical predicate needs that box to reproduce pair sets and cell ids bit for bit."""
    for ax in "xyz":
        ora.set_option("virtual_box_" + ax, ctx.counter("virtual_box_" + ax))


def check_sets(ctx, ora):
    sync_canon(ctx, ora)
    for t in range(ctx.num_pair_terms()):
        pa, pb = ctx.neighbor_pairs(t), ora.neighbor_pairs(t)
        assert pa.shape == pb.shape and np.array_equal(pa, pb), f"pair set of term {t} differs"


def test_sh3_golden(engine, oracle):
    spec = load_golden_spec("sh3")
    gold = load_golden_values("sh3")
    ctx, ora = ti.make_context(engine, spec), ti.make_context(oracle, spec)
    a, _ = check_state(ctx, ora, spec)
    check_sets(ctx, ora)
    for k, gid in spec.group_ids.items():
        ref = gold["energy_by_group_kj"][k]
        assert abs(a["energy_by_group"][gid] - ref) <= ENERGY_TOL * max(abs(ref), 0.1)
    assert force_error(a["forces"], np.array(gold["forces"])).max() < FORCE_TOL
    for t in range(ctx.num_pair_terms()):
        assert ctx.neighbor_pairs(t).tolist() == gold["pairs"][str(t)]


@pytest.mark.parametrize("n_chains", [3, 12])
def test_hps_slab_small(engine, oracle, n_chains):
    spec = synth.hps_slab(n_chains=n_chains)
    ctx, ora = ti.make_context(engine, spec), ti.make_context(oracle, spec)
    check_state(ctx, ora, spec)
    check_sets(ctx, ora)


def test_hps_c2_full_size(engine, oracle):
    """BASELINE config C2 (16 300 beads): forces, energies, and the 2.5 M-pair set bit-exact."""
    spec = synth.hps_slab(n_chains=100)
    ctx, ora = ti.make_context(engine, spec), ti.make_context(oracle, spec)
    check_state(ctx, ora, spec)
    check_sets(ctx, ora)
    # Newton's third law on the fixed-point accumulator: total force is zero to rounding
    f = ctx.get_state(forces=True)["forces"]
    assert np.abs(f.sum(0)).max() < 1e-3 * np.abs(f).max()


def test_c2_equilibrated_bench_state(engine, oracle):
    """The state bench.py times (synth.c2_equilibrated: the compacted condensate, ~4.8 M in-cutoff Ashbaugh-Hatch pairs, closer
    contacts than the serpentine start -> the hard case for FP32 r^-12): forces, per-group energies, pair sets bit-exact."""
    spec = synth.c2_equilibrated()
    ctx, ora = ti.make_context(engine, spec), ti.make_context(oracle, spec)
    check_state(ctx, ora, spec)
    check_sets(ctx, ora)
    assert len(ctx.neighbor_pairs(1)) > 4_000_000


def test_fus100_reference_full_size(engine):
    """sample/toml_input/FUS_100_HPS.toml at full size (16 300 beads) against the REFERENCE's own answer at the same relaxed
    coordinates (tests/golden/make_ref_fus100.py: oracle/_ref, the reference's reader, generators and expression strings)."""
    import os

    from conftest import GOLDEN
    spec = load_golden_spec("ref_fus100")
    with np.load(os.path.join(GOLDEN, "ref_fus100.npz")) as d:
        fref, names, eg, etot = d["forces"], [str(s) for s in d["group_names"]], d["group_energies_kj"], float(d["potential_energy_kj"])
        assert np.array_equal(d["positions"], spec.positions)
    ctx = ti.make_context(engine, spec)
    a = ctx.get_state(forces=True, energy=True)
    assert force_error(a["forces"], fref).max() < FORCE_TOL
    escale = max(np.abs(eg).max(), 1.0)
    for name, ref in zip(names, eg):
        got = a["energy_by_group"][spec.group_ids[name]]
        assert abs(got - ref) <= ENERGY_TOL * max(abs(ref), 1e-3 * escale), (name, got, ref)
    assert a["potential_energy"] == pytest.approx(etot, rel=ENERGY_TOL)


def test_c5_tiling_reproduces_the_c2_forces(engine):
    """Config C5 (the compacted C2 state tiled 8 x 8: 1 043 200 beads) on one GPU: exact periodic images, so every copy must
    feel the C2 forces (SURVEY.md 8d "free correctness check"); the C2 forces themselves are pinned to the oracle above."""
    base = synth.c2_equilibrated()
    f1 = ti.make_context(engine, base).get_state(forces=True)["forces"]
    big = synth.c2_equilibrated(tile=(8, 8))
    assert big.mass.size == 1_043_200
    ctx = ti.make_context(engine, big)
    f = ctx.get_state(forces=True)["forces"].reshape(64, -1, 3)
    frms = np.sqrt((f1 ** 2).sum(1).mean())
    worst = max(np.abs(f[c] - f1).max() for c in range(64))
    # FP32 pair arithmetic in a box 8 x larger (coarser fixed-point grid): copies agree to ~1e-6 of the rms force
    assert worst < 1e-5 * frms, worst / frms
    ctx.step(20)
    assert np.isfinite(ctx.get_state(energy=True)["potential_energy"])


def check_da_sets(ctx, ora):
    assert ctx.num_da_terms() == ora.num_da_terms()
    for t in range(ctx.num_da_terms()):
        assert np.array_equal(ctx.da_pairs(t), ora.da_pairs(t)), f"donor-acceptor set of term {t} differs"


@pytest.mark.parametrize("name", ["dna_3spn2", "nucleosome"])
def test_dna_samples(engine, oracle, name):
    """3SPN.2 base stacking / base pairing / cross stacking, 3SPN.2 excluded volume, PDNS (config C3): the engine's
    analytic forces against the oracle's automatic differentiation of the reference's expression strings."""
    spec = load_golden_spec(name)
    gold = load_golden_values(name)
    ctx, ora = ti.make_context(engine, spec), ti.make_context(oracle, spec)
    a, _ = check_state(ctx, ora, spec)
    check_sets(ctx, ora)
    check_da_sets(ctx, ora)
    for k, gid in spec.group_ids.items():
        ref = gold["energy_by_group_kj"][k]
        assert abs(a["energy_by_group"][gid] - ref) <= ENERGY_TOL * max(abs(ref), 0.1)
    assert force_error(a["forces"], np.array(gold["forces"])).max() < FORCE_TOL
    for t in range(ctx.num_da_terms()):
        assert ctx.da_pairs(t).tolist() == gold["da_pairs"][str(t)]


def test_nucleosome_trajectory(engine, oracle):
    spec = load_golden_spec("nucleosome")
    ctx, ora = ti.make_context(engine, spec), ti.make_context(oracle, spec)
    ctx.step(12); ora.step(12)
    a = ctx.get_state(positions=True); b = ora.get_state(positions=True)
    assert np.abs(a["positions"] - b["positions"]).max() < 2e-5
    check_state(ctx, ora, spec, ftol=5e-3, etol=5e-3)


@pytest.mark.parametrize("kind", ["3spn2_base_pair", "3spn2_cross_stacking", "pdns"])
@pytest.mark.parametrize("periodic", [False, True])
def test_donor_acceptor_random(engine, oracle, kind, periodic):
    from test_oracle_kat import _da_system
    n = 16
    box = np.array([3.1, 2.9, 3.3]) if periodic else None
    for seed in range(4):
        ctxs = []
        for lib in (engine, oracle):
            rng = np.random.default_rng(100 + seed)
            s = System(lib, np.full(n, 60.0), box)
            pos = _da_system(s, rng, kind, periodic, n)
            c = Context(s, 300.0, 0.2, 0.002, seed=3)
            c.set_positions(pos)
            ctxs.append(c)
        ctx, ora = ctxs

        class Spec:
            group_ids = {"0": 0}
        check_state(ctx, ora, Spec)
        check_da_sets(ctx, ora)
        ctx.step(10); ora.step(10)
        a = ctx.get_state(positions=True); b = ora.get_state(positions=True)
        dev = np.abs(a["positions"] - b["positions"]).max()
        assert dev < 5e-5, f"seed {seed}: trajectories deviate by {dev:.3e} nm after 10 steps"
        check_state(ctx, ora, Spec, ftol=2e-3, etol=2e-3)


def test_replica_batch(engine, oracle):
    """Config C4: a replica ensemble in one context.  Every replica matches the oracle's batch (which equals
    stand-alone runs, tests/test_oracle_kat.py) in forces, per-replica energies, pair sets and trajectory."""
    spec = load_golden_spec("sh3")
    R, n = 7, spec.mass.size
    ctx, ora = ti.make_context(engine, spec, replicas=R), ti.make_context(oracle, spec, replicas=R)
    assert ctx.n_replicas == R
    check_state(ctx, ora, spec)
    check_sets(ctx, ora)
    ctx.step(20); ora.step(20)
    a = ctx.get_state(positions=True, velocities=True); b = ora.get_state(positions=True, velocities=True)
    assert np.abs(a["positions"] - b["positions"]).max() < 2e-5
    assert np.abs(a["positions"][:n] - a["positions"][n:2 * n]).max() > 1e-4   # replicas have their own noise
    ea, eb = ctx.replica_energies(), ora.replica_energies()
    for r in range(R):
        assert ea["potential_energy"][r] == pytest.approx(eb["potential_energy"][r], rel=5e-4, abs=0.05)
        assert ea["kinetic_energy"][r] == pytest.approx(eb["kinetic_energy"][r], rel=5e-4)
    tot = ctx.get_state(energy=True)
    assert tot["potential_energy"] == pytest.approx(ea["potential_energy"].sum(), rel=1e-9)
    # a longer run keeps the replicas apart from each other's pair lists: forces still match replica by replica
    ctx.step(200)
    ora.set_positions(ctx.get_state(positions=True)["positions"])
    fa, fb = ctx.get_state(forces=True)["forces"], ora.get_state(forces=True)["forces"]
    assert force_error(fa, fb).max() < FORCE_TOL


def test_replica_batch_large(engine):
    """512 SH3 replicas (the full C4 ensemble) step and stay finite; replica 0 equals a 1-replica batch."""
    spec = load_golden_spec("sh3")
    n = spec.mass.size
    big = ti.make_context(engine, spec, replicas=512)
    one = ti.make_context(engine, spec, replicas=1)
    big.step(50); one.step(50)
    a = big.get_state(positions=True)["positions"]; b = one.get_state(positions=True)["positions"]
    assert np.isfinite(a).all()
    assert np.abs(a[:n] - b).max() < 2e-5


@pytest.mark.parametrize("replicas", [None, 5])
def test_whole_step_replica_kernel_option(engine, oracle, replicas):
    """option replica_kernel = 1 (the default): the whole MD step of a small system in one persistent kernel, one block per
    replica (replica_kernels.cuh).  Same noise streams, so it tracks the oracle like the general path, step by step and across
    launches."""
    spec = load_golden_spec("sh3")
    ctx, ora = ti.make_context(engine, spec, replicas=replicas), ti.make_context(oracle, spec, replicas=replicas)
    ctx.set_option("replica_kernel", 1)
    assert ctx.get_option("replica_kernel") == 1
    ctx.step(7); ctx.step(13); ora.step(20)
    assert ctx.counter("replica_kernel_launches") == 2 and ctx.current_step == 20
    a = ctx.get_state(positions=True, velocities=True); b = ora.get_state(positions=True, velocities=True)
    assert np.abs(a["positions"] - b["positions"]).max() < 2e-5
    assert np.abs(a["velocities"] - b["velocities"]).max() < 2e-3
    check_state(ctx, ora, spec, ftol=5e-3, etol=5e-3)  # states differ by the accumulated FP32 drift


def test_async_observer_matches_blocking_and_overlaps(engine, oracle):
    """ocm_get_state_begin / _end (side-stream download from pinned staging, energy kernels on the lists of the last window) against
    the blocking ocm_get_state, with a step issued between begin and end; ocm_set_state_async against set_positions/velocities."""
    spec = synth.hps_slab(n_chains=12)
    a, b = ti.make_context(engine, spec), ti.make_context(engine, spec)
    a.step(100); b.step(100)
    req = a.get_state_begin(positions=True, velocities=True, forces=True, energy=True)
    a.step(50)                                   # runs while the frame of step 100 is on its way to the host
    sa = a.get_state_end(req)
    sb = b.get_state(positions=True, velocities=True, forces=True, energy=True)
    assert np.array_equal(sa["positions"], sb["positions"]) and np.array_equal(sa["velocities"], sb["velocities"])
    # (the blocking call rebuilds the lists, the asynchronous one reuses the last window's: same pair set, other summation order)
    assert force_error(sa["forces"], sb["forces"]).max() < 1e-6
    assert np.allclose(sa["energy_by_group"], sb["energy_by_group"], rtol=1e-7, atol=1e-6)
    assert sa["kinetic_energy"] == pytest.approx(sb["kinetic_energy"], rel=1e-7)
    assert a.counter("observer_rebuilds_skipped") >= 1 and a.counter("async_states") == 1
    b.step(50)
    assert np.array_equal(a.get_state(positions=True)["positions"], b.get_state(positions=True)["positions"])  # the observer left no trace
    # two frames in flight, ended in order
    r1 = a.get_state_begin(positions=True, energy=True); a.step(10)
    r2 = a.get_state_begin(positions=True, energy=True); a.step(10)
    f1, f2 = a.get_state_end(r1), a.get_state_end(r2)
    assert np.abs(f1["positions"] - f2["positions"]).max() > 0
    # upload through pinned staging: same trajectory as the blocking setters, and the oracle's restatement agrees
    st = b.get_state(positions=True, velocities=True)
    c, d, o = ti.make_context(engine, spec), ti.make_context(engine, spec), ti.make_context(oracle, spec)
    c.set_state_async(st["positions"], st["velocities"])
    d.set_positions(st["positions"]); d.set_velocities(st["velocities"])
    o.set_state_async(st["positions"], st["velocities"])
    c.step(20); d.step(20); o.step(20)
    pc = c.get_state(positions=True)["positions"]
    assert np.array_equal(pc, d.get_state(positions=True)["positions"])
    assert np.abs(pc - o.get_state_end(o.get_state_begin(positions=True))["positions"]).max() < 2e-5


def test_cell_assignment_bit_exact(engine, oracle):
    import ctypes as C
    for spec in (synth.hps_slab(n_chains=12), load_golden_spec("sh3")):
        ctx, ora = ti.make_context(engine, spec), ti.make_context(oracle, spec)
        ctx.get_state(energy=True)  # lists sized: the virtual box of the unlimited-boundary case is laid out
        sync_canon(ctx, ora)
        for t in range(ctx.num_pair_terms()):
            # the grid rule has two inputs, which the engine reports (rows: half the list radius; tiles: a sixth of it, capped)
            width, cap = ctx.counter(f"cell_width_nm.{t}"), int(ctx.counter(f"cell_cap.{t}"))
            ids, dims = ctx.cell_ids(t)
            oid = np.zeros(spec.mass.size, np.int32)
            od = np.zeros(3, np.int32)
            fn = oracle.dll.oracle_get_cell_ids_grid
            fn.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_longlong, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
            assert fn(ora.handle, t, width, cap, oid.ctypes.data_as(C.POINTER(C.c_int32)), od.ctypes.data_as(C.POINTER(C.c_int32))) == 0
            assert np.array_equal(dims, od), (dims, od)
            assert np.array_equal(ids, oid)


def test_positions_across_periodic_boundary(engine, oracle):
    """Unwrapped coordinates many box lengths away give the same forces (fixed-point box fractions)."""
    spec = synth.hps_slab(n_chains=6)
    rng = np.random.default_rng(3)
    shift = rng.integers(-3, 4, size=(spec.mass.size, 3)) * spec.box
    chain = np.repeat(np.arange(6), 163)
    spec.positions = spec.positions + shift[chain * 163]  # whole chains moved by lattice vectors
    ctx, ora = ti.make_context(engine, spec), ti.make_context(oracle, spec)
    check_state(ctx, ora, spec)
    check_sets(ctx, ora)


def test_trajectory_tracks_oracle(engine, oracle):
    """Same Philox noise on both sides: a short Langevin trajectory stays together."""
    spec = synth.hps_slab(n_chains=3)
    ctx, ora = ti.make_context(engine, spec), ti.make_context(oracle, spec)
    ctx.step(25); ora.step(25)
    a = ctx.get_state(positions=True, velocities=True)
    b = ora.get_state(positions=True, velocities=True)
    assert np.abs(a["positions"] - b["positions"]).max() < 2e-5
    assert np.abs(a["velocities"] - b["velocities"]).max() < 2e-3
    assert ctx.current_step == 25


def test_sh3_trajectory_and_split_steps(engine, oracle):
    spec = load_golden_spec("sh3")
    ctx, ora = ti.make_context(engine, spec), ti.make_context(oracle, spec)
    ctx.step(7); ctx.step(13); ora.step(20)
    a = ctx.get_state(positions=True); b = ora.get_state(positions=True)
    assert np.abs(a["positions"] - b["positions"]).max() < 2e-5
    check_state(ctx, ora, spec, ftol=5e-3, etol=5e-3)  # states differ by the accumulated FP32 drift


def test_velocity_initialisation(engine, oracle):
    spec = synth.hps_slab(n_chains=3)
    ctx, ora = ti.make_context(engine, spec), ti.make_context(oracle, spec)
    ctx.set_velocities_to_temperature(300.0, 99); ora.set_velocities_to_temperature(300.0, 99)
    va = ctx.get_state(velocities=True)["velocities"]; vb = ora.get_state(velocities=True)["velocities"]
    assert np.abs(va - vb).max() < 1e-5
    kT = 8.31446261815324e-3 * 300
    assert (spec.mass[:, None] * vb ** 2).mean() == pytest.approx(kT, rel=0.1)


def test_nve_limit_energy_conservation(engine):
    """gamma = 0 turns the Langevin step into leap-frog: total energy is conserved (drift matched to the
    integrator's O(dt^2) fluctuation, no secular loss through the Verlet buffer or rollbacks)."""
    spec = synth.hps_slab(n_chains=12)
    spec.friction = 0.0
    ctx = ti.make_context(engine, spec)
    ctx.step(200)
    e = []
    for _ in range(10):
        ctx.step(100)
        st = ctx.get_state(energy=True)
        e.append(st["potential_energy"] + st["kinetic_energy"])
    e = np.array(e)
    ke = st["kinetic_energy"]
    assert np.abs(e - e.mean()).max() < 5e-3 * ke
    assert abs(e[-1] - e[0]) < 5e-3 * ke


def test_rebuild_interval_independence(engine):
    """Forced list rebuild cadence changes only the FP32 summation order."""
    spec = synth.hps_slab(n_chains=6)
    out = []
    for nst in (1, 5, 0):
        ctx = ti.make_context(engine, spec)
        ctx.set_option("nstlist", nst)
        ctx.step(60)
        out.append(ctx.get_state(positions=True)["positions"])
    assert np.abs(out[0] - out[1]).max() < 1e-4
    assert np.abs(out[0] - out[2]).max() < 1e-4


def test_run_to_run_determinism(engine):
    spec = synth.hps_slab(n_chains=6)
    res = []
    for _ in range(2):
        ctx = ti.make_context(engine, spec)
        ctx.set_option("nstlist", 10)
        ctx.step(100)
        res.append(ctx.get_state(positions=True, velocities=True))
    assert np.array_equal(res[0]["positions"], res[1]["positions"])
    assert np.array_equal(res[0]["velocities"], res[1]["velocities"])


def test_equipartition_long_langevin(engine):
    spec = synth.hps_slab(n_chains=12)
    ctx = ti.make_context(engine, spec)
    ctx.step(3000)
    kes = []
    for _ in range(10):
        ctx.step(200)
        kes.append(ctx.get_state(energy=True)["kinetic_energy"])
    kT = 8.31446261815324e-3 * spec.temperature
    assert np.mean(kes) == pytest.approx(1.5 * spec.mass.size * kT, rel=0.03)


def test_options_and_counters(engine):
    spec = synth.hps_slab(n_chains=3)
    ctx = ti.make_context(engine, spec)
    ctx.set_option("skin_nm", 0.25)
    assert ctx.get_option("skin_nm") == pytest.approx(0.25)
    for tpb in (4, 8, 16, 32):
        ctx.set_option("threads_per_bead", tpb)
        f = ctx.get_state(forces=True)["forces"]
        if tpb == 4:
            f0 = f
        assert np.abs(f - f0).max() < 1e-3
    ctx.step(30)
    assert ctx.counter("kernel_launches") > 60
    assert ctx.counter("rebuilds") >= 1
    assert ctx.counter("pair_evals_per_step") > 0
    assert ctx.counter("num_pair_lists") == 2  # LJ-rep + LJ-att fused into one Ashbaugh-Hatch list, + DH
    with pytest.raises(Exception):
        ctx.set_option("no_such_option", 1)


def all_kinds_contexts(engine, oracle):
    """Random small system exercising all nine bonded kinds, EXV with interaction groups, 3SPN.2 EXV: (engine, oracle) contexts."""
    from opencafemol_b200 import tables
    rng = np.random.default_rng(5)
    n = 40
    # self-avoiding walk: no unphysical overlaps (forces stay inside the fixed-point accumulator's
    # +-2^31 kJ/mol/nm range, the same limit OpenMM's CUDA platform has)
    pos = np.zeros((n, 3))
    i = 1
    while i < n:
        d = rng.normal(size=3); d /= np.linalg.norm(d)
        cand = pos[i - 1] + 0.45 * d
        if i < 2 or np.linalg.norm(pos[:i - 1] - cand, axis=1).min() > 0.5:
            pos[i] = cand
            i += 1
    ij = np.array([[i, i + 1] for i in range(n - 1)])
    ijk = np.array([[i, i + 1, i + 2] for i in range(n - 2)])
    ijkl = np.array([[i, i + 1, i + 2, i + 3] for i in range(n - 3)])
    m2, m3, m4 = len(ij), len(ijk), len(ijkl)
    radius = rng.uniform(0.15, 0.25, n)
    radius[::7] = np.nan
    has = ~np.isnan(radius)
    grp = ["A" if i < 20 else "B" for i in range(n)]
    ig = ti.extract_interaction_group(has, [("A", "A")], grp)
    excl = np.array([[i, i + d] for d in (1, 2) for i in range(n - d)])
    ctxs = []
    for lib in (engine, oracle):
        s = System(lib, rng.uniform(50, 150, n) * 0 + 100.0)
        s.add_harmonic_bond(0, ij, np.full(m2, 0.45), np.full(m2, 5000.0))
        s.add_gaussian_bond(1, ijk[:, [0, 2]], -np.full(m3, 2.0), np.full(m3, 0.7), np.full(m3, 0.1))
        s.add_go_contact(2, ijkl[:, [0, 3]], np.full(m4, 1.3), np.full(m4, 0.9))
        s.add_3spn2_bond(3, ij, np.full(m2, 200.0), np.full(m2, 2e4), np.full(m2, 0.45))
        s.add_harmonic_angle(4, ijk, np.full(m3, 2.0), np.full(m3, 40.0))
        s.add_flexible_local_angle(5, ijk, np.full(m3, 4.184), tables.FLA_SPLINE["GLY"], tables.FLA_MIN_THETA,
                                   tables.FLA_MAX_THETA, 0.5)
        s.add_gaussian_dihedral(6, ijkl, -np.full(m4, 1.8), np.linspace(-3, 3, m4), np.full(m4, 0.3))
        s.add_cosine_dihedral(7, ijkl, np.full(m4, 2.0), np.linspace(-3, 3, m4), (np.arange(m4) % 3 + 1).astype(np.int32))
        s.add_flexible_local_dihedral(8, ijkl, np.full(m4, 4.184), tables.FLD_FOURIER[("ALA", "R3")], 1.2)
        s.add_excluded_volume(9, 2.5, 0.9, radius, PairTopology(excl=excl, igroups=ig))
        s.add_3spn2_excluded_volume(10, 1.4, np.full(n, 1.0), rng.uniform(0.45, 0.7, n) * 0 + np.linspace(0.45, 0.7, n),
                                    PairTopology(excl=excl))
        c = Context(s, 300.0, 0.2, 0.002, seed=11)
        c.set_positions(pos)
        ctxs.append(c)
    return ctxs


class AllKindsSpec:
    group_ids = {str(g): g for g in range(11)}


def test_every_bonded_kind_and_interaction_groups(engine, oracle):
    ctx, ora = all_kinds_contexts(engine, oracle)
    check_state(ctx, ora, AllKindsSpec)
    check_sets(ctx, ora)


def test_whole_step_replica_kernel_all_kinds(engine, oracle):
    """The whole-step replica kernel (default for small unlimited-boundary systems) on every bonded kind it restates in FP32
    (replica_math.hpp) and on two pair lists with interaction groups, against the oracle under the same noise, and against
    the general path (option replica_kernel = 0)."""
    ctx, ora = all_kinds_contexts(engine, oracle)
    gen, _ = all_kinds_contexts(engine, oracle)
    assert ctx.get_option("replica_kernel") == 1
    gen.set_option("replica_kernel", 0)
    for c in (ctx, ora, gen):
        c.set_velocities_to_temperature(300.0, 3)
    ctx.step(1); ora.step(1); gen.step(1)
    a, b, g = (c.get_state(positions=True, velocities=True) for c in (ctx, ora, gen))
    # one step from identical states under identical noise: the velocity difference is dt/m times the force difference
    assert np.abs(a["velocities"] - b["velocities"]).max() < 2e-6
    assert np.abs(g["velocities"] - b["velocities"]).max() < 2e-6
    ctx.step(29); ora.step(29); gen.step(29)
    a, b, g = (c.get_state(positions=True, velocities=True) for c in (ctx, ora, gen))
    assert ctx.counter("replica_kernel_launches") == 2 and gen.counter("replica_kernel_launches") == 0
    assert np.abs(a["positions"] - b["positions"]).max() < 2e-5
    assert np.abs(g["positions"] - b["positions"]).max() < 2e-5
    check_state(ctx, ora, AllKindsSpec, ftol=5e-3, etol=5e-3)   # (own states: differ by the accumulated FP32 drift)


def test_whole_step_replica_kernel_is_reproducible_and_rebuilds_its_lists(engine, oracle):
    """Two runs of the replica kernel give the same bits (fixed summation order, no atomics on data), and a run long enough for
    beads to cross the kernel's Verlet skin (in-kernel list rebuilds) still tracks a run that scans all pairs every step
    (skin 0: option replica_kernel = 0 is the general path with its own lists)."""
    spec = load_golden_spec("sh3")
    a, b = ti.make_context(engine, spec, replicas=3), ti.make_context(engine, spec, replicas=3)
    for c in (a, b):
        c.step(1500); c.step(2500)   # (the same calls: a launch re-centres its FP32 frame at its first step)
    xa, xb = a.get_state(positions=True)["positions"], b.get_state(positions=True)["positions"]
    assert np.array_equal(xa, xb)
    # forces at the end state against the oracle (the kernel's own lists played no part in them) and energies finite
    ora = ti.make_context(oracle, spec, replicas=3)
    ora.set_positions(xa)
    fa, fb = a.get_state(forces=True)["forces"], ora.get_state(forces=True)["forces"]
    assert force_error(fa, fb).max() < FORCE_TOL
    # the general path from the same start, same noise: chaotic divergence over 4000 steps is bounded by the thermal scale, but
    # the first 200 steps must agree closely
    c, d = ti.make_context(engine, spec, replicas=3), ti.make_context(engine, spec, replicas=3)
    d.set_option("replica_kernel", 0)
    c.step(200); d.step(200)
    assert np.abs(c.get_state(positions=True)["positions"] - d.get_state(positions=True)["positions"]).max() < 1e-4


def test_c3_nucleosome_grid(engine, oracle):
    """config C3: nucleosome_1kx5 on a 2 x 3 grid (11 160 beads: 3SPN.2C DNA, AICG2+ protein, PDNS, Debye-Hueckel)"""
    one = load_golden_spec("nucleosome")
    spec = synth.replicate(one, [[30.0 * i, 30.0 * j, 0.0] for i in range(2) for j in range(3)])
    ctx, ora = ti.make_context(engine, spec), ti.make_context(oracle, spec)
    a, b = check_state(ctx, ora, spec)
    single = ti.make_context(oracle, one).get_state(energy=True)
    assert a["potential_energy"] == pytest.approx(6.0 * single["potential_energy"], rel=1e-5)
    for t in range(ctx.num_da_terms()):
        assert np.array_equal(ctx.da_pairs(t), ora.da_pairs(t))
    ctx.step(200)
    assert np.isfinite(ctx.get_state(energy=True)["potential_energy"])


def test_periodic_tiling_reproduces_the_forces(engine):
    """a periodic system tiled 2 x 2 in x, y into a box twice as large has, copy by copy, the forces
    of the original (exact periodic images); the fixed-point grid is twice as coarse, hence 1e-5 and not bit-exact"""
    base = synth.hps_slab(n_chains=12)
    lx, ly, lz = base.box
    tiled = synth.replicate(base, [[i * lx, j * ly, 0.0] for i in range(2) for j in range(2)], box=[2 * lx, 2 * ly, lz])
    a = ti.make_context(engine, base).get_state(forces=True, energy=True)
    b = ti.make_context(engine, tiled).get_state(forces=True, energy=True)
    n = base.mass.size
    for c in range(4):
        assert force_error(b["forces"][c * n:(c + 1) * n], a["forces"]).max() < FORCE_TOL
    assert b["potential_energy"] == pytest.approx(4.0 * a["potential_energy"], rel=1e-6)


@pytest.mark.parametrize("periodic", [False, True])
def test_wca_random(engine, oracle, periodic):
    """Weeks-Chandler-Andersen (geometric epsilon): same kernel path as LJ-repulsive, its own coefficient table"""
    rng = np.random.default_rng(21)
    n = 600
    box = np.array([6.0, 5.5, 6.5]) if periodic else None
    g = np.stack(np.meshgrid(np.arange(8), np.arange(8), np.arange(10), indexing="ij"), -1).reshape(-1, 3)[:n]
    pos = g * 0.62 + rng.uniform(-0.08, 0.08, (n, 3)) + 0.3
    # (a handful of distinct parameter sets, as real force fields have: the engine keys its coefficient table on them)
    eps, sig = rng.choice([0.5, 0.9, 1.4, 2.0], n), rng.choice([0.5, 0.56, 0.62, 0.7, 0.75], n)
    eps[::17] = np.nan
    sig[::17] = np.nan
    has = ~np.isnan(sig)
    two = np.sort(sig[has])[-2:]
    rc = two.sum() * 0.5 * 2.0 ** (1.0 / 6.0)
    excl = np.array([[i, i + 1] for i in range(0, n - 1, 3) if has[i] and has[i + 1]], np.int32)
    part = np.nonzero(has)[0]

    def make(lib):
        s = System(lib, np.full(n, 100.0), box)
        s.add_wca(0, rc, eps, sig, PairTopology(excl=excl, igroups=[(part, part)], periodic=periodic))
        c = Context(s, 300.0, 0.2, 0.002, seed=5)
        c.set_positions(pos)
        c.set_velocities_to_temperature(300.0, 3)
        return c

    ctx, ora = make(engine), make(oracle)
    a, b = ctx.get_state(forces=True, energy=True), ora.get_state(forces=True, energy=True)
    assert force_error(a["forces"], b["forces"]).max() < FORCE_TOL
    assert a["potential_energy"] == pytest.approx(b["potential_energy"], rel=ENERGY_TOL)
    assert b["potential_energy"] > 0
    check_sets(ctx, ora)


@pytest.mark.parametrize("periodic", [False, True])
@pytest.mark.parametrize("kind", ["isolf_attractive", "trigonometric"])
def test_attractive_wells_random(engine, oracle, kind, periodic):
    """iSoLFAttractive / Trigonometric (the attractive wells of the lipid model): a pair form of their own (Verlet rows,
    well_dedr in kernels.cuh) against the oracle, which is pinned to the reference's expression strings by the membrane fixture.
    Unequal sigmas and omegas; for Trigonometric the generator's cutoff lies inside the widest well."""
    rng = np.random.default_rng(33)
    n = 600
    box = np.array([6.0, 5.5, 6.5]) if periodic else None
    g = np.stack(np.meshgrid(np.arange(8), np.arange(8), np.arange(10), indexing="ij"), -1).reshape(-1, 3)[:n]
    pos = g * 0.62 + rng.uniform(-0.12, 0.12, (n, 3)) + 0.3
    eps, sig, om = rng.choice([0.5, 0.9, 1.7], n), rng.choice([0.5, 0.56, 0.7], n), rng.choice([0.3, 0.45, 0.6], n)
    for a in (eps, sig, om):
        a[::13] = np.nan
    has = ~np.isnan(sig)
    s2, o2, r2 = np.sort(sig[has])[-2:], np.sort(om[has])[-2:], np.sort((sig + om)[has])[-2:]
    rc = s2.sum() * 0.5 * 2.0 ** (1.0 / 6.0) + o2.sum() * 0.5 if kind == "isolf_attractive" else r2.sum() * 0.5
    excl = np.array([[i, i + 1] for i in range(0, n - 1, 3) if has[i] and has[i + 1]], np.int32)
    part = np.nonzero(has)[0]

    def make(lib):
        s = System(lib, np.full(n, 100.0), box)
        getattr(s, "add_" + kind)(0, rc, eps, sig, om, PairTopology(excl=excl, igroups=[(part, part)], periodic=periodic))
        s.add_harmonic_bond(1, np.array([[i, i + 1] for i in range(0, n - 1, 2)]), np.full(n // 2, 0.62), np.full(n // 2, 500.0), periodic=periodic)
        c = Context(s, 300.0, 0.2, 0.002, seed=5)
        c.set_positions(pos)
        c.set_velocities_to_temperature(300.0, 3)
        return c

    ctx, ora = make(engine), make(oracle)
    a, b = ctx.get_state(forces=True, energy=True), ora.get_state(forces=True, energy=True)
    assert force_error(a["forces"], b["forces"]).max() < FORCE_TOL
    assert a["energy_by_group"][0] == pytest.approx(b["energy_by_group"][0], rel=ENERGY_TOL)
    assert b["energy_by_group"][0] < 0
    check_sets(ctx, ora)
    ctx.step(20); ora.step(20)
    assert np.abs(ctx.get_state(positions=True)["positions"] - ora.get_state(positions=True)["positions"]).max() < 2e-5


def test_capped_go_contact(engine, oracle):
    """Go contacts with the linear short-range cap (CappedGoContactForceFieldGenerator): both branches, against the oracle"""
    rng = np.random.default_rng(8)
    n = 40
    pos = np.cumsum(rng.normal(scale=0.22, size=(n, 3)), axis=0) + 5.0
    ij = np.array([[i, j] for i in range(n) for j in range(i + 3, n, 5)], np.int32)
    r = np.linalg.norm(pos[ij[:, 0]] - pos[ij[:, 1]], axis=1)
    r0 = r * rng.choice([0.8, 1.0, 1.2, 1.6, 2.0], len(ij))  # r/r0 from 0.5 (on the cap) to 1.25
    cap = 0.75
    slope, icpt = -60 * (cap ** -13 - cap ** -11), 65 * cap ** -12 - 66 * cap ** -10
    assert ((r / r0) <= cap).any() and ((r / r0) > cap).any()

    def make(lib):
        s = System(lib, np.full(n, 100.0), None)
        s.add_capped_go_contact(0, ij, rng_k, r0, np.full(len(ij), slope), np.full(len(ij), icpt), np.full(len(ij), cap))
        s.add_go_contact(1, ij[::3], rng_k[::3], r0[::3])
        c = Context(s, 300.0, 0.2, 0.002, seed=5)
        c.set_positions(pos)
        c.set_velocities_to_temperature(300.0, 3)
        return c

    rng_k = rng.uniform(0.3, 1.5, len(ij))
    ctx, ora = make(engine), make(oracle)
    a, b = ctx.get_state(forces=True, energy=True), ora.get_state(forces=True, energy=True)
    assert force_error(a["forces"], b["forces"]).max() < FORCE_TOL
    for g in (0, 1):
        assert a["energy_by_group"][g] == pytest.approx(b["energy_by_group"][g], rel=ENERGY_TOL)


@pytest.mark.parametrize("periodic", [False, True])
@pytest.mark.parametrize("form", ["position_restraint", "pulling", "exv_rectangular_box", "cylindrical_restraint", "harmonic_com_pulling",
                                  "combinatorial_go_contact"])
def test_external_terms_random(engine, oracle, form, periodic):
    """The reference's external generators (PositionRestraint, Pulling, EXVRectangularBox, CylindricalRestraint, HarmonicCoMPulling)
    and CombinatorialGoContact on random chains, both boundaries: forces and group energies at 1e-5 against the oracle's dual-number
    restatement of the reference's expression strings, then a short trajectory (the terms live in the captured step graph) and the
    combination sets of the contact term."""
    from test_oracle_kat import external_terms
    if periodic and form == "exv_rectangular_box":
        pytest.skip("the generator throws under a periodic boundary")
    n = 24
    box = np.array([3.4, 3.1, 3.6]) if periodic else None
    masses = np.random.default_rng(5).uniform(20.0, 120.0, n)
    for seed in range(3):
        ctxs = []
        for lib in (engine, oracle):
            rng = np.random.default_rng(300 + seed)
            s = System(lib, masses, box)
            pos = external_terms(s, rng, n, form, periodic)
            # a chain of harmonic bonds keeps the beads together during the trajectory part
            ij = np.array([[i, i + 1] for i in range(n - 1)], np.int32)
            s.add_harmonic_bond(0, ij, np.full(n - 1, 0.38), np.full(n - 1, 2000.0), periodic)
            c = Context(s, 300.0, 0.2, 0.002, seed=3)
            c.set_positions(pos)
            c.set_velocities_to_temperature(300.0, 17)
            ctxs.append(c)
        ctx, ora = ctxs

        class Spec:
            group_ids = {"bond": 0, "term": 2, "term_b": 3}
        a, _ = check_state(ctx, ora, Spec)
        assert abs(a["energy_by_group"][2]) > 1e-6, "vacuous check: the term has no energy"
        if form == "combinatorial_go_contact":
            pa, pb = ctx.da_pairs(0), ora.da_pairs(0)
            assert len(pa) > 0 and np.array_equal(pa, pb)
        ctx.step(25); ora.step(25)
        pa, pb = ctx.get_state(positions=True)["positions"], ora.get_state(positions=True)["positions"]
        assert np.abs(pa - pb).max() < 5e-5, f"{form} seed {seed}: trajectories deviate by {np.abs(pa - pb).max():.3e} nm after 25 steps"
        check_state(ctx, ora, Spec, ftol=2e-3, etol=2e-3)


def test_external_terms_refused_where_unsupported(engine):
    """centre-of-mass springs couple beads of different slabs / replicas: refused loudly instead of being dropped"""
    s = System(engine, np.full(8, 50.0), None)
    s.add_harmonic_com_pulling(0, 10.0, 0.5, [0, 1, 2], [5, 6, 7])
    with pytest.raises(RuntimeError, match="centre-of-mass"):
        Context(s, 300.0, 0.2, 0.002, replica_seeds=[1, 2, 3])


def test_external_terms_in_a_replica_batch_take_the_general_path(engine, oracle):
    """one-bead external terms are not part of the whole-step replica kernel: a batch that has them must run (correctly) on the
    general path"""
    from test_oracle_kat import external_terms
    n, R = 12, 4
    ctxs = []
    for lib in (engine, oracle):
        rng = np.random.default_rng(77)
        s = System(lib, np.full(n, 60.0), None)
        pos = external_terms(s, rng, n, "position_restraint")
        ij = np.array([[i, i + 1] for i in range(n - 1)], np.int32)
        s.add_harmonic_bond(0, ij, np.full(n - 1, 0.38), np.full(n - 1, 2000.0))
        c = Context(s, 300.0, 0.2, 0.002, replica_seeds=[11 + r for r in range(R)])
        c.set_positions(np.tile(pos, (R, 1)))
        c.set_velocities_to_temperature(300.0, 5)
        ctxs.append(c)
    ctx, ora = ctxs
    assert ctx.get_option("replica_kernel") == 0

    class Spec:
        group_ids = {"bond": 0, "term": 2}
    check_state(ctx, ora, Spec)
    ctx.step(20); ora.step(20)
    pa, pb = ctx.get_state(positions=True)["positions"], ora.get_state(positions=True)["positions"]
    assert np.abs(pa - pb).max() < 5e-5
    ea, eb = ctx.replica_energies(), ora.replica_energies()
    assert np.allclose(ea["potential_energy"], eb["potential_energy"], rtol=2e-3, atol=1e-3)
