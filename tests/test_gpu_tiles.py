"""The cluster-pair ("tile") pair lists and the warp-tiled Newton-3 force kernel (opencafemol_b200/csrc/tile_kernels.cuh),
forced on with option pair_kernel = 1 (by default only lists of >= 131 072 beads use them), against the CPU oracle:
same bars as tests/test_gpu_parity.py — pair sets and cell ids bit-exact, forces and per-group energies 1e-5."""
import numpy as np
import pytest

from conftest import load_golden_spec
from opencafemol_b200 import synth, toml_input as ti
from test_gpu_parity import check_da_sets, check_sets, check_state, force_error, sync_canon
from test_gpu_slab import SlabRun

pytestmark = pytest.mark.gpu


def tiled(lib, spec, **kw):
    ctx = ti.make_context(lib, spec, **kw)
    ctx.set_option("pair_kernel", 1)
    return ctx


def n_tiled(ctx):
    return sum(int(ctx.counter(f"tiled.{t}")) for t in range(ctx.num_pair_terms()))


@pytest.mark.parametrize("n_chains", [7, 12, 40])
def test_tiles_hps_slab(engine, oracle, n_chains):
    """Ashbaugh-Hatch (fused LJ-repulsive + switched LJ-attractive): ragged last supercluster, exclusions inside and across
    superclusters, the i < j rule on the diagonal entries."""
    spec = synth.hps_slab(n_chains=n_chains)
    ctx, ora = tiled(engine, spec), ti.make_context(oracle, spec)
    check_state(ctx, ora, spec)
    check_sets(ctx, ora)
    assert n_tiled(ctx) >= 2  # both members of the fused list report tiled
    f = ctx.get_state(forces=True)["forces"]
    assert np.abs(f.sum(0)).max() < 1e-3 * np.abs(f).max()  # Newton's third law through the fixed-point accumulator


def test_tiles_c2_bench_state(engine, oracle):
    spec = synth.c2_equilibrated()
    ctx, ora = tiled(engine, spec), ti.make_context(oracle, spec)
    check_state(ctx, ora, spec)
    check_sets(ctx, ora)
    evals, pairs = ctx.counter("pair_evals_per_step"), len(ctx.neighbor_pairs(1))
    assert 0.35 < pairs / evals < 0.7, "lane efficiency of the tiles (in-cutoff pairs per evaluated lane pair)"


@pytest.mark.parametrize("name", ["nucleosome", "dna_3spn2"])
def test_tiles_dna_samples(engine, oracle, name):
    """Excluded volume (protein), 3SPN.2 excluded volume (geometric-mean epsilon), Debye-Hueckel, interaction classes
    (protein-DNA groups), exclusions up to three bonds, in the virtual box of an unlimited-boundary system."""
    spec = load_golden_spec(name)
    ctx, ora = tiled(engine, spec), ti.make_context(oracle, spec)
    check_state(ctx, ora, spec)
    check_sets(ctx, ora)
    check_da_sets(ctx, ora)


def test_tiles_c3_grid_uses_tiles(engine, oracle):
    """config C3 (the nucleosome sample on a 2 x 3 grid, 11 160 beads): every large pair list in tile form"""
    base = load_golden_spec("nucleosome")
    spec = synth.replicate(base, [[30.0 * i, 30.0 * j, 0.0] for i in range(2) for j in range(3)])
    ctx, ora = tiled(engine, spec), ti.make_context(oracle, spec)
    check_state(ctx, ora, spec)
    check_sets(ctx, ora)
    assert n_tiled(ctx) >= 2


def test_tiles_wca_and_lj_forms(engine, oracle):
    """the remaining members of the two-zone 12-6 family on tile lists: WCA (geometric-mean epsilon), LJ-repulsive alone,
    LJ-attractive alone (inner zone without force, switched tail), 63 parameter types"""
    from opencafemol_b200.capi import Context, PairTopology, System

    rng = np.random.default_rng(11)
    n = 1500
    box = np.array([9.0, 8.5, 9.5])
    g = np.stack(np.meshgrid(np.arange(12), np.arange(12), np.arange(12), indexing="ij"), -1).reshape(-1, 3)[:n]
    pos = g * 0.72 + rng.uniform(-0.12, 0.12, (n, 3)) + 0.2
    excl = np.stack([np.arange(n - 1), np.arange(1, n)], 1).astype(np.int32)
    # (at most 255 distinct per-bead parameter sets per pair term: the per-type energy tables of the observer side)
    sig = rng.choice(np.linspace(0.45, 0.6, 7), n)
    eps = rng.choice(np.linspace(0.3, 1.2, 9), n)
    for kind in ("wca", "lj_repulsive", "lj_attractive"):
        ctxs = []
        for lib in (engine, oracle):
            s = System(lib, np.full(n, 100.0), box)
            topo = PairTopology(excl=excl, periodic=True)
            if kind == "wca":
                s.add_wca(0, 2.0, eps, sig, topo)
            elif kind == "lj_repulsive":
                s.add_lj_repulsive(0, 2.0, eps, sig, topo)
            else:
                s.add_lj_attractive(0, 2.0, 1.8, eps, sig, topo)
            c = Context(s, 300.0, 0.5, 0.002, seed=5)
            if lib is engine:
                c.set_option("pair_kernel", 1)
            c.set_positions(pos)
            ctxs.append(c)

        class Spec:
            group_ids = {"0": 0}
        check_state(ctxs[0], ctxs[1], Spec)
        check_sets(ctxs[0], ctxs[1])
        assert ctxs[0].counter("tiled.0") == 1


def test_tiles_cell_assignment_bit_exact(engine, oracle):
    """the fine Hilbert-ordered grid of the tile lists: linear cell ids and grid dimensions against the oracle's restatement"""
    import ctypes as C
    for spec in (synth.hps_slab(n_chains=12), load_golden_spec("nucleosome")):
        ctx, ora = tiled(engine, spec), ti.make_context(oracle, spec)
        ctx.get_state(energy=True)
        sync_canon(ctx, ora)
        for t in range(ctx.num_pair_terms()):
            width, cap = ctx.counter(f"cell_width_nm.{t}"), int(ctx.counter(f"cell_cap.{t}"))
            ids, dims = ctx.cell_ids(t)
            oid = np.zeros(spec.mass.size, np.int32)
            od = np.zeros(3, np.int32)
            fn = oracle.dll.oracle_get_cell_ids_grid
            fn.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_longlong, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
            assert fn(ora.handle, t, width, cap, oid.ctypes.data_as(C.POINTER(C.c_int32)), od.ctypes.data_as(C.POINTER(C.c_int32))) == 0
            assert np.array_equal(dims, od), (dims, od)
            assert np.array_equal(ids, oid)


def test_tiles_trajectory_tracks_oracle_and_rows(engine, oracle):
    spec = synth.hps_slab(n_chains=8)
    ctx, ora, rows = tiled(engine, spec), ti.make_context(oracle, spec), ti.make_context(engine, spec)
    rows.set_option("pair_kernel", 0)
    ctx.step(25); ora.step(25); rows.step(25)
    a, b, c = (x.get_state(positions=True, velocities=True) for x in (ctx, ora, rows))
    assert np.abs(a["positions"] - b["positions"]).max() < 2e-5
    assert np.abs(a["velocities"] - b["velocities"]).max() < 2e-3
    assert np.abs(a["positions"] - c["positions"]).max() < 2e-5
    # several rebuilds, adaptive interval, no rollback storms
    ctx.step(600)
    assert np.isfinite(ctx.get_state(energy=True)["potential_energy"]) and ctx.counter("rollbacks") <= 2


def test_tiles_run_to_run_determinism(engine):
    """deterministic lists (ordered compaction), fixed warp -> entry assignment, fixed-point accumulation: bit-reproducible"""
    spec = synth.hps_slab(n_chains=10)
    res = []
    for _ in range(2):
        ctx = tiled(engine, spec)
        ctx.set_option("nstlist", 10)
        ctx.step(100)
        res.append(ctx.get_state(positions=True, velocities=True, forces=True))
    for k in ("positions", "velocities", "forces"):
        assert np.array_equal(res[0][k], res[1][k]), k


def test_tiles_nve_energy_conservation(engine):
    spec = synth.hps_slab(n_chains=12)
    spec.friction = 0.0
    ctx = tiled(engine, spec)
    ctx.step(200)
    e = []
    for _ in range(10):
        ctx.step(100)
        st = ctx.get_state(energy=True)
        e.append(st["potential_energy"] + st["kinetic_energy"])
    e = np.array(e)
    assert np.abs(e - e.mean()).max() < 5e-3 * st["kinetic_energy"]


@pytest.mark.parametrize("world,axis", [(2, 1), (3, 2)])
def test_tiles_slab_decomposition(engine, world, axis):
    """tile lists under the slab decomposition (in-process transport): every rank's forces on the beads it owns equal the
    single-domain forces to FP32 summation order (the clusters differ between the decompositions, so not bit for bit),
    energies add up, and a trajectory with migrations tracks the single-domain run"""
    spec = synth.hps_slab(n_chains=30)
    ref = tiled(engine, spec)
    ref.set_option("nstlist", 10)
    a = ref.get_state(forces=True, energy=True)
    run = SlabRun(engine, spec, world, axis, nstlist=10)
    try:
        for c in run.ctx:
            c.set_option("pair_kernel", 1)
        for st in run.each(lambda c: c.get_state(forces=True, energy=True)):
            assert force_error(st["forces"], a["forces"]).max() < 2e-6
            assert np.allclose(st["energy_by_group"], a["energy_by_group"], rtol=2e-6, atol=1e-4)
        ref.step(200)
        run.each(lambda c: c.step(200))
        b = ref.get_state(positions=True)
        for st in run.each(lambda c: c.get_state(positions=True)):
            assert np.abs(st["positions"] - b["positions"]).max() < 1e-4
    finally:
        run.close()
