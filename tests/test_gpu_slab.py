"""Slab domain decomposition (config C5) checked on ONE GPU: the ranks are threads of this process
(in-process transport), every rank on cuda:0, and the decomposed run must reproduce the single-domain run bit for
bit — same ownership as the oracle's integer rule, same pair set, same fixed-point forces, same trajectory.
The NCCL transport (one process per GPU) is exercised by tests/test_gpu_slab_nccl.py on boxes with >= 2 GPUs."""
import threading

import numpy as np
import pytest

from opencafemol_b200 import synth, toml_input as ti
from opencafemol_b200.capi import Comm

pytestmark = pytest.mark.gpu


def run_ranks(world, fn):
    """fn(rank) on `world` host threads (every engine call of a slab context is collective)."""
    out, err = [None] * world, [None] * world

    def body(r):
        try:
            out[r] = fn(r)
        except BaseException as e:  # noqa: BLE001
            err[r] = e

    th = [threading.Thread(target=body, args=(r,)) for r in range(world)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=600)
    for e in err:
        if e is not None:
            raise e
    return out


class SlabRun:
    def __init__(self, lib, spec, world, axis, nstlist=None, skin=None):
        self.world = world
        self.comm = Comm.local(lib, world)

        def make(r):
            ctx = ti.make_context(lib, spec, device=0, comm=self.comm, rank=r, axis=axis)
            if skin is not None:
                ctx.set_option("skin_nm", skin)
            if nstlist is not None:
                ctx.set_option("nstlist", nstlist)
            return ctx

        self.ctx = run_ranks(world, make)

    def each(self, fn):
        return run_ranks(self.world, lambda r: fn(self.ctx[r]))

    def close(self):
        for c in self.ctx:
            c.close()
        self.comm.close()


@pytest.mark.parametrize("world,axis,n_chains", [(2, 1, 24), (4, 2, 40), (3, 2, 12)])
def test_slab_forces_energies_roles(engine, oracle, world, axis, n_chains):
    spec = synth.hps_slab(n_chains=n_chains)
    ref = ti.make_context(engine, spec)
    a = ref.get_state(positions=True, forces=True, energy=True)
    run = SlabRun(engine, spec, world, axis)
    try:
        states = run.each(lambda c: c.get_state(positions=True, velocities=True, forces=True, energy=True))
        for st in states:
            assert np.array_equal(st["positions"], a["positions"])
            # fixed-point accumulation + identical row order: the decomposed forces are bit-identical
            assert np.array_equal(st["forces"], a["forces"])
            assert np.allclose(st["energy_by_group"], a["energy_by_group"], rtol=1e-12, atol=1e-9)
            assert st["kinetic_energy"] == pytest.approx(a["kinetic_energy"], rel=1e-12)
        # ownership and halo membership: the oracle's restatement of the integer rule, bit-exact
        roles = run.each(lambda c: c.roles())
        info = run.ctx[0].slab_info()
        assert info["world"] == world and info["axis"] == axis and info["halo_nm"] > 0
        ocomm = Comm.local(oracle, world)
        for r in range(world):
            ora = ti.make_context(oracle, spec, comm=ocomm, rank=r, axis=axis)
            ora.set_option("slab_halo_nm", info["halo_nm"])
            assert np.array_equal(roles[r], ora.roles()), f"roles of rank {r}"
        owned = np.stack([ro == 1 for ro in roles])
        assert (owned.sum(0) == 1).all(), "every bead has exactly one owner"
        # pair sets: the union over ranks is the single-domain set, no pair found twice
        for t in range(ref.num_pair_terms()):
            parts = run.each(lambda c, t=t: c.neighbor_pairs(t))
            merged = np.concatenate(parts)
            merged = merged[np.lexsort((merged[:, 1], merged[:, 0]))]
            assert np.array_equal(merged, ref.neighbor_pairs(t)), f"pair set of term {t}"
    finally:
        run.close()


@pytest.mark.parametrize("world,axis", [(2, 1), (4, 2)])
def test_slab_trajectory_bit_exact(engine, world, axis):
    """Several rebuilds with beads migrating between slabs: positions and velocities equal the single-domain run
    bit for bit (counter-based noise, fixed-point forces, deterministic lists, same rebuild cadence)."""
    spec = synth.hps_slab(n_chains=30)
    ref = ti.make_context(engine, spec)
    ref.set_option("nstlist", 10)
    run = SlabRun(engine, spec, world, axis, nstlist=10)
    try:
        owners0 = np.stack([ro == 1 for ro in run.each(lambda c: c.roles())]).argmax(0)
        for nsteps in (35, 400):
            ref.step(nsteps)
            run.each(lambda c, n=nsteps: c.step(n))
            a = ref.get_state(positions=True, velocities=True, energy=True)
            for st in run.each(lambda c: c.get_state(positions=True, velocities=True, energy=True)):
                assert np.array_equal(st["positions"], a["positions"])
                assert np.array_equal(st["velocities"], a["velocities"])
                assert st["potential_energy"] == pytest.approx(a["potential_energy"], rel=1e-12)
        owners1 = np.stack([ro == 1 for ro in run.each(lambda c: c.roles())]).argmax(0)
        if axis == 1:  # rows of the start structure run along y, through the slab faces
            assert (owners0 != owners1).sum() > 0, "the run was long enough for beads to change slabs"
        assert run.ctx[0].counter("halo_exchanges") > 400
    finally:
        run.close()


def test_slab_adaptive_rebuild_and_message_growth(engine):
    """Adaptive rebuild interval (the all-reduced flags keep every rank on the same cadence) and halo messages that
    start too small (forced overflow -> roll back, grow, replay) still track the single-domain run."""
    spec = synth.hps_slab(n_chains=30)
    ref = ti.make_context(engine, spec)
    ref.step(300)
    a = ref.get_state(positions=True)
    run = SlabRun(engine, spec, 2, 1)
    try:
        for c in run.ctx:
            c.set_option("halo_capacity", 256)
        run.each(lambda c: c.step(300))
        assert all(c.get_option("halo_capacity") > 256 for c in run.ctx)
        for st in run.each(lambda c: c.get_state(positions=True)):
            # different rebuild cadence changes only the FP32 summation order
            assert np.abs(st["positions"] - a["positions"]).max() < 1e-4
        nst = {c.counter("nstlist") for c in run.ctx}
        assert len(nst) == 1
    finally:
        run.close()


def test_slab_rejects_unsupported(engine):
    from opencafemol_b200.capi import OcmError

    spec = synth.hps_slab(n_chains=6)
    comm = Comm.local(engine, 8)
    with pytest.raises(OcmError, match="wider than the halo"):
        ti.make_context(engine, spec, comm=comm, rank=0, axis=1)  # 17.7 nm / 8 slabs < halo
    comm.close()
