"""Monte Carlo barostats: MonteCarloAnisotropicBarostat (ensemble "NPT") and
MonteCarloMembraneBarostat (ensemble "NPgammaT") as the reference configures them (MonteCarloAnisotropicBarostatGenerator.cpp:9-17,
MonteCarloMembraneBarostatGenerator.cpp:6-14, SystemGenerator.cpp:89-108, ReadTOMLInput.cpp:77-175).

Fixtures (tests/golden/make_ref_barostat.py): the two POPC samples as shipped run through the reference's own reader, generators
and Context on oracle/ref_shim, whose barostat restates OpenMM 8.0's algorithm (OpenMM is not under /root/reference: this part of
the parity chain is pinned to that restatement, not to OpenMM binaries).  CPU tests pin the oracle's independent restatement to
it; GPU tests compare the engine with the oracle trial by trial."""
import gzip
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_golden_spec

CASES = ["membrane_npt", "membrane_npgt"]


def load_ref(name):
    with gzip.open(os.path.join(GOLDEN, f"ref_{name}.json.gz"), "rt") as fh:
        return json.load(fh)


def context_of(lib, name):
    from opencafemol_b200 import toml_input as ti

    spec = load_golden_spec(name)
    assert spec.ensemble is not None
    spec.velocities = None
    spec.vel_seed = 2027
    return ti.make_context(lib, spec), spec


@pytest.mark.parametrize("name", CASES)
def test_reader_takes_the_ensemble_of_the_input(name):
    spec = load_golden_spec(name)
    ref = load_ref(name)
    assert spec.ensemble == ref["ensemble"]
    if name == "membrane_npt":
        assert spec.ensemble == {"type": "NPT", "pressure": [0.0, 0.0, 0.0], "scale_axis": [True, True, False], "frequency": 5}
    else:
        # the generator does not pass the input's frequency on: OpenMM's default 25 (MonteCarloMembraneBarostatGenerator.cpp:8-12)
        assert spec.ensemble["type"] == "NPgammaT" and spec.ensemble["xymode"] == "Isotropic" and spec.ensemble["zmode"] == "Fixed"
        assert spec.ensemble["frequency"] == 25


@pytest.mark.parametrize("name", CASES)
def test_oracle_barostat_reproduces_the_reference_run(oracle, name):
    """box edges and potential energy after each of 12 windows of `frequency` steps (one trial each): the oracle's own
    restatement, fed by the Python reader, against the reference pipeline on oracle/ref_shim; both in double precision with the
    same SplitMix64 stream, so they agree to rounding"""
    ref = load_ref(name)
    ctx, spec = context_of(oracle, name)
    assert ctx.barostat_info()["molecules"] == 128
    box = [ctx.box()]
    epot = [ctx.get_state(energy=True)["potential_energy"]]
    for _ in range(12):
        ctx.step(ref["window"])
        box.append(ctx.box())
        epot.append(ctx.get_state(energy=True)["potential_energy"])
    box, gbox = np.array(box), np.array(ref["box_nm"])
    assert np.abs(box - gbox).max() <= 1e-9
    assert len({tuple(b) for b in box.tolist()}) >= 6          # the box did change
    assert np.abs(np.array(epot) - np.array(ref["potential_energy_kj"])).max() <= 1e-6 * np.abs(ref["potential_energy_kj"]).max()
    info = ctx.barostat_info()
    assert info["trials"] == 12 and 0 < info["accepted"] <= 12
    x = ctx.get_state(positions=True)["positions"]
    assert np.abs(x - np.array(ref["final_positions"])).max() <= 1e-8
    if name == "membrane_npt":
        assert np.all(box[:, 2] == 20.0)                        # scale_axis = [true, true, false]
    else:
        assert np.array_equal(box[:, 0], box[:, 1]) and np.all(box[:, 2] == 20.0)   # isotropic XY, fixed Z


def test_oracle_barostat_schedule_and_modes(oracle):
    """OpenMM calls the barostat before every MD step and it acts at every frequency-th call: trials come before steps
    f, 2f, ... whatever the split into step() calls; constant-volume mode keeps the volume, anisotropic XY moves one edge"""
    ctx, spec = context_of(oracle, "membrane_npgt")
    ctx.set_membrane_barostat(1.0, 50.0, spec.temperature, "Anisotropic", "ConstantVolume", frequency=4, seed=5)
    v0 = np.prod(ctx.box())
    for n, want in ((3, 0), (1, 1), (3, 1), (2, 2), (7, 4)):
        ctx.step(n)
        assert ctx.barostat_info()["trials"] == want
        assert abs(np.prod(ctx.box()) / v0 - 1.0) < 1e-12
    assert ctx.barostat_info()["accepted"] > 0
    b = ctx.box()
    assert b[0] != 6.68 or b[1] != 6.68
    ctx.remove_barostat()
    ctx.step(8)
    assert ctx.barostat_info()["trials"] == 4
    ctx.set_membrane_barostat(1.0, 0.0, spec.temperature, "Isotropic", "Free", frequency=1, seed=6)
    axes = set()
    for _ in range(24):
        ctx.step(1)
        axes.add(ctx.barostat_info()["axis"])
    assert axes == {0, 2}                                       # isotropic XY never draws Y alone; free Z is drawn


def test_barostat_needs_a_periodic_box(oracle):
    from opencafemol_b200 import toml_input as ti
    from opencafemol_b200.capi import OcmError

    spec = load_golden_spec("sh3")
    ctx = ti.make_context(oracle, spec)
    with pytest.raises(OcmError, match="periodic boundary"):
        ctx.set_anisotropic_barostat([1.0, 1.0, 1.0], 300.0)
    assert np.all(ctx.box() == 0.0)


# ---- engine ------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_engine_barostat_tracks_the_oracle_trial_by_trial(engine, oracle, name):
    """same input, seeds and random stream: the engine (FP32 pair arithmetic) must draw the same axis and volume change, see the
    same energies to FP32 accuracy, take the same decision and end every window in the same box as the oracle"""
    ref = load_ref(name)
    ce, spec = context_of(engine, name)
    co, _ = context_of(oracle, name)
    assert ce.barostat_info()["molecules"] == co.barostat_info()["molecules"] == 128
    for k in range(12):
        ce.step(ref["window"])
        co.step(ref["window"])
        ie, io = ce.barostat_info(), co.barostat_info()
        assert ie["trials"] == io["trials"] == k + 1
        assert ie["axis"] == io["axis"] and ie["last_accepted"] == io["last_accepted"], (k, ie, io)
        assert abs(ie["dV"] - io["dV"]) <= 1e-12 * max(1.0, abs(io["dV"]))
        scale = max(abs(io["e_before"]), 1.0)
        assert abs(ie["e_before"] - io["e_before"]) <= 5e-5 * scale and abs(ie["e_after"] - io["e_after"]) <= 5e-5 * scale, (k, ie, io)
        assert abs(ie["w"] - io["w"]) <= 1e-4 * scale
        assert np.abs(ce.box() - co.box()).max() <= 1e-12
        assert np.abs(ce.box() - np.array(ref["box_nm"][k + 1])).max() <= 1e-9
    xe = ce.get_state(positions=True)["positions"]
    xo = co.get_state(positions=True)["positions"]
    assert np.abs(xe - xo).max() <= 2e-4                        # 60 / 300 steps of an FP32 trajectory
    se, so = ce.get_state(energy=True), co.get_state(energy=True)
    assert abs(se["potential_energy"] - so["potential_energy"]) <= 2e-3 * abs(so["potential_energy"])


@pytest.mark.gpu
def test_engine_set_box_rescales_the_layout(engine, oracle):
    """Context::setPeriodicBoxVectors: positions kept, cell grids / fixed-point positions / lists laid out again; energies and forces
    in the new box against the oracle in the same box; a box below twice the cutoff is refused like OpenMM refuses it"""
    from opencafemol_b200.capi import OcmError

    ce, spec = context_of(engine, "membrane_npt")
    co, _ = context_of(oracle, "membrane_npt")
    ce.remove_barostat(); co.remove_barostat()
    # (edges that keep the bilayer's periodic images apart: the positions stay where they are)
    for box in ([6.67, 6.9, 19.0], [7.4, 6.675, 25.0], [6.68, 6.68, 20.0]):
        ce.set_box(box); co.set_box(box)
        assert np.array_equal(ce.box(), np.array(box))
        se = ce.get_state(forces=True, energy=True)
        so = co.get_state(forces=True, energy=True)
        assert abs(se["potential_energy"] - so["potential_energy"]) <= 1e-5 * abs(so["potential_energy"])
        assert np.abs(se["forces"] - so["forces"]).max() <= 1e-5 * np.abs(so["forces"]).max()
        ce.step(7); co.step(7)
        xe, xo = ce.get_state(positions=True)["positions"], co.get_state(positions=True)["positions"]
        assert np.abs(xe - xo).max() <= 2e-5
        ce.set_positions(xo)
    with pytest.raises(OcmError, match="less than twice the nonbonded cutoff"):
        ce.set_box([6.68, 3.0, 20.0])


@pytest.mark.gpu
def test_engine_barostat_schedule_and_refusals(engine):
    from opencafemol_b200 import toml_input as ti
    from opencafemol_b200.capi import OcmError

    ctx, spec = context_of(engine, "membrane_npgt")
    ctx.set_membrane_barostat(1.0, 50.0, spec.temperature, "Anisotropic", "ConstantVolume", frequency=4, seed=5)
    v0 = np.prod(ctx.box())
    for n, want in ((3, 0), (1, 1), (3, 1), (2, 2), (7, 4)):
        ctx.step(n)
        assert ctx.barostat_info()["trials"] == want
        assert abs(np.prod(ctx.box()) / v0 - 1.0) < 1e-12
    assert ctx.current_step == 16
    # a long NPT run stays finite and keeps the acceptance in OpenMM's tuning band once the step size has adapted
    z0 = ctx.box()[2]
    ctx.set_anisotropic_barostat([1.0, 1.0, 1.0], spec.temperature, (True, True, False), frequency=5, seed=9)
    ctx.step(1500)
    info = ctx.barostat_info()
    assert info["trials"] == 300 and 0.15 <= info["accepted"] / info["trials"] <= 0.85
    b = ctx.box()
    assert b[2] == z0 and 5.5 < b[0] < 7.5 and 5.5 < b[1] < 7.5
    assert np.isfinite(ctx.get_state(energy=True)["potential_energy"])
    sh3 = ti.make_context(engine, load_golden_spec("sh3"))
    with pytest.raises(OcmError, match="periodic boundary"):
        sh3.set_anisotropic_barostat([1.0, 1.0, 1.0], 300.0)
