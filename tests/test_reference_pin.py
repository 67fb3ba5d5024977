"""Parity pinned by the REFERENCE ITSELF.

tests/golden/ref_*.json.gz were produced by tests/golden/make_ref_golden.py from oracle/_ref/libocm_ref.so: the
reference's own reader, Topology, force-field generators, SystemGenerator, Simulator and Observers, compiled
unmodified (oracle/Makefile, target `ref`) against stand-in OpenMM / toml11 headers whose Custom*Force classes
evaluate the reference's own energy-expression strings (oracle/ref_shim).

  * CPU (`-m "not gpu"`): the oracle restatement against those vectors (energies 1e-10, forces 1e-9, 20-step Langevin
    trajectory 1e-8); and, where oracle/_ref and /root/reference exist (the build container), live checks of the
    parameter pipeline: exclusion sets, interaction groups, cutoffs, donors/acceptors bit-exact between this
    repository's reader and the reference's, plus the TOML reader against Python's tomllib.
  * GPU (`-m gpu`): the CUDA engine through the C ABI against the same vectors at the north-star tolerance 1e-5.
"""
import glob
import gzip
import hashlib
import json
import os
import sys
import tomllib

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, load_golden_spec
from opencafemol_b200 import toml_input as ti

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_binding as rb  # noqa: E402

# "membrane": the POPC sample in the NVT ensemble with a Trigonometric term added (tests/golden/make_ref_membrane.py)
# "external" / "external_pbc": the SH3 sample and 10 FUS chains with the external generators (PositionRestraint, Pulling,
# EXVRectangularBox, CylindricalRestraint, HarmonicCoMPulling) and a CombinatorialGoContact added (tests/golden/make_ref_external.py)
CASES = ["sh3", "dna_3spn2", "nucleosome", "fus10", "membrane", "external", "external_pbc"]
REF_SAMPLES = "/root/reference/sample/toml_input"
PAIR_KINDS = ("excluded_volume", "debye_huckel", "lj_repulsive", "lj_attractive", "3spn2_excluded_volume", "wca", "isolf_attractive",
              "trigonometric")
DA_KINDS = ("3spn2_base_pair", "3spn2_cross_stacking", "pdns", "combinatorial_go_contact")


def load_ref(name):
    with gzip.open(os.path.join(GOLDEN, f"ref_{name}.json.gz"), "rt") as fh:
        return json.load(fh)


def force_error(f, fref):
    frms = np.sqrt((fref ** 2).sum(1).mean())
    return np.linalg.norm(f - fref, axis=1) / np.maximum(np.linalg.norm(fref, axis=1), frms)


def context_at_ref_state(lib, name, gold):
    spec = load_golden_spec(name)
    ctx = ti.make_context(lib, spec)
    ctx.set_positions(np.array(gold["positions"]))
    if gold["vel_seed_used"] is not None:
        ctx.set_velocities_to_temperature(spec.temperature, gold["vel_seed_used"])
    return spec, ctx


# ---------------------------------------------------------------------------------------------------------
# oracle vs the reference's vectors (CPU)
@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_vectors(oracle, name):
    gold = load_ref(name)
    spec, ctx = context_at_ref_state(oracle, name, gold)
    assert spec.group_ids == gold["group_map"]  # SystemGenerator::add_ff_generator group numbering
    st = ctx.get_state(forces=True, energy=True)
    escale = max(max(abs(v) for v in gold["energy_by_group_kj"].values()), 1.0)
    for k, g in spec.group_ids.items():
        ref = gold["energy_by_group_kj"][k]
        assert abs(st["energy_by_group"][g] - ref) <= 1e-10 * max(abs(ref), 1e-3 * escale), (k, st["energy_by_group"][g], ref)
    assert st["potential_energy"] == pytest.approx(gold["potential_energy_kj"], rel=1e-10)
    assert force_error(st["forces"], np.array(gold["forces"])).max() < 1e-9
    assert st["kinetic_energy"] == pytest.approx(gold["kinetic_energy_kj"], rel=1e-9)
    integ = gold["integrator"]
    assert spec.dt == pytest.approx(integ["dt"], rel=1e-14)
    assert spec.friction == pytest.approx(integ["friction"], rel=1e-14)
    assert spec.temperature == integ["temperature"] and spec.seed == integ["seed"]


def test_oracle_matches_reference_vectors_name_pair_tables(oracle):
    """"wcamix" (tests/golden/make_ref_golden.py: per-particle WCA, the name-pair table forms UniformWeeksChandlerAndersen and
    UniformLennardJonesAttractive, capped Go contacts): this repository's reader on the input the reference ran, the oracle against the
    reference's energies and forces"""
    gold = load_ref("wcamix")
    with gzip.open(os.path.join(GOLDEN, "ref_wcamix_run.toml.gz"), "rt") as fh:
        spec = ti.read_toml(tomllib.loads(fh.read()))
    assert spec.group_ids == gold["group_map"]
    ctx = ti.make_context(oracle, spec)
    ctx.set_positions(np.array(gold["positions"]))
    st = ctx.get_state(forces=True, energy=True)
    for k, g in spec.group_ids.items():
        assert st["energy_by_group"][g] == pytest.approx(gold["energy_by_group_kj"][k], rel=1e-10, abs=1e-9), k
    assert force_error(st["forces"], np.array(gold["forces"])).max() < 1e-9
    # the digest of the Forces the reference built: one CustomNonbondedForce per table entry, same cutoffs and interaction groups
    ref_nb = [f for f in gold["force_digest"] if f["class"] == "CustomNonbondedForce"]
    mine_nb = [t for t in spec.terms if t.kind in PAIR_KINDS]
    assert len(ref_nb) == len(mine_nb) == 7

    def key(cut, groups):
        return (round(cut, 9), tuple(sorted(groups)))
    ref_keys = sorted(key(r["cutoff"], [(a, b, sha) for a, b, sha in r["interaction_groups"]]) for r in ref_nb)
    mine_keys = []
    for t in mine_nb:
        groups = []
        for a, b in t.args["topo"].igroups:
            a, b = sorted(map(int, a)), sorted(map(int, b))
            groups.append((len(a), len(b), hashlib.sha1(np.asarray(a + b, np.int32).tobytes()).hexdigest()))
        mine_keys.append(key(t.args["rc"], groups))
    assert sorted(mine_keys) == ref_keys   # (the reference walks its table in unordered_map order: compare as sets)


def test_oracle_matches_reference_vectors_local_dna_terms(oracle):
    """"dna_localize" (sample/toml_input/dna_3SPN2_localize.toml: 3SPN2BasePairLocal / 3SPN2CrossStackingLocal, explicit nucleotide pairs):
    this repository's reader on the input the reference ran, the oracle against the reference's energies and forces"""
    gold = load_ref("dna_localize")
    with gzip.open(os.path.join(GOLDEN, "ref_dna_localize_run.toml.gz"), "rt") as fh:
        spec = ti.read_toml(tomllib.loads(fh.read()))
    assert spec.group_ids == gold["group_map"]
    ctx = ti.make_context(oracle, spec)
    ctx.set_positions(np.array(gold["positions"]))
    st = ctx.get_state(forces=True, energy=True)
    escale = max(max(abs(v) for v in gold["energy_by_group_kj"].values()), 1.0)
    for k, g in spec.group_ids.items():
        ref = gold["energy_by_group_kj"][k]
        assert abs(st["energy_by_group"][g] - ref) <= 1e-10 * max(abs(ref), 1e-3 * escale), (k, st["energy_by_group"][g], ref)
    assert force_error(st["forces"], np.array(gold["forces"])).max() < 1e-9


def test_oracle_matches_reference_fus100_full_size(oracle):
    """sample/toml_input/FUS_100_HPS.toml at full size (16 300 beads, config C2's sample): the oracle against the reference's own
    energies and forces at relaxed coordinates (tests/golden/make_ref_fus100.py)."""
    spec = load_golden_spec("ref_fus100")
    with np.load(os.path.join(GOLDEN, "ref_fus100.npz")) as d:
        fref, names, eg = d["forces"], [str(s) for s in d["group_names"]], d["group_energies_kj"]
        assert spec.group_ids == json.loads(str(d["group_map"]))
    st = ti.make_context(oracle, spec).get_state(forces=True, energy=True)
    assert force_error(st["forces"], fref).max() < 1e-9
    for name, ref in zip(names, eg):
        assert st["energy_by_group"][spec.group_ids[name]] == pytest.approx(ref, rel=1e-10, abs=1e-7)


@pytest.mark.parametrize("name", ["sh3", "dna_3spn2", "nucleosome", "membrane", "external", "external_pbc"])
def test_oracle_trajectory_matches_reference_run(oracle, name):
    """20 leap-frog Langevin steps of the reference's LangevinIntegrator semantics with the shared Philox noise."""
    gold = load_ref(name)
    spec, ctx = context_at_ref_state(oracle, name, gold)
    ctx.step(gold["trajectory"]["steps"])
    st = ctx.get_state(positions=True, velocities=True)
    assert np.abs(st["positions"] - np.array(gold["trajectory"]["positions"])).max() < 1e-8
    assert np.abs(st["velocities"] - np.array(gold["trajectory"]["velocities"])).max() < 1e-6


def test_reader_matches_reference_force_digest():
    """counts, cutoffs, switching distances and exclusion hashes of every Force the reference built"""
    for name in CASES:
        gold = load_ref(name)
        spec = load_golden_spec(name)
        ref_nb = [f for f in gold["force_digest"] if f["class"] == "CustomNonbondedForce"]
        ref_hb = [f for f in gold["force_digest"] if f["class"] == "CustomHbondForce"]
        mine_nb = [t for t in spec.terms if t.kind in PAIR_KINDS]
        mine_hb = [t for t in spec.terms if t.kind in DA_KINDS]
        assert len(ref_nb) == len(mine_nb) and len(ref_hb) == len(mine_hb), name
        for r, t in zip(ref_nb, mine_nb):
            topo = t.args["topo"]
            assert r["group"] == t.group
            assert t.args["rc"] == pytest.approx(r["cutoff"], rel=1e-13), (name, t.name)
            assert (r["method"] == 2) == bool(topo.periodic)
            ex = np.asarray(sorted((min(a, b), max(a, b)) for a, b in np.asarray(topo.excl).reshape(-1, 2).tolist()), np.int32)
            assert len(ex) == r["n_exclusions"], (name, t.name)
            assert hashlib.sha1(ex.tobytes()).hexdigest() == r["exclusions_sha1"], (name, t.name)
            assert len(topo.igroups) == len(r["interaction_groups"]), (name, t.name)
            for (a, b), (na, nb, sha) in zip(topo.igroups, r["interaction_groups"]):
                a, b = sorted(map(int, a)), sorted(map(int, b))
                assert (len(a), len(b)) == (na, nb)
                assert hashlib.sha1(np.asarray(a + b, np.int32).tobytes()).hexdigest() == sha
            if r.get("use_switch"):
                assert t.args["r_switch"] == pytest.approx(r["switch_distance"], rel=1e-13)
        for r, t in zip(ref_hb, mine_hb):
            topo = t.args["topo"]
            assert r["group"] == t.group
            assert topo.cutoff == pytest.approx(r["cutoff"], rel=1e-13)
            assert len(topo.donors) == r["n_donors"] and len(topo.acceptors) == r["n_acceptors"]
            ex = np.asarray(sorted(map(tuple, np.asarray(topo.excl).reshape(-1, 2).tolist())), np.int32)
            assert len(ex) == r["n_exclusions"]
            assert hashlib.sha1(ex.tobytes()).hexdigest() == r["exclusions_sha1"], (name, t.name)


# ---------------------------------------------------------------------------------------------------------
# live checks against oracle/_ref (build container only: needs the compiled reference and its sample inputs)
needs_ref = pytest.mark.skipif(not (rb.available() and os.path.isdir(REF_SAMPLES)),
                               reason="oracle/_ref (make -C oracle ref) and /root/reference are only present in the build container")


@needs_ref
def test_toml_reader_matches_tomllib():
    R = rb.RefLibrary()

    def norm(v):
        if isinstance(v, dict):
            return {k: norm(x) for k, x in v.items()}
        if isinstance(v, list):
            return [norm(x) for x in v]
        if isinstance(v, float) and v != v:
            return "nan"
        return v

    files = sorted(glob.glob(os.path.join(REF_SAMPLES, "*.toml")))
    assert len(files) >= 5
    for path in files:
        with open(path, "rb") as fh:
            want = tomllib.load(fh)
        got = R.toml_to_json(path)
        assert norm(got) == norm(want), path
        # and the toml library the reference itself was compiled with (its toml11 submodule) reads the same
        assert norm(R.toml_to_json(path, reader="toml11")) == norm(want), path


@needs_ref
@pytest.mark.parametrize("sample,name", [("sh3_AICG2+.toml", "sh3"), ("dna_3SPN2.toml", "dna_3spn2"), ("nucleosome_1kx5.toml", "nucleosome")])
def test_reader_tables_match_reference_live(oracle, sample, name):
    """every per-term parameter the reference hands to OpenMM vs what this repository's reader hands to the C ABI,
    compared through the energies of each Force group and the full description of the pair forces"""
    R = rb.RefLibrary()
    rs = R.open_toml(os.path.join(REF_SAMPLES, sample))
    spec = ti.read_toml(os.path.join(REF_SAMPLES, sample))
    desc = rs.describe()
    assert desc["num_particles"] == spec.mass.size
    assert np.allclose(desc["masses"], spec.mass, rtol=0, atol=0)
    ref_nb = [f for f in desc["forces"] if f["class"] == "CustomNonbondedForce"]
    mine_nb = [t for t in spec.terms if t.kind in PAIR_KINDS]
    assert len(ref_nb) == len(mine_nb)
    for r, t in zip(ref_nb, mine_nb):
        ex_ref = sorted({(min(a, b), max(a, b)) for a, b in r["exclusions"]})
        ex_mine = sorted({(min(a, b), max(a, b)) for a, b in np.asarray(t.args["topo"].excl).reshape(-1, 2).tolist()})
        assert ex_ref == ex_mine, t.name
        assert [(sorted(a), sorted(b)) for a, b in r["interaction_groups"]] == [(sorted(map(int, a)), sorted(map(int, b))) for a, b in t.args["topo"].igroups]
        # per-particle parameters: NaN marks a non-participant on both sides
        per = np.array([[np.nan if v is None else v for v in p] for p in r["particles"]], float)
        for col, pname in enumerate(r["per_particle_parameters"]):
            key = {"sigma": "radius" if t.kind == "excluded_volume" else "sigma", "q": "charge", "epsilon": "eps"}[pname.split("_")[-1]]
            mine = np.asarray(t.args[key], float)
            assert np.array_equal(np.isnan(mine), np.isnan(per[:, col])), (t.name, pname)
            assert np.allclose(mine[~np.isnan(mine)], per[~np.isnan(mine), col], rtol=1e-14, atol=0), (t.name, pname)
    # a perturbed geometry (not the native minimum, where many derivatives vanish)
    rng = np.random.default_rng(11)
    x = spec.positions + rng.normal(scale=0.02, size=spec.positions.shape)
    rs.set_positions(x)
    ctx = ti.make_context(oracle, spec)
    ctx.set_positions(x)
    st = ctx.get_state(forces=True, energy=True)
    eg = rs.energy_by_group()
    escale = max(max(abs(v) for v in eg.values()), 1.0)
    for k, g in spec.group_ids.items():
        assert abs(st["energy_by_group"][g] - eg[k]) <= 1e-10 * max(abs(eg[k]), 1e-3 * escale), (k, st["energy_by_group"][g], eg[k])
    _, f = rs.energy(forces=True)
    assert force_error(st["forces"], f).max() < 1e-9


@needs_ref
def test_wca_reader_and_oracle_match_reference_live(oracle, tmp_path):
    """per-particle Weeks-Chandler-Andersen: this repository's reader + oracle against the
    reference's reader + generator on the same input"""
    sys.path.insert(0, GOLDEN)
    import make_ref_golden as mk

    d = tomllib.loads(mk.wcamix_toml())
    d["forcefields"][0]["global"] = [g for g in d["forcefields"][0]["global"] if "table" not in g]
    path = tmp_path / "wca.toml"
    path.write_text(mk.dump_toml(d))
    rs = rb.RefLibrary().open_toml(str(path))
    spec = ti.read_toml(str(path))
    assert spec.group_ids == rs.group_map()
    ctx = ti.make_context(oracle, spec)
    rng = np.random.default_rng(2)
    x = spec.positions + rng.normal(scale=0.03, size=spec.positions.shape)
    rs.set_positions(x)
    ctx.set_positions(x)
    st = ctx.get_state(forces=True, energy=True)
    eg = rs.energy_by_group()
    assert eg["WeeksChandlerAndersen"] > 0
    for k, g in spec.group_ids.items():
        assert st["energy_by_group"][g] == pytest.approx(eg[k], rel=1e-11, abs=1e-9)
    _, f = rs.energy(forces=True)
    assert force_error(st["forces"], f).max() < 1e-9
    wca_ref = [fo for fo in rs.describe()["forces"] if fo["class"] == "CustomNonbondedForce"][0]
    wca_mine = [t for t in spec.terms if t.kind == "wca"][0]
    assert wca_mine.args["rc"] == pytest.approx(wca_ref["cutoff"], rel=1e-14)


@needs_ref
def test_reference_unit_tests_pass_on_the_checker():
    """the reference's own googletest file (test/test_read_toml_ff_generator.cpp) built by `make -C oracle ref`"""
    import subprocess

    exe = os.path.join(ROOT, "oracle", "_ref", "ref_tests")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "[  PASSED  ] 32 tests." in r.stdout, r.stdout[-2000:]


@needs_ref
def test_reference_forces_are_energy_gradients():
    """self-check of the stand-in's automatic differentiation: central differences of the reference's energy"""
    R = rb.RefLibrary()
    rs = R.open_toml(os.path.join(REF_SAMPLES, "dna_3SPN2.toml"))
    rng = np.random.default_rng(5)
    x = rs.positions() + rng.normal(scale=0.01, size=(rs.n, 3))
    rs.set_positions(x)
    _, f = rs.energy(forces=True)
    h = 1e-6
    for i, c in [(0, 0), (17, 1), (64, 2), (101, 0), (189, 1)]:
        xp, xm = x.copy(), x.copy()
        xp[i, c] += h
        xm[i, c] -= h
        rs.set_positions(xp)
        ep = rs.energy()
        rs.set_positions(xm)
        em = rs.energy()
        assert -(ep - em) / (2 * h) == pytest.approx(f[i, c], rel=2e-6, abs=1e-5)


# ---------------------------------------------------------------------------------------------------------
# the CUDA engine vs the reference's vectors
@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_engine_matches_reference_vectors(engine, name):
    gold = load_ref(name)
    spec, ctx = context_at_ref_state(engine, name, gold)
    st = ctx.get_state(forces=True, energy=True)
    assert force_error(st["forces"], np.array(gold["forces"])).max() < 1e-5
    escale = max(max(abs(v) for v in gold["energy_by_group_kj"].values()), 1.0)
    for k, g in spec.group_ids.items():
        ref = gold["energy_by_group_kj"][k]
        assert abs(st["energy_by_group"][g] - ref) <= 1e-5 * max(abs(ref), 1e-3 * escale), (k, st["energy_by_group"][g], ref)
    assert st["kinetic_energy"] == pytest.approx(gold["kinetic_energy_kj"], rel=1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["sh3", "dna_3spn2", "membrane", "external", "external_pbc"])
def test_engine_trajectory_tracks_reference_run(engine, name):
    """FP32 pair arithmetic and FP32 velocities against the reference's double-precision run, same noise"""
    gold = load_ref(name)
    spec, ctx = context_at_ref_state(engine, name, gold)
    ctx.step(gold["trajectory"]["steps"])
    st = ctx.get_state(positions=True)
    assert np.abs(st["positions"] - np.array(gold["trajectory"]["positions"])).max() < 2e-5
