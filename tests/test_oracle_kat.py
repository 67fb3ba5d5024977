"""Pins the CPU oracle: closed-form known answers of the reference's own SH3 sample (SURVEY.md
Appendix D), the committed golden fixture, finite differences F = -dE/dx for every term kind, and the
cell-list / brute-force agreement.  CPU only."""
import numpy as np
import pytest

from conftest import load_golden_spec, load_golden_values
from opencafemol_b200 import synth, toml_input as ti
from opencafemol_b200.capi import Context, DATopology, PairTopology, System

KCAL = 4.184


def test_sh3_known_answers(oracle):
    spec = load_golden_spec("sh3")
    ctx = ti.make_context(oracle, spec)
    e = ctx.get_state(energy=True)["energy_by_group"]
    g = spec.group_ids
    # sample is at its parameter minima: sum of k's (kcal/mol)
    assert abs(e[g["GoContact"]] / KCAL - (-112.5303)) < 1e-6
    assert abs(e[g["GaussianBond"]] / KCAL - (-140.674)) < 1e-5
    assert abs(e[g["GaussianDihedral"]] / KCAL - (-40.7812)) < 1e-5  # pins the dihedral sign convention
    assert abs(e[g["HarmonicBond"]] / KCAL) < 1e-8


def test_sh3_matches_golden(oracle):
    spec = load_golden_spec("sh3")
    gold = load_golden_values("sh3")
    ctx = ti.make_context(oracle, spec)
    st = ctx.get_state(forces=True, energy=True)
    for k, gid in spec.group_ids.items():
        assert st["energy_by_group"][gid] == pytest.approx(gold["energy_by_group_kj"][k], rel=1e-10, abs=1e-10)
    assert np.allclose(st["forces"], np.array(gold["forces"]), atol=1e-7)
    assert st["kinetic_energy"] == pytest.approx(gold["kinetic_energy_kj"], rel=1e-10)
    for t in range(ctx.num_pair_terms()):
        assert ctx.neighbor_pairs(t).tolist() == gold["pairs"][str(t)]


def test_sh3_exclusion_counts():
    spec = load_golden_spec("sh3")
    exv = [t for t in spec.terms if t.kind == "excluded_volume"][0]
    # bond<=3 (162 pairs) U contact<=1 (215 pairs), disjoint
    assert len(exv.args["topo"].excl) == 377
    assert exv.args["rc"] == pytest.approx((0.2348 + 0.2265) * 2.0, abs=1e-9)


def _fd_check(lib, build, n, box=None, seed=0, h=1e-6, tol=2e-6, masses=None):
    rng = np.random.default_rng(seed)
    sysm = System(lib, np.full(n, 50.0) if masses is None else masses, box)
    pos = build(sysm, rng)
    ctx = Context(sysm, 300.0, 0.2, 0.001, 1)
    ctx.set_positions(pos)
    f = ctx.get_state(forces=True)["forces"]
    fd = np.zeros_like(f)
    for i in range(n):
        for k in range(3):
            p = pos.copy(); p[i, k] += h; ctx.set_positions(p)
            ep = ctx.get_state(energy=True)["potential_energy"]
            p[i, k] -= 2 * h; ctx.set_positions(p)
            em = ctx.get_state(energy=True)["potential_energy"]
            fd[i, k] = -(ep - em) / (2 * h)
    scale = max(np.abs(f).max(), 1.0)
    assert np.abs(f).max() > 1e-3, "vacuous check: the configuration exerts no force"
    assert np.abs(f - fd).max() / scale < tol, (np.abs(f - fd).max(), scale)


def _chain(n, rng, bond=0.38):
    p = np.zeros((n, 3))
    for i in range(1, n):
        d = rng.normal(size=3); d /= np.linalg.norm(d)
        p[i] = p[i - 1] + bond * d
    return p + 3.0


BONDED = ["harmonic_bond", "gaussian_bond", "go_contact", "3spn2_bond", "harmonic_angle", "flexible_local_angle",
          "gaussian_dihedral", "cosine_dihedral", "flexible_local_dihedral", "capped_go_contact"]


@pytest.mark.parametrize("kind", BONDED)
@pytest.mark.parametrize("periodic", [False, True])
def test_bonded_finite_difference(oracle, kind, periodic):
    from opencafemol_b200 import tables
    n = 6
    box = np.array([2.5, 2.2, 2.0]) if periodic else None  # smaller than the chain: bonds cross the boundary

    def build(s, rng):
        ij = np.array([[i, i + 1] for i in range(n - 1)])
        ijk = np.array([[i, i + 1, i + 2] for i in range(n - 2)])
        ijkl = np.array([[i, i + 1, i + 2, i + 3] for i in range(n - 3)])
        m2, m3, m4 = len(ij), len(ijk), len(ijkl)
        if kind == "harmonic_bond":
            s.add_harmonic_bond(0, ij, rng.uniform(0.3, 0.4, m2), rng.uniform(100, 1000, m2), periodic)
        elif kind == "gaussian_bond":
            s.add_gaussian_bond(0, ij, -rng.uniform(1, 5, m2), rng.uniform(0.3, 0.5, m2), rng.uniform(0.05, 0.2, m2), periodic)
        elif kind == "go_contact":
            s.add_go_contact(0, ij, rng.uniform(0.5, 2, m2), rng.uniform(0.3, 0.45, m2), periodic)
        elif kind == "capped_go_contact":
            # bonds are 0.38 nm: with r0 in 0.4..0.6 some contacts sit on the linear cap (r/r0 <= 0.75), the others on the 12-10 curve
            cap = 0.75
            slope, icpt = -60 * (cap ** -13 - cap ** -11), 65 * cap ** -12 - 66 * cap ** -10
            s.add_capped_go_contact(0, ij, rng.uniform(0.5, 2, m2), np.array([0.4, 0.6, 0.45, 0.55, 0.42])[:m2], np.full(m2, slope), np.full(m2, icpt),
                                    np.full(m2, cap), periodic)
        elif kind == "3spn2_bond":
            s.add_3spn2_bond(0, ij, rng.uniform(100, 300, m2), rng.uniform(1e4, 3e4, m2), rng.uniform(0.3, 0.45, m2), periodic)
        elif kind == "harmonic_angle":
            s.add_harmonic_angle(0, ijk, rng.uniform(1.5, 2.5, m3), rng.uniform(10, 100, m3), periodic)
        elif kind == "flexible_local_angle":
            s.add_flexible_local_angle(0, ijk, rng.uniform(2, 5, m3), tables.FLA_SPLINE["PRO"], tables.FLA_MIN_THETA,
                                       tables.FLA_MAX_THETA, 0.6, periodic)
        elif kind == "gaussian_dihedral":
            s.add_gaussian_dihedral(0, ijkl, -rng.uniform(1, 3, m4), rng.uniform(-3, 3, m4), rng.uniform(0.15, 0.6, m4), periodic)
        elif kind == "cosine_dihedral":
            s.add_cosine_dihedral(0, ijkl, rng.uniform(1, 3, m4), rng.uniform(-3, 3, m4), rng.integers(1, 4, m4), periodic)
        else:
            s.add_flexible_local_dihedral(0, ijkl, rng.uniform(2, 5, m4), tables.FLD_FOURIER[("GLY", "PRO")], 1.0, periodic)
        return _chain(n, rng)

    _fd_check(oracle, build, n, box, seed=BONDED.index(kind))


def external_terms(s, rng, n, form, periodic=False):
    """One external / centre-of-mass / combinatorial term of the given form on an n-bead chain (shared by the oracle's
    finite-difference test and the GPU parity test); returns the positions."""
    from opencafemol_b200.capi import DATopology
    pos = _chain(n, rng)
    idx = np.arange(n)
    if form == "combinatorial_go_contact":
        # a random walk can bring a donor and an acceptor almost on top of each other; (r0/r)^12 then exceeds every force a run
        # would see (and the range of the engine's fixed-point accumulators, 2e9 kJ/mol/nm): draw again until none is closer than 0.3 nm
        while np.linalg.norm(pos[: n // 2, None, :] - pos[None, n // 2:, :], axis=2).min() < 0.3:
            pos = _chain(n, rng)
    if form == "position_restraint":
        s.add_position_restraint(2, idx, pos + rng.normal(scale=0.3, size=(n, 3)), rng.uniform(10, 100, n), rng.uniform(0.0, 0.4, n))
    elif form == "pulling":
        s.add_pulling(2, idx[::2], rng.normal(scale=20.0, size=(len(idx[::2]), 3)))
    elif form == "exv_rectangular_box":
        s.add_exv_rectangular_box(2, 0.6, pos.min(0) - 0.35, pos.max(0) + 0.3, idx, rng.uniform(0.2, 0.3, n))
    elif form == "cylindrical_restraint":
        axis = rng.normal(size=(n, 3)); axis /= np.linalg.norm(axis, axis=1)[:, None]
        shift = rng.normal(scale=2.0, size=(n, 3)); shift -= (shift * axis).sum(1)[:, None] * axis
        s.add_cylindrical_restraint(2, idx, axis, shift, rng.uniform(10, 100, n), rng.uniform(0.0, 0.5, n), periodic)
    elif form == "harmonic_com_pulling":
        s.add_harmonic_com_pulling(2, 350.0, 0.25, idx[: n // 2], idx[n // 2:], periodic)
        s.add_harmonic_com_pulling(3, 120.0, 1.5, idx[1::3], idx[::3], periodic)   # (interleaved groups)
    else:
        don = np.full((n // 2, 3), -1); don[:, 0] = idx[: n // 2]
        acc = np.full((n - n // 2, 3), -1); acc[:, 0] = idx[n // 2:]
        s.add_combinatorial_go_contact(2, DATopology(don, acc, 0.9, np.array([[0, 0], [1, 2]]), periodic), 1.3, 0.45)
    return pos


EXTERNAL = ["position_restraint", "pulling", "exv_rectangular_box", "cylindrical_restraint", "harmonic_com_pulling",
            "combinatorial_go_contact"]


@pytest.mark.parametrize("form", EXTERNAL)
@pytest.mark.parametrize("periodic", [False, True])
def test_external_finite_difference(oracle, form, periodic):
    """The reference's external generators (PositionRestraint, Pulling, EXVRectangularBox, CylindricalRestraint,
    HarmonicCoMPulling) and CombinatorialGoContact: F = -dE/dx of the oracle's restatement."""
    if periodic and form == "exv_rectangular_box":
        pytest.skip("the generator throws under a periodic boundary")
    n = 8
    box = np.array([2.5, 2.2, 2.0]) if periodic else None
    masses = np.random.default_rng(5).uniform(20.0, 120.0, n)

    def build(s, rng):
        return external_terms(s, rng, n, form, periodic)

    _fd_check(oracle, build, n, box, seed=EXTERNAL.index(form), masses=masses)


def test_external_argument_checks(oracle):
    s = System(oracle, np.full(4, 50.0), np.array([3.0, 3.0, 3.0]))
    with pytest.raises(RuntimeError, match="periodic"):
        s.add_exv_rectangular_box(0, 0.5, np.zeros(3), np.ones(3), [0], [0.2])
    with pytest.raises(RuntimeError, match="out of range"):
        s.add_pulling(0, [7], np.ones((1, 3)))
    with pytest.raises(RuntimeError, match="non-empty"):
        s.add_harmonic_com_pulling(0, 1.0, 0.0, [], [1])


@pytest.mark.parametrize("periodic", [False, True])
def test_base_stacking_finite_difference(oracle, periodic):
    n = 9
    box = np.array([2.5, 2.2, 2.0]) if periodic else None

    def build(s, rng):
        ijk = np.array([[i, i + 1, i + 2] for i in range(n - 2)])
        m = len(ijk)
        # K_BS (theta - theta0) spread over all three branches of the modulating function, dr of both signs
        s.add_3spn2_base_stacking(0, ijk, rng.uniform(10, 15, m), rng.uniform(0.3, 0.45, m), rng.uniform(1.2, 2.4, m), 30.0, 6.0, periodic)
        return _chain(n, rng)

    _fd_check(oracle, build, n, box, seed=31, tol=5e-6)


DA_KINDS = ["3spn2_base_pair", "3spn2_cross_stacking", "pdns"]


def _da_system(s, rng, kind, periodic, n):
    """two rows of beads facing each other: donors on one row, acceptors on the other.  The angular stiffnesses K
    are softer than 3SPN.2's so that random geometries populate all three branches of the modulating function."""
    half = n // 2
    don = np.array([[i, i + 1, i + 2] for i in range(0, half - 2)])
    acc = np.array([[half + i, half + i + 1, half + i + 2] for i in range(0, half - 2)])
    excl = np.array([[0, 0], [1, 2]])
    nd, na = len(don), len(acc)
    if kind == "3spn2_base_pair":
        t = DATopology(donors=np.c_[don[:, :2], -np.ones(nd, int)], acceptors=np.c_[acc[:, :2], -np.ones(na, int)], cutoff=0.9, excl=excl,
                       periodic=periodic)
        s.add_3spn2_base_pair(0, t, rng.uniform(10, 20, nd), rng.uniform(0.45, 0.6, nd), rng.uniform(1.5, 2.6, nd),
                              rng.uniform(1.5, 2.6, nd), rng.uniform(-1, 1, nd), 6.0, 2.0)
    elif kind == "3spn2_cross_stacking":
        t = DATopology(donors=don, acceptors=acc, cutoff=0.9, excl=excl, periodic=periodic)
        s.add_3spn2_cross_stacking(0, t, rng.uniform(2, 5, na), rng.uniform(0.45, 0.65, na), rng.uniform(1.5, 2.6, na),
                                   rng.uniform(1.5, 2.6, na), 2.0, 1.5, 8.0)
    else:
        t = DATopology(donors=don, acceptors=np.c_[acc[:, :2], -np.ones(na, int)], cutoff=0.9, periodic=periodic)
        s.add_pdns(0, t, rng.uniform(1, 3, nd), rng.uniform(0.45, 0.6, nd), rng.uniform(1.2, 2.2, nd), rng.uniform(1.2, 2.2, nd),
                   0.1, 0.4)
    a = _chain(half, rng, 0.4)
    b = a + np.array([0.0, 0.0, 0.55]) + rng.uniform(-0.08, 0.08, a.shape)
    return np.concatenate([a, b])


@pytest.mark.parametrize("kind", DA_KINDS)
@pytest.mark.parametrize("periodic", [False, True])
def test_donor_acceptor_finite_difference(oracle, kind, periodic):
    """The oracle differentiates the reference's expression strings by forward-mode AD; central differences of
    the energy check that machinery (pairs that cross the hard cutoff during the stencil are avoided by h)."""
    n = 16
    box = np.array([3.1, 2.9, 3.3]) if periodic else None
    _fd_check(oracle, lambda s, rng: _da_system(s, rng, kind, periodic, n), n, box, seed=40 + DA_KINDS.index(kind), h=1e-6, tol=2e-5)


def test_dna_3spn2_known_answers(oracle):
    """dna_3SPN2.toml is the ideal B-DNA of its own parameter set: every Watson-Crick pair
    sits at dr, dtheta1, dtheta2, dphi ~ 1e-4, so E_BP = -eps per pair (A-T 16.73, G-C 21.18 kJ/mol; this also
    pins the first-wins GC angles of quirk C1 and the (d2,d1,a1,a2) ordering), every stacking step at its
    minimum (E_BS = -sum eps), bonds and angles at theirs."""
    spec = load_golden_spec("dna_3spn2")
    ctx = ti.make_context(oracle, spec)
    e = ctx.get_state(energy=True)["energy_by_group"]
    g = spec.group_ids
    seq_pairs = {"AT": 0, "GC": 0}
    names = spec.names
    bases = [x for x in names if x in "ATGC"]
    half = len(bases) // 2
    for k in range(half):
        kind = "".join(sorted(bases[k] + bases[len(bases) - 1 - k]))
        seq_pairs["AT" if kind == "AT" else "GC"] += 1
    # non-native donor-acceptor combinations inside the 1.8 nm cutoff contribute < 0.2 % (angle-modulated to ~0)
    assert e[g["3SPN2BasePair (A-T)"]] == pytest.approx(-16.73 * seq_pairs["AT"], rel=2e-3)
    assert e[g["3SPN2BasePair (G-C)"]] == pytest.approx(-21.18 * seq_pairs["GC"], rel=2e-3)
    bs = [t for t in spec.terms if t.kind == "3spn2_base_stacking"][0]
    assert e[g["3SPN2BaseStacking"]] == pytest.approx(-np.sum(bs.args["eps"]), rel=1e-3)
    assert abs(e[g["3SPN2Bond"]]) < 0.05 and abs(e[g["HarmonicAngle"]]) < 0.05
    assert e[g["3SPN2ExcludedVolume"]] == 0.0


@pytest.mark.parametrize("name", ["dna_3spn2", "nucleosome"])
def test_dna_samples_match_golden(oracle, name):
    spec = load_golden_spec(name)
    gold = load_golden_values(name)
    ctx = ti.make_context(oracle, spec)
    st = ctx.get_state(forces=True, energy=True)
    for k, gid in spec.group_ids.items():
        assert st["energy_by_group"][gid] == pytest.approx(gold["energy_by_group_kj"][k], rel=1e-9, abs=1e-9)
    assert np.allclose(st["forces"], np.array(gold["forces"]), atol=1e-6)
    for t in range(ctx.num_da_terms()):
        assert ctx.da_pairs(t).tolist() == gold["da_pairs"][str(t)]


PAIRS = ["excluded_volume", "debye_huckel", "lj_repulsive", "lj_attractive", "3spn2_excluded_volume", "wca"]


@pytest.mark.parametrize("kind", PAIRS)
@pytest.mark.parametrize("periodic", [False, True])
def test_pair_finite_difference(oracle, kind, periodic):
    n = 24
    box = np.array([2.6, 2.4, 2.8]) if periodic else None

    def build(s, rng):
        topo = PairTopology(excl=np.array([[0, 1], [2, 5]]), periodic=periodic)
        if kind == "excluded_volume":
            s.add_excluded_volume(0, 2.5, 1.0, rng.uniform(0.15, 0.25, n), topo)
        elif kind == "debye_huckel":
            s.add_debye_huckel(0, 1.83, 0.3, 1.1, np.exp(-1.1 / 0.3) / 1.1, rng.choice([-1.0, 1.0, 0.5], n), topo)
        elif kind == "lj_repulsive":
            s.add_lj_repulsive(0, 1.2, rng.uniform(0.5, 1.0, n), rng.uniform(0.4, 0.6, n), topo)
        elif kind == "lj_attractive":
            s.add_lj_attractive(0, 1.2, 1.08, rng.uniform(0.5, 1.0, n), rng.uniform(0.4, 0.6, n), topo)
        elif kind == "wca":
            s.add_wca(0, 0.75 * 2.0 ** (1.0 / 6.0), rng.uniform(0.5, 1.0, n), rng.uniform(0.55, 0.75, n), topo)
        else:
            s.add_3spn2_excluded_volume(0, 1.2, rng.uniform(0.8, 1.2, n), rng.uniform(0.4, 0.7, n), topo)
        # jittered lattice: no overlaps, many pairs in every zone
        g = np.stack(np.meshgrid(np.arange(2), np.arange(3), np.arange(4), indexing="ij"), -1).reshape(-1, 3)
        return g * 0.62 + rng.uniform(-0.06, 0.06, (n, 3)) + 0.3

    _fd_check(oracle, build, n, box, seed=10 + PAIRS.index(kind), tol=5e-6)


def test_cell_list_matches_brute_force(oracle):
    spec = synth.hps_slab(n_chains=14)  # > 1500 beads: cell-list path
    ctx = ti.make_context(oracle, spec)
    a = ctx.get_state(forces=True, energy=True)
    ctx.set_option("brute_force", 1)
    b = ctx.get_state(forces=True, energy=True)
    assert np.abs(a["forces"] - b["forces"]).max() < 1e-8
    assert np.allclose(a["energy_by_group"], b["energy_by_group"], rtol=1e-11, atol=1e-9)
    ctx.set_option("brute_force", 0)
    p1 = ctx.neighbor_pairs(1)
    ctx.set_option("brute_force", 1)
    assert np.array_equal(p1, ctx.neighbor_pairs(1))
    # unwrapped master positions (beads that left the box by whole periods): the cell path wraps and shifts per
    # cell pair, the brute-force path takes nearbyint minimum images; same forces as before the shift
    rng = np.random.default_rng(4)
    x = ctx.get_state(positions=True)["positions"]
    ctx.set_positions(x + rng.integers(-2, 3, x.shape) * np.asarray(spec.box))
    ctx.set_option("brute_force", 0)
    c = ctx.get_state(forces=True, energy=True)
    ctx.set_option("brute_force", 1)
    d = ctx.get_state(forces=True, energy=True)
    assert np.abs(c["forces"] - d["forces"]).max() < 1e-8 and np.abs(c["forces"] - a["forces"]).max() < 1e-7
    assert np.allclose(c["energy_by_group"], d["energy_by_group"], rtol=1e-11, atol=1e-9)


def test_interaction_groups_and_errors(oracle):
    n = 8
    s = System(oracle, np.full(n, 10.0))
    q = np.array([1.0, np.nan, -1.0, np.nan, 1.0, np.nan, np.nan, -1.0])
    charged = [0, 2, 4, 7]
    s.add_debye_huckel(0, 1.8, 0.9, 5.0, 0.0, q, PairTopology(igroups=[(charged, charged)]))
    ctx = Context(s, 300, 0.2, 0.001)
    ctx.set_positions(np.arange(3 * n).reshape(n, 3) * 0.1)
    pairs = ctx.neighbor_pairs(0)
    assert sorted(map(tuple, pairs.tolist())) == [(a, b) for i, a in enumerate(charged) for b in charged[i + 1:]]
    # NaN parameter without an interaction group is an error
    s2 = System(oracle, np.full(n, 10.0))
    with pytest.raises(Exception):
        s2.add_debye_huckel(0, 1.8, 0.9, 5.0, 0.0, q, PairTopology())


def test_langevin_equipartition(oracle):
    """Free particles under the Langevin step reach kT per degree of freedom (B.1 coefficients)."""
    n = 2000
    s = System(oracle, np.full(n, 80.0))
    s.add_harmonic_bond(0, np.zeros((0, 2), np.int32), [], [])
    ctx = Context(s, 300.0, 5.0, 0.05, seed=7)
    ctx.set_positions(np.zeros((n, 3)))
    ctx.step(60)
    v = ctx.get_state(velocities=True)["velocities"]
    kT = 8.31446261815324e-3 * 300
    assert (80.0 * (v ** 2).mean()) == pytest.approx(kT, rel=0.05)


def test_replica_batch_equals_standalone_runs(oracle):
    """Config C4 semantics: replica r of a batch follows exactly the trajectory of a stand-alone context with seed
    base + r (own noise key, local bead counter, no cross-replica pairs although all replicas overlap in space)."""
    spec = load_golden_spec("sh3")
    R = 3
    batch = ti.make_context(oracle, spec, replicas=R)
    singles = [ti.make_context(oracle, spec, seed=spec.seed + r) for r in range(R)]
    n = spec.mass.size
    assert batch.n_replicas == R and batch.n == R * n
    batch.step(15)
    for c in singles:
        c.step(15)
    a = batch.get_state(positions=True, velocities=True, forces=True, energy=True)
    re = batch.replica_energies()
    for r, c in enumerate(singles):
        b = c.get_state(positions=True, velocities=True, forces=True, energy=True)
        sl = slice(r * n, (r + 1) * n)
        # (same arithmetic, but the batch's Verlet rows are shared out to the OpenMP threads differently: the per-thread partial
        # sums may round differently in the last bit)
        assert np.allclose(a["positions"][sl], b["positions"], rtol=0, atol=1e-12)
        assert np.allclose(a["velocities"][sl], b["velocities"], rtol=0, atol=1e-11)
        assert np.allclose(a["forces"][sl], b["forces"], rtol=0, atol=1e-9)
        assert re["potential_energy"][r] == pytest.approx(b["potential_energy"], rel=1e-12)
        assert re["kinetic_energy"][r] == pytest.approx(b["kinetic_energy"], rel=1e-12)
        assert np.allclose(re["energy_by_group"][r], b["energy_by_group"], rtol=1e-12, atol=1e-12)
    assert a["potential_energy"] == pytest.approx(re["potential_energy"].sum(), rel=1e-12)
    # replicas diverge from each other (different noise)
    assert np.abs(a["positions"][:n] - a["positions"][n:2 * n]).max() > 1e-4
    # pair set of the batch = union of the replicas' sets, shifted
    for t in range(batch.num_pair_terms()):
        want = np.concatenate([c.neighbor_pairs(t) + r * n for r, c in enumerate(singles)])
        assert np.array_equal(batch.neighbor_pairs(t), want)


def test_replicate_tiles_a_periodic_system_exactly(oracle):
    """synth.replicate with a tiled box: every copy feels the forces of the original (periodic images), 4x the energy"""
    from opencafemol_b200 import synth, toml_input as ti

    base = synth.hps_slab(n_chains=3)
    lx, ly, lz = base.box
    tiled = synth.replicate(base, [[i * lx, j * ly, 0.0] for i in range(2) for j in range(2)], box=[2 * lx, 2 * ly, lz])
    a = ti.make_context(oracle, base).get_state(forces=True, energy=True)
    b = ti.make_context(oracle, tiled).get_state(forces=True, energy=True)
    n = base.mass.size
    for c in range(4):
        assert np.abs(b["forces"][c * n:(c + 1) * n] - a["forces"]).max() < 1e-8 * max(1.0, np.abs(a["forces"]).max())
    assert b["potential_energy"] == pytest.approx(4.0 * a["potential_energy"], rel=1e-12)


def test_committed_c2_state_is_a_valid_compacted_condensate(oracle):
    """opencafemol_b200/data/c2_state.npz (tests/golden/make_c2_state.py), the start state of both bench arms: same system as
    hps_slab(100), chains intact, finite energy, and the pair count its metadata states (bit-exact set size)."""
    import os

    if not os.path.exists(synth.C2_STATE):
        pytest.skip("no committed C2 state")
    spec = synth.c2_equilibrated()
    ref = synth.hps_slab(n_chains=100)
    assert spec.positions.shape == ref.positions.shape == (16300, 3) and spec.velocities.shape == (16300, 3)
    assert np.isfinite(spec.positions).all() and np.isfinite(spec.velocities).all()
    bonds = next(t for t in spec.terms if t.kind == "harmonic_bond").args["ij"]
    d = spec.positions[bonds[:, 0]] - spec.positions[bonds[:, 1]]
    d -= spec.box * np.round(d / spec.box)
    r = np.linalg.norm(d, axis=1)
    assert 0.30 < r.min() and r.max() < 0.46  # k = 92 383 kJ/mol/nm^2 around 0.38 nm at 300 K
    # thermal velocities: kinetic temperature within 3 % of 300 K
    ke = 0.5 * (spec.mass[:, None] * spec.velocities ** 2).sum()
    T = 2.0 * ke / (3 * 16300 * 0.0083144626)
    assert abs(T - 300.0) < 9.0
    ctx = ti.make_context(oracle, spec)
    with np.load(synth.C2_STATE) as meta:
        n_pairs = int(meta["pairs"])
    # float32 storage moves a handful of pairs across the cutoff
    assert abs(len(ctx.neighbor_pairs(1)) - n_pairs) < 200 and n_pairs > 4_000_000
    e = ctx.get_state(energy=True)["energy_by_group"]
    assert np.isfinite(e).all()
    # tiling: exact periodic images, so the forces of the 2 x 1 tiling are the C2 forces twice (SURVEY.md 8d, C5 input)
    f1 = ctx.get_state(forces=True)["forces"]
    two = synth.c2_equilibrated(tile=(2, 1))
    assert two.positions.shape == (32600, 3) and np.allclose(two.box, spec.box * np.array([2, 1, 1]))
    f2 = ti.make_context(oracle, two).get_state(forces=True)["forces"]
    assert np.abs(f2 - np.concatenate([f1, f1])).max() < 1e-7 * max(1.0, np.abs(f1).max())


def test_cpu_arm_verlet_lists_reproduce_the_cell_list_path(oracle):
    """oracle_step on large systems walks Verlet half lists with a skin (the CPU arm of bench.py); same pair set as the per-step
    cell list, so the same trajectory up to the order of double-precision additions, across several list rebuilds"""
    from opencafemol_b200 import synth, toml_input as ti

    spec = synth.hps_slab(n_chains=12)
    out = []
    for verlet in (0, 1):
        ctx = ti.make_context(oracle, spec)
        ctx.set_option("cpu_verlet", verlet)
        ctx.set_option("cpu_skin_nm", 0.05)   # a thin skin: several rebuilds within the run
        ctx.step(40)
        out.append(ctx.get_state(positions=True, velocities=True))
        if verlet:
            assert ctx.get_option("cpu_verlet_builds") >= 3
    assert np.abs(out[0]["positions"] - out[1]["positions"]).max() < 1e-10
    assert np.abs(out[0]["velocities"] - out[1]["velocities"]).max() < 1e-9


def test_external_terms_closed_form_statistics(oracle):
    """Physics known answers for two external generators through the whole Langevin path (sign and scale of the forces):
    a bead in the PositionRestraint well k |x - x0|^2 holds 3/2 kT of potential energy on average (equipartition), and a bead
    under a constant Pulling force f drifts with the terminal velocity f / (m gamma)."""
    kT = 8.31446261815324e-3 * 300.0
    n = 64   # independent beads = independent samples
    s = System(oracle, np.full(n, 40.0), None)
    x0 = np.random.default_rng(1).normal(scale=3.0, size=(n, 3))
    s.add_position_restraint(0, np.arange(n), x0, np.full(n, 25.0), np.zeros(n))
    c = Context(s, 300.0, 2.0, 0.01, seed=11)
    c.set_positions(x0 + 0.05)
    c.set_velocities_to_temperature(300.0, 5)
    c.step(2000)
    e = []
    for _ in range(150):
        c.step(40)
        e.append(c.get_state(energy=True)["potential_energy"] / n)
    e = np.array(e)
    assert e.mean() == pytest.approx(1.5 * kT, rel=0.06), (e.mean(), 1.5 * kT)

    m, gamma, f = 40.0, 2.0, np.array([30.0, -12.0, 0.0])
    s = System(oracle, np.full(n, m), None)
    s.add_pulling(0, np.arange(n), np.tile(f, (n, 1)))
    c = Context(s, 300.0, gamma, 0.01, seed=12)
    c.set_positions(np.zeros((n, 3)))
    c.set_velocities_to_temperature(300.0, 6)
    c.step(1000)                     # several 1 / gamma: the drift has reached its terminal value
    a = c.get_state(positions=True)["positions"].mean(0)
    c.step(4000)
    b = c.get_state(positions=True)["positions"].mean(0)
    v = (b - a) / (4000 * 0.01)
    vt = f / (m * gamma)
    # thermal spread of the mean displacement over 40 ps: sqrt(2 D t / n), D = kT / (m gamma)
    tol = 4.0 * np.sqrt(2.0 * kT / (m * gamma) * 40.0 / n) / 40.0
    assert np.all(np.abs(v - vt) < tol + 0.02 * np.abs(vt)), (v, vt, tol)
