"""TOML (Mjolnir-style) input -> typed term tables in OpenMM units.

Python mirror of the reference's reader layer, used by the Python front end, the tests and the
benchmark: ``read_toml_system`` (src/input/ReadTOMLInput.cpp:21-875), the per-potential
``read_toml_*_ff_generator`` functions (src/input/ReadTOMLForceFieldGenerator.cpp), ``Topology``
(src/Topology.cpp), ``extract_interaction_group`` (src/forcefield/InteractionGroup.hpp:19-101) and the
force-group bookkeeping of ``SystemGenerator::add_ff_generator`` (src/SystemGenerator.cpp:4-32).

Where the reference hands OpenMM an expression string plus parameters, this module emits a ``Term`` with
the same parameters (same unit conversions) that ``build_system`` feeds to the C ABI.
"""
from __future__ import annotations

import math
import os
import tomllib
from dataclasses import dataclass, field

import numpy as np

from . import tables
from .capi import DATopology, Context, Library, PairTopology, System

NM_PER_ANGSTROM = 0.1
ANGSTROM_PER_NM = 10.0
KJ_PER_KCAL = 4.184
CAFETIME = math.sqrt(1.0 / KJ_PER_KCAL) * 0.1  # ps, src/util/Constants.hpp:23


# ------------------------------------------------------------------------------------------------
class Topology:
    """Bond graph with typed edges (src/Topology.hpp, src/Topology.cpp).

    The depth-limited DFS with a shared visited list is reproduced verbatim in behaviour
    (SURVEY.md Appendix C4): in cyclic graphs it can miss nodes a BFS would find.
    """

    def __init__(self, n):
        self.nodes = [[] for _ in range(n)]  # (neighbour, edge_type) in insertion order
        self.molecules = None

    def size(self):
        return len(self.nodes)

    def add_edges(self, index_sets, edge_type):
        for ids in index_sets:
            ids = list(ids)
            for a in range(len(ids) - 1):
                for b in range(a + 1, len(ids)):
                    self.nodes[ids[a]].append((ids[b], edge_type))
                    self.nodes[ids[b]].append((ids[a], edge_type))
        self.molecules = None

    def make_molecule(self, edge_type):
        n = len(self.nodes)
        seen = [False] * n
        mols = []
        for start in range(n):
            if seen[start]:
                continue
            stack = [start]
            mol = []
            while stack:
                s = stack.pop()
                mol.append(s)  # duplicates possible exactly as in the reference
                seen[s] = True
                for nb, et in self.nodes[s]:
                    if et == edge_type and not seen[nb]:
                        stack.append(nb)
            mols.append(mol)
        self.molecules = mols

    def ignore_list_within_molecule(self):
        assert self.molecules is not None
        out = set()
        for mol in self.molecules:
            for a in range(len(mol) - 1):
                for b in range(a + 1, len(mol)):
                    out.add((mol[a], mol[b]))
        return out

    def _adjacent_within(self, node, dist, edge_type, acc):
        acc.append(node)
        if dist == 0:
            return
        for nb, et in self.nodes[node]:
            if et == edge_type and nb not in acc:
                self._adjacent_within(nb, dist - 1, edge_type, acc)

    def ignore_list_within_edge(self, dist, edge_type):
        out = set()
        for i in range(len(self.nodes)):
            if not any(et == edge_type for _, et in self.nodes[i]):
                continue
            acc = []
            self._adjacent_within(i, dist, edge_type, acc)
            for j in acc:
                if j != i:
                    out.add((min(i, j), max(i, j)))
        return out


# ------------------------------------------------------------------------------------------------
@dataclass
class Term:
    kind: str  # C-ABI adder suffix, e.g. "harmonic_bond"
    name: str  # reference generator name() -> .ene column / force group
    args: dict
    group: int = -1


@dataclass
class SystemSpec:
    mass: np.ndarray
    box: np.ndarray | None
    positions: np.ndarray  # nm
    velocities: np.ndarray | None  # nm/ps
    names: list
    groups: list
    terms: list = field(default_factory=list)
    group_ids: dict = field(default_factory=dict)  # name -> force group
    temperature: float = 300.0
    friction: float = 0.01 / CAFETIME
    dt: float = 0.0
    seed: int = 0
    vel_seed: int = 0
    total_step: int = 0
    save_step: int = 1
    delta_t_raw: float = 0.0
    energy_minimization: bool = False
    periodic: bool = False
    output: dict = field(default_factory=dict)
    platform: dict = field(default_factory=dict)
    state: str = ""  # provenance of positions/velocities when they are not the input file's (synth.c2_equilibrated)
    ensemble: dict | None = None  # None = NVT; {"type": "NPT" | "NPgammaT", ...} = a Monte Carlo barostat (apply_ensemble)

    def add_term(self, term: Term):
        """SystemGenerator::add_ff_generator group rule (src/SystemGenerator.cpp:4-32): the group id is
        the map size at first insertion of the name; all FlexibleLocalAngle (XXX) share one group,
        likewise FlexibleLocalDihedral; duplicate names share the first id."""
        key = term.name
        if "FlexibleLocalAngle" in key:
            key = "FlexibleLocalAngle"
        elif "FlexibleLocalDihedral" in key:
            key = "FlexibleLocalDihedral"
        if key not in self.group_ids:
            self.group_ids[key] = len(self.group_ids)
        term.group = self.group_ids[key]
        self.terms.append(term)


def _either(tab, k1, k2):
    if k1 in tab and k2 in tab:
        raise ValueError(f"key duplicates: {k1} and {k2}")
    if k1 in tab:
        return tab[k1]
    if k2 in tab:
        return tab[k2]
    raise KeyError(f"both keys, {k1!r} and {k2!r}, are not found")


def _param(p, env, key, default=None):
    """Utility::find_parameter(_or): string values are looked up in `env` (src/util/Utility.hpp:47-158)."""
    if key not in p:
        if default is None:
            raise KeyError(f"value {key} does not exists")
        return default
    v = p[key]
    if isinstance(v, str):
        if env is None or v not in env:
            if default is not None:
                return default
            raise KeyError(f'named variable "{v}" does not exists')
        return env[v]
    return v


def _param_either(p, env, k1, k2):
    if k1 in p and k2 in p:
        raise ValueError("key duplicates")
    if k1 in p:
        return _param(p, env, k1)
    return _param(p, env, k2)


def _indices(p, env, n, who, nsys):
    idx = list(_param(p, env, "indices"))
    off = _param(p, env, "offset", 0)
    if isinstance(off, (list, tuple)):  # add_offset, src/input/Utility.cpp:82-131
        idx = [a + b for a, b in zip(idx, off)]
    else:
        idx = [a + off for a in idx]
    assert len(idx) == n
    for a in idx:
        if a >= nsys:
            raise RuntimeError(f"[error] {who} : index {a} exceeds the system's largest index {nsys - 1}.")
    return idx


def _index(p, env, who, nsys):
    i = _param(p, env, "index") + _param(p, env, "offset", 0)
    if i >= nsys:
        raise RuntimeError(f"[error] {who} : index {i} exceeds the system's largest index {nsys - 1}.")
    return i


def expand_include(tab, base_dir):
    """`include = "file"` / `include = [..]` expansion at any table (src/input/Utility.cpp:41-80)."""
    if isinstance(tab, dict):
        if "include" in tab:
            inc = tab.pop("include")
            files = [inc] if isinstance(inc, str) else list(inc)
            for f in files:
                with open(os.path.join(base_dir, f), "rb") as fh:
                    sub = tomllib.load(fh)
                for k, v in sub.items():  # merge_toml_tables: existing keys win
                    if k not in tab:
                        tab[k] = v
        for v in tab.values():
            expand_include(v, base_dir)
    elif isinstance(tab, list):
        for v in tab:
            expand_include(v, base_dir)
    return tab


# ------------------------------------------------------------------------------------------------
def _read_ignore(ff, topo: Topology):
    """read_ignore_molecule_and_particles_within + read_ignore_group
    (src/input/ReadTOMLForceFieldGenerator.cpp:963-1051)."""
    excl = set()
    gpairs = []
    ign = ff.get("ignore")
    if ign is None:
        return np.zeros((0, 2), np.int32), gpairs
    if "molecule" in ign:
        name = ign["molecule"]
        if name in ("Intra", "Self"):
            excl |= {(min(a, b), max(a, b)) if a != b else (a, b) for a, b in topo.ignore_list_within_molecule()}
        elif name in ("Others", "Inter"):
            raise RuntimeError('[error] ignore molecule do not support "Others" or "Inter".')
    if "particles_within" in ign:
        for etype, dist in ign["particles_within"].items():
            excl |= topo.ignore_list_within_edge(int(dist), etype)
    if "group" in ign:
        g = ign["group"]
        for a, b in g.get("inter", []):
            gpairs.append((a, b))
        for a in g.get("intra", []):
            gpairs.append((a, a))
    arr = np.array(sorted(excl), np.int32).reshape(-1, 2)
    return arr, gpairs


def extract_interaction_group(has_param, ignore_group_pairs, group_vec):
    """src/forcefield/InteractionGroup.hpp:19-101 — returns [(set1, set2), ...] or [] when every
    particle takes part and no group rule applies."""
    n = len(has_param)
    if not ignore_group_pairs:
        if not all(has_param):
            part = [i for i in range(n) if has_param[i]]
            return [(part, part)]
        return []
    related = sorted({g for pr in ignore_group_pairs for g in pr})
    gmap = {g: [] for g in related}
    others = []
    for i in range(n):
        if has_param[i]:
            if group_vec[i] is not None and group_vec[i] in gmap:
                gmap[group_vec[i]].append(i)
            else:
                others.append(i)

    def ignored(a, b):
        return any((p == a and q == b) or (p == b and q == a) for p, q in ignore_group_pairs)

    out = [(others, others)]
    for ia, a in enumerate(related):
        out.append((gmap[a], others))
        for b in related[ia:]:
            if not ignored(a, b):
                out.append((gmap[a], gmap[b]))
    return out


def _two_largest(vals, strict):
    """The reference's max / second-max trackers (ExcludedVolume uses '<', LJ uses '<=';
    SURVEY.md Appendix C12)."""
    mx = sm = np.finfo(np.float64).tiny
    for v in vals:
        if (mx < v) if strict else (mx <= v):
            sm, mx = mx, v
        elif (sm < v) if strict else (sm <= v):
            sm = v
    return mx, sm


# ------------------------------------------------------------------------------------------------
def fla_min_energy(aa):
    """E_min scan of FlexibleLocalAngleForceFieldGenerator::generate (…cpp:11-19) with the
    reference's own tabulated-second-derivative spline (…cpp:47-68)."""
    y = tables.FLA_SPLINE[aa]
    y2 = tables.FLA_SPLINE_D2[aa]
    thetas = tables.FLA_THETAS
    tmin, tmax = tables.FLA_MIN_THETA, tables.FLA_MAX_THETA
    dth = (tmax - tmin) / 9.0
    rd = 1.0 / dth
    emin = y[0]
    th = tmin
    while th < tmax:
        n = min(8, int(math.floor((th - tmin) * rd)))
        if thetas[n + 1] < th:
            n += 1
        elif th < thetas[n]:
            n -= 1
        a = (thetas[n + 1] - th) * rd
        b = (th - thetas[n]) * rd
        e = a * y[n] + b * y[n + 1] + ((a * a * a - a) * y2[n] + (b * b * b - b) * y2[n + 1]) * dth * dth / 6.0
        emin = min(emin, e)
        th += 1e-4
    return emin


def fld_min_energy(c):
    """E_min scan of FlexibleLocalDihedralForceFieldGenerator::generate (…cpp:7-26)."""
    n = int(math.ceil(2 * math.pi / 1e-4)) + 2
    # reproduce the accumulating loop phi += 1e-4 exactly
    phi = np.empty(n)
    p = -math.pi
    k = 0
    while p < math.pi:
        phi[k] = p
        p += 1e-4
        k += 1
    phi = phi[:k]
    s1, c1 = np.sin(phi), np.cos(phi)
    ssq, csq = s1 * s1, c1 * c1
    s2, c2 = 2 * s1 * c1, csq - ssq
    s3, c3 = s1 * (3 - 4 * ssq), c1 * (4 * csq - 3)
    e = c[1] * c1 + c[2] * s1 + c[3] * c2 + c[4] * s2 + c[5] * c3 + c[6] * s3 + c[0]
    return float(e.min())


_fla_cache = {}
_fld_cache = {}


# ------------------------------------------------------------------------------------------------
def read_toml(path_or_dict, base_dir=None) -> SystemSpec:
    if isinstance(path_or_dict, dict):
        data = path_or_dict
        base_dir = base_dir or "."
    else:
        with open(path_or_dict, "rb") as fh:
            data = tomllib.load(fh)
        base_dir = os.path.dirname(os.path.abspath(path_or_dict))
    expand_include(data, base_dir)

    systems = data["systems"]
    particles = systems[0]["particles"]
    n = len(particles)
    mass = np.array([float(_either(p, "m", "mass")) for p in particles])
    groups = [p.get("group") for p in particles]
    names = [p.get("name", "") for p in particles]
    pos = np.array([_either(p, "pos", "position") for p in particles], float) * NM_PER_ANGSTROM
    vel = None
    if "vel" in particles[0] or "velocity" in particles[0]:
        for p in particles:
            if "vel" not in p and "velocity" not in p:
                raise RuntimeError("[error] All particle need to have velocity, because first particle in system has velocity.")
        vel = np.array([_either(p, "vel", "velocity") for p in particles], float) * NM_PER_ANGSTROM / CAFETIME

    sim = data["simulator"]
    btype = sim["boundary_type"]
    periodic = btype in ("Periodic", "PeriodicCuboid")
    box = None
    if periodic:
        shape = systems[0]["boundary_shape"]
        lo, up = np.array(shape["lower"], float), np.array(shape["upper"], float)
        box = (up - lo) * NM_PER_ANGSTROM
    attr = systems[0]["attributes"]
    temperature = float(attr["temperature"])
    ensemble = None
    if "ensemble" in attr and attr["ensemble"]["type"] != "NVT":
        # ReadTOMLInput.cpp:77-175: NPT = MonteCarloAnisotropicBarostat, NPgammaT = MonteCarloMembraneBarostat (the generator does
        # not hand the input's frequency to the membrane barostat: MonteCarloMembraneBarostatGenerator.cpp:8-12, OpenMM's default 25)
        ens = attr["ensemble"]
        if not periodic:
            raise RuntimeError("barostat should be used with periodic boundary condition.")
        if ens["type"] == "NPT":
            ensemble = {"type": "NPT", "pressure": [float(v) for v in ens["pressure"]], "scale_axis": [bool(v) for v in ens["scale_axis"]],
                        "frequency": int(ens.get("frequency", 25))}
        elif ens["type"] in ("NPgammaT", "NPγT"):
            if ens["xymode"] not in ("Isotropic", "Anisotropic"):
                raise RuntimeError("[error] unknown xymode " + ens["xymode"] + " was specified.")
            if ens["zmode"] not in ("Free", "Fixed", "ConstantVolume"):
                raise RuntimeError("[error] unknown zmode " + ens["xymode"] + " was specified.")
            ensemble = {"type": "NPgammaT", "pressure": float(ens["pressure"]), "surface_tension": float(ens["surface_tension"]),
                        "xymode": ens["xymode"], "zmode": ens["zmode"], "frequency": 25}
        else:
            raise RuntimeError("[error] unknown ensemble type " + ens["type"] + " was specified.")

    spec = SystemSpec(mass=mass, box=box, positions=pos, velocities=vel, names=names, groups=groups)
    spec.periodic = periodic
    spec.ensemble = ensemble
    spec.temperature = temperature
    spec.delta_t_raw = float(sim["delta_t"])
    spec.dt = spec.delta_t_raw * CAFETIME
    spec.friction = float(sim.get("gamma_t", 0.01)) / CAFETIME
    spec.seed = int(sim.get("seed", 0))
    spec.vel_seed = int(systems[0].get("vel_seed", 0)) if isinstance(systems[0], dict) else 0
    spec.total_step = int(sim["total_step"])
    spec.save_step = int(sim["save_step"])
    spec.energy_minimization = bool(sim.get("energy_minimization", False))
    spec.output = dict(data.get("files", {}).get("output", {}))
    spec.platform = dict(data.get("platform", {}))

    topo = Topology(n)
    ff = data["forcefields"][0]
    for lf in ff.get("local", []):
        _read_local(spec, lf, topo, periodic)
    topo.make_molecule("bond")
    for gf in ff.get("global", []):
        _read_global(spec, gf, topo, periodic, attr)
    for xf in ff.get("external", []):
        _read_external(spec, xf, topo, periodic)
    spec.topology = topo
    return spec


def _read_external(spec, xf, topo, periodic):
    """[[forcefields.external]] (src/input/ReadTOMLInput.cpp:763-870 and the readers of
    ReadTOMLForceFieldGenerator.cpp:2496-2790); unit handling as the readers do it, quirks included."""
    inter = xf["interaction"]
    env = xf.get("env")
    n = topo.size()
    if inter == "CoMPullingForce":
        if xf["potential"] != "Harmonic":
            raise RuntimeError(f"[error] invalid potential {xf['potential']} found.")
        for p in xf["parameters"]:
            # read_toml_harmonic_com_pulling_ff_generator (:2622-2659): k and v0 are handed over unconverted
            g1, g2 = p["indices_pair"]
            spec.add_term(Term("harmonic_com_pulling", "HamonicCoMPulling",
                               dict(k=float(_param(p, env, "k")), v0=float(_param(p, env, "v0")), group1=list(g1), group2=list(g2), periodic=periodic)))
    elif inter == "PullingForce":
        # read_toml_pulling_ff_generator (:2506-2548): kcal/(mol A) -> kJ/(mol nm)
        idx = [_index(p, env, "read_toml_pulling_ff_generator", n) for p in xf["parameters"]]
        f = [[c * KJ_PER_KCAL * 10.0 for c in _param(p, env, "force")] for p in xf["parameters"]]
        spec.add_term(Term("pulling", "Pulling", dict(index=idx, force=f)))
    elif inter == "PositionRestraint":
        if xf["potential"] != "Harmonic":
            raise RuntimeError(f"[error] invalid potential {xf['potential']} found.")
        # read_toml_position_restraint_ff_generator (:2559-2611)
        prm = xf["parameters"]
        idx = [_index(p, env, "read_toml_position_restraint_ff_generator", n) for p in prm]
        pos = [[c * NM_PER_ANGSTROM for c in _param(p, env, "position")] for p in prm]
        k = [_param(p, env, "k") * KJ_PER_KCAL * 100.0 for p in prm]
        v0 = [_param(p, env, "v0") * NM_PER_ANGSTROM for p in prm]
        spec.add_term(Term("position_restraint", "PositionRestraint", dict(index=idx, position=pos, k=k, v0=v0)))
    elif inter == "CylindricalRestraint":
        if xf["potential"] != "Harmonic":
            raise RuntimeError(f"[error] invalid potential {xf['potential']} found.")
        # read_toml_cylindrical_restraint_ff_generator (:2720-2775) and the generator's constructor (…Generator.hpp:20-51):
        # the axis is normalised, the shift made orthogonal to it
        prm = xf["parameters"]
        idx = [_index(p, env, "read_toml_position_restraint_ff_generator", n) for p in prm]
        axis = np.array([[c * NM_PER_ANGSTROM for c in _param(p, env, "axis")] for p in prm], float).reshape(-1, 3)
        shift = np.array([[c * NM_PER_ANGSTROM for c in _param(p, env, "shift")] for p in prm], float).reshape(-1, 3)
        axis = axis / np.sqrt((axis * axis).sum(1))[:, None]
        shift = shift - (axis * shift).sum(1)[:, None] * axis
        k = [_param(p, env, "k") * KJ_PER_KCAL * 100.0 for p in prm]
        v0 = [_param(p, env, "v0") * NM_PER_ANGSTROM for p in prm]
        spec.add_term(Term("cylindrical_restraint", "CylinderRestraint", dict(index=idx, axis=axis, shift=shift, k=k, v0=v0, periodic=periodic)))
    elif inter == "RectangularBox":
        if xf["potential"] != "ExcludedVolumeWall":
            raise RuntimeError(f"[error] invalid potential {xf['potential']} found.")
        if periodic:
            raise RuntimeError("[error] read_toml_rectangular_box_ff_generator : RectangularBox interaction does not work under "
                               "periodic boundary condition. Use unlimited boundary.")
        # read_toml_exv_rectangular_box_ff_generator (:2665-2710): epsilon is handed over unconverted
        prm = xf["parameters"]
        lo = [c * NM_PER_ANGSTROM for c in _param(xf["box"], env, "lower")]
        hi = [c * NM_PER_ANGSTROM for c in _param(xf["box"], env, "upper")]
        idx = [_param(p, env, "index") + _param(p, env, "offset", 0) for p in prm]
        rad = [_param(p, env, "radius") * NM_PER_ANGSTROM for p in prm]
        spec.add_term(Term("exv_rectangular_box", "EXVRectangularBox", dict(eps=float(_param(xf, env, "epsilon")), lower=lo, upper=hi, index=idx, radius=rad)))
    # (any other interaction name is skipped without a message, as ReadTOMLInput.cpp:767-868 does)


def _read_local(spec, lf, topo, periodic):
    inter, pot = lf["interaction"], lf["potential"]
    env = lf.get("env")
    params = lf["parameters"]
    n = topo.size()
    topo_name = lf.get("topology")

    def add_topology(idx_list):
        if "topology" in lf:
            topo.add_edges(idx_list, topo_name)

    if inter == "BondLength" and pot == "Harmonic":
        idx = [_indices(p, env, 2, "read_toml_harmonic_bond_ff_generator", n) for p in params]
        r0 = [_param(p, env, "v0") * NM_PER_ANGSTROM for p in params]
        k = [_param(p, env, "k") * KJ_PER_KCAL * ANGSTROM_PER_NM * ANGSTROM_PER_NM * 2.0 for p in params]
        add_topology(idx)
        spec.add_term(Term("harmonic_bond", "HarmonicBond", dict(ij=idx, r0=r0, k=k, periodic=periodic)))
    elif inter == "BondLength" and pot == "Gaussian":
        idx = [_indices(p, env, 2, "read_toml_gaussian_bond_ff_generator", n) for p in params]
        k = [_param(p, env, "k") * KJ_PER_KCAL for p in params]
        r0 = [_param(p, env, "v0") * NM_PER_ANGSTROM for p in params]
        sg = [_param_either(p, env, "sigma", "σ") * NM_PER_ANGSTROM for p in params]
        add_topology(idx)
        spec.add_term(Term("gaussian_bond", "GaussianBond", dict(ij=idx, k=k, r0=r0, sigma=sg, periodic=periodic)))
    elif inter == "BondLength" and pot == "GoContact":
        idx = [_indices(p, env, 2, "read_toml_go_contact_ff_generator", n) for p in params]
        k = [_param(p, env, "k") * KJ_PER_KCAL for p in params]
        r0 = [_param(p, env, "v0") * NM_PER_ANGSTROM for p in params]
        add_topology(idx)
        spec.add_term(Term("go_contact", "GoContact", dict(ij=idx, k=k, r0=r0, periodic=periodic)))
    elif inter == "BondLength" and pot == "CappedGoContact":
        # read_toml_capped_go_contact_ff_generator (ReadTOMLForceFieldGenerator.cpp:226-275) and the generator's tangent
        # (CappedGoContactForceFieldGenerator.cpp:19-20)
        cap = float(lf.get("capping_ratio", 0.75))
        idx = [_indices(p, env, 2, "read_toml_capped_go_contact_ff_generator", n) for p in params]
        k = [_param(p, env, "k") * KJ_PER_KCAL for p in params]
        r0 = [_param(p, env, "v0") * NM_PER_ANGSTROM for p in params]
        slope = -60.0 * ((1.0 / cap) ** 13.0 - (1.0 / cap) ** 11.0)
        intercept = 65.0 * (1.0 / cap) ** 12.0 - 66.0 * (1.0 / cap) ** 10.0
        add_topology(idx)
        m = len(idx)
        spec.add_term(Term("capped_go_contact", "CappedGoContact",
                           dict(ij=idx, k=k, r0=r0, slope=[slope] * m, intercept=[intercept] * m, cap=[cap] * m, periodic=periodic)))
    elif inter == "BondLength" and pot == "3SPN2Bond":
        idx = [_indices(p, env, 2, "read_toml_3spn2_bond_ff_generator", n) for p in params]
        r0 = [_param(p, env, "v0") * NM_PER_ANGSTROM for p in params]
        k2 = [_param(p, env, "k") * KJ_PER_KCAL * 100.0 for p in params]
        # the reference's expression multiplies its k4 (= k * 4.184e4) by a further 100
        k4 = [100.0 * _param(p, env, "k") * KJ_PER_KCAL * 1.0e4 for p in params]
        add_topology(idx)
        spec.add_term(Term("3spn2_bond", "3SPN2Bond", dict(ij=idx, k2=k2, k4=k4, r0=r0, periodic=periodic)))
    elif inter == "BondAngle" and pot == "Harmonic":
        idx = [_indices(p, env, 3, "read_toml_harmonic_angle_ff_generator", n) for p in params]
        th0 = []
        for p in params:
            v0 = _param(p, env, "v0")
            if v0 < 0 or 2 * math.pi < v0:
                raise RuntimeError("[error] read_toml_harmonic_angle_ff_generator: v0 must be between 0 and 2pi")
            if math.pi < v0:
                v0 = 2.0 * math.pi - v0
            th0.append(v0)
        k = [_param(p, env, "k") * KJ_PER_KCAL * 2.0 for p in params]
        add_topology(idx)
        spec.add_term(Term("harmonic_angle", "HarmonicAngle", dict(ijk=idx, theta0=th0, k=k, periodic=periodic)))
    elif inter == "BondAngle" and pot == "FlexibleLocalAngle":
        # one Force per residue type, alphabetical std::map order (ReadTOMLInput.cpp:237-246)
        for aa in sorted(tables.FLA_SPLINE):
            sel = [p for p in params if p["y"][3:6] == aa]
            idx = [_indices(p, env, 3, "read_toml_flexible_local_angle_ff_generator", n) for p in sel]
            k = [_param(p, env, "k") * KJ_PER_KCAL for p in sel]
            add_topology(idx)
            if aa not in _fla_cache:
                _fla_cache[aa] = fla_min_energy(aa)
            spec.add_term(
                Term(
                    "flexible_local_angle", f"FlexibleLocalAngle ({aa})",
                    dict(ijk=idx, k=k, y10=tables.FLA_SPLINE[aa], theta_min=tables.FLA_MIN_THETA,
                         theta_max=tables.FLA_MAX_THETA, e_min=_fla_cache[aa], periodic=periodic),
                )
            )
    elif inter == "DihedralAngle" and pot == "Gaussian":
        _gaussian_dihedral(spec, params, env, topo, lf, periodic)
    elif inter == "DihedralAngle" and pot == "Cosine":
        _cosine_dihedral(spec, params, env, topo, lf, periodic)
    elif inter == "DihedralAngle" and pot in ("Gaussian+Cosine", "Cosine+Gaussian"):
        a1, a2 = [], []
        for e in params:
            v1, v2 = dict(e["Gaussian"]), dict(e["Cosine"])
            v1["indices"] = e["indices"]
            v2["indices"] = e["indices"]
            a1.append(v1)
            a2.append(v2)
        # the synthesised tables carry only "parameters": no topology edges, no env (ReadTOMLInput.cpp:262-290)
        _gaussian_dihedral(spec, a1, None, topo, {}, periodic)
        _cosine_dihedral(spec, a2, None, topo, {}, periodic)
    elif inter == "DihedralAngle" and pot == "FlexibleLocalDihedral":
        for key in sorted(tables.FLD_FOURIER):  # std::map<pair<string,string>> order
            first, second = key
            sel = []
            for p in params:
                coef = p["coef"]
                if second == "GLY":
                    ok = coef[4:7] == "GLY"
                elif second == "PRO":
                    # NB: the R2-PRO branch also captures "GLY-PRO" entries (reference behaviour,
                    # ReadTOMLForceFieldGenerator.cpp:696-717), which are then counted by both Forces
                    ok = (coef == "GLY-PRO") if first == "GLY" else (coef[4:7] == "PRO")
                else:
                    ok = coef[4:7] not in ("GLY", "PRO") and coef[0:3] == first
                if ok:
                    sel.append(p)
            idx = [_indices(p, env, 4, "read_toml_flexible_local_dihedral_ff_generator", n) for p in sel]
            k = [_param(p, env, "k") * KJ_PER_KCAL for p in sel]
            add_topology(idx)
            if key not in _fld_cache:
                _fld_cache[key] = fld_min_energy(tables.FLD_FOURIER[key])
            spec.add_term(
                Term(
                    "flexible_local_dihedral", f"FlexibleLocalDihedral ({first}-{second})",
                    dict(ijkl=idx, k=k, coef7=tables.FLD_FOURIER[key], e_min=_fld_cache[key], periodic=periodic),
                )
            )
    elif inter == "3SPN2BaseStacking" and pot in ("3SPN2", "3SPN2C"):
        _base_stacking(spec, lf, params, env, topo, periodic, pot)
    elif inter == "3SPN2BasePairLocal" and pot in ("3SPN2", "3SPN2C"):
        _base_pair_local(spec, lf, params, env, topo, periodic, pot)
    elif inter == "3SPN2CrossStackingLocal" and pot in ("3SPN2", "3SPN2C"):
        _cross_stacking_local(spec, lf, params, env, topo, periodic, pot)
    else:
        raise RuntimeError(f"[error] invalid local force field: interaction {inter} with potential {pot} is none of the reference's (ReadTOMLInput.cpp:125-420)")


_T3 = None


def tables_3spn2():
    """3SPN.2 / 3SPN.2C parameter tables (opencafemol_b200/tables_3spn2.json, generated from the reference
    headers by tools/extract_3spn2_tables.py; units nm, rad, kJ/mol)."""
    global _T3
    if _T3 is None:
        import json
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "tables_3spn2.json")) as fh:
            _T3 = json.load(fh)
    return _T3


def _nucleotides(params, env):
    """{strand, P?, S, B, Base} records of the 3SPN.2 tables (ReadTOMLForceFieldGenerator.cpp:801-858)."""
    out = []
    for p in params:
        base = p["Base"]
        if base not in "ATGC" or len(base) != 1:
            raise RuntimeError(f'[error] invalid base type {base} here. One of the "A", "T", "C", "G" is expected.')
        out.append(dict(P=_param(p, env, "P") if "P" in p else None, S=_param(p, env, "S"), B=_param(p, env, "B"),
                        strand=_param(p, env, "strand"), base=base))
    return out


def _base_stacking(spec, lf, params, env, topo, periodic, pot):
    """read_toml_3spn2_base_stacking_ff_generator (ReadTOMLForceFieldGenerator.cpp:769-958)."""
    tab = tables_3spn2()["ThreeSPN2BaseStackingPotentialParameter" if pot == "3SPN2" else "ThreeSPN2CBaseStackingPotentialParameter"]
    nucs = _nucleotides(params, env)
    idx, eps, r0, th0 = [], [], [], []
    for b5, b3 in zip(nucs[:-1], nucs[1:]):
        if b5["strand"] != b3["strand"]:
            continue
        kind = b5["base"] + b3["base"]
        idx.append([b5["S"], b5["B"], b3["B"]])
        eps.append(tab["maps"]["epsilon_BS"][kind])
        r0.append(tab["maps"]["r0_BS"][kind] * NM_PER_ANGSTROM)
        th0.append(tab["maps"]["theta0_BS"][kind] * math.pi / 180.0)
    if "topology" in lf:
        for b5, b3 in zip(nucs[:-1], nucs[1:]):
            if b5["strand"] != b3["strand"]:
                continue
            edges = [(b5["S"], b3["S"]), (b5["S"], b3["B"]), (b5["B"], b3["S"]), (b5["B"], b3["B"])]
            if b5["P"] is not None:
                edges += [(b5["P"], b3["S"]), (b5["P"], b3["B"])]
            if b3["P"] is not None:
                edges += [(b5["S"], b3["P"]), (b5["B"], b3["P"])]
            if b5["P"] is not None and b3["P"] is not None:
                edges.append((b5["P"], b3["P"]))
            topo.add_edges(edges, lf["topology"])
    spec.add_term(Term("3spn2_base_stacking", "3SPN2BaseStacking",
                       dict(ijk=idx, eps=eps, r0=r0, theta0=th0, alpha=tab["scalars"]["alpha_BS"] / NM_PER_ANGSTROM,
                            K_BS=tab["scalars"]["K_BS"], periodic=periodic)))


def _da_exclusions(donors, acceptors, ignore):
    """donor x acceptor exclusions: base-particle pairs found in the ignore list
    (ThreeSPN2BasePairForceFieldGenerator.cpp:76-96, ThreeSPN2CrossStackingForceFieldGenerator.cpp:84-100)."""
    ign = {(int(a), int(b)) for a, b in ignore}
    out = []
    for di, d in enumerate(donors):
        for ai, a in enumerate(acceptors):
            if (min(d[0], a[0]), max(d[0], a[0])) in ign:
                out.append((di, ai))
    return np.array(out, np.int32).reshape(-1, 2)


def _pad3(rows):
    out = np.full((len(rows), 3), -1, np.int32)
    for k, r in enumerate(rows):
        out[k, :len(r)] = r
    return out


def _base_pair(spec, gf, topo, periodic, pot):
    """read_toml_3spn2_base_pair_ff_generator x {A-T, G-C} (ReadTOMLForceFieldGenerator.cpp:1645-1771,
    ReadTOMLInput.cpp:716-760)."""
    tab = tables_3spn2()["ThreeSPN2BasePairPotentialParameter" if pot == "3SPN2" else "ThreeSPN2CBasePairPotentialParameter"]
    env = gf.get("env")
    nucs = _nucleotides(gf["parameters"], env)
    ignore, _ = _read_ignore(gf, topo)
    for donor, acceptor in (("A", "T"), ("G", "C")):
        don = [[n["B"], n["S"]] for n in nucs if n["base"] == donor]
        acc = [[n["B"], n["S"]] for n in nucs if n["base"] == acceptor]
        kind = donor + acceptor
        m, sc = tab["maps"], tab["scalars"]
        nd = len(don)
        t = DATopology(donors=_pad3(don), acceptors=_pad3(acc), cutoff=sc["cutoff"], excl=_da_exclusions(don, acc, ignore),
                       periodic=periodic)
        spec.add_term(Term("3spn2_base_pair", f"{pot}BasePair ({donor}-{acceptor})",
                           dict(topo=t, eps=np.full(nd, m["epsilon_BP"][kind]), r0=np.full(nd, m["r0"][kind]),
                                theta0_1=np.full(nd, m["theta0_1"][kind]), theta0_2=np.full(nd, m["theta0_2"][kind]),
                                phi0=np.full(nd, m["phi0"][kind]), alpha=sc["alpha_BP"], K_BP=sc["K_BP"])))


def _cross_stacking(spec, gf, topo, periodic, pot):
    """read_toml_3spn2_cross_stacking_ff_generator x {A-T, T-A, G-C, C-G} x {sense, antisense}
    (ReadTOMLForceFieldGenerator.cpp:1904-2101, ReadTOMLInput.cpp:662-714)."""
    tab = tables_3spn2()["ThreeSPN2CrossStackingPotentialParameter" if pot == "3SPN2" else "ThreeSPN2CCrossStackingPotentialParameter"]
    env = gf.get("env")
    nucs = _nucleotides(gf["parameters"], env)
    ignore, _ = _read_ignore(gf, topo)
    m, sc = tab["maps"], tab["scalars"]
    for first, second in (("A", "T"), ("T", "A"), ("G", "C"), ("C", "G")):
        for strand_kind in ("sense", "antisense"):
            don, acc, acc_kind = [], [], []
            for i, nuc in enumerate(nucs):
                nxt = nucs[i + 1] if i + 1 < len(nucs) and nucs[i + 1]["strand"] == nuc["strand"] else None
                prv = nucs[i - 1] if i > 0 and nucs[i - 1]["strand"] == nuc["strand"] else None
                if strand_kind == "sense":
                    if nuc["base"] == first and nxt is not None:
                        don.append([nuc["B"], nuc["S"], nxt["B"]])
                    if nuc["base"] == second and prv is not None:
                        acc.append([nuc["B"], nuc["S"], prv["B"]])
                        acc_kind.append(first + prv["base"] + "5")
                else:
                    if nuc["base"] == first and nxt is not None:
                        acc.append([nuc["B"], nuc["S"], nxt["B"]])
                        acc_kind.append(second + nxt["base"] + "3")
                    if nuc["base"] == second and prv is not None:
                        don.append([nuc["B"], nuc["S"], prv["B"]])
            t = DATopology(donors=_pad3(don), acceptors=_pad3(acc), cutoff=sc["cutoff"], excl=_da_exclusions(don, acc, ignore),
                           periodic=periodic)
            spec.add_term(Term("3spn2_cross_stacking", f"{pot}CrossStacking ({first}-{second})",
                               dict(topo=t, eps=[m["epsilon_CS"][k] for k in acc_kind], r0=[m["r0_CS"][k] for k in acc_kind],
                                    theta0_3=[m["theta3_0"][first + second]] * len(acc), theta0_CS=[m["theta_CS_0"][k] for k in acc_kind],
                                    K_BP=sc["K_BP"], K_CS=sc["K_CS"], alpha=sc["alpha_CS"])))


# The *Local forms: the same energies over an explicit list of nucleotide pairs (CustomCompoundBondForce in the reference, no
# cutoff).  On the C ABI they are donor x acceptor terms in which entry i contributes donor i and acceptor i and every other
# combination is excluded, evaluated inside LOCAL_PAIR_RANGE (beyond r0 + 1.9 nm the energy is below eps 2 exp(-alpha 1.9 nm)
# < 1e-15 kJ/mol): the mapping the C++ host layer uses (openmm_compat.cpp, kLocalPairRange).
LOCAL_PAIR_RANGE = 2.5


def _local_da_topology(donors, acceptors, periodic):
    m = len(donors)
    excl = np.array([[d, a] for d in range(m) for a in range(m) if a != d], np.int32).reshape(-1, 2)
    return DATopology(donors=np.array(donors, np.int32).reshape(-1, 3), acceptors=np.array(acceptors, np.int32).reshape(-1, 3),
                      cutoff=LOCAL_PAIR_RANGE, excl=excl, periodic=periodic)


def _reserve_group(spec, name):
    """a generator without entries still takes its force-group number (SystemGenerator::add_ff_generator is unconditional)"""
    spec.group_ids.setdefault(name, len(spec.group_ids))


def _base_pair_local(spec, lf, params, env, topo, periodic, pot):
    """read_toml_3spn2_base_pair_local_ff_generator x {A-T, G-C} (ReadTOMLForceFieldGenerator.cpp:1565-1643, ReadTOMLInput.cpp:313-353)
    and ThreeSPN2BasePairLocalForceFieldGenerator.cpp: particles (p1, p2, p3, p4) = (sugar, base, paired base, its sugar)"""
    tab = tables_3spn2()["ThreeSPN2BasePairPotentialParameter" if pot == "3SPN2" else "ThreeSPN2CBasePairPotentialParameter"]
    m, sc = tab["maps"], tab["scalars"]
    for b0, b1 in (("A", "T"), ("G", "C")):
        quads = []
        for p in params:
            pair = p["nucleotide_pair"]
            if len(pair) != 2:
                raise RuntimeError("[error] invalid toml array size.The array size of nucleotide_pair must be 2.")
            s0, s1 = pair[0]["Base"], pair[1]["Base"]
            if s0 == b0 and s1 == b1:
                quads.append([_param(pair[0], env, "S"), _param(pair[0], env, "B"), _param(pair[1], env, "B"), _param(pair[1], env, "S")])
            elif s0 == b1 and s1 == b0:
                quads.append([_param(pair[1], env, "S"), _param(pair[1], env, "B"), _param(pair[0], env, "B"), _param(pair[0], env, "S")])
        if "topology" in lf:
            topo.add_edges(quads, lf["topology"])
        name, kind, nd = f"{pot}BasePairLocal ({b0}-{b1})", b0 + b1, len(quads)
        if not quads:
            _reserve_group(spec, name)
            continue
        t = _local_da_topology([[q[1], q[0], -1] for q in quads], [[q[2], q[3], -1] for q in quads], periodic)
        spec.add_term(Term("3spn2_base_pair", name,
                           dict(topo=t, eps=np.full(nd, m["epsilon_BP"][kind]), r0=np.full(nd, m["r0"][kind]),
                                theta0_1=np.full(nd, m["theta0_1"][kind]), theta0_2=np.full(nd, m["theta0_2"][kind]),
                                phi0=np.full(nd, m["phi0"][kind]), alpha=sc["alpha_BP"], K_BP=sc["K_BP"])))


def _cross_stacking_local(spec, lf, params, env, topo, periodic, pot):
    """read_toml_3spn2_cross_stacking_local_ff_generator x {A-T, T-A, G-C, C-G} x {sense, antisense}
    (ReadTOMLForceFieldGenerator.cpp:1774-1901, ReadTOMLInput.cpp:355-410) and ThreeSPN2CrossStackingLocalForceFieldGenerator.cpp:
    particles (p1..p6) = (sugar, base, paired base, its sugar, neighbouring base, cross base); nucleotide_group = [reference,
    base-pairing, neighbouring, cross-stacking nucleotide]"""
    tab = tables_3spn2()["ThreeSPN2CrossStackingPotentialParameter" if pot == "3SPN2" else "ThreeSPN2CCrossStackingPotentialParameter"]
    m, sc = tab["maps"], tab["scalars"]

    def six(g, order):
        a, b, c, d, e, f = order   # nucleotide numbers giving S, B, B, S, B, B
        return [_param(g[a], env, "S"), _param(g[b], env, "B"), _param(g[c], env, "B"), _param(g[d], env, "S"), _param(g[e], env, "B"),
                _param(g[f], env, "B")]

    for first, second in (("A", "T"), ("T", "A"), ("G", "C"), ("C", "G")):
        for strand_kind in ("sense", "antisense"):
            rows, kinds = [], []
            for p in params:
                g = p["nucleotide_group"]
                if len(g) != 4:
                    raise RuntimeError("[error] invalid toml array size.The array size of nucleotide_group must be 4.")
                b_0, b_p, b_n, b_c = (x["Base"] for x in g)
                if strand_kind == "sense":
                    if b_0 == first and b_p == second:
                        rows.append(six(g, (0, 0, 1, 1, 2, 3))); kinds.append(b_0 + b_c + "5")
                    if b_c == first and b_n == second:
                        rows.append(six(g, (3, 3, 2, 2, 1, 0))); kinds.append(b_c + b_0 + "5")
                else:
                    if b_0 == first and b_p == second:
                        rows.append(six(g, (1, 1, 0, 0, 3, 2))); kinds.append(b_p + b_n + "3")
                    if b_c == first and b_n == second:
                        rows.append(six(g, (2, 2, 3, 3, 0, 1))); kinds.append(b_n + b_p + "3")
            if "topology" in lf:
                topo.add_edges(rows, lf["topology"])
            name = f"{pot}CrossStackingLocal ({first}-{second})"
            if not rows:
                _reserve_group(spec, name)
                continue
            t = _local_da_topology([[r[1], r[0], -1] for r in rows], [[r[2], r[3], r[5]] for r in rows], periodic)
            spec.add_term(Term("3spn2_cross_stacking", name,
                               dict(topo=t, eps=[m["epsilon_CS"][k] for k in kinds], r0=[m["r0_CS"][k] for k in kinds],
                                    theta0_3=[m["theta3_0"][first + second]] * len(rows), theta0_CS=[m["theta_CS_0"][k] for k in kinds],
                                    K_BP=sc["K_BP"], K_CS=sc["K_CS"], alpha=sc["alpha_CS"])))


def _pdns(spec, gf, periodic):
    """read_toml_protein_dna_non_specific_ff_generator (ReadTOMLForceFieldGenerator.cpp:2413-2493) and the
    generator's cutoff rule (ProteinDNANonSpecificForceFieldGenerator.cpp:124-135)."""
    env = gf.get("env")
    n = spec.mass.size
    sigma = _param(gf, env, "sigma", 1.0) * NM_PER_ANGSTROM
    delta = _param(gf, env, "delta", 0.17453)
    ratio = _param(gf, env, "cutoff", 10.0)
    unit = _param(gf, env, "energy_unit", 0.593) * KJ_PER_KCAL
    dna, pro, k, r0, th0, ph0 = [], [], [], [], [], []
    for p in gf["parameters"]:
        index = _index(p, env, "read_toml_protein_dna_non_specific_ff_generator", n)
        if p["kind"] == "DNA":
            dna.append([index, _param(p, env, "S3"), -1])
        if p["kind"] == "Protein":
            # donor slots: d1 = CA of the N-side neighbour, d2 = CA, d3 = CA of the C-side neighbour
            pro.append([_param(p, env, "PN"), index, _param(p, env, "PC")])
            k.append(_param(p, env, "k") * unit)
            r0.append(_param(p, env, "r0") * NM_PER_ANGSTROM)
            th0.append(_param(p, env, "theta0") * math.pi / 180.0)
            ph0.append(_param(p, env, "phi0") * math.pi / 180.0)
    cutoff = max([0.0] + [r + ratio * sigma for r in r0])
    t = DATopology(donors=np.array(pro, np.int32).reshape(-1, 3), acceptors=np.array(dna, np.int32).reshape(-1, 3), cutoff=cutoff,
                   periodic=periodic)
    spec.add_term(Term("pdns", "PDNS", dict(topo=t, k=k, r0=r0, theta0=th0, phi0=ph0, sigma=sigma, delta=delta)))


def _gaussian_dihedral(spec, params, env, topo, lf, periodic):
    n = topo.size()
    idx = [_indices(p, env, 4, "read_toml_gaussian_dihedral_ff_generator", n) for p in params]
    k = [_param(p, env, "k") * KJ_PER_KCAL for p in params]
    phi0 = [_param(p, env, "v0") for p in params]
    sg = [_param_either(p, env, "sigma", "σ") for p in params]
    if "topology" in lf:
        topo.add_edges(idx, lf["topology"])
    spec.add_term(Term("gaussian_dihedral", "GaussianDihedral", dict(ijkl=idx, k=k, phi0=phi0, sigma=sg, periodic=periodic)))


def _cosine_dihedral(spec, params, env, topo, lf, periodic):
    n = topo.size()
    idx = [_indices(p, env, 4, "read_toml_cosine_dihedral_ff_generator", n) for p in params]
    k = [_param(p, env, "k") * KJ_PER_KCAL for p in params]
    t0 = [_param(p, env, "v0") + math.pi for p in params]
    nper = [int(_param(p, env, "n")) for p in params]
    if "topology" in lf:
        topo.add_edges(idx, lf["topology"])
    spec.add_term(Term("cosine_dihedral", "CosineDihedral", dict(ijkl=idx, k=k, t0=t0, nper=nper, periodic=periodic)))


def debye_huckel_constants(temperature, ionic_strength, cutoff_ratio):
    """DebyeHuckelForceFieldGenerator ctor (…cpp:10-50) and calc_dielectric_water (…hpp:36-41)."""
    Na, kB = 6.02214076e23, 1.380649e-23
    eps0_c = 8.854187817e-12 * 1.0e3 * 1.0e-9 / Na
    echarge = 1.6021766208e-19
    T, Cs = temperature, ionic_strength
    epsk = (249.4 - 0.788 * T + 7.2e-4 * T * T) * (1.0 - 2.551e-1 * Cs + 5.151e-2 * Cs * Cs - 6.889e-3 * Cs * Cs * Cs)
    eps0 = eps0_c / echarge / echarge
    K = 1.0 / (4.0 * 3.1415926535897932385 * eps0 * epsk)
    if not Cs > 0:
        raise RuntimeError("[error] ionic strength must be larger than zero")
    ion = Cs * 1.0e-24
    debye = math.sqrt((eps0 * epsk * kB * 1.0e-3 * T) / (2.0 * ion))
    rc = debye * cutoff_ratio
    cc = math.exp(-cutoff_ratio) / rc
    return K, debye, rc, cc


def _read_global(spec, gf, topo, periodic, attr):
    pot = gf["potential"]
    inter = gf.get("interaction")
    if inter == "3SPN2CrossStacking":
        if pot not in ("3SPN2", "3SPN2C"):
            raise RuntimeError(f"[error] invalid potential {pot} found.")
        return _cross_stacking(spec, gf, topo, periodic, pot)
    if inter == "3SPN2BasePair":
        if pot not in ("3SPN2", "3SPN2C"):
            raise RuntimeError(f"[error] invalid potential {pot} found.")
        return _base_pair(spec, gf, topo, periodic, pot)
    if pot == "PDNS":
        return _pdns(spec, gf, periodic)
    if pot == "CombinatorialGoContact":
        return _combinatorial_go(spec, gf, topo, periodic)
    env = gf.get("env")
    n = topo.size()
    params = gf.get("parameters", [])
    groups = spec.groups

    def topo_for(has):
        excl, gpairs = _read_ignore(gf, topo)
        ig = extract_interaction_group(has, gpairs, groups)
        return PairTopology(excl=excl, igroups=ig, periodic=periodic)

    if pot == "ExcludedVolume":
        eps = float(gf["epsilon"]) * KJ_PER_KCAL
        cutoff = float(gf.get("cutoff", 2.0))
        rad = np.full(n, np.nan)
        for p in params:
            rad[_index(p, env, "read_toml_excluded_volume_ff_generator", n)] = _param(p, env, "radius") * NM_PER_ANGSTROM
        has = ~np.isnan(rad)
        mx, sm = _two_largest(rad[has], strict=True)
        rc = (mx + sm) * cutoff
        spec.add_term(Term("excluded_volume", "ExcludedVolume", dict(eps=eps, rc=rc, radius=rad, topo=topo_for(has))))
    elif pot == "3SPN2ExcludedVolume":
        # read_toml_3spn2_excluded_volume_ff_generator (ReadTOMLForceFieldGenerator.cpp:1136-1224) and the
        # generator's cutoff rule (ThreeSPN2ExcludedVolumeForceFieldGenerator.cpp:19-60)
        tab = tables_3spn2()["ThreeSPN2ExcludedVolumePotentialParameter"]
        cutoff = float(gf.get("cutoff", 2.0))
        sig = np.full(n, np.nan)
        kinds = []
        for p in params:
            i = _index(p, env, "read_toml_3spn2_excluded_volume_ff_generator", n)
            sig[i] = tab["maps"]["sigma"][p["kind"]]
            kinds.append((i, p["kind"]))
        has = ~np.isnan(sig)
        t = topo_for(has)
        comp = {("A", "T"), ("T", "A"), ("G", "C"), ("C", "G")}
        extra = [(min(i, j), max(i, j)) for a, (i, ki) in enumerate(kinds) for (j, kj) in kinds[a + 1:] if (ki, kj) in comp]
        allx = {tuple(e) for e in np.asarray(t.excl).reshape(-1, 2).tolist()} | set(extra)
        t.excl = np.array(sorted(allx), np.int32).reshape(-1, 2)
        mx, sm = _two_largest(sig[has], strict=True)
        eps = np.where(has, tab["scalars"]["epsilon"], np.nan)
        spec.add_term(Term("3spn2_excluded_volume", "3SPN2ExcludedVolume", dict(rc=(mx + sm) * cutoff, eps=eps, sigma=sig, topo=t)))
    elif pot == "DebyeHuckel":
        ratio = float(gf.get("cutoff", 5.5))
        q = np.full(n, np.nan)
        for p in params:
            q[_index(p, env, "read_toml_debye_huckel_ff_generator", n)] = _param(p, env, "charge")
        K, debye, rc, cc = debye_huckel_constants(spec.temperature, float(attr["ionic_strength"]), ratio)
        spec.add_term(
            Term("debye_huckel", "DebyeHuckel",
                 dict(K=K, debye=debye, rc=rc, cutoff_correction=cc, charge=q, topo=topo_for(~np.isnan(q))))
        )
    elif pot in ("LennardJonesRepulsive", "LennardJonesAttractive"):
        if "table" in gf:
            if pot != "LennardJonesAttractive":
                raise RuntimeError(f"[error] invalid potential {pot} with a parameter table found.")
            return _uniform_pair(spec, gf, topo, periodic, "lj_attractive")
        ratio = float(_param(gf, env, "cutoff", 2.5))
        eps = np.full(n, np.nan)
        sig = np.full(n, np.nan)
        for p in params:
            i = _param(p, env, "index") + _param(p, env, "offset", 0)
            sig[i] = _param(p, env, "sigma") * NM_PER_ANGSTROM
            eps[i] = _param(p, env, "epsilon") * KJ_PER_KCAL
        has = ~np.isnan(sig)
        mx, sm = _two_largest(sig[has], strict=False)
        rc = (mx + sm) * 0.5 * ratio
        if pot == "LennardJonesRepulsive":
            spec.add_term(Term("lj_repulsive", pot, dict(rc=rc, eps=eps, sigma=sig, topo=topo_for(has))))
        else:
            spec.add_term(Term("lj_attractive", pot, dict(rc=rc, r_switch=0.9 * rc, eps=eps, sigma=sig, topo=topo_for(has))))
    elif pot == "WCA":
        # read_toml_weeks_chandler_andersen_ff_generator (ReadTOMLForceFieldGenerator.cpp:1238-1293) and the generator's
        # cutoff (WeeksChandlerAndersenForceFieldGenerator.cpp:49-96): (max + 2nd max sigma)/2 * 2^(1/6), "<=" updates.
        # The name-pair table form (UniformWeeksChandlerAndersen): _uniform_pair.
        if "table" in gf:
            return _uniform_pair(spec, gf, topo, periodic, "wca")
        eps = np.full(n, np.nan)
        sig = np.full(n, np.nan)
        for p in params:
            i = _param(p, env, "index") + _param(p, env, "offset", 0)
            if i >= n:
                raise RuntimeError(f"[error] read_toml_weeks_chandler_andersen_ff_generator : index {i} exceeds the system's largest index {n - 1}.")
            sig[i] = _param_either(p, env, "sigma", "σ") * NM_PER_ANGSTROM
            eps[i] = _param_either(p, env, "epsilon", "ε") * KJ_PER_KCAL
        has = ~np.isnan(sig)
        mx, sm = _two_largest(sig[has], strict=False)
        rc = (mx + sm) * 0.5 * 2.0 ** (1.0 / 6.0)
        spec.add_term(Term("wca", "WeeksChandlerAndersen", dict(rc=rc, eps=eps, sigma=sig, topo=topo_for(has))))
    elif pot in ("iSoLFAttractive", "Trigonometric"):
        # read_toml_isolf_attractive_ff_generator (ReadTOMLForceFieldGenerator.cpp:1434-1492) /
        # read_toml_trigonometric_ff_generator (:2246-2303); cutoffs of the generators
        # (iSoLFAttractiveForceFieldGenerator.cpp:129-134: mean of the two largest sigmas * 2^(1/6) + mean of the two largest
        # omegas; TrigonometricForceFieldGenerator.cpp:119-123: mean of the two largest sigma + omega), "<=" trackers
        who = "read_toml_isolf_attractive_ff_generator" if pot == "iSoLFAttractive" else "read_toml_trigonometric_ff_generator"
        eps, sig, om = np.full(n, np.nan), np.full(n, np.nan), np.full(n, np.nan)
        for p in params:
            i = _param(p, env, "index") + _param(p, env, "offset", 0)
            if i >= n:
                raise RuntimeError(f"[error] {who} : index {i} exceeds the system's largest index {n - 1}.")
            sig[i] = _param(p, env, "sigma") * NM_PER_ANGSTROM
            eps[i] = _param(p, env, "epsilon") * KJ_PER_KCAL
            om[i] = _param(p, env, "omega") * NM_PER_ANGSTROM
        has = ~np.isnan(sig)
        if pot == "iSoLFAttractive":
            ms, ss = _two_largest(sig[has], strict=False)
            mo, so = _two_largest(om[has], strict=False)
            rc = (ms + ss) * 0.5 * 2.0 ** (1.0 / 6.0) + (mo + so) * 0.5
            spec.add_term(Term("isolf_attractive", pot, dict(rc=rc, eps=eps, sigma=sig, omega=om, topo=topo_for(has))))
        else:
            mr, sr = _two_largest((sig + om)[has], strict=False)
            spec.add_term(Term("trigonometric", pot, dict(rc=(mr + sr) * 0.5, eps=eps, sigma=sig, omega=om, topo=topo_for(has))))
    else:
        raise RuntimeError(f"[error] invalid global force field: potential {pot} is none of the reference's (ReadTOMLInput.cpp:422-760)")


def _uniform_groups(former, latter, ignore_group_pairs, group_vec):
    """interaction groups of the uniform (name-pair table) pair potentials: the constructor of
    UniformWeeksChandlerAndersenForceFieldGenerator.cpp:9-131 (UniformLennardJonesAttractive...cpp has the same code)"""
    if not ignore_group_pairs:
        return [(sorted(set(former)), sorted(set(latter)))]
    related = sorted({g for pr in ignore_group_pairs for g in pr})

    def split(members):
        by_group, others = {g: set() for g in related}, set()
        for i in members:
            g = group_vec[i]
            if g is not None and g in by_group:
                by_group[g].add(i)
            else:
                others.add(i)
        return by_group, others

    fmap, fothers = split(former)
    lmap, lothers = split(latter)
    out = []
    if fothers:
        for g in related:
            if lmap[g]:
                out.append((sorted(fothers), sorted(lmap[g])))
        if lothers:
            out.append((sorted(fothers), sorted(lothers)))
    pairs = {tuple(pr) for pr in ignore_group_pairs}
    for g1 in related:
        if not fmap[g1]:
            continue
        if lothers:
            out.append((sorted(fmap[g1]), sorted(lothers)))
        for g2 in related:
            if (g1, g2) not in pairs and lmap[g2]:
                out.append((sorted(fmap[g1]), sorted(lmap[g2])))
    return out


def _uniform_pair(spec, gf, topo, periodic, kind):
    """WCA / LennardJonesAttractive with a residue-name pair table: one uniform Force per table entry, acting between the beads
    named `first` and the beads named `second` (ReadTOMLInput.cpp:437-499 and :548-610; readers ReadTOMLForceFieldGenerator.cpp:1306-1358
    and :1508-1562).  All entries share one generator name, hence one force group, whatever order the reference's (unordered) table
    is walked in; an entry without beads on either side is skipped."""
    env = gf.get("env")
    n = topo.size()
    excl, gpairs = _read_ignore(gf, topo) if "ignore" in gf else (np.zeros((0, 2), np.int32), [])
    who = ("read_toml_uniform_weeks_chandler_andersen_ff_generator" if kind == "wca"
           else "read_toml_uniform_lennard_jones_attractive_ff_generator")
    named = [(p["name"], _index(p, env, who, n)) for p in gf["parameters"]]
    seen = set()
    for first, sub in gf["table"].items():
        for second, tab in sub.items():
            if (first, second) in seen:
                raise RuntimeError(f"[error] parameter table for {first}-{second} is defined in dupulicate.")
            sigma = float(_either(tab, "sigma", "σ")) * NM_PER_ANGSTROM
            eps = float(_either(tab, "epsilon", "ε")) * KJ_PER_KCAL
            former = [i for nm, i in named if nm == first]
            latter = [i for nm, i in named if nm == second]
            if not former or not latter:
                continue
            seen.add((first, second))
            ig = _uniform_groups(former, latter, gpairs, spec.groups)
            sg, ep = np.full(n, np.nan), np.full(n, np.nan)
            for a, b in ig:
                for i in list(a) + list(b):
                    sg[i], ep[i] = sigma, eps
            tp = PairTopology(excl=excl, igroups=ig, periodic=periodic)
            if kind == "wca":
                spec.add_term(Term("wca", "UniformWeeksChadlerAndersen", dict(rc=sigma * 2.0 ** (1.0 / 6.0), eps=ep, sigma=sg, topo=tp)))
            else:
                rc = sigma * float(_param(gf, env, "cutoff", 2.5))
                spec.add_term(Term("lj_attractive", "UniformLennardJonesAttractive", dict(rc=rc, r_switch=0.9 * rc, eps=ep, sigma=sg, topo=tp)))


def _combinatorial_go(spec, gf, topo, periodic):
    """read_toml_combinatorial_go_contact_ff_generators (ReadTOMLForceFieldGenerator.cpp:2307-2402), one term per
    parameter row.  As the reference runs it: the `offset` is computed and then dropped (the un-offset index
    lists reach the generator, :2336-2339), and the ignore list, pairs of PARTICLE indices, is handed to
    CustomHbondForce::addExclusion, which takes (donor number, acceptor number) (CombinatorialGoContact...cpp:37-41); pairs that
    are not valid numbers there make OpenMM (and this engine) refuse the system."""
    env = gf.get("env")
    cutoff_ratio = float(gf.get("cutoff", 1.8))
    excl = np.zeros((0, 2), np.int32)
    if "ignore" in gf:
        excl, _ = _read_ignore(gf, topo)
    for p in gf["parameters"]:
        first, second = _param(p, env, "indices_pair")
        k = _param(p, env, "k") * KJ_PER_KCAL
        r0 = _param(p, env, "v0") * NM_PER_ANGSTROM
        don = np.full((len(first), 3), -1, np.int32); don[:, 0] = first
        acc = np.full((len(second), 3), -1, np.int32); acc[:, 0] = second
        da = DATopology(don, acc, r0 * cutoff_ratio, np.asarray(excl, np.int32).reshape(-1, 2), periodic)
        spec.add_term(Term("combinatorial_go_contact", "CombinatorialGoContact", dict(topo=da, k=k, r0=r0)))


# ------------------------------------------------------------------------------------------------
def build_system(lib: Library, spec: SystemSpec) -> System:
    """SystemGenerator::generate (src/SystemGenerator.cpp:34-111): masses, box, forces with groups."""
    sysm = System(lib, spec.mass, spec.box)
    for t in spec.terms:
        getattr(sysm, "add_" + t.kind)(t.group, **t.args)
    return sysm


def apply_ensemble(ctx: Context, spec: SystemSpec, seed=None):
    """SystemGenerator.cpp:89-108: the barostat of the input's ensemble, if any (its random stream is seeded with the integrator
    seed; OpenMM seeds its own from the clock)"""
    ens = spec.ensemble
    if not ens:
        return
    sd = spec.seed if seed is None else seed
    if ens["type"] == "NPT":
        ctx.set_anisotropic_barostat(ens["pressure"], spec.temperature, ens["scale_axis"], ens["frequency"], sd)
    else:
        ctx.set_membrane_barostat(ens["pressure"], ens["surface_tension"], spec.temperature, ens["xymode"], ens["zmode"], ens["frequency"], sd)


def make_context(lib: Library, spec: SystemSpec, device=0, seed=None, replicas=None, comm=None, rank=0, axis=1) -> Context:
    """Simulator ctor + initialize + set_velocity (src/Simulator.cpp:4-47, ReadTOMLInput.cpp:1094-1121).
    replicas = R builds a replica ensemble (config C4): R copies of the system in one context, replica r with the
    integrator seed `seed + r` (and velocity seed `vel_seed + r`), all starting from the spec's coordinates.
    comm/rank/axis build one rank of a slab decomposition (config C5); every rank passes the same spec."""
    sysm = build_system(lib, spec)
    base_seed = spec.seed if seed is None else seed
    if replicas is None:
        ctx = Context(sysm, spec.temperature, spec.friction, spec.dt, base_seed, device, comm=comm, rank=rank, axis=axis)
        ctx.set_positions(spec.positions)
        if spec.velocities is not None:
            ctx.set_velocities(spec.velocities)
        else:
            ctx.set_velocities_to_temperature(spec.temperature, spec.vel_seed)
        apply_ensemble(ctx, spec, base_seed)
    else:
        ctx = Context(sysm, spec.temperature, spec.friction, spec.dt, device=device,
                      replica_seeds=[base_seed + r for r in range(replicas)])
        ctx.set_positions(np.tile(spec.positions, (replicas, 1)))
        if spec.velocities is not None:
            ctx.set_velocities(np.tile(spec.velocities, (replicas, 1)))
        else:
            ctx.set_velocities_to_temperature(spec.temperature, spec.vel_seed)
    ctx.group_ids = dict(spec.group_ids)
    return ctx


# ------------------------------------------------------------------------------------------------
# SystemSpec <-> plain dict (JSON-able): golden fixtures travel to machines without the reference tree
def spec_to_dict(spec: SystemSpec) -> dict:
    def conv(v):
        if isinstance(v, PairTopology):
            return {"__topo__": True, "excl": np.asarray(v.excl).tolist(),
                    "igroups": [[list(map(int, a)), list(map(int, b))] for a, b in v.igroups], "periodic": bool(v.periodic)}
        if isinstance(v, DATopology):
            return {"__da__": True, "donors": np.asarray(v.donors).tolist(), "acceptors": np.asarray(v.acceptors).tolist(),
                    "cutoff": float(v.cutoff), "excl": np.asarray(v.excl).tolist(), "periodic": bool(v.periodic)}
        if isinstance(v, np.ndarray):
            return {"__nd__": True, "v": [None if (isinstance(x, float) and math.isnan(x)) else x for x in v.reshape(-1).tolist()],
                    "shape": list(v.shape)}
        if isinstance(v, (list, tuple)):
            return [conv(x) for x in v]
        if isinstance(v, (np.floating, np.integer)):
            return v.item()
        return v

    return {
        "mass": spec.mass.tolist(), "box": None if spec.box is None else spec.box.tolist(),
        "positions": spec.positions.tolist(), "velocities": None if spec.velocities is None else spec.velocities.tolist(),
        "names": spec.names, "groups": spec.groups, "group_ids": spec.group_ids,
        "temperature": spec.temperature, "friction": spec.friction, "dt": spec.dt, "seed": spec.seed,
        "vel_seed": spec.vel_seed, "total_step": spec.total_step, "save_step": spec.save_step,
        "delta_t_raw": spec.delta_t_raw, "energy_minimization": spec.energy_minimization, "periodic": spec.periodic,
        "ensemble": spec.ensemble,
        "terms": [{"kind": t.kind, "name": t.name, "group": t.group, "args": {k: conv(v) for k, v in t.args.items()}}
                  for t in spec.terms],
    }


def spec_from_dict(d: dict) -> SystemSpec:
    def conv(v):
        if isinstance(v, dict) and v.get("__topo__"):
            return PairTopology(excl=np.array(v["excl"], np.int32).reshape(-1, 2),
                                igroups=[(a, b) for a, b in v["igroups"]], periodic=v["periodic"])
        if isinstance(v, dict) and v.get("__da__"):
            return DATopology(donors=np.array(v["donors"], np.int32).reshape(-1, 3), acceptors=np.array(v["acceptors"], np.int32).reshape(-1, 3),
                              cutoff=v["cutoff"], excl=np.array(v["excl"], np.int32).reshape(-1, 2), periodic=v["periodic"])
        if isinstance(v, dict) and v.get("__nd__"):
            return np.array([np.nan if x is None else x for x in v["v"]], float).reshape(v["shape"])
        return v

    spec = SystemSpec(mass=np.array(d["mass"], float), box=None if d["box"] is None else np.array(d["box"], float),
                      positions=np.array(d["positions"], float),
                      velocities=None if d["velocities"] is None else np.array(d["velocities"], float),
                      names=d["names"], groups=d["groups"])
    for k in ("temperature", "friction", "dt", "seed", "vel_seed", "total_step", "save_step", "delta_t_raw",
              "energy_minimization", "periodic"):
        setattr(spec, k, d[k])
    spec.group_ids = dict(d["group_ids"])
    spec.ensemble = d.get("ensemble")
    for t in d["terms"]:
        spec.terms.append(Term(t["kind"], t["name"], {k: conv(v) for k, v in t["args"].items()}, t["group"]))
    return spec
