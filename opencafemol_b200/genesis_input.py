"""GENESIS input path (src/input/ReadGenesisInput.cpp, ReadGenesisForceFieldGenerator.cpp): `.inp` control file, GROMACS-style
`.top` / `.itp` topology with `#include`, `.gro` coordinates -> the same SystemSpec the TOML reader produces.

Restated as the reference runs it, fixed-column parsing and quirks included:
  * numbers are read with std::stoi / std::stod from fixed columns whose lengths overrun (`substr(25, 42)`): the leading number of the
    slice counts;
  * bond coefficients of the .itp are halved (ReadGenesisForceFieldGenerator.cpp:22);
  * the run starts from zero velocities (no set_velocity call on this path), never minimises, always writes DCD + .ene, and dumps the
    initial structure to `pdbfile` when the control file names one (ReadGenesisInput.cpp:404-435);
  * the time step and the friction are in ps and 1/ps (no cafetime conversion on this path);
  * excluded volume ignores bonds within 3 and contacts within 1 whatever `nrexcl` says (…ForceFieldGenerator.cpp:279-287); all
    `[ atomtypes ]` rows must carry the same eps.
"""
from __future__ import annotations

import os
import re

import numpy as np

from . import tables
from . import toml_input as ti
from .capi import PairTopology

_SECTION = re.compile(r"\s*\[\s*(\S.*?)\s*]\s*")
_KEYVAL = re.compile(r"\s*(\S+)\s*=\s*(\S+)\s*")
_INCLUDE = re.compile(r'^#include "(\S*)"\s*')
_NUMBER = re.compile(r"\s*([+-]?(\d+\.?\d*|\.\d+)([eE][+-]?\d+)?)")
_INT = re.compile(r"\s*([+-]?\d+)")


def _stod(s):
    m = _NUMBER.match(s)
    if not m:
        raise ValueError("stod: no conversion in " + repr(s))
    return float(m.group(1))


def _stoi(s):
    m = _INT.match(s)
    if not m:
        raise ValueError("stoi: no conversion in " + repr(s))
    return int(m.group(1))


def read_inp_file(path):
    """[SECTION] / key = value with # comments (ReadGenesisInput.cpp:18-70); the first value of a repeated key wins (map::insert)"""
    data, section = {}, None
    with open(path) as fh:
        for line in fh.read().splitlines():
            m = _SECTION.fullmatch(line)
            if m:
                section = m.group(1)
                data.setdefault(section, {})
            elif section is not None and line.strip():
                kv = _KEYVAL.fullmatch(line.split("#", 1)[0])
                if kv:
                    data[section].setdefault(kv.group(1), kv.group(2))
    return data


def _preprocess_top(name, base):
    """comment stripping and #include expansion (ReadGenesisInput.cpp:72-113); includes are relative to the .inp's directory"""
    out = []
    with open(os.path.join(base, name)) as fh:
        for line in fh.read().splitlines():
            if not line.strip() or re.match(r"^\s*;", line):
                continue
            if ";" in line:
                out.append(line.split(";", 1)[0])
                continue
            m = _INCLUDE.fullmatch(line)
            if m:
                out.extend(_preprocess_top(m.group(1), base))
            else:
                out.append(line)
    return out


def read_top_file(name, base):
    """section name -> its data lines, sections of the same name concatenated (ReadGenesisInput.cpp:115-168)"""
    top, section = {}, None
    for line in _preprocess_top(name, base):
        m = _SECTION.fullmatch(line)
        if m:
            section = m.group(1)
            top.setdefault(section, [])
        elif section is None:
            raise RuntimeError(f'[error] the line "{line}" does not belong to any section in {name} file.')
        else:
            top[section].append(line)
    return top


def read_genesis(inp_path) -> ti.SystemSpec:
    """make_simulator_from_genesis_inputs (ReadGenesisInput.cpp:215-499) up to the Simulator: system, force fields, run parameters"""
    base = os.path.dirname(os.path.abspath(inp_path))
    inp = read_inp_file(inp_path)
    top = read_top_file(inp["INPUT"]["grotopfile"], base)
    if "atoms" not in top:
        raise RuntimeError("[error] There is no atoms section. Top file must contain one atoms section.")
    atoms = top["atoms"]
    n = len(atoms)
    mass = np.array([_stod(a[49:58]) for a in atoms])
    res_names = [a[27:30] for a in atoms]
    atom_names = [a[31:35] for a in atoms]
    aa = [a[10:15].replace(" ", "") for a in atoms]
    if inp["BOUNDARY"]["type"] == "PBC":
        # the reference marks its forces periodic but never gives the System a box (SystemGenerator(mass_vec) only): OpenMM then
        # refuses the cutoff against its default 2 nm box
        raise RuntimeError("GENESIS input with a periodic boundary: the reference sets no box on this path")
    periodic = False

    with open(os.path.join(base, inp["INPUT"]["grocrdfile"])) as fh:
        gro = fh.read().splitlines()
    ngro = _stoi(gro[1])
    pos = np.array([[_stod(g[20:29]), _stod(g[29:38]), _stod(g[38:47])] for g in gro[2:2 + ngro]])

    spec = ti.SystemSpec(mass=mass, box=None, positions=pos, velocities=np.zeros((n, 3)), names=res_names, groups=[None] * n)
    spec.periodic = False
    topo = ti.Topology(n)

    if "bonds" in top:
        ij = [[_stoi(b[0:10]) - 1, _stoi(b[10:20]) - 1] for b in top["bonds"]]
        if ij:
            topo.add_edges(ij, "bond")
            spec.add_term(ti.Term("harmonic_bond", "HarmonicBond", dict(ij=ij, r0=[_stod(b[25:]) for b in top["bonds"]],
                                                                        k=[_stod(b[43:]) * 0.5 for b in top["bonds"]], periodic=periodic)))
    if "angles" in top:
        ang = top["angles"]
        g = [a for a in ang if _stoi(a[30:35]) == 21]   # AICG2+ 1-3 Gaussian
        if g:
            spec.add_term(ti.Term("gaussian_bond", "GaussianBond",
                                  dict(ij=[[_stoi(a[0:10]) - 1, _stoi(a[20:30]) - 1] for a in g], k=[-_stod(a[50:65]) for a in g],
                                       r0=[_stod(a[35:50]) for a in g], sigma=[_stod(a[65:80]) for a in g], periodic=periodic)))
        for name in sorted(tables.FLA_SPLINE):
            idx = [[_stoi(a[0:10]) - 1, _stoi(a[10:20]) - 1, _stoi(a[20:30]) - 1] for a in ang
                   if _stoi(a[30:35]) == 22 and aa[_stoi(a[10:20]) - 1] == name]
            if not idx:
                continue
            if name not in ti._fla_cache:
                ti._fla_cache[name] = ti.fla_min_energy(name)
            spec.add_term(ti.Term("flexible_local_angle", f"FlexibleLocalAngle ({name})",
                                  dict(ijk=idx, k=[ti.KJ_PER_KCAL] * len(idx), y10=tables.FLA_SPLINE[name], theta_min=tables.FLA_MIN_THETA,
                                       theta_max=tables.FLA_MAX_THETA, e_min=ti._fla_cache[name], periodic=periodic)))
    if "dihedrals" in top:
        dih = top["dihedrals"]
        g = [d for d in dih if _stoi(d[40:45]) == 41]
        if g:
            spec.add_term(ti.Term("gaussian_dihedral", "GaussianDihedral",
                                  dict(ijkl=[[_stoi(d[0:10]) - 1, _stoi(d[10:20]) - 1, _stoi(d[20:30]) - 1, _stoi(d[30:40]) - 1] for d in g],
                                       k=[-_stod(d[60:75]) for d in g], phi0=[_stod(d[45:60]) / 180.0 * np.pi for d in g],
                                       sigma=[_stod(d[75:90]) for d in g], periodic=periodic)))
        fld = [[_stoi(d[0:10]) - 1, _stoi(d[10:20]) - 1, _stoi(d[20:30]) - 1, _stoi(d[30:40]) - 1] for d in dih if _stoi(d[40:45]) == 52]
        for key in sorted(tables.FLD_FOURIER):
            first, second = key
            if first == "R1":
                idx = [q for q in fld if aa[q[2]] == "GLY"]
            elif first == "R2":
                idx = [q for q in fld if aa[q[1]] != "GLY" and aa[q[2]] == "PRO"]
            elif first == "GLY" and second == "PRO":
                idx = [q for q in fld if aa[q[1]] == "GLY" and aa[q[2]] == "PRO"]
            else:
                idx = [q for q in fld if aa[q[1]] == first]   # (whatever the third residue is: …ForceFieldGenerator.cpp:206-216)
            if not idx:
                continue
            if key not in ti._fld_cache:
                ti._fld_cache[key] = ti.fld_min_energy(tables.FLD_FOURIER[key])
            spec.add_term(ti.Term("flexible_local_dihedral", f"FlexibleLocalDihedral ({first}-{second})",
                                  dict(ijkl=idx, k=[ti.KJ_PER_KCAL] * len(idx), coef7=tables.FLD_FOURIER[key], e_min=ti._fld_cache[key],
                                       periodic=periodic)))
    if "pairs" in top:
        pr = top["pairs"]
        ij = [[_stoi(p[0:10]) - 1, _stoi(p[10:20]) - 1] for p in pr]
        if ij:
            topo.add_edges(ij, "contact")
            spec.add_term(ti.Term("go_contact", "GoContact", dict(ij=ij, k=[_stod(p[45:60]) for p in pr], r0=[_stod(p[30:45]) for p in pr],
                                                                  periodic=periodic)))
    if "atomtypes" not in top:
        raise RuntimeError("[error] There is no `[ atomtypes ]` section. Genesis input mode needs this section.")
    if "moleculetype" not in top:
        raise RuntimeError("[error] There is no `[ moleculetype ]` section. Genesis input mode needs this section.")
    nrexcl = [_stoi(m[13:23]) for m in top["moleculetype"]]
    if any(v != nrexcl[0] for v in nrexcl):
        raise RuntimeError("[error] all nrexcl need to be same in `[moleculetype]` section.")
    rmin, eps_str = {}, ""
    for line in top["atomtypes"]:
        name = line[0:5].replace(" ", "")
        if name in rmin:
            raise RuntimeError("[error] `" + name + "` key was duplicated in `[ atomtypes ]` section.")
        rmin[name] = _stod(line[30:40]) * 0.5
        e = line[41:50].replace(" ", "")
        if not eps_str:
            eps_str = e
        elif eps_str != e:
            raise RuntimeError("[error] ExcludedVolumeForceField only support uniform eps for all particle case. Some different value were "
                               "defined in `[ atomtypes ]` table.")
    radius = np.array([rmin[t] for t in aa])
    excl = sorted(topo.ignore_list_within_edge(3, "bond") | topo.ignore_list_within_edge(1, "contact"))
    mx, sm = ti._two_largest(radius, strict=True)
    spec.add_term(ti.Term("excluded_volume", "ExcludedVolume",
                          dict(eps=_stod(eps_str), rc=(mx + sm) * 2.0, radius=radius,
                               topo=PairTopology(excl=np.array(excl, np.int32).reshape(-1, 2), periodic=periodic))))

    dyn, ens = inp["DYNAMICS"], inp["ENSEMBLE"]
    spec.total_step, spec.save_step = _stoi(dyn["nsteps"]), _stoi(dyn["crdout_period"])
    spec.delta_t_raw = spec.dt = _stod(dyn["timestep"])          # ps
    spec.seed = _stoi(dyn["iseed"]) if "iseed" in dyn else 0
    spec.temperature, spec.friction = _stod(ens["temperature"]), _stod(ens["gamma_t"])   # K, 1/ps
    spec.energy_minimization = False
    dcd = inp["OUTPUT"]["dcdfile"]
    cut = dcd.rfind(".")
    spec.output = {"path": "./", "prefix": dcd if cut < 0 else dcd[:cut], "format": "dcd"}
    if "pdbfile" in inp["OUTPUT"]:
        spec.output["initial_pdb"] = {"file": inp["OUTPUT"]["pdbfile"], "atom_names": atom_names, "res_names": res_names}
    spec.platform = {k: v for k, v in inp.get("PLATFORM", {}).items()}
    spec.topology = topo
    return spec
