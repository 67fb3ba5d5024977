"""Synthetic workloads of the shapes BASELINE.json names (no reference file is needed at run time).

``hps_slab``  — HPS / Ashbaugh-Hatch IDR condensate slab: chains of the 163-residue FUS low-complexity
sequence (UniProt P35637 1-163, the sequence of the reference's sample/toml_input/FUS_100_HPS.toml), HPS
residue parameters (Dignon et al. 2018 as tabulated in that sample: sigma, lambda*0.2 kcal/mol, charge),
harmonic bonds k = 110.4 kcal/mol/A^2, LennardJonesRepulsive + LennardJonesAttractive (cutoff 6.19) with
bond<=1 exclusions and DebyeHuckel (I = 0.1 M, bond<=3 exclusions).  100 chains in the sample's box
(153.06 x 176.62 x 1138.2 A) reproduce config C2 (16 300 beads); other sizes tile the box in x/y.

The start structure is an overlap-free serpentine at the sample's bead density (1.12 nm^-3): rows of
beads 3.8 A apart along x, rows and layers 15.2 A apart joined by 4-bond connectors, one empty site
between consecutive chains.  It is deterministic (no RNG).
"""
from __future__ import annotations

import os

import numpy as np

from .capi import PairTopology
from .toml_input import (ANGSTROM_PER_NM, CAFETIME, KJ_PER_KCAL, NM_PER_ANGSTROM, SystemSpec, Term,
                         debye_huckel_constants)

FUS_LC = (
    "MASNDYTQQATQSYGAYPTQPGQGYSQQSNQPYGQQSYSGYGQSADTSGYGQSSYGSSYGQTQNTGYGTQSAPQGYGSTGGYGSSQSSQSSYGQQSSYPGYGQQPAPSS"
    "TSGSYGGSSQSSSYGQPQSGGYGQQSGYGGQQQSYGQQQSSYNPPQGYGQQNQY"
)
ONE_TO_THREE = {"A": "ALA", "R": "ARG", "N": "ASN", "D": "ASP", "C": "CYS", "Q": "GLN", "E": "GLU", "G": "GLY",
                "H": "HIS", "I": "ILE", "L": "LEU", "K": "LYS", "M": "MET", "F": "PHE", "P": "PRO", "S": "SER",
                "T": "THR", "W": "TRP", "Y": "TYR", "V": "VAL"}
# residue -> (mass amu, sigma A, lambda*0.2 kcal/mol, charge e); HPS model, values as in the sample
HPS = {
    "ALA": (71.09, 5.04, 0.00022, 0.0), "ASN": (114.11, 5.68, 0.08767, 0.0), "ASP": (115.09, 5.58, 0.00584, -1.0),
    "GLN": (128.14, 6.02, 0.06536, 0.0), "GLY": (57.05, 4.50, 0.14025, 0.0), "MET": (131.19, 6.18, 0.14918, 0.0),
    "PRO": (97.12, 5.56, 0.07459, 0.0), "SER": (87.08, 5.18, 0.09297, 0.0), "THR": (101.11, 5.62, 0.10760, 0.0),
    "TYR": (163.18, 6.46, 0.19900, 0.0),
}
EPS_REP = 0.2  # kcal/mol
C2_BOX_A = (153.06, 176.62, 1138.2)
BOND_A = 3.8


def serpentine_sites(n_sites, box_a, z_center_a=None):
    """Overlap-free space-filling path: consecutive sites BOND_A apart."""
    lx, ly, lz = box_a
    nrow = int((lx - 8.0) // BOND_A)  # sites per row; leaves > 2 bonds to the periodic image
    rows_per_layer = int((ly - 8.0) // (4 * BOND_A)) + 1
    per_layer = rows_per_layer * nrow + (rows_per_layer - 1) * 3 + 3
    n_layers = -(-n_sites // per_layer)
    height = (n_layers - 1) * 4 * BOND_A
    assert height < lz - 8.0, "slab does not fit in z"
    z0 = (lz - height) * 0.5 if z_center_a is None else z_center_a - height * 0.5
    x0 = (lx - (nrow - 1) * BOND_A) * 0.5
    y0 = (ly - (rows_per_layer - 1) * 4 * BOND_A) * 0.5
    out = np.empty((n_layers * per_layer, 3))
    k = 0
    xdir = 1
    for layer in range(n_layers):
        ydir = 1 if layer % 2 == 0 else -1
        z = z0 + layer * 4 * BOND_A
        for r in range(rows_per_layer):
            row = r if ydir == 1 else rows_per_layer - 1 - r
            y = y0 + row * 4 * BOND_A
            xs = np.arange(nrow) if xdir == 1 else np.arange(nrow - 1, -1, -1)
            out[k:k + nrow, 0] = x0 + xs * BOND_A
            out[k:k + nrow, 1] = y
            out[k:k + nrow, 2] = z
            k += nrow
            xend = out[k - 1, 0]
            if r < rows_per_layer - 1:
                for c in range(1, 4):
                    out[k] = (xend, y + ydir * c * BOND_A, z)
                    k += 1
            else:
                for c in range(1, 4):
                    out[k] = (xend, y, z + c * BOND_A)
                    k += 1
            xdir = -xdir
    return out[:n_sites]


def hps_slab(n_chains=100, tile=(1, 1), seq=FUS_LC, temperature=300.0, ionic_strength=0.1, delta_t=0.1,
             seed=12345, vel_seed=12345, cutoff=6.19) -> SystemSpec:
    """Config C2 (tile=(1,1), n_chains=100 -> 16 300 beads) or a tiled multiple of it."""
    box_a = (C2_BOX_A[0] * tile[0], C2_BOX_A[1] * tile[1], C2_BOX_A[2])
    L = len(seq)
    res = [ONE_TO_THREE[c] for c in seq]
    n = n_chains * L
    sites = serpentine_sites(n_chains * (L + 1), box_a)
    keep = (np.arange(n_chains * (L + 1)) % (L + 1)) < L
    pos = sites[keep] * NM_PER_ANGSTROM
    par = np.array([HPS[r] for r in res])
    mass = np.tile(par[:, 0], n_chains)
    sigma = np.tile(par[:, 1], n_chains) * NM_PER_ANGSTROM
    eps_att = np.tile(par[:, 2], n_chains) * KJ_PER_KCAL
    charge_chain = par[:, 3].copy()
    charge_chain[0] += 1.0   # N-terminus
    charge_chain[-1] -= 1.0  # C-terminus
    q = np.tile(charge_chain, n_chains)
    box = np.array(box_a) * NM_PER_ANGSTROM

    spec = SystemSpec(mass=mass, box=box, positions=pos, velocities=None, names=res * n_chains,
                      groups=["FUS"] * n)
    spec.periodic = True
    spec.temperature = temperature
    spec.delta_t_raw = delta_t
    spec.dt = delta_t * CAFETIME
    spec.seed = seed
    spec.vel_seed = vel_seed
    spec.total_step = 1_000_000
    spec.save_step = 1000

    first = (np.arange(n_chains)[:, None] * L + np.arange(L - 1)[None, :]).reshape(-1)
    bonds = np.stack([first, first + 1], 1).astype(np.int32)
    nb = bonds.shape[0]
    spec.add_term(Term("harmonic_bond", "HarmonicBond",
                       dict(ij=bonds, r0=np.full(nb, BOND_A * NM_PER_ANGSTROM),
                            k=np.full(nb, 110.4 * KJ_PER_KCAL * ANGSTROM_PER_NM * ANGSTROM_PER_NM * 2.0), periodic=True)))

    def within(k):
        out = []
        for d in range(1, k + 1):
            a = (np.arange(n_chains)[:, None] * L + np.arange(L - d)[None, :]).reshape(-1)
            out.append(np.stack([a, a + d], 1))
        e = np.concatenate(out).astype(np.int32)
        return e[np.lexsort((e[:, 1], e[:, 0]))]

    smax = np.sort(np.unique(sigma))[-2:] if np.unique(sigma).size > 1 else np.array([sigma.max()] * 2)
    two = np.sort(sigma)[-2:]
    rc = (two[0] + two[1]) * 0.5 * cutoff
    eps_rep = np.full(n, EPS_REP * KJ_PER_KCAL)
    spec.add_term(Term("lj_repulsive", "LennardJonesRepulsive",
                       dict(rc=rc, eps=eps_rep, sigma=sigma, topo=PairTopology(excl=within(1), periodic=True))))
    spec.add_term(Term("lj_attractive", "LennardJonesAttractive",
                       dict(rc=rc, r_switch=0.9 * rc, eps=eps_att, sigma=sigma,
                            topo=PairTopology(excl=within(1), periodic=True))))
    K, debye, rc_dh, cc = debye_huckel_constants(temperature, ionic_strength, 5.5)
    qq = np.where(q != 0.0, q, np.nan)
    charged = np.nonzero(q != 0.0)[0]
    spec.add_term(Term("debye_huckel", "DebyeHuckel",
                       dict(K=K, debye=debye, rc=rc_dh, cutoff_correction=cc, charge=qq,
                            topo=PairTopology(excl=within(3), igroups=[(charged, charged)], periodic=True))))
    return spec


C2_STATE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "c2_state.npz")


def c2_equilibrated(tile=(1, 1)) -> SystemSpec:
    """Config C2 (hps_slab, 100 chains) started from the committed compacted state data/c2_state.npz instead of the
    serpentine (tests/golden/make_c2_state.py: 40 000 Langevin steps; the slab compacts from 2.5 M to ~4.8 M in-cutoff pairs
    over that time, and the timed step cost follows the pair count).  Positions nm, velocities nm/ps.
    tile = (tx, ty): that state repeated tx x ty times in a box tx x ty times as large (tile (8, 8) = config C5,
    1 043 200 beads): exact periodic images, so the forces at t = 0 are the C2 forces copy by copy; the copies decorrelate
    through the per-bead Langevin noise."""
    tx, ty = tile
    spec = hps_slab(n_chains=100 * tx * ty, tile=(tx, ty))
    with np.load(C2_STATE) as d:
        x, v = d["x"].astype(np.float64), d["v"].astype(np.float64)
        assert x.shape[0] * tx * ty == spec.positions.shape[0]
        lx, ly = C2_BOX_A[0] * NM_PER_ANGSTROM, C2_BOX_A[1] * NM_PER_ANGSTROM
        spec.positions = np.concatenate([x + np.array([i * lx, j * ly, 0.0]) for i in range(tx) for j in range(ty)])
        spec.velocities = np.tile(v, (tx * ty, 1))
        spec.state = "compacted condensate: %d Langevin steps from the serpentine start, %d in-cutoff Ashbaugh-Hatch pairs " \
                     "per 16 300 beads (start: 2.5 M)" % (int(d["steps"]), int(d["pairs"]))
    return spec


def replicate(spec, shifts_nm, box=None):
    """`len(shifts_nm)` copies of a system side by side in one system (config C3: the nucleosome sample on a
    2 x 3 grid; with `box`: a periodic system tiled into a larger box, whose forces must equal the original's copy
    by copy).  Bead indices of copy c are offset by c*N, positions shifted by shifts_nm[c]; per-term and per-bead
    tables are tiled; exclusions, donor/acceptor tables and interaction-group sets are offset per copy (sets become
    the union over copies, so copies interact like any other beads)."""
    import copy as _copy

    from .capi import DATopology, PairTopology
    from .toml_input import SystemSpec, Term

    shifts = np.asarray(shifts_nm, float).reshape(-1, 3)
    m, n = shifts.shape[0], spec.mass.size
    assert (spec.box is None) == (box is None), "give the tiled box of a periodic system (and only of a periodic one)"
    out = SystemSpec(mass=np.tile(spec.mass, m), box=None if box is None else np.asarray(box, float),
                     positions=np.concatenate([spec.positions + s for s in shifts]),
                     velocities=None if spec.velocities is None else np.tile(spec.velocities, (m, 1)),
                     names=list(spec.names) * m, groups=list(spec.groups) * m)
    for k in ("temperature", "friction", "dt", "seed", "vel_seed", "total_step", "save_step", "delta_t_raw", "energy_minimization", "periodic"):
        setattr(out, k, getattr(spec, k))
    out.group_ids = dict(spec.group_ids)
    index_keys = {"ij": 2, "ijk": 3, "ijkl": 4}
    for t in spec.terms:
        a = {}
        nterm = None
        for key, w in index_keys.items():
            if key in t.args:
                nterm = np.asarray(t.args[key]).reshape(-1, w).shape[0]
        topo = t.args.get("topo")
        nd = na = None
        if isinstance(topo, DATopology):
            nd, na = np.asarray(topo.donors).reshape(-1, 3).shape[0], np.asarray(topo.acceptors).reshape(-1, 3).shape[0]
        for key, v in t.args.items():
            if key in index_keys:
                idx = np.asarray(v, np.int64).reshape(-1, index_keys[key])
                a[key] = np.concatenate([idx + c * n for c in range(m)])
            elif isinstance(v, PairTopology):
                ex = np.asarray(v.excl, np.int64).reshape(-1, 2)
                a[key] = PairTopology(excl=np.concatenate([ex + c * n for c in range(m)]).astype(np.int32),
                                      igroups=[(np.concatenate([np.asarray(s1, np.int64) + c * n for c in range(m)]),
                                                np.concatenate([np.asarray(s2, np.int64) + c * n for c in range(m)])) for s1, s2 in v.igroups],
                                      periodic=v.periodic)
            elif isinstance(v, DATopology):
                def off(sites):
                    sites = np.asarray(sites, np.int64).reshape(-1, 3)
                    return np.concatenate([np.where(sites >= 0, sites + c * n, -1) for c in range(m)]).astype(np.int32)
                ex = np.asarray(v.excl, np.int64).reshape(-1, 2)
                a[key] = DATopology(donors=off(v.donors), acceptors=off(v.acceptors), cutoff=v.cutoff,
                                    excl=np.concatenate([ex + np.array([c * nd, c * na]) for c in range(m)]).astype(np.int32), periodic=v.periodic)
            elif isinstance(v, (list, np.ndarray)) and key not in ("y10", "coef7") and np.asarray(v).ndim == 1 and \
                    np.asarray(v).size in {x for x in (nterm, n, nd, na) if x is not None}:
                a[key] = np.tile(np.asarray(v), m)
            else:
                a[key] = _copy.deepcopy(v)
        out.terms.append(Term(t.kind, t.name, a, t.group))
    return out
