"""ctypes binding of the C ABI declared in include/ocm_b200.h.

The binder is generic over (shared library path, symbol prefix): the product library
``opencafemol_b200/lib/libocm_b200.so`` exports ``ocm_*``.  Tests bind the CPU oracle
(``oracle/libocm_oracle.so``, prefix ``oracle_``) through the same class so that both sides are fed by
identical term tables; nothing in this package ever loads the oracle.

This mirrors how the reference drives OpenMM: build a System from force-field generators
(src/SystemGenerator.cpp:34-111), create a Context with a LangevinIntegrator (src/Simulator.cpp:14-17),
set positions / velocities (src/Simulator.cpp:41, src/Simulator.hpp:49-65), ``step`` (src/Simulator.cpp:60)
and ``getState`` (src/Observer.cpp:34,107,113,340,344).
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field

import numpy as np

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
PRODUCT_LIB = os.path.join(_PKG_DIR, "lib", "libocm_b200.so")

_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)
_f64p = C.POINTER(C.c_double)


class OcmError(RuntimeError):
    pass


class DATopologyStruct(C.Structure):
    _fields_ = [
        ("n_donors", C.c_int64),
        ("donors", C.POINTER(C.c_int32)),
        ("n_acceptors", C.c_int64),
        ("acceptors", C.POINTER(C.c_int32)),
        ("n_excl", C.c_int64),
        ("excl_da", C.POINTER(C.c_int32)),
        ("cutoff", C.c_double),
        ("periodic", C.c_int32),
    ]


class PairTopologyStruct(C.Structure):
    _fields_ = [
        ("n_excl", C.c_int64),
        ("excl_ij", _i32p),
        ("n_igroups", C.c_int32),
        ("set1_off", _i64p),
        ("set1", _i32p),
        ("set2_off", _i64p),
        ("set2", _i32p),
        ("periodic", C.c_int32),
    ]


@dataclass
class PairTopology:
    """Exclusions + interaction groups of one pair term (see ocm_pair_topology)."""

    excl: np.ndarray = field(default_factory=lambda: np.zeros((0, 2), np.int32))
    igroups: list = field(default_factory=list)  # list of (set1, set2) index arrays
    periodic: bool = False

    def to_struct(self):
        keep = []
        excl = np.ascontiguousarray(np.asarray(self.excl, np.int32).reshape(-1, 2))
        keep.append(excl)
        s = PairTopologyStruct()
        s.n_excl = excl.shape[0]
        s.excl_ij = excl.ctypes.data_as(_i32p)
        ng = len(self.igroups)
        s.n_igroups = ng
        off1 = np.zeros(ng + 1, np.int64)
        off2 = np.zeros(ng + 1, np.int64)
        m1, m2 = [], []
        for g, (a, b) in enumerate(self.igroups):
            a = np.asarray(sorted(a), np.int32)
            b = np.asarray(sorted(b), np.int32)
            m1.append(a)
            m2.append(b)
            off1[g + 1] = off1[g] + a.size
            off2[g + 1] = off2[g] + b.size
        set1 = np.ascontiguousarray(np.concatenate(m1) if m1 else np.zeros(0, np.int32), np.int32)
        set2 = np.ascontiguousarray(np.concatenate(m2) if m2 else np.zeros(0, np.int32), np.int32)
        keep += [off1, off2, set1, set2]
        s.set1_off = off1.ctypes.data_as(_i64p)
        s.set2_off = off2.ctypes.data_as(_i64p)
        s.set1 = set1.ctypes.data_as(_i32p)
        s.set2 = set2.ctypes.data_as(_i32p)
        s.periodic = 1 if self.periodic else 0
        return s, keep


@dataclass
class DATopology:
    """Donors, acceptors, exclusions and cutoff of one donor x acceptor term (see ocm_da_topology)."""

    donors: np.ndarray      # (nd, 3) bead indices, -1 = unused
    acceptors: np.ndarray   # (na, 3)
    cutoff: float
    excl: np.ndarray = field(default_factory=lambda: np.zeros((0, 2), np.int32))  # (donor index, acceptor index)
    periodic: bool = False

    def to_struct(self):
        d = np.ascontiguousarray(np.asarray(self.donors, np.int32).reshape(-1, 3))
        a = np.ascontiguousarray(np.asarray(self.acceptors, np.int32).reshape(-1, 3))
        e = np.ascontiguousarray(np.asarray(self.excl, np.int32).reshape(-1, 2))
        s = DATopologyStruct()
        s.n_donors, s.donors = d.shape[0], d.ctypes.data_as(_i32p)
        s.n_acceptors, s.acceptors = a.shape[0], a.ctypes.data_as(_i32p)
        s.n_excl, s.excl_da = e.shape[0], e.ctypes.data_as(_i32p)
        s.cutoff = float(self.cutoff)
        s.periodic = 1 if self.periodic else 0
        return s, [d, a, e]


def _f64(a):
    return np.ascontiguousarray(np.asarray(a, np.float64))


def _i32(a):
    return np.ascontiguousarray(np.asarray(a, np.int32))


class Library:
    """A loaded engine library (product or oracle) with typed entry points."""

    # name -> (restype, argtypes); every symbol include/ocm_b200.h declares
    SIGNATURES = {
        "last_error": (C.c_char_p, []),
        "version": (C.c_char_p, []),
        "device_count": (C.c_int, []),
        "system_create": (C.c_int, [C.c_int32, _f64p, _f64p, C.POINTER(C.c_void_p)]),
        "system_destroy": (None, [C.c_void_p]),
        "system_num_particles": (C.c_int, [C.c_void_p]),
        "system_add_harmonic_bond": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int64, _i32p, _f64p, _f64p]),
        "system_add_gaussian_bond": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int64, _i32p, _f64p, _f64p, _f64p]),
        "system_add_go_contact": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int64, _i32p, _f64p, _f64p]),
        "system_add_capped_go_contact": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int64, _i32p, _f64p, _f64p, _f64p, _f64p, _f64p]),
        "system_add_3spn2_bond": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int64, _i32p, _f64p, _f64p, _f64p]),
        "system_add_harmonic_angle": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int64, _i32p, _f64p, _f64p]),
        "system_add_flexible_local_angle": (
            C.c_int,
            [C.c_void_p, C.c_int, C.c_int, C.c_int64, _i32p, _f64p, _f64p, C.c_double, C.c_double, C.c_double],
        ),
        "system_add_gaussian_dihedral": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int64, _i32p, _f64p, _f64p, _f64p]),
        "system_add_cosine_dihedral": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int64, _i32p, _f64p, _f64p, _i32p]),
        "system_add_flexible_local_dihedral": (
            C.c_int,
            [C.c_void_p, C.c_int, C.c_int, C.c_int64, _i32p, _f64p, _f64p, C.c_double],
        ),
        "system_add_excluded_volume": (
            C.c_int,
            [C.c_void_p, C.c_int, C.c_double, C.c_double, _f64p, C.POINTER(PairTopologyStruct)],
        ),
        "system_add_debye_huckel": (
            C.c_int,
            [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, _f64p, C.POINTER(PairTopologyStruct)],
        ),
        "system_add_lj_repulsive": (C.c_int, [C.c_void_p, C.c_int, C.c_double, _f64p, _f64p, C.POINTER(PairTopologyStruct)]),
        "system_add_wca": (C.c_int, [C.c_void_p, C.c_int, C.c_double, _f64p, _f64p, C.POINTER(PairTopologyStruct)]),
        "system_add_isolf_attractive": (C.c_int, [C.c_void_p, C.c_int, C.c_double, _f64p, _f64p, _f64p, C.POINTER(PairTopologyStruct)]),
        "system_add_trigonometric": (C.c_int, [C.c_void_p, C.c_int, C.c_double, _f64p, _f64p, _f64p, C.POINTER(PairTopologyStruct)]),
        "system_add_lj_attractive": (
            C.c_int,
            [C.c_void_p, C.c_int, C.c_double, C.c_double, _f64p, _f64p, C.POINTER(PairTopologyStruct)],
        ),
        "system_add_3spn2_excluded_volume": (
            C.c_int,
            [C.c_void_p, C.c_int, C.c_double, _f64p, _f64p, C.POINTER(PairTopologyStruct)],
        ),
        "system_add_3spn2_base_stacking": (
            C.c_int,
            [C.c_void_p, C.c_int, C.c_int, C.c_int64, _i32p, _f64p, _f64p, _f64p, C.c_double, C.c_double],
        ),
        "system_add_3spn2_base_pair": (
            C.c_int,
            [C.c_void_p, C.c_int, C.POINTER(DATopologyStruct), _f64p, _f64p, _f64p, _f64p, _f64p, C.c_double, C.c_double],
        ),
        "system_add_3spn2_cross_stacking": (
            C.c_int,
            [C.c_void_p, C.c_int, C.POINTER(DATopologyStruct), _f64p, _f64p, _f64p, _f64p, C.c_double, C.c_double, C.c_double],
        ),
        "system_add_pdns": (
            C.c_int,
            [C.c_void_p, C.c_int, C.POINTER(DATopologyStruct), _f64p, _f64p, _f64p, _f64p, C.c_double, C.c_double],
        ),
        "system_add_position_restraint": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, _i32p, _f64p, _f64p, _f64p]),
        "system_add_pulling": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, _i32p, _f64p]),
        "system_add_exv_rectangular_box": (C.c_int, [C.c_void_p, C.c_int, C.c_double, _f64p, _f64p, C.c_int64, _i32p, _f64p]),
        "system_add_cylindrical_restraint": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int64, _i32p, _f64p, _f64p, _f64p, _f64p]),
        "system_add_harmonic_com_pulling": (
            C.c_int,
            [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int64, _i32p, C.c_int64, _i32p],
        ),
        "system_add_combinatorial_go_contact": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(DATopologyStruct), C.c_double, C.c_double]),
        "num_da_terms": (C.c_int, [C.c_void_p]),
        "get_da_pairs": (C.c_int, [C.c_void_p, C.c_int, _i32p, C.c_int64, _i64p]),
        "context_create": (
            C.c_int,
            [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_uint64, C.c_int, C.POINTER(C.c_void_p)],
        ),
        "context_destroy": (None, [C.c_void_p]),
        "context_create_batch": (
            C.c_int,
            [C.c_void_p, C.c_int32, C.c_double, C.c_double, C.c_double, C.POINTER(C.c_uint64), C.c_int, C.POINTER(C.c_void_p)],
        ),
        "num_replicas": (C.c_int, [C.c_void_p]),
        "get_replica_energies": (C.c_int, [C.c_void_p, C.c_uint32, _f64p, _f64p, _f64p]),
        "comm_unique_id": (C.c_int, [C.c_char_p]),
        "comm_create_nccl": (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
        "comm_create_local": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
        "comm_destroy": (None, [C.c_void_p]),
        "context_create_slab": (
            C.c_int,
            [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_uint64, C.c_int, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)],
        ),
        "slab_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), _f64p]),
        "get_roles": (C.c_int, [C.c_void_p, C.POINTER(C.c_ubyte)]),
        "context_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_double]),
        "context_get_option": (C.c_int, [C.c_void_p, C.c_char_p, _f64p]),
        "set_positions": (C.c_int, [C.c_void_p, _f64p]),
        "set_velocities": (C.c_int, [C.c_void_p, _f64p]),
        "set_velocities_to_temperature": (C.c_int, [C.c_void_p, C.c_double, C.c_uint64]),
        "step": (C.c_int, [C.c_void_p, C.c_int64]),
        "current_step": (C.c_int64, [C.c_void_p]),
        "minimize": (C.c_int, [C.c_void_p, C.c_double, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double)]),
        "context_set_barostat": (C.c_int, [C.c_void_p, C.c_int, _f64p, C.c_double, C.c_double, _i32p, C.c_int32, C.c_uint64]),
        "get_box": (C.c_int, [C.c_void_p, _f64p]),
        "set_box": (C.c_int, [C.c_void_p, _f64p]),
        "get_barostat_info": (C.c_int, [C.c_void_p, _f64p]),
        "get_state": (C.c_int, [C.c_void_p, C.c_uint32, _f64p, _f64p, _f64p, _f64p, _f64p, _f64p]),
        "get_state_begin": (C.c_int, [C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
        "get_state_end": (C.c_int, [C.c_void_p, _f64p, _f64p, _f64p, _f64p, _f64p, _f64p]),
        "set_state_async": (C.c_int, [C.c_void_p, _f64p, _f64p]),
        "num_pair_terms": (C.c_int, [C.c_void_p]),
        "get_neighbor_pairs": (C.c_int, [C.c_void_p, C.c_int, _i32p, C.c_int64, _i64p]),
        "get_cell_ids": (C.c_int, [C.c_void_p, C.c_int, _i32p, _i32p]),
        "get_counter": (C.c_int, [C.c_void_p, C.c_char_p, _f64p]),
        "profile_kernel": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int, _f64p]),
        "measure_fp32_peak": (C.c_int, [C.c_int, _f64p]),
    }

    def __init__(self, path: str = PRODUCT_LIB, prefix: str = "ocm_"):
        if not os.path.exists(path):
            raise OcmError(
                f"engine library not found: {path} — build it first (python -c 'import __graft_entry__ as g; g.build()')"
            )
        self.path = path
        self.prefix = prefix
        self.dll = C.CDLL(path)
        self.fn = {}
        for name, (res, args) in self.SIGNATURES.items():
            f = getattr(self.dll, prefix + name)
            f.restype = res
            f.argtypes = args
            self.fn[name] = f

    def check(self, status: int):
        if status != 0:
            raise OcmError(self.fn["last_error"]().decode("utf-8", "replace"))

    def version(self) -> str:
        return self.fn["version"]().decode()

    def device_count(self) -> int:
        return int(self.fn["device_count"]())

    def measure_fp32_peak(self, device=0) -> float:
        v = C.c_double()
        self.check(self.fn["measure_fp32_peak"](int(device), C.byref(v)))
        return v.value


_product = None


def product_library() -> Library:
    """The CUDA engine.  Fails loudly when the extension has not been built."""
    global _product
    if _product is None:
        _product = Library(PRODUCT_LIB, "ocm_")
    return _product


def _p(a, typ):
    return a.ctypes.data_as(typ) if a is not None else None


class System:
    """Term-table builder (OpenMM::System + Force objects in the reference)."""

    def __init__(self, lib: Library, mass, box=None):
        self.lib = lib
        mass = _f64(mass)
        self.n = int(mass.size)
        self.box = None if box is None else _f64(box)
        h = C.c_void_p()
        lib.check(lib.fn["system_create"](self.n, _p(mass, _f64p), _p(self.box, _f64p), C.byref(h)))
        self.handle = h
        self.pair_terms = []  # names, in order of addition (term index for introspection)

    def close(self):
        if self.handle:
            self.lib.fn["system_destroy"](self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- bonded -----------------------------------------------------------------------------
    def _bonded(self, name, group, periodic, idx, width, *cols, extra=()):
        idx = _i32(idx).reshape(-1, width)
        n = idx.shape[0]
        arrs = []
        for c in cols:
            if c.dtype == np.int32:
                assert c.size == n
                arrs.append(_p(c, _i32p))
            else:
                assert c.size == n, f"{name}: parameter column length {c.size} != {n}"
                arrs.append(_p(c, _f64p))
        self.lib.check(self.lib.fn[name](self.handle, int(group), int(bool(periodic)), n, _p(idx, _i32p), *arrs, *extra))

    def add_harmonic_bond(self, group, ij, r0, k, periodic=False):
        self._bonded("system_add_harmonic_bond", group, periodic, ij, 2, _f64(r0), _f64(k))

    def add_gaussian_bond(self, group, ij, k, r0, sigma, periodic=False):
        self._bonded("system_add_gaussian_bond", group, periodic, ij, 2, _f64(k), _f64(r0), _f64(sigma))

    def add_go_contact(self, group, ij, k, r0, periodic=False):
        self._bonded("system_add_go_contact", group, periodic, ij, 2, _f64(k), _f64(r0))

    def add_capped_go_contact(self, group, ij, k, r0, slope, intercept, cap, periodic=False):
        self._bonded("system_add_capped_go_contact", group, periodic, ij, 2, _f64(k), _f64(r0), _f64(slope), _f64(intercept), _f64(cap))

    def add_3spn2_bond(self, group, ij, k2, k4, r0, periodic=False):
        self._bonded("system_add_3spn2_bond", group, periodic, ij, 2, _f64(k2), _f64(k4), _f64(r0))

    def add_harmonic_angle(self, group, ijk, theta0, k, periodic=False):
        self._bonded("system_add_harmonic_angle", group, periodic, ijk, 3, _f64(theta0), _f64(k))

    def add_flexible_local_angle(self, group, ijk, k, y10, theta_min, theta_max, e_min, periodic=False):
        y10 = _f64(y10)
        assert y10.size == 10
        idx = _i32(ijk).reshape(-1, 3)
        k = _f64(k)
        self.lib.check(
            self.lib.fn["system_add_flexible_local_angle"](
                self.handle, int(group), int(bool(periodic)), idx.shape[0], _p(idx, _i32p), _p(k, _f64p),
                _p(y10, _f64p), float(theta_min), float(theta_max), float(e_min),
            )
        )

    def add_gaussian_dihedral(self, group, ijkl, k, phi0, sigma, periodic=False):
        self._bonded("system_add_gaussian_dihedral", group, periodic, ijkl, 4, _f64(k), _f64(phi0), _f64(sigma))

    def add_cosine_dihedral(self, group, ijkl, k, t0, nper, periodic=False):
        self._bonded("system_add_cosine_dihedral", group, periodic, ijkl, 4, _f64(k), _f64(t0), _i32(nper))

    def add_flexible_local_dihedral(self, group, ijkl, k, coef7, e_min, periodic=False):
        coef7 = _f64(coef7)
        assert coef7.size == 7
        idx = _i32(ijkl).reshape(-1, 4)
        k = _f64(k)
        self.lib.check(
            self.lib.fn["system_add_flexible_local_dihedral"](
                self.handle, int(group), int(bool(periodic)), idx.shape[0], _p(idx, _i32p), _p(k, _f64p),
                _p(coef7, _f64p), float(e_min),
            )
        )

    def add_3spn2_base_stacking(self, group, ijk, eps, r0, theta0, alpha, K_BS, periodic=False):
        self._bonded("system_add_3spn2_base_stacking", group, periodic, ijk, 3, _f64(eps), _f64(r0), _f64(theta0),
                     extra=(float(alpha), float(K_BS)))

    # ---- donor x acceptor -------------------------------------------------------------------
    def _da(self, name, group, topo: DATopology, cols, count, scalars):
        s, keep = topo.to_struct()
        arrs = []
        for c in cols:
            c = _f64(c)
            assert c.size == count, f"{name}: parameter column length {c.size} != {count}"
            keep.append(c)
            arrs.append(_p(c, _f64p))
        self.lib.check(self.lib.fn[name](self.handle, int(group), C.byref(s), *arrs, *[float(x) for x in scalars]))

    def add_3spn2_base_pair(self, group, topo, eps, r0, theta0_1, theta0_2, phi0, alpha, K_BP):
        nd = np.asarray(topo.donors).reshape(-1, 3).shape[0]
        self._da("system_add_3spn2_base_pair", group, topo, [eps, r0, theta0_1, theta0_2, phi0], nd, (alpha, K_BP))

    def add_3spn2_cross_stacking(self, group, topo, eps, r0, theta0_3, theta0_CS, K_BP, K_CS, alpha):
        na = np.asarray(topo.acceptors).reshape(-1, 3).shape[0]
        self._da("system_add_3spn2_cross_stacking", group, topo, [eps, r0, theta0_3, theta0_CS], na, (K_BP, K_CS, alpha))

    def add_pdns(self, group, topo, k, r0, theta0, phi0, sigma, delta):
        nd = np.asarray(topo.donors).reshape(-1, 3).shape[0]
        self._da("system_add_pdns", group, topo, [k, r0, theta0, phi0], nd, (sigma, delta))

    def add_combinatorial_go_contact(self, group, topo, k, r0):
        self._da("system_add_combinatorial_go_contact", group, topo, [], 0, (k, r0))

    # ---- external (one bead each) and centre-of-mass terms ------------------------------------
    def _rows(self, a, n, width, what):
        a = _f64(a).reshape(-1)
        assert a.size == n * width, f"{what}: expected {n} x {width} values, got {a.size}"
        return a

    def add_position_restraint(self, group, index, position, k, v0):
        idx = _i32(index).reshape(-1)
        n = idx.size
        pos, k, v0 = self._rows(position, n, 3, "position"), self._rows(k, n, 1, "k"), self._rows(v0, n, 1, "v0")
        self.lib.check(self.lib.fn["system_add_position_restraint"](self.handle, int(group), n, _p(idx, _i32p), _p(pos, _f64p), _p(k, _f64p), _p(v0, _f64p)))

    def add_pulling(self, group, index, force):
        idx = _i32(index).reshape(-1)
        f = self._rows(force, idx.size, 3, "force")
        self.lib.check(self.lib.fn["system_add_pulling"](self.handle, int(group), idx.size, _p(idx, _i32p), _p(f, _f64p)))

    def add_exv_rectangular_box(self, group, eps, lower, upper, index, radius):
        idx = _i32(index).reshape(-1)
        lo, hi, r = self._rows(lower, 1, 3, "lower"), self._rows(upper, 1, 3, "upper"), self._rows(radius, idx.size, 1, "radius")
        self.lib.check(self.lib.fn["system_add_exv_rectangular_box"](self.handle, int(group), float(eps), _p(lo, _f64p), _p(hi, _f64p), idx.size,
                                                                      _p(idx, _i32p), _p(r, _f64p)))

    def add_cylindrical_restraint(self, group, index, axis, shift, k, v0, periodic=False):
        idx = _i32(index).reshape(-1)
        n = idx.size
        ax, sh = self._rows(axis, n, 3, "axis"), self._rows(shift, n, 3, "shift")
        k, v0 = self._rows(k, n, 1, "k"), self._rows(v0, n, 1, "v0")
        self.lib.check(self.lib.fn["system_add_cylindrical_restraint"](self.handle, int(group), int(bool(periodic)), n, _p(idx, _i32p), _p(ax, _f64p),
                                                                        _p(sh, _f64p), _p(k, _f64p), _p(v0, _f64p)))

    def add_harmonic_com_pulling(self, group, k, v0, group1, group2, periodic=False):
        g1, g2 = _i32(group1).reshape(-1), _i32(group2).reshape(-1)
        self.lib.check(self.lib.fn["system_add_harmonic_com_pulling"](self.handle, int(group), int(bool(periodic)), float(k), float(v0), g1.size,
                                                                       _p(g1, _i32p), g2.size, _p(g2, _i32p)))

    # ---- pair -------------------------------------------------------------------------------
    def _per_bead(self, a):
        a = _f64(a)
        assert a.size == self.n, "per-bead parameter arrays must have one entry per particle (NaN = none)"
        return a

    def add_excluded_volume(self, group, eps, rc, radius, topo: PairTopology):
        s, keep = topo.to_struct()
        r = self._per_bead(radius)
        self.lib.check(self.lib.fn["system_add_excluded_volume"](self.handle, int(group), float(eps), float(rc), _p(r, _f64p), C.byref(s)))
        self.pair_terms.append("ExcludedVolume")

    def add_debye_huckel(self, group, K, debye, rc, cutoff_correction, charge, topo: PairTopology):
        s, keep = topo.to_struct()
        q = self._per_bead(charge)
        self.lib.check(
            self.lib.fn["system_add_debye_huckel"](
                self.handle, int(group), float(K), float(debye), float(rc), float(cutoff_correction), _p(q, _f64p), C.byref(s)
            )
        )
        self.pair_terms.append("DebyeHuckel")

    def add_lj_repulsive(self, group, rc, eps, sigma, topo: PairTopology):
        s, keep = topo.to_struct()
        e, sg = self._per_bead(eps), self._per_bead(sigma)
        self.lib.check(self.lib.fn["system_add_lj_repulsive"](self.handle, int(group), float(rc), _p(e, _f64p), _p(sg, _f64p), C.byref(s)))
        self.pair_terms.append("LennardJonesRepulsive")

    def add_wca(self, group, rc, eps, sigma, topo: PairTopology):
        s, keep = topo.to_struct()
        e, sg = self._per_bead(eps), self._per_bead(sigma)
        self.lib.check(self.lib.fn["system_add_wca"](self.handle, int(group), float(rc), _p(e, _f64p), _p(sg, _f64p), C.byref(s)))
        self.pair_terms.append("WeeksChandlerAndersen")

    def add_isolf_attractive(self, group, rc, eps, sigma, omega, topo: PairTopology):
        s, keep = topo.to_struct()
        e, sg, om = self._per_bead(eps), self._per_bead(sigma), self._per_bead(omega)
        self.lib.check(self.lib.fn["system_add_isolf_attractive"](self.handle, int(group), float(rc), _p(e, _f64p), _p(sg, _f64p), _p(om, _f64p), C.byref(s)))
        self.pair_terms.append("iSoLFAttractive")

    def add_trigonometric(self, group, rc, eps, sigma, omega, topo: PairTopology):
        s, keep = topo.to_struct()
        e, sg, om = self._per_bead(eps), self._per_bead(sigma), self._per_bead(omega)
        self.lib.check(self.lib.fn["system_add_trigonometric"](self.handle, int(group), float(rc), _p(e, _f64p), _p(sg, _f64p), _p(om, _f64p), C.byref(s)))
        self.pair_terms.append("Trigonometric")

    def add_lj_attractive(self, group, rc, r_switch, eps, sigma, topo: PairTopology):
        s, keep = topo.to_struct()
        e, sg = self._per_bead(eps), self._per_bead(sigma)
        self.lib.check(
            self.lib.fn["system_add_lj_attractive"](
                self.handle, int(group), float(rc), float(r_switch), _p(e, _f64p), _p(sg, _f64p), C.byref(s)
            )
        )
        self.pair_terms.append("LennardJonesAttractive")

    def add_3spn2_excluded_volume(self, group, rc, eps, sigma, topo: PairTopology):
        s, keep = topo.to_struct()
        e, sg = self._per_bead(eps), self._per_bead(sigma)
        self.lib.check(
            self.lib.fn["system_add_3spn2_excluded_volume"](self.handle, int(group), float(rc), _p(e, _f64p), _p(sg, _f64p), C.byref(s))
        )
        self.pair_terms.append("3SPN2ExcludedVolume")


class Comm:
    """Communicator of a slab decomposition (include/ocm_b200.h): `Comm.local(lib, world)` for ranks that are threads
    of this process, `Comm.nccl(lib, uid, world, rank, device)` for one process per GPU (uid = Comm.unique_id(lib) on
    rank 0, broadcast by the host program)."""

    def __init__(self, lib: Library, handle, world, kind):
        self.lib, self.handle, self.world, self.kind = lib, handle, int(world), kind

    @staticmethod
    def unique_id(lib: Library) -> bytes:
        buf = C.create_string_buffer(128)
        lib.check(lib.fn["comm_unique_id"](buf))
        return buf.raw

    @classmethod
    def nccl(cls, lib: Library, uid: bytes, world, rank, device):
        h = C.c_void_p()
        assert len(uid) == 128
        lib.check(lib.fn["comm_create_nccl"](C.create_string_buffer(uid, 128), int(world), int(rank), int(device), C.byref(h)))
        return cls(lib, h, world, "nccl")

    @classmethod
    def local(cls, lib: Library, world):
        h = C.c_void_p()
        lib.check(lib.fn["comm_create_local"](int(world), C.byref(h)))
        return cls(lib, h, world, "local")

    def close(self):
        if self.handle:
            self.lib.fn["comm_destroy"](self.handle)
            self.handle = None


class Context:
    """OpenMM::Context + LangevinIntegrator equivalent."""

    def __init__(self, system: System, temperature, friction, dt, seed=0, device=0, replica_seeds=None, comm: Comm = None,
                 rank=0, axis=1):
        self.lib = system.lib
        h = C.c_void_p()
        self.comm = comm
        if comm is not None:
            # one rank of a slab decomposition (collective calls: every rank makes the same sequence)
            self.n = system.n
            self.n_replicas = 1
            self.lib.check(
                self.lib.fn["context_create_slab"](system.handle, float(temperature), float(friction), float(dt), int(seed), int(device),
                                                   comm.handle, int(rank), int(axis), C.byref(h))
            )
        elif replica_seeds is None:
            self.n = system.n
            self.n_replicas = 1
            self.lib.check(
                self.lib.fn["context_create"](system.handle, float(temperature), float(friction), float(dt), int(seed), int(device), C.byref(h))
            )
        else:
            # replica ensemble: one context, len(replica_seeds) independent copies of the system
            seeds = np.ascontiguousarray(np.asarray(replica_seeds, np.uint64))
            self.n_replicas = int(seeds.size)
            self.n = system.n * self.n_replicas
            self.lib.check(
                self.lib.fn["context_create_batch"](system.handle, self.n_replicas, float(temperature), float(friction), float(dt),
                                                    seeds.ctypes.data_as(C.POINTER(C.c_uint64)), int(device), C.byref(h))
            )
        self.handle = h
        self.pair_terms = list(system.pair_terms)

    def close(self):
        if self.handle:
            self.lib.fn["context_destroy"](self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, key, value):
        self.lib.check(self.lib.fn["context_set_option"](self.handle, key.encode(), float(value)))

    def get_option(self, key):
        v = C.c_double()
        self.lib.check(self.lib.fn["context_get_option"](self.handle, key.encode(), C.byref(v)))
        return v.value

    def set_positions(self, xyz):
        a = _f64(xyz).reshape(-1)
        assert a.size == 3 * self.n
        self.lib.check(self.lib.fn["set_positions"](self.handle, _p(a, _f64p)))

    def set_velocities(self, xyz):
        a = _f64(xyz).reshape(-1)
        assert a.size == 3 * self.n
        self.lib.check(self.lib.fn["set_velocities"](self.handle, _p(a, _f64p)))

    def set_velocities_to_temperature(self, temperature, seed):
        self.lib.check(self.lib.fn["set_velocities_to_temperature"](self.handle, float(temperature), int(seed)))

    def step(self, nsteps):
        self.lib.check(self.lib.fn["step"](self.handle, int(nsteps)))

    def minimize(self, tolerance=10.0, max_iterations=0):
        """LocalEnergyMinimizer::minimize (src/Simulator.cpp:43-46); returns (iterations, final potential energy)"""
        it, e = C.c_int(), C.c_double()
        self.lib.check(self.lib.fn["minimize"](self.handle, float(tolerance), int(max_iterations), C.byref(it), C.byref(e)))
        return it.value, e.value

    # ---- Monte Carlo barostats (OpenMM::MonteCarloAnisotropicBarostat / MonteCarloMembraneBarostat) ----
    def set_anisotropic_barostat(self, pressure_bar, temperature, scale_axis=(True, True, True), frequency=25, seed=0):
        """ensemble "NPT" of the reference's input (ReadTOMLInput.cpp:88-109)"""
        p = _f64(np.broadcast_to(np.asarray(pressure_bar, dtype=np.float64), (3,)).copy())
        m = _i32([1 if a else 0 for a in scale_axis])
        self.lib.check(self.lib.fn["context_set_barostat"](self.handle, 1, _p(p, _f64p), 0.0, float(temperature), _p(m, _i32p),
                                                           int(frequency), int(seed)))

    def set_membrane_barostat(self, pressure_bar, surface_tension_bar_nm, temperature, xymode="Isotropic", zmode="Free", frequency=25, seed=0):
        """ensemble "NPgammaT" of the reference's input (ReadTOMLInput.cpp:110-169)"""
        p = _f64([float(pressure_bar), 0.0, 0.0])
        m = _i32([{"Isotropic": 0, "Anisotropic": 1}[xymode], {"Free": 0, "Fixed": 1, "ConstantVolume": 2}[zmode], 0])
        self.lib.check(self.lib.fn["context_set_barostat"](self.handle, 2, _p(p, _f64p), float(surface_tension_bar_nm), float(temperature),
                                                           _p(m, _i32p), int(frequency), int(seed)))

    def remove_barostat(self):
        p = _f64([0.0, 0.0, 0.0]); m = _i32([0, 0, 0])
        self.lib.check(self.lib.fn["context_set_barostat"](self.handle, 0, _p(p, _f64p), 0.0, 0.0, _p(m, _i32p), 0, 0))

    def box(self):
        b = np.zeros(3)
        self.lib.check(self.lib.fn["get_box"](self.handle, _p(b, _f64p)))
        return b

    def set_box(self, box_nm):
        b = _f64(box_nm)
        self.lib.check(self.lib.fn["set_box"](self.handle, _p(b, _f64p)))

    def barostat_info(self):
        o = np.zeros(12)
        self.lib.check(self.lib.fn["get_barostat_info"](self.handle, _p(o, _f64p)))
        return {"trials": int(o[0]), "accepted": int(o[1]), "axis": int(o[2]), "dV": o[3], "e_before": o[4], "e_after": o[5], "w": o[6],
                "last_accepted": bool(o[7]), "volume_step": o[8:11].copy(), "molecules": int(o[11])}

    @property
    def current_step(self):
        return int(self.lib.fn["current_step"](self.handle))

    def get_state(self, positions=False, velocities=False, forces=False, energy=False, groups=0xFFFFFFFF):
        out = {}
        pos = np.empty((self.n, 3)) if positions else None
        vel = np.empty((self.n, 3)) if velocities else None
        frc = np.empty((self.n, 3)) if forces else None
        ep, ek = C.c_double(), C.c_double()
        ebg = np.zeros(32) if energy else None
        self.lib.check(
            self.lib.fn["get_state"](
                self.handle, int(groups) & 0xFFFFFFFF, _p(pos, _f64p), _p(vel, _f64p), _p(frc, _f64p),
                C.byref(ep) if energy else None, C.byref(ek) if energy else None, _p(ebg, _f64p),
            )
        )
        if positions:
            out["positions"] = pos
        if velocities:
            out["velocities"] = vel
        if forces:
            out["forces"] = frc
        if energy:
            out["potential_energy"] = ep.value
            out["kinetic_energy"] = ek.value
            out["energy_by_group"] = ebg
        return out

    def get_state_begin(self, positions=False, velocities=False, forces=False, energy=False, groups=0xFFFFFFFF):
        """Observer without stalls (Observer.cpp:103-118, 324-346): queues the frame's energy kernels behind the issued steps and
        its download on a side stream, returns at once; `step` may follow immediately.  Finish with get_state_end(request)."""
        h = C.c_void_p()
        self.lib.check(self.lib.fn["get_state_begin"](self.handle, int(groups) & 0xFFFFFFFF, int(positions), int(velocities), int(forces),
                                                       int(energy), C.byref(h)))
        return (h, positions, velocities, forces, energy)

    def get_state_end(self, request):
        h, positions, velocities, forces, energy = request
        out = {}
        pos = np.empty((self.n, 3)) if positions else None
        vel = np.empty((self.n, 3)) if velocities else None
        frc = np.empty((self.n, 3)) if forces else None
        ep, ek = C.c_double(), C.c_double()
        ebg = np.zeros(32) if energy else None
        self.lib.check(self.lib.fn["get_state_end"](h, _p(pos, _f64p), _p(vel, _f64p), _p(frc, _f64p), C.byref(ep) if energy else None,
                                                     C.byref(ek) if energy else None, _p(ebg, _f64p)))
        if positions:
            out["positions"] = pos
        if velocities:
            out["velocities"] = vel
        if forces:
            out["forces"] = frc
        if energy:
            out["potential_energy"] = ep.value
            out["kinetic_energy"] = ek.value
            out["energy_by_group"] = ebg
        return out

    def set_state_async(self, positions, velocities=None):
        """positions (and velocities) through pinned staging without a host synchronisation"""
        a = _f64(positions).reshape(-1)
        b = _f64(velocities).reshape(-1) if velocities is not None else None
        self.lib.check(self.lib.fn["set_state_async"](self.handle, _p(a, _f64p), _p(b, _f64p)))

    def slab_info(self):
        r, w, a, h = C.c_int(), C.c_int(), C.c_int(), C.c_double()
        self.lib.check(self.lib.fn["slab_info"](self.handle, C.byref(r), C.byref(w), C.byref(a), C.byref(h)))
        return {"rank": r.value, "world": w.value, "axis": a.value, "halo_nm": h.value}

    def roles(self):
        out = np.zeros(self.n, np.uint8)
        self.lib.check(self.lib.fn["get_roles"](self.handle, out.ctypes.data_as(C.POINTER(C.c_ubyte))))
        return out

    def replica_energies(self, groups=0xFFFFFFFF):
        r = self.n_replicas
        ep, ek, ebg = np.zeros(r), np.zeros(r), np.zeros((r, 32))
        self.lib.check(self.lib.fn["get_replica_energies"](self.handle, int(groups) & 0xFFFFFFFF, _p(ep, _f64p), _p(ek, _f64p), _p(ebg, _f64p)))
        return {"potential_energy": ep, "kinetic_energy": ek, "energy_by_group": ebg}

    def num_pair_terms(self):
        return int(self.lib.fn["num_pair_terms"](self.handle))

    def neighbor_pairs(self, term):
        n = C.c_int64()
        self.lib.check(self.lib.fn["get_neighbor_pairs"](self.handle, int(term), None, 0, C.byref(n)))
        ij = np.zeros((max(n.value, 1), 2), np.int32)
        self.lib.check(self.lib.fn["get_neighbor_pairs"](self.handle, int(term), _p(ij, _i32p), ij.shape[0], C.byref(n)))
        return ij[: n.value]

    def num_neighbor_pairs(self, term):
        """number of in-cutoff pairs of pair term `term` (the counting pass of neighbor_pairs, nothing is copied)"""
        n = C.c_int64()
        self.lib.check(self.lib.fn["get_neighbor_pairs"](self.handle, int(term), None, 0, C.byref(n)))
        return int(n.value)

    def num_da_terms(self):
        return int(self.lib.fn["num_da_terms"](self.handle))

    def da_pairs(self, term):
        n = C.c_int64()
        self.lib.check(self.lib.fn["get_da_pairs"](self.handle, int(term), None, 0, C.byref(n)))
        da = np.zeros((max(n.value, 1), 2), np.int32)
        self.lib.check(self.lib.fn["get_da_pairs"](self.handle, int(term), _p(da, _i32p), da.shape[0], C.byref(n)))
        return da[: n.value]

    def cell_ids(self, term):
        ids = np.zeros(self.n, np.int32)
        dims = np.zeros(3, np.int32)
        self.lib.check(self.lib.fn["get_cell_ids"](self.handle, int(term), _p(ids, _i32p), _p(dims, _i32p)))
        return ids, dims

    def counter(self, key):
        v = C.c_double()
        self.lib.check(self.lib.fn["get_counter"](self.handle, key.encode(), C.byref(v)))
        return v.value

    def profile_kernel(self, which, reps=50):
        """average device milliseconds of one launch of `which` ("force", "integrate", "rebuild")"""
        v = C.c_double()
        self.lib.check(self.lib.fn["profile_kernel"](self.handle, which.encode(), int(reps), C.byref(v)))
        return v.value
