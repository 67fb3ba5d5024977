"""The C-ABI library loads without a GPU and exports every symbol include/ocm_b200.h declares; the oracle
exports the same surface under its own prefix.  CPU only (no compute calls)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT
from opencafemol_b200.capi import Library, OcmError, PairTopology, System


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "ocm_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(ocm_[a-z0-9_]+)\s*\(", hdr)))


def test_engine_exports_every_declared_symbol(built):
    dll = ctypes.CDLL(built.ENGINE_SO)
    names = declared_symbols()
    assert len(names) >= 34
    for n in names:
        assert hasattr(dll, n), n


def test_binder_covers_header(built):
    declared = {n[len("ocm_"):] for n in declared_symbols()}
    assert declared == set(Library.SIGNATURES), declared ^ set(Library.SIGNATURES)


def test_oracle_exports_same_surface(built):
    dll = ctypes.CDLL(built.ORACLE_SO)
    for n in declared_symbols():
        assert hasattr(dll, "oracle_" + n[len("ocm_"):]), n


def test_host_side_validation_without_gpu(built):
    """Term-table building is host code: argument errors surface through ocm_last_error."""
    lib = Library(built.ENGINE_SO, "ocm_")
    assert "sm_100a" in lib.version()
    s = System(lib, np.full(4, 12.0), box=[3.0, 3.0, 3.0])
    with pytest.raises(OcmError, match="out of range"):
        s.add_harmonic_bond(0, [[0, 9]], [0.38], [100.0])
    with pytest.raises(OcmError, match="group"):
        s.add_harmonic_bond(40, [[0, 1]], [0.38], [100.0])
    with pytest.raises(OcmError, match="half the box"):
        s.add_excluded_volume(0, 1.0, 2.0, np.full(4, 0.2), PairTopology(periodic=True))
    with pytest.raises(OcmError, match="r_switch"):
        s.add_lj_attractive(0, 1.0, 1.2, np.full(4, 0.2), np.full(4, 0.5), PairTopology(periodic=True))
    s.add_harmonic_bond(0, [[0, 1]], [0.38], [100.0], periodic=True)
    if lib.device_count() == 0:
        from opencafemol_b200.capi import Context
        with pytest.raises(OcmError, match="no CUDA device"):
            Context(s, 300.0, 0.2, 0.001)


def test_header_is_valid_c_and_links_from_c(built, tmp_path):
    """include/ocm_b200.h compiled as C99 and linked from a C program (no torch, no C++ types at the boundary)"""
    import subprocess

    exe = tmp_path / "abi_smoke"
    libdir = os.path.dirname(built.ENGINE_SO)
    cmd = ["/usr/bin/gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c", "abi_smoke.c"),
           "-o", str(exe), "-L", libdir, "-locm_b200", "-Wl,-rpath," + libdir]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert "abi ok" in r.stdout
