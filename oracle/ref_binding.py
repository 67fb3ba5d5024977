"""ctypes binding of oracle/_ref/libocm_ref.so — the reference's own sources compiled against the stand-in
OpenMM / toml11 headers of oracle/ref_shim (see oracle/Makefile, target `ref`).

TEST INFRASTRUCTURE: imported by tests/ and tests/golden/make_ref_golden.py only.
"""
from __future__ import annotations

import ctypes as C
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libocm_ref.so")
REF_BIN = os.path.join(HERE, "_ref", "open_cafemol")


def available() -> bool:
    return os.path.exists(REF_SO)


class RefError(RuntimeError):
    pass


_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


class RefLibrary:
    def __init__(self, path: str = REF_SO):
        self.lib = C.CDLL(path)
        L = self.lib
        L.ref_last_error.restype = C.c_char_p
        L.ref_open_toml.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
        L.ref_open_simulator.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
        L.ref_run.argtypes = [C.c_void_p]
        L.ref_close.argtypes = [C.c_void_p]
        L.ref_close.restype = None
        L.ref_num_particles.argtypes = [C.c_void_p]
        L.ref_group_map.argtypes = [C.c_void_p, C.c_char_p, C.c_long]
        L.ref_describe.argtypes = [C.c_void_p, C.c_char_p, C.c_long]
        L.ref_describe.restype = C.c_long
        L.ref_set_positions.argtypes = [C.c_void_p, _dp]
        L.ref_set_velocities.argtypes = [C.c_void_p, _dp]
        L.ref_set_velocities_to_temperature.argtypes = [C.c_void_p, C.c_double, C.c_int]
        L.ref_get_state.argtypes = [C.c_void_p, C.c_uint, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_potential_energy.argtypes = [C.c_void_p, C.c_uint, C.POINTER(C.c_double), C.c_void_p]
        L.ref_step.argtypes = [C.c_void_p, C.c_int]
        if hasattr(L, "ref_get_box"):
            L.ref_get_box.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_minimize.argtypes = [C.c_void_p, C.c_double, C.c_int]
        L.ref_integrator.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int)]
        L.ref_toml_to_json.argtypes = [C.c_char_p, C.c_char_p, C.c_long]
        L.ref_toml_to_json.restype = C.c_long
        L.ref_toml11_to_json.argtypes = [C.c_char_p, C.c_char_p, C.c_long]
        L.ref_toml11_to_json.restype = C.c_long

    def check(self, rc):
        if rc != 0:
            raise RefError(self.lib.ref_last_error().decode(errors="replace"))

    def toml_to_json(self, path: str, reader="lite"):
        """`path` parsed by this repository's TOML reader ("lite") or by the toml library the reference was compiled
        with ("toml11": its own submodule when checked out), as a JSON-able object"""
        fn = self.lib.ref_toml_to_json if reader == "lite" else self.lib.ref_toml11_to_json
        n = fn(path.encode(), None, 0)
        if n < 0:
            raise RefError(self.lib.ref_last_error().decode(errors="replace"))
        buf = C.create_string_buffer(n)
        fn(path.encode(), buf, n)
        return json.loads(buf.value.decode())

    def open_toml(self, path: str) -> "RefSystem":
        """reference reader -> generators -> System, in the input's directory (relative includes)"""
        h = C.c_void_p()
        cwd = os.getcwd()
        try:
            os.chdir(os.path.dirname(os.path.abspath(path)) or ".")
            self.check(self.lib.ref_open_toml(os.path.abspath(path).encode(), C.byref(h)))
        finally:
            os.chdir(cwd)
        return RefSystem(self, h)

    def open_simulator(self, path: str) -> "RefSystem":
        """the reference's full front end (read_toml_input / read_genesis_input); run from the current directory"""
        h = C.c_void_p()
        self.check(self.lib.ref_open_simulator(path.encode(), C.byref(h)))
        return RefSystem(self, h)


class RefSystem:
    def __init__(self, lib: RefLibrary, h):
        self.L, self.h = lib, h
        self.n = lib.lib.ref_num_particles(h)

    def close(self):
        if self.h:
            self.L.lib.ref_close(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def group_map(self) -> dict:
        buf = C.create_string_buffer(1 << 16)
        self.L.check(self.L.lib.ref_group_map(self.h, buf, len(buf)))
        out = {}
        for item in buf.value.decode().split(";"):
            if item:
                k, v = item.rsplit("=", 1)
                out[k] = int(v)
        return out

    def describe(self) -> dict:
        n = self.L.lib.ref_describe(self.h, None, 0)
        if n < 0:
            raise RefError(self.L.lib.ref_last_error().decode(errors="replace"))
        buf = C.create_string_buffer(n)
        self.L.lib.ref_describe(self.h, buf, n)
        return json.loads(buf.value.decode())

    def set_positions(self, x):
        self.L.check(self.L.lib.ref_set_positions(self.h, np.ascontiguousarray(x, dtype=np.float64)))

    def set_velocities(self, v):
        self.L.check(self.L.lib.ref_set_velocities(self.h, np.ascontiguousarray(v, dtype=np.float64)))

    def set_velocities_to_temperature(self, T, seed):
        self.L.check(self.L.lib.ref_set_velocities_to_temperature(self.h, float(T), int(seed)))

    def positions(self):
        x = np.empty((self.n, 3))
        self.L.check(self.L.lib.ref_get_state(self.h, 0xFFFFFFFF, x.ctypes.data, None, None, None, None))
        return x

    def velocities(self):
        v = np.empty((self.n, 3))
        self.L.check(self.L.lib.ref_get_state(self.h, 0xFFFFFFFF, None, v.ctypes.data, None, None, None))
        return v

    def kinetic_energy(self):
        e, k = C.c_double(), C.c_double()
        self.L.check(self.L.lib.ref_get_state(self.h, 0xFFFFFFFF, None, None, None, C.addressof(e), C.addressof(k)))
        return k.value

    def energy(self, groups=0xFFFFFFFF, forces=False):
        e = C.c_double()
        f = np.empty((self.n, 3)) if forces else None
        self.L.check(self.L.lib.ref_potential_energy(self.h, groups & 0xFFFFFFFF, C.byref(e), f.ctypes.data if forces else None))
        return (e.value, f) if forces else e.value

    def energy_by_group(self) -> dict:
        return {name: self.energy(1 << g) for name, g in self.group_map().items()}

    def step(self, n):
        self.L.check(self.L.lib.ref_step(self.h, int(n)))

    def box(self):
        b = np.empty(3)
        self.L.check(self.L.lib.ref_get_box(self.h, b.ctypes.data))
        return b

    def minimize(self, tolerance=10.0, max_iterations=0):
        self.L.check(self.L.lib.ref_minimize(self.h, float(tolerance), int(max_iterations)))

    def run(self):
        self.L.check(self.L.lib.ref_run(self.h))

    def integrator(self) -> dict:
        T, g, dt, s = C.c_double(), C.c_double(), C.c_double(), C.c_int()
        self.L.check(self.L.lib.ref_integrator(self.h, C.byref(T), C.byref(g), C.byref(dt), C.byref(s)))
        return {"temperature": T.value, "friction": g.value, "dt": dt.value, "seed": s.value}
