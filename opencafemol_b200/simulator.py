"""The reference's run loop and observers on the C ABI, in Python: `python -m opencafemol_b200 input.toml | input.inp` runs a TOML or
GENESIS input end to end on the B200 engine and writes the files the reference writes (.ene, _position.dcd / _velocity.dcd or .pdb), without
the reference tree (the drop-in build `bin/open_cafemol_b200` compiles the reference's own front end and needs it).

What is restated here, with the reference file:line it follows:
  * Simulator::initialize / run (src/Simulator.cpp:39-76): positions, optional LocalEnergyMinimizer, then per frame
    `observer->output(frame * save_step)` for every observer followed by `integrator->step(save_step)`; frames = floor(total / save);
  * the observer set of read_toml_input (src/input/ReadTOMLInput.cpp:1023-1052): a trajectory observer by `files.output.format`
    ("pdb" or "dcd", anything else is an error) and the energy observer; output path / prefix rules of :987-1000;
  * EnergyObserver (src/Observer.cpp:310-346), DCDObserver (:64-298), PDBObserver (:6-62): byte-for-byte the same layouts.

The engine is asked once per frame for everything the observers of that frame need (`Context.get_state`): one device evaluation
of all group energies instead of the reference's G + 1 `getState` calls.
"""
from __future__ import annotations

import math
import os
import struct
import sys

import numpy as np

from . import toml_input as ti
from .capi import Library

KCAL_PER_KJ = 1.0 / 4.184
ANGSTROM_PER_NM = 10.0


class EnergyObserver:
    """`<prefix>.ene`: one column per force-field name in std::map (byte-wise alphabetical) order, then the kinetic energy;
    kcal/mol, six decimals (src/Observer.cpp:310-346)."""

    def __init__(self, prefix, group_ids):
        self.path = prefix + ".ene"
        self.columns = sorted(group_ids.items(), key=lambda kv: kv[0].encode())
        open(self.path, "w").close()

    def initialize(self, n):
        with open(self.path, "w") as fh:
            fh.write("# unit of energy : kcal/mol\n")
            fh.write("# timestep" + "".join(" %24s" % name for name, _ in self.columns) + " kinetic_energy\n")

    needs = dict(energy=True)

    def output(self, step, state, box):
        with open(self.path, "a") as fh:
            fh.write("%-10d" % step)
            for _, gid in self.columns:
                fh.write(" %24.6f" % (state["energy_by_group"][gid] * KCAL_PER_KJ))
            fh.write(" %14.6f\n" % (state["kinetic_energy"] * KCAL_PER_KJ))


class DCDObserver:
    """`<prefix>_position.dcd` and `<prefix>_velocity.dcd` (src/Observer.cpp:64-298): CHARMM layout, version 24, three header blocks,
    per frame an optional unit-cell block (periodic boundary, position file only) and one float32 block per axis; positions in
    Angstrom, velocities in Angstrom per cafetime; the header's delta_t is the raw input value."""

    def __init__(self, prefix, total_step, save_step, delta_t, periodic):
        self.pos_path, self.vel_path = prefix + "_position.dcd", prefix + "_velocity.dcd"
        self.total_step, self.save_step, self.delta_t, self.periodic = int(total_step), int(save_step), float(delta_t), bool(periodic)
        open(self.pos_path, "wb").close()
        open(self.vel_path, "wb").close()

    def header(self, n):
        b = struct.pack("<i4s", 84, b"CORD")
        b += struct.pack("<5i", int(math.floor(self.total_step / self.save_step)), 0, self.save_step, self.total_step, 0)
        b += struct.pack("<4i", 0, 0, 0, 0)
        b += struct.pack("<f", self.delta_t)
        b += struct.pack("<i", 1 if self.periodic else 0)
        b += struct.pack("<8i", *([0] * 8))
        b += struct.pack("<2i", 24, 84)
        comment = b"OpenMoCa -- written by Yutaka Murata 2022"
        b += struct.pack("<2i", 84, 1) + comment + b"\0" * (80 - len(comment)) + struct.pack("<i", 84)
        b += struct.pack("<3i", 4, n, 4)
        return b

    def initialize(self, n):
        for p in (self.pos_path, self.vel_path):
            with open(p, "ab") as fh:
                fh.write(self.header(n))

    needs = dict(positions=True, velocities=True)

    @staticmethod
    def _planes(fh, a):
        a = np.asarray(a, np.float64)
        size = struct.pack("<i", 4 * a.shape[0])
        for k in range(3):
            fh.write(size + a[:, k].astype("<f4").tobytes() + size)

    def output(self, step, state, box):
        with open(self.pos_path, "ab") as fh:
            if self.periodic:
                x, y, z = (float(v) * ANGSTROM_PER_NM for v in box)
                fh.write(struct.pack("<i6di", 48, x, 90.0, y, 90.0, 90.0, z, 48))
            self._planes(fh, state["positions"] * ANGSTROM_PER_NM)
        with open(self.vel_path, "ab") as fh:
            self._planes(fh, state["velocities"] * ANGSTROM_PER_NM * ti.CAFETIME)


class PDBObserver:
    """`<prefix>.pdb`: one MODEL per frame, CA atoms, residue name = particle name or UNK, Angstrom with three decimals
    (src/Observer.cpp:6-62)."""

    def __init__(self, prefix, names):
        self.path = prefix + ".pdb"
        self.names = [nm if nm else "UNK" for nm in names]
        open(self.path, "w").close()

    def initialize(self, n):
        pass

    needs = dict(positions=True)

    def output(self, step, state, box):
        pos = state["positions"] * ANGSTROM_PER_NM
        with open(self.path, "a") as fh:
            fh.write("MODEL     %d\n" % step)
            for i, (p, name) in enumerate(zip(pos, self.names)):
                fh.write("ATOM  %5d   CA %3s     1    %8.3f%8.3f%8.3f  1.00  0.00\n" % (i + 1, name, p[0], p[1], p[2]))
            fh.write("ENDMDL\n")


class Simulator:
    """Simulator of the reference (src/Simulator.cpp) for one TOML input on a given engine library."""

    def __init__(self, lib: Library, spec: ti.SystemSpec, workdir=".", device=0):
        out = spec.output
        for key in ("prefix", "format", "path"):
            if key not in out:
                raise KeyError(f"[files.output] {key} is missing")   # toml::find throws in the reference
        path = out["path"] or "./"
        if not path.endswith("/"):
            path += "/"
        prefix = os.path.join(workdir, path, out["prefix"])   # (an absolute output path wins over workdir)
        self.spec = spec
        self.total_step, self.save_step = int(spec.total_step), int(spec.save_step)
        fmt = out["format"]
        if fmt == "pdb":
            self.observers = [PDBObserver(prefix, spec.names)]
        elif fmt == "dcd":
            self.observers = [DCDObserver(prefix, self.total_step, self.save_step, spec.delta_t_raw, spec.periodic)]
        else:
            raise RuntimeError("[error] output file format `" + str(fmt) + "` is not supported.")
        self.observers.append(EnergyObserver(prefix, spec.group_ids))
        if "initial_pdb" in out:
            # GENESIS path: the input structure goes to `pdbfile` before the run (src/input/ReadGenesisInput.cpp:404-435)
            ip = out["initial_pdb"]
            with open(os.path.join(workdir, ip["file"]), "w") as fh:
                fh.write("MODEL     0\n")
                for i, p in enumerate(spec.positions * ANGSTROM_PER_NM):
                    fh.write("ATOM  %5d %s %s     1    %8.3f%8.3f%8.3f  1.00  0.00\n" % (i + 1, ip["atom_names"][i], ip["res_names"][i], p[0], p[1], p[2]))
                fh.write("ENDMDL\n")
        # Simulator ctor + initialize + set_velocity (make_context); the minimiser runs between positions and velocities in the
        # reference, which is the same thing: it does not touch velocities
        self.ctx = ti.make_context(lib, spec, device=device)
        for ob in self.observers:
            ob.initialize(spec.mass.size)
        if spec.energy_minimization:
            self.ctx.minimize()   # LocalEnergyMinimizer::minimize(context): tolerance 10 kJ/mol/nm, until converged
        self.needs = {}
        for ob in self.observers:
            self.needs.update(ob.needs)

    def run(self, progress=None):
        frames = int(math.floor(self.total_step / self.save_step))
        for f in range(frames):
            step = f * self.save_step
            state = self.ctx.get_state(**self.needs)
            box = self.ctx.box() if self.spec.periodic else None
            for ob in self.observers:
                ob.output(step, state, box)
            self.ctx.step(self.save_step)
            if progress:
                progress(step, self.total_step)
        return frames


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if len(argv) != 1 or argv[0] in ("-h", "--help"):
        print("usage: python -m opencafemol_b200 <input.toml> or <input.inp>", file=sys.stderr)
        return 2
    from .capi import product_library

    lib = product_library()   # raises when the CUDA extension is missing; ocm_context_create refuses to run without a GPU
    if argv[0].endswith(".toml"):
        spec = ti.read_toml(argv[0])
    elif argv[0].endswith(".inp"):
        from .genesis_input import read_genesis

        spec = read_genesis(argv[0])
    else:
        print("usage: python -m opencafemol_b200 <input.toml> or <input.inp>", file=sys.stderr)
        return 2
    device = int(spec.platform.get("properties", {}).get("DeviceIndex", 0)) if isinstance(spec.platform, dict) else 0
    sim = Simulator(lib, spec, workdir=".", device=device)
    print(f"system of {spec.mass.size} beads, {sim.total_step} steps, a frame every {sim.save_step}", file=sys.stderr)
    sim.run()
    return 0
