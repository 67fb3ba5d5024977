"""Golden vectors for the Monte Carlo barostats from the REFERENCE'S OWN PROGRAM, like make_ref_membrane.py
(run in the build container after `make -C oracle ref`):

    python tests/golden/make_ref_barostat.py

Inputs: the two lipid-membrane samples as shipped, ensembles included — sample/toml_input/POPC_flat_membrane_128.toml (NPT,
MonteCarloAnisotropicBarostat scaling X and Y) and POPC_flat_membrane_128_MembraneBarostat.toml (NPgammaT,
MonteCarloMembraneBarostat, isotropic XY, fixed Z) — with fixed seeds, lattice coordinates jittered by 0.3 A (on the perfect
lattice every force is zero) and, for the NPT case, a trial every 5 steps so that a 40-step program run holds 8 of them.

What is pinned and what is not: the reference's reader, generators, SystemGenerator, Simulator and Observers run unmodified
(MonteCarloAnisotropicBarostatGenerator.cpp:9-17, MonteCarloMembraneBarostatGenerator.cpp:6-14, SystemGenerator.cpp:89-108,
Observer.cpp:190-213 for the unit cell of every DCD frame).  OpenMM 8.0.0 itself is not under /root/reference: the barostat
behind those calls is oracle/ref_shim's restatement of its published algorithm, with SplitMix64 of the integrator seed as the
random stream.  So these fixtures pin the engine, the oracle and the host layer to ONE independent restatement driven by the
reference's own call sites; they cannot pin that restatement to OpenMM's binaries (oracle header: parity unpinned for this part).

Writes tests/golden/ref_membrane_npt.json.gz, ref_membrane_npgt.json.gz, the two *_spec.json.gz and *_run.toml.gz."""
import gzip
import json
import os
import shutil
import sys
import tempfile
import tomllib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_ref_golden as G  # noqa: E402
from make_ref_golden import rb, ti  # noqa: E402

CASES = {"membrane_npt": ("POPC_flat_membrane_128.toml", 5), "membrane_npgt": ("POPC_flat_membrane_128_MembraneBarostat.toml", None)}


def case_toml(sample, frequency):
    with open(os.path.join(G.REF, sample), "rb") as fh:
        d = tomllib.load(fh)
    if frequency is not None:
        d["systems"][0]["attributes"]["ensemble"]["frequency"] = frequency
    d["simulator"]["total_step"] = 40
    d["simulator"]["save_step"] = 10
    d["simulator"]["seed"] = 90210
    d["systems"][0]["vel_seed"] = 2027
    d["platform"] = {"name": "Reference"}
    rng = np.random.default_rng(12)
    for p in d["systems"][0]["particles"]:
        key = "position" if "position" in p else "pos"
        p[key] = [float(v + rng.normal(scale=0.3)) for v in p[key]]
    return G.dump_toml(d)


def main():
    assert rb.available(), "build oracle/_ref first: make -C oracle ref"
    R = rb.RefLibrary()
    for name, (sample, freq) in CASES.items():
        tmp = tempfile.mkdtemp(prefix="refbaro_")
        text = case_toml(sample, freq)
        path = os.path.join(tmp, name + ".toml")
        with open(path, "w") as fh:
            fh.write(text)
        spec = ti.read_toml(path)
        with gzip.open(os.path.join(HERE, f"{name}_spec.json.gz"), "wt") as fh:
            json.dump(ti.spec_to_dict(spec), fh)
        # the library view of the same run: the reference pipeline up to its Context (ref_open_toml) stepped from outside, box and energies after every
        # window of `frequency` steps (each holds exactly one trial)
        rs = R.open_toml(path)
        out = {"source": sample, "n": rs.n, "group_map": rs.group_map(), "integrator": rs.integrator(), "ensemble": spec.ensemble}
        rs.set_velocities_to_temperature(spec.temperature, 2027)
        window = spec.ensemble["frequency"]
        boxes, epots = [rs.box().tolist()], [rs.energy()]
        for _ in range(12):
            rs.step(window)
            boxes.append(rs.box().tolist())
            epots.append(rs.energy())
        out["window"] = window
        out["box_nm"] = boxes
        out["potential_energy_kj"] = epots
        out["final_positions"] = rs.positions().tolist()
        rs.close()
        files = {}
        for fmt in ("pdb", "dcd"):
            wd = os.path.join(tmp, "run_" + fmt)
            os.makedirs(wd, exist_ok=True)
            got, used = G.run_program(text, wd, fmt)
            for fn, b in got.items():
                if fn.endswith((".ene", ".pdb")):
                    files[fn] = {"text": b.decode()} if len(b) < 200000 else {"sha256": G.hashlib.sha256(b).hexdigest(), "head": b[:2000].decode(), "size": len(b)}
                else:
                    files[fn] = {"sha256": G.hashlib.sha256(b).hexdigest(), "size": len(b), "head_hex": b[:400].hex()}
                    if len(b) < 65536:
                        files[fn]["hex"] = b.hex()
            if fmt == "dcd":
                with gzip.open(os.path.join(HERE, f"ref_{name}_run.toml.gz"), "wt") as fh:
                    fh.write(used)
        out["program_files"] = files
        with gzip.open(os.path.join(HERE, f"ref_{name}.json.gz"), "wt") as fh:
            json.dump(out, fh)
        print(name, "N", out["n"], "box", boxes[0], "->", boxes[-1], "distinct boxes", len({tuple(b) for b in boxes}))
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
