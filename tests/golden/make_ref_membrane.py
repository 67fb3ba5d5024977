"""Golden vectors for the attractive-well pair potentials of the lipid model (iSoLFAttractive,
Trigonometric) from the REFERENCE ITSELF, like make_ref_golden.py (run in the build container after `make -C oracle ref`):

    python tests/golden/make_ref_membrane.py

Input: sample/toml_input/POPC_flat_membrane_128.toml (128 lipids, 640 beads, periodic box; harmonic bonds and angles, per-particle
WCA, iSoLFAttractive) with three changes — the ensemble is NVT (the Monte Carlo barostats are out of scope), the lattice
coordinates are jittered by 0.3 A (on the perfect lattice every force is zero), and a Trigonometric term on the head beads is
added so that the reference's own expression string of that generator is evaluated too (unequal sigmas: its cutoff then lies
inside the well of the largest pair, as the generator's cutoff rule has it).
Writes tests/golden/ref_membrane.json.gz, membrane_spec.json.gz and ref_membrane_run.toml.gz."""
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


def membrane_toml():
    with open(os.path.join(G.REF, "POPC_flat_membrane_128.toml"), "rb") as fh:
        d = tomllib.load(fh)
    d["systems"][0]["attributes"].pop("ensemble", None)
    d["simulator"]["total_step"] = 40
    d["simulator"]["save_step"] = 10
    d["simulator"]["seed"] = 4242
    d["systems"][0]["vel_seed"] = 2026
    d["platform"] = {"name": "Reference"}
    rng = np.random.default_rng(11)
    parts = d["systems"][0]["particles"]
    for p in parts:
        key = "position" if "position" in p else "pos"
        p[key] = [float(v + rng.normal(scale=0.3)) for v in p[key]]
    heads = [i for i, p in enumerate(parts) if str(p.get("group", "")).startswith("H")]
    d["forcefields"][0]["global"].append({
        "interaction": "Pair", "potential": "Trigonometric", "ignore": {"particles_within": {"bond": 1, "angle": 1}},
        "parameters": [{"index": i, "epsilon": 0.2 + 0.05 * (k % 3), "sigma": 6.0 + 0.4 * (k % 4), "omega": 3.0 + 0.5 * (k % 2)}
                       for k, i in enumerate(heads)]})
    return G.dump_toml(d)


def main():
    assert rb.available(), "build oracle/_ref first: make -C oracle ref"
    R = rb.RefLibrary()
    tmp = tempfile.mkdtemp(prefix="refmemb_")
    text = membrane_toml()
    path = os.path.join(tmp, "membrane.toml")
    with open(path, "w") as fh:
        fh.write(text)
    rs = R.open_toml(path)
    spec = ti.read_toml(path)
    with gzip.open(os.path.join(HERE, "membrane_spec.json.gz"), "wt") as fh:
        json.dump(ti.spec_to_dict(spec), fh)
    out = {"source": "membrane", "n": rs.n, "group_map": rs.group_map(), "integrator": rs.integrator()}
    out["positions"] = rs.positions().tolist()
    out["energy_by_group_kj"] = rs.energy_by_group()
    e, f = rs.energy(forces=True)
    out["potential_energy_kj"] = e
    out["forces"] = f.tolist()
    out["force_digest"] = G.force_digest(rs.describe())
    rs.set_velocities_to_temperature(spec.temperature, 2026)
    out["vel_seed_used"] = 2026
    out["kinetic_energy_kj"] = rs.kinetic_energy()
    rs.step(20)
    out["trajectory"] = {"steps": 20, "positions": rs.positions().tolist(), "velocities": rs.velocities().tolist()}
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
            with gzip.open(os.path.join(HERE, "ref_membrane_run.toml.gz"), "wt") as fh:
                fh.write(used)
    out["program_files"] = files
    with gzip.open(os.path.join(HERE, "ref_membrane.json.gz"), "wt") as fh:
        json.dump(out, fh)
    print("membrane N", out["n"], {k: round(v / 4.184, 6) for k, v in out["energy_by_group_kj"].items()})
    shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
