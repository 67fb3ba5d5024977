"""Golden vectors for the reference's external force-field generators and the combinatorial Go contact (PositionRestraint, Pulling,
EXVRectangularBox, CylindricalRestraint, HarmonicCoMPulling: [[forcefields.external]], ReadTOMLInput.cpp:763-870;
CombinatorialGoContact: [[forcefields.global]], :640-652) from the REFERENCE ITSELF, like make_ref_golden.py (run in the build
container after `make -C oracle ref`):

    python tests/golden/make_ref_external.py

No shipped sample uses these generators, so the inputs are the shipped samples with such tables added:
  * "external":     sample/toml_input/sh3_AICG2+.toml (56 beads, unlimited boundary) + all five external generators + a
                    CombinatorialGoContact between the two ends of the chain;
  * "external_pbc": the first 10 chains of sample/toml_input/FUS_100_HPS.toml (1 630 beads, periodic box) + the generators that
                    exist under a periodic boundary (CylindricalRestraint with periodicdistance, HarmonicCoMPulling and
                    CombinatorialGoContact with the minimum image, PositionRestraint, Pulling).
The reference's reader, generators and SystemGenerator run unmodified; the stand-in OpenMM evaluates their expression strings.
Writes tests/golden/ref_<name>.json.gz, <name>_spec.json.gz and ref_<name>_run.toml.gz."""
import gzip
import json
import os
import shutil
import sys
import tempfile
import tomllib

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_ref_golden as G  # noqa: E402
from make_ref_golden import rb, ti  # noqa: E402


def external_toml():
    with open(os.path.join(G.REF, "sh3_AICG2+.toml"), "rb") as fh:
        d = tomllib.load(fh)
    d["simulator"]["energy_minimization"] = False
    parts = d["systems"][0]["particles"]
    pos = [p["pos"] if "pos" in p else p["position"] for p in parts]
    lo = [min(p[a] for p in pos) - 2.6 for a in range(3)]
    hi = [max(p[a] for p in pos) + 2.3 for a in range(3)]
    n = len(parts)
    ff = d["forcefields"][0]
    ff["global"].append({
        "interaction": "CombinatorialGoContact", "potential": "CombinatorialGoContact", "cutoff": 1.8,
        "parameters": [{"indices_pair": [list(range(0, 10)), list(range(40, n))], "k": 0.31, "v0": 5.6},
                       {"indices_pair": [[12, 14, 16, 18], [30, 32, 34]], "k": "K2", "v0": 5.2}],
        "env": {"K2": 0.12}})
    ff["external"] = [
        {"interaction": "PullingForce",
         "parameters": [{"index": 0, "force": [0.9, -0.2, 0.1]}, {"index": 1, "offset": n - 2, "force": [-0.9, 0.2, -0.1]}]},
        {"interaction": "PositionRestraint", "potential": "Harmonic",
         "parameters": [{"index": i, "position": [pos[i][0] + 1.5, pos[i][1] - 1.0, pos[i][2] + 0.5], "k": 0.4 + 0.1 * k, "v0": 0.8 * k}
                        for k, i in enumerate((10, 20, 30))]},
        {"interaction": "CylindricalRestraint", "potential": "Harmonic",
         "parameters": [{"index": i, "axis": [1.0 + 0.2 * k, 0.5, -0.3 * k], "shift": [2.0, 1.0 * k, 4.0], "k": 0.05 + 0.02 * k, "v0": "R0"}
                        for k, i in enumerate((3, 17, 25, 44, 51))],
         "env": {"R0": 6.0}},
        {"interaction": "RectangularBox", "potential": "ExcludedVolumeWall", "box": {"lower": lo, "upper": hi}, "epsilon": 0.6,
         "parameters": [{"index": i, "radius": 2.0 + 0.05 * (i % 5)} for i in range(n)]},
        {"interaction": "CoMPullingForce", "potential": "Harmonic",
         "parameters": [{"k": 35.0, "v0": 1.2, "indices_pair": [list(range(0, 10)), list(range(46, n))]},
                        {"k": 12.0, "v0": 0.4, "indices_pair": [list(range(20, 30, 2)), list(range(21, 31, 2))]}]},
    ]
    return G.dump_toml(d)


def external_pbc_toml():
    d = tomllib.loads(G.fus_subset(10))
    parts = d["systems"][0]["particles"]
    # the shipped coordinates overlap: start from the minimised ones of the fus10 fixture (make_ref_golden.py)
    with gzip.open(os.path.join(HERE, "fus10_spec.json.gz"), "rt") as fh:
        x0 = json.load(fh)["positions"]
    for p, xyz in zip(parts, x0):
        p.pop("position", None)
        p["pos"] = [float(v) * 10.0 for v in xyz]
    d["systems"][0]["vel_seed"] = 2027
    pos = [p["pos"] for p in parts]
    L = 163
    ff = d["forcefields"][0]
    ff["global"].append({
        "interaction": "CombinatorialGoContact", "potential": "CombinatorialGoContact", "cutoff": 2.0,
        "parameters": [{"indices_pair": [list(range(2 * L, 3 * L, 3)), list(range(2 * L + 1, 3 * L, 3))], "k": 0.25, "v0": 4.4}]})
    ff["external"] = [
        {"interaction": "PullingForce", "parameters": [{"index": 5 * L, "force": [0.0, 0.0, 0.7]}]},
        {"interaction": "PositionRestraint", "potential": "Harmonic",
         "parameters": [{"index": 6 * L + 7 * k, "position": [pos[6 * L + 7 * k][0] + 2.0, pos[6 * L + 7 * k][1], pos[6 * L + 7 * k][2] - 1.0], "k": 0.3, "v0": 0.0}
                        for k in range(4)]},
        # the axis points along z through a line near the box edge: for most beads the nearest image of the line is not the line itself
        {"interaction": "CylindricalRestraint", "potential": "Harmonic",
         "parameters": [{"index": i, "axis": [0.0, 0.0, 2.0], "shift": [3.0, 170.0, 0.0], "k": 0.02, "v0": 20.0} for i in range(0, L, 5)]},
        {"interaction": "CoMPullingForce", "potential": "Harmonic",
         "parameters": [{"k": 20.0, "v0": 3.0, "indices_pair": [list(range(7 * L, 8 * L)), list(range(8 * L, 9 * L))]}]},
    ]
    return G.dump_toml(d)


def generate(R, tmp, name, text, steps):
    path = os.path.join(tmp, name + ".toml")
    with open(path, "w") as fh:
        fh.write(text)
    rs = R.open_toml(path)
    spec = ti.read_toml(path)
    with gzip.open(os.path.join(HERE, f"{name}_spec.json.gz"), "wt") as fh:
        json.dump(ti.spec_to_dict(spec), fh)
    out = {"source": name, "n": rs.n, "group_map": rs.group_map(), "integrator": rs.integrator()}
    out["positions"] = rs.positions().tolist()
    out["energy_by_group_kj"] = rs.energy_by_group()
    e, f = rs.energy(forces=True)
    out["potential_energy_kj"] = e
    out["forces"] = f.tolist()
    out["force_digest"] = G.force_digest(rs.describe())
    vs = 2027
    rs.set_velocities_to_temperature(spec.temperature, vs)
    out["vel_seed_used"] = vs
    out["kinetic_energy_kj"] = rs.kinetic_energy()
    rs.step(steps)
    out["trajectory"] = {"steps": steps, "positions": rs.positions().tolist(), "velocities": rs.velocities().tolist()}
    rs.close()
    files = {}
    for fmt in ("pdb", "dcd"):
        wd = os.path.join(tmp, f"run_{name}_{fmt}")
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
    print(name, "N", out["n"], {k: round(v / 4.184, 6) for k, v in out["energy_by_group_kj"].items()})


def main():
    assert rb.available(), "build oracle/_ref first: make -C oracle ref"
    R = rb.RefLibrary()
    tmp = tempfile.mkdtemp(prefix="refext_")
    generate(R, tmp, "external", external_toml(), 20)
    generate(R, tmp, "external_pbc", external_pbc_toml(), 20)
    shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
