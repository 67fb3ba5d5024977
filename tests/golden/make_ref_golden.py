"""Golden vectors from the REFERENCE ITSELF (run in the build container, where /root/reference is mounted):

    make -C oracle ref && python tests/golden/make_ref_golden.py

oracle/_ref/libocm_ref.so is the reference's own reader, Topology, force-field generators, SystemGenerator,
Simulator and Observers compiled unmodified against the stand-in OpenMM / toml11 headers of oracle/ref_shim
(see oracle/Makefile).  For each sample input this script stores, under tests/golden/ref_<name>.json.gz:

  * per-force-group energies (kJ/mol), forces, kinetic energy at the sample's coordinates (FUS: after an energy
    minimisation, because the shipped coordinates overlap);
  * a digest of every OpenMM::Force the reference built (class, group, counts, cutoff, exclusion hash);
  * a short Langevin trajectory (shared Philox noise) for the small systems;
  * the files the reference's own main program writes (.ene lines, PDB head, DCD bytes) for a short run.

The fixtures travel to the GPU box; /root/reference does not.
"""
import gzip
import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile
import tomllib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_binding as rb  # noqa: E402
from opencafemol_b200 import toml_input as ti  # noqa: E402

REF = "/root/reference/sample/toml_input"
HERE = os.path.dirname(os.path.abspath(__file__))


# ---- a minimal TOML writer (enough for the sample inputs) ---------------------------------------------
def _inline(v):
    if isinstance(v, bool):
        return "true" if v else "false"
    if isinstance(v, (int, float)):
        return repr(v)
    if isinstance(v, str):
        return json.dumps(v)
    if isinstance(v, list):
        return "[" + ", ".join(_inline(x) for x in v) + "]"
    if isinstance(v, dict):
        return "{" + ", ".join(f"{json.dumps(k)} = {_inline(x)}" for k, x in v.items()) + "}"
    raise TypeError(type(v))


def dump_toml(d, prefix=""):
    out = []
    for k, v in d.items():
        if not isinstance(v, dict) and not (isinstance(v, list) and v and all(isinstance(x, dict) for x in v) and k in ("systems", "forcefields", "local", "global")):
            if isinstance(v, list) and v and all(isinstance(x, dict) for x in v):
                out.append(f"{k} = [\n" + "".join("  " + _inline(x) + ",\n" for x in v) + "]")
            else:
                out.append(f"{k} = {_inline(v)}")
    for k, v in d.items():
        if isinstance(v, dict):
            out.append(f"[{prefix}{k}]")
            out.append(dump_toml(v, prefix + k + "."))
        elif isinstance(v, list) and v and all(isinstance(x, dict) for x in v) and k in ("systems", "forcefields", "local", "global"):
            for x in v:
                out.append(f"[[{prefix}{k}]]")
                out.append(dump_toml(x, prefix + k + "."))
    return "\n".join(out)


def fus_subset(n_chains=10, chain_len=163):
    """FUS_100_HPS.toml restricted to its first n_chains chains (same box, same parameters)."""
    with open(os.path.join(REF, "FUS_100_HPS.toml"), "rb") as fh:
        d = tomllib.load(fh)
    n = n_chains * chain_len
    d["systems"][0]["particles"] = d["systems"][0]["particles"][:n]
    ff = d["forcefields"][0]
    for key in ("local", "global"):
        for f in ff.get(key, []):
            keep = []
            for p in f["parameters"]:
                idx = p.get("indices", p.get("index"))
                idx = idx if isinstance(idx, list) else [idx]
                if all(i < n for i in idx):
                    keep.append(p)
            f["parameters"] = keep
    d["simulator"]["energy_minimization"] = False
    return dump_toml(d)


def force_digest(desc):
    out = []
    for f in desc["forces"]:
        e = {"class": f["class"], "group": f["group"]}
        for key in ("bonds", "angles", "torsions", "particles", "donors", "acceptors"):
            if key in f:
                e["n_" + key] = len(f[key])
        if "cutoff" in f:
            e["cutoff"] = f["cutoff"]
            e["method"] = f["method"]
        if "use_switch" in f:
            e["use_switch"] = f["use_switch"]
            e["switch_distance"] = f["switch_distance"]
        if "exclusions" in f:
            ex = sorted((min(a, b), max(a, b)) for a, b in f["exclusions"]) if f["class"] == "CustomNonbondedForce" else sorted(map(tuple, f["exclusions"]))
            e["n_exclusions"] = len(ex)
            e["exclusions_sha1"] = hashlib.sha1(np.asarray(ex, np.int32).tobytes()).hexdigest()
        if "interaction_groups" in f:
            e["interaction_groups"] = [[len(a), len(b), hashlib.sha1(np.asarray(a + b, np.int32).tobytes()).hexdigest()] for a, b in f["interaction_groups"]]
        if "expression" in f:
            e["expression"] = f["expression"]
            e["globals"] = f["globals"]
        out.append(e)
    return out


def run_program(toml_text, workdir, fmt):
    """the reference's main program on `toml_text`; returns the bytes of the files it wrote"""
    d = tomllib.loads(toml_text)
    d["files"]["output"]["format"] = fmt
    d["files"]["output"]["path"] = "./out/"
    d["files"]["output"]["prefix"] = "run"
    d["simulator"]["energy_minimization"] = False
    d["simulator"]["total_step"] = 40
    d["simulator"]["save_step"] = 10
    os.makedirs(os.path.join(workdir, "out"), exist_ok=True)
    with open(os.path.join(workdir, "in.toml"), "w") as fh:
        fh.write(dump_toml(d))
    subprocess.run([rb.REF_BIN, "in.toml"], cwd=workdir, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    files = {}
    for fn in sorted(os.listdir(os.path.join(workdir, "out"))):
        with open(os.path.join(workdir, "out", fn), "rb") as fh:
            files[fn] = fh.read()
    return files, dump_toml(d)


def main():
    assert rb.available(), "build oracle/_ref first: make -C oracle ref"
    R = rb.RefLibrary()
    tmp = tempfile.mkdtemp(prefix="refgold_")
    samples = {"sh3": open(os.path.join(REF, "sh3_AICG2+.toml")).read(),
               "dna_3spn2": open(os.path.join(REF, "dna_3SPN2.toml")).read(),
               "nucleosome": open(os.path.join(REF, "nucleosome_1kx5.toml")).read(),
               "fus10": fus_subset(10), "wcamix": wcamix_toml(),
               # 3SPN2BasePairLocal / 3SPN2CrossStackingLocal (explicit nucleotide pairs as compound bonds)
               "dna_localize": open(os.path.join(REF, "dna_3SPN2_localize.toml")).read()}
    for name, text in samples.items():
        path = os.path.join(tmp, name + ".toml")
        with open(path, "w") as fh:
            fh.write(text)
        rs = R.open_toml(path)
        # (the name-pair table forms and the *Local forms are read by the reference's reader only)
        spec = ti.read_toml(path) if name not in ("wcamix", "dna_localize") else None
        out = {"source": name, "n": rs.n, "group_map": rs.group_map(), "integrator": rs.integrator()}
        if name == "fus10":
            rs.minimize(tolerance=10.0, max_iterations=400)
            # the spec fixture for this case is written here (the other samples have theirs from make_golden.py)
            spec.positions = rs.positions()
            with gzip.open(os.path.join(HERE, "fus10_spec.json.gz"), "wt") as fh:
                json.dump(ti.spec_to_dict(spec), fh)
        x0 = rs.positions()
        out["positions"] = x0.tolist()
        out["energy_by_group_kj"] = rs.energy_by_group()
        e, f = rs.energy(forces=True)
        out["potential_energy_kj"] = e
        out["forces"] = f.tolist()
        out["force_digest"] = force_digest(rs.describe())
        # short Langevin run with the documented Philox noise (same streams as oracle/ocm_oracle.cpp and the engine)
        if spec is not None and spec.velocities is not None:
            rs.set_velocities(spec.velocities)
        else:
            rs.set_velocities_to_temperature(300.0 if spec is None else spec.temperature, 2026)
        out["vel_seed_used"] = None if (spec is not None and spec.velocities is not None) else 2026
        out["kinetic_energy_kj"] = rs.kinetic_energy()
        nst = 20
        rs.step(nst)
        out["trajectory"] = {"steps": nst, "positions": rs.positions().tolist(), "velocities": rs.velocities().tolist()}
        rs.close()
        # what the reference's own main program writes
        if name in ("sh3", "dna_3spn2", "fus10", "wcamix", "dna_localize"):
            files = {}
            for fmt in ("pdb", "dcd"):
                wd = os.path.join(tmp, name + "_" + fmt)
                os.makedirs(wd, exist_ok=True)
                if name == "fus10":
                    # start the program from the minimised coordinates
                    d = tomllib.loads(text)
                    for p, xyz in zip(d["systems"][0]["particles"], x0 * 10.0):
                        p["pos"] = [float(v) for v in xyz]
                    d["systems"][0]["vel_seed"] = 2026  # the sample has neither velocities nor a seed: make the run reproducible
                    text_run = dump_toml(d)
                elif spec is None:
                    d = tomllib.loads(text)
                    d["systems"][0].setdefault("vel_seed", 2026)
                    text_run = dump_toml(d)
                elif spec.velocities is None:
                    # neither velocities nor a velocity seed in the sample: make the program run reproducible
                    d = tomllib.loads(text)
                    d["systems"][0]["vel_seed"] = 2026
                    text_run = dump_toml(d)
                else:
                    text_run = text
                got, used = run_program(text_run, wd, fmt)
                for fn, b in got.items():
                    if fn.endswith((".ene", ".pdb")):
                        files[fn] = {"text": b.decode()} if len(b) < 200000 else {"sha256": hashlib.sha256(b).hexdigest(), "head": b[:2000].decode(), "size": len(b)}
                    else:
                        files[fn] = {"sha256": hashlib.sha256(b).hexdigest(), "size": len(b), "head_hex": b[:400].hex()}
                        if len(b) < 65536:
                            files[fn]["hex"] = b.hex()
                if fmt == "dcd":
                    with gzip.open(os.path.join(HERE, f"ref_{name}_run.toml.gz"), "wt") as fh:
                        fh.write(used)
            out["program_files"] = files
        with gzip.open(os.path.join(HERE, f"ref_{name}.json.gz"), "wt") as fh:
            json.dump(out, fh)
        print(name, "N", out["n"], {k: round(v / 4.184, 6) for k, v in out["energy_by_group_kj"].items()})
    genesis_fixture(tmp)
    shutil.rmtree(tmp, ignore_errors=True)


def wcamix_toml():
    """a small synthetic system for the potentials of SURVEY.md 8f rank 4: per-particle WCA, name-pair table WCA
    (UniformWeeksChandlerAndersen), name-pair table LennardJonesAttractive (UniformLennardJonesAttractive), capped Go
    contacts; 4 chains of 12 beads, unlimited boundary"""
    rng = np.random.default_rng(7)
    nch, L = 4, 12
    n = nch * L
    grid = np.stack(np.meshgrid(np.arange(4), np.arange(4), np.arange(3), indexing="ij"), -1).reshape(-1, 3)
    pos = grid * 5.6 + rng.uniform(-0.5, 0.5, (n, 3))
    order = np.lexsort((grid[:, 0] if False else grid[:, 2], grid[:, 1], grid[:, 0]))
    pos = pos[order]
    names = ["A" if (i // 3) % 2 == 0 else "B" for i in range(n)]
    particles = [{"m": 100.0 + i % 5, "pos": [float(v) for v in pos[i]], "name": names[i], "group": "G%d" % (i // L)} for i in range(n)]
    bonds = [{"indices": [c * L + k, c * L + k + 1], "v0": float(np.linalg.norm(pos[c * L + k] - pos[c * L + k + 1])), "k": 5.0}
             for c in range(nch) for k in range(L - 1)]
    d = {
        "files": {"output": {"prefix": "run", "path": "./out/", "format": "dcd"}},
        "units": {"length": "angstrom", "energy": "kcal/mol"},
        "simulator": {"type": "MolecularDynamics", "boundary_type": "Unlimited", "precision": "double", "delta_t": 0.1, "total_step": 40,
                      "save_step": 10, "seed": 777, "integrator": {"type": "BAOABLangevin", "gammas": []}},
        "systems": [{"attributes": {"temperature": 300.0}, "vel_seed": 2026, "particles": particles}],
        "forcefields": [{
            "local": [{"interaction": "BondLength", "potential": "Harmonic", "topology": "bond", "parameters": bonds},
                      # contacts between chains: with v0 = 7.5 or 9.5 A and neighbours 5.6 A apart, some sit on the linear cap (r/v0 <= 0.75)
                      {"interaction": "BondLength", "potential": "CappedGoContact", "topology": "contact", "capping_ratio": 0.75,
                       "parameters": [{"indices": [i, i + L], "v0": 7.5 + 2.0 * (i % 2), "k": 0.3} for i in range(0, n - L, 2)]}],
            "global": [
                {"interaction": "Pair", "potential": "WCA", "ignore": {"particles_within": {"bond": 1}},
                 "parameters": [{"index": i, "sigma": 4.6 + 0.2 * (i % 4), "epsilon": 0.3 + 0.05 * (i % 3)} for i in range(0, n, 2)]},
                {"interaction": "Pair", "potential": "WCA", "ignore": {"particles_within": {"bond": 2}},
                 "table": {"A": {"A": {"sigma": 5.2, "epsilon": 0.4}, "B": {"sigma": 4.8, "epsilon": 0.6}}, "B": {"B": {"sigma": 5.0, "epsilon": 0.5}}},
                 "parameters": [{"index": i, "name": names[i]} for i in range(n)]},
                {"interaction": "Pair", "potential": "LennardJonesAttractive", "cutoff": 2.5, "ignore": {"particles_within": {"bond": 1}},
                 "table": {"A": {"A": {"sigma": 5.0, "epsilon": 0.25}, "B": {"sigma": 5.4, "epsilon": 0.15}}, "B": {"B": {"sigma": 4.9, "epsilon": 0.3}}},
                 "parameters": [{"index": i, "name": names[i]} for i in range(n)]},
            ]}],
    }
    return dump_toml(d)


def genesis_fixture(tmp):
    """the GENESIS sample (sample/genesis_input/1SRL_cg.*, run shortened to 40 steps) as an input fixture, and the files
    the reference's own program writes for it"""
    import tarfile

    src = "/root/reference/sample/genesis_input"
    work = os.path.join(tmp, "genesis_1SRL")
    os.makedirs(work)
    for f in ("1SRL_cg.gro", "1SRL_cg.inp", "1SRL_cg.itp", "1SRL_cg.top"):
        shutil.copy(os.path.join(src, f), work)
    shutil.copytree(os.path.join(src, "param"), os.path.join(work, "param"))
    inp = os.path.join(work, "1SRL_cg.inp")
    t = open(inp).read()
    for a, b in (("nsteps              = 1000", "nsteps              = 40"), ("eneout_period       = 100", "eneout_period       = 10"),
                 ("crdout_period       = 100", "crdout_period       = 10"), ("rstout_period       = 100", "rstout_period       = 10")):
        assert a in t
        t = t.replace(a, b)
    open(inp, "w").write(t)
    with tarfile.open(os.path.join(HERE, "genesis_1SRL_input.tar.gz"), "w:gz") as tf:
        tf.add(work, arcname="genesis_1SRL")
    os.makedirs(os.path.join(work, "output"))
    subprocess.run([rb.REF_BIN, "1SRL_cg.inp"], cwd=work, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    out = {}
    for fn in sorted(os.listdir(os.path.join(work, "output"))):
        b = open(os.path.join(work, "output", fn), "rb").read()
        out[fn] = {"text": b.decode()} if fn.endswith((".ene", ".pdb")) else {"hex": b.hex(), "size": len(b)}
    with gzip.open(os.path.join(HERE, "ref_genesis_1SRL.json.gz"), "wt") as fh:
        json.dump(out, fh)
    print("genesis_1SRL", sorted(out))


if __name__ == "__main__":
    main()
