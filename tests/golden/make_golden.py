"""Regenerates the golden fixtures from the reference's own sample inputs (run in the build container,
where /root/reference is mounted):

    python tests/golden/make_golden.py

For each sample: <name>_spec.json.gz = the parsed term tables (what the reader hands to the C ABI) and
<name>_oracle.json = oracle energies per force group, forces, pair sets at the sample's coordinates.
The closed-form known answers of SURVEY.md Appendix D are asserted here and again in tests/test_oracle_kat.py.
"""
import gzip
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from opencafemol_b200 import toml_input as ti  # noqa: E402
from opencafemol_b200.capi import Library  # noqa: E402

REF = "/root/reference/sample/toml_input"
SAMPLES = {"sh3": "sh3_AICG2+.toml", "dna_3spn2": "dna_3SPN2.toml", "nucleosome": "nucleosome_1kx5.toml"}
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    lib = Library(os.path.join(ROOT, "oracle", "libocm_oracle.so"), "oracle_")
    for name, fn in SAMPLES.items():
        spec = ti.read_toml(os.path.join(REF, fn))
        with gzip.open(os.path.join(HERE, name + "_spec.json.gz"), "wt") as fh:
            json.dump(ti.spec_to_dict(spec), fh)
        ctx = ti.make_context(lib, spec)
        st = ctx.get_state(forces=True, energy=True)
        out = {
            "source": fn,
            "energy_by_group_kj": {k: st["energy_by_group"][g] for k, g in spec.group_ids.items()},
            "kinetic_energy_kj": st["kinetic_energy"],
            "forces": np.round(st["forces"], 7).tolist(),
            "pairs": {str(t): ctx.neighbor_pairs(t).tolist() for t in range(ctx.num_pair_terms())} if spec.mass.size < 500 else {},
            "da_pairs": {str(t): ctx.da_pairs(t).tolist() for t in range(ctx.num_da_terms())},
        }
        with open(os.path.join(HERE, name + "_oracle.json"), "w") as fh:
            json.dump(out, fh)
        print(name, {k: round(v / 4.184, 6) for k, v in out["energy_by_group_kj"].items()})


if __name__ == "__main__":
    main()
