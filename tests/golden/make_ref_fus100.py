"""Full-size reference fixture for config C2: sample/toml_input/FUS_100_HPS.toml (16 300 beads) evaluated by the
REFERENCE ITSELF (oracle/_ref/libocm_ref.so: the reference's reader, generators and expression strings, see
make_ref_golden.py), run in the build container where /root/reference is mounted:

    make -C oracle ref && python tests/golden/make_ref_fus100.py

The shipped coordinates overlap (the sample minimises before it runs), so the state is first relaxed: 200 L-BFGS
iterations plus 2 000 Langevin steps of the OpenMP oracle (any overlap-free state would do; the oracle only
produces the coordinates).  What is stored in ref_fus100.npz is the reference's answer AT those coordinates:
per-force-group energies and per-particle forces, plus the in-cutoff pair counts.  The GPU test
(tests/test_gpu_parity.py::test_fus100_reference_full_size) reads the sample-derived spec through the repo's
reader fixture ref_fus100_spec.json.gz and compares the engine with these numbers at 1e-5.
"""
import gzip
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_binding as rb  # noqa: E402
from opencafemol_b200 import toml_input as ti  # noqa: E402
from opencafemol_b200.capi import Library  # noqa: E402

SAMPLE = "/root/reference/sample/toml_input/FUS_100_HPS.toml"
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    assert rb.available(), "build oracle/_ref first: make -C oracle ref"
    spec = ti.read_toml(SAMPLE)
    ora = ti.make_context(Library(os.path.join(ROOT, "oracle", "libocm_oracle.so"), "oracle_"), spec)
    t0 = time.time()
    ora.minimize(10.0, 200)
    ora.set_velocities_to_temperature(spec.temperature, 2026)
    ora.step(2000)
    x = ora.get_state(positions=True)["positions"]
    print("relaxed in %.0f s" % (time.time() - t0))
    R = rb.RefLibrary()
    rs = R.open_toml(SAMPLE)
    rs.set_positions(x)
    t0 = time.time()
    eg = rs.energy_by_group()
    e, f = rs.energy(forces=True)
    print("reference evaluated in %.0f s: E = %.6f kJ/mol" % (time.time() - t0, e), eg)
    spec.positions = x
    spec.velocities = None
    with gzip.open(os.path.join(HERE, "ref_fus100_spec.json.gz"), "wt") as fh:
        json.dump(ti.spec_to_dict(spec), fh)
    np.savez_compressed(os.path.join(HERE, "ref_fus100.npz"), positions=x, forces=f, potential_energy_kj=e,
                        group_names=np.array(list(eg.keys())), group_energies_kj=np.array(list(eg.values())),
                        group_map=json.dumps(rs.group_map()))
    rs.close()


if __name__ == "__main__":
    main()
