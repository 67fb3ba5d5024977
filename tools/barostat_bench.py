"""NPT / NPgammaT throughput of the lipid-membrane fixtures against the same system in NVT (one GPU):
    python tools/barostat_bench.py [steps]
prints MD steps/s with and without the Monte Carlo barostat, the trial count and the acceptance."""
import gzip
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from opencafemol_b200 import toml_input as ti  # noqa: E402
from opencafemol_b200.capi import product_library  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    lib = product_library()
    for name in ("membrane_npt", "membrane_npgt"):
        with gzip.open(os.path.join(ROOT, "tests", "golden", name + "_spec.json.gz"), "rt") as fh:
            spec = ti.spec_from_dict(json.load(fh))
        for ens in (None, spec.ensemble):
            s = ti.spec_from_dict(ti.spec_to_dict(spec))
            s.ensemble = ens
            ctx = ti.make_context(lib, s)
            ctx.step(2000)
            t0 = time.perf_counter()
            ctx.step(steps)
            dt = time.perf_counter() - t0
            info = ctx.barostat_info()
            print(f"{name:14s} {'NVT' if ens is None else ens['type']:9s} {steps / dt:10.0f} MD steps/s  trials {info['trials']:5d} accepted {info['accepted']:5d} "
                  f"box {ctx.box().round(4).tolist()}", flush=True)


if __name__ == "__main__":
    main()
