"""CPU check of the FP32 bonded terms of the whole-step replica kernel (opencafemol_b200/csrc/replica_math.hpp):
the header compiles for the host, so its double instantiation is checked against finite differences of its own energy
and its float instantiation against the double one (tests/c/replica_math_check.cpp).  The GPU parity of the kernel that
uses it is tests/test_gpu_parity.py::test_whole_step_replica_kernel*."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fp32_bonded_terms_match_finite_differences_and_double(tmp_path):
    exe = tmp_path / "replica_math_check"
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", os.path.join(ROOT, "tests", "c", "replica_math_check.cpp"), "-o", str(exe)],
                   check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    m = re.search(r"max_fd_err (\S+) max_f32_err (\S+)", out)
    assert m, out
    # analytic gradients (double) vs central differences: 1e-5 of the term's largest force component
    assert float(m.group(1)) < 1e-5, out
    # float vs double, relative to the TERM's own largest force component (the force bar of the engine, 1e-5, is relative
    # to the total force on a bead, which is one to two orders of magnitude larger than a soft angle / dihedral term)
    assert float(m.group(2)) < 5e-4, out
    for kind, a in re.findall(r"kind (\d): f32 - f64 max abs (\S+)", out):
        if kind != "2":   # (the 12-10 contact is sampled into its clash region, forces of 1e9: relative bar only)
            assert float(a) < 2e-2, out
