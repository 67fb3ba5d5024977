"""One launch of the whole-step replica kernel for an ncu capture: python tools/replica_profile_run.py [replicas] [steps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_golden_spec  # noqa: E402

from opencafemol_b200 import capi, toml_input  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 512
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
ctx = toml_input.make_context(capi.product_library(), load_golden_spec("sh3"), replicas=reps if reps > 1 else None)
ctx.set_option("replica_kernel", 1)
ctx.step(steps)
ctx.step(steps)
print("done", ctx.counter("replica_kernel_launches"))
