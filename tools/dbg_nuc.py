import sys, traceback
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from conftest import load_golden_spec
from opencafemol_b200 import toml_input as ti
from opencafemol_b200.capi import product_library
lib=product_library()
spec=load_golden_spec("nucleosome")
try:
    ctx=ti.make_context(lib,spec)
    st=ctx.get_state(forces=True,energy=True)
    print("ok", st["potential_energy"])
    ctx.step(20)
    print("stepped", ctx.counter("nstlist"))
except Exception as e:
    traceback.print_exc()
