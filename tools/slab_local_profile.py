"""Two slab ranks as host threads on ONE GPU (in-process transport), or the single-domain system of the same per-rank size:
the workload behind the per-kernel ncu comparisons of profiles/r01d_summary.md (slab rank vs single domain).
    ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:force_kernel -s 60 -c 1 python tools/slab_local_profile.py dd|single"""
import sys, threading
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from opencafemol_b200 import synth, toml_input as ti
from opencafemol_b200.capi import Comm, product_library
from test_gpu_slab import SlabRun
lib = product_library()
mode = sys.argv[1]
if mode == "dd":
    spec = synth.hps_slab(n_chains=1600, tile=(8, 2))
    run = SlabRun(lib, spec, 2, 1, nstlist=28)
    run.each(lambda c: c.step(280))
    run.each(lambda c: c.step(56))
    print("dd done", [c.counter("owned_beads") for c in run.ctx])
else:
    spec = synth.hps_slab(n_chains=800, tile=(8, 1))
    ctx = ti.make_context(lib, spec)
    ctx.set_option("nstlist", 28)
    ctx.set_option("phase_timing", 1)
    ctx.step(280); ctx.step(56)
    print("single done")
