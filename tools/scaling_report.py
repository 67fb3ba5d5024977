#!/usr/bin/env python
"""Throughput of the configurations of BASELINE.json (ns/day and MD steps/s vs bead count and GPU count).

One GPU:   python tools/scaling_report.py sweep      # HPS slabs of 1 630 / 16 300 / 130 400 / 1 043 200 beads
           python tools/scaling_report.py c1 c3      # SH3 (56 beads), nucleosome 2 x 3 grid (11 160 beads)
N GPUs:    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port 29520 \
               tools/scaling_report.py c4            # 512 SH3 replicas, 512/N per GPU, no collective per step
           (C5, the decomposed 1 043 200-bead condensate: tools/slab_run.py)
Each configuration prints one JSON line on rank 0.  Device time = CUDA events on the engine's stream, max over ranks.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def timed(ctx, warm, steps):
    ctx.step(warm)
    ctx.step(steps)
    return ctx.counter("last_step_ms") * 1e-3


def line(name, beads, steps, dev_s, dt_ps, world=1, units=1, extra=None):
    out = {"config": name, "n_gpus": world, "beads": int(beads), "md_steps": steps, "device_s": dev_s,
           "md_steps_per_s": steps / dev_s, "ns_per_day": units * steps * dt_ps * 1e-3 / dev_s * 86400.0,
           "bead_steps_per_s": units * beads * steps / dev_s if units == 1 else beads * steps / dev_s}
    if extra:
        out.update(extra)
    print(json.dumps(out), flush=True)


def main():
    import torch

    from conftest import load_golden_spec
    from opencafemol_b200 import synth, toml_input as ti
    from opencafemol_b200.capi import product_library

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    lib = product_library()
    modes = sys.argv[1:] or ["sweep"]

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    for mode in modes:
        if mode == "sweep" and rank == 0:
            for chains, tile, warm, steps in ((10, (1, 1), 4000, 20000), (100, (1, 1), 6000, 10000), (800, (8, 1), 1500, 2000), (6400, (8, 8), 300, 300)):
                spec = synth.hps_slab(n_chains=chains, tile=tile)
                ctx = ti.make_context(lib, spec, device=local)
                dev = timed(ctx, warm, steps)
                line("HPS slab sweep", spec.mass.size, steps, dev, spec.dt,
                     extra={"list_entries_per_step": ctx.counter("pair_evals_per_step"), "nstlist": ctx.counter("nstlist"),
                            "list_entries_per_s": ctx.counter("pair_evals_per_step") * steps / dev})
                ctx.close()
        elif mode == "c1" and rank == 0:
            spec = load_golden_spec("sh3")
            ctx = ti.make_context(lib, spec, device=local)
            dev = timed(ctx, 20000, 200000)
            line("C1 SH3 AICG2+ (sample/toml_input/sh3_AICG2+.toml)", spec.mass.size, 200000, dev, spec.dt)
            ctx.close()
        elif mode == "c3" and rank == 0:
            one = load_golden_spec("nucleosome")
            spec = synth.replicate(one, [[30.0 * i, 30.0 * j, 0.0] for i in range(2) for j in range(3)])
            ctx = ti.make_context(lib, spec, device=local)
            dev = timed(ctx, 2000, 10000)
            line("C3 nucleosome_1kx5 on a 2 x 3 grid (3SPN.2C DNA + AICG2+ protein + PDNS)", spec.mass.size, 10000, dev, spec.dt,
                 extra={"nstlist": ctx.counter("nstlist")})
            ctx.close()
        elif mode == "c4":
            from opencafemol_b200 import sharding

            total = 512
            first, per = sharding.replica_shard(total, world, rank)
            spec = load_golden_spec("sh3")
            ctx = ti.make_context(lib, spec, device=local, seed=spec.seed + first, replicas=per)
            steps = 20000
            ctx.step(5000)
            if world > 1:
                dist.barrier()
            ctx.step(steps)
            dev = max_over_ranks(ctx.counter("last_step_ms") * 1e-3)
            if rank == 0:
                line("C4 replica ensemble: 512 SH3 replicas, %d per GPU" % per, spec.mass.size, steps, dev, spec.dt, world=world, units=total,
                     extra={"replicas": total, "replica_steps_per_s": total * steps / dev, "collective_per_step": "none"})
            ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
