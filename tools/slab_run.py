#!/usr/bin/env python
"""One rank of a slab-decomposed run (config C5 and smaller), one process per GPU under torchrun:

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/slab_run.py --tile 8 8 --chains 6400 --steps 2000 [--check] [--equilibrated]

Prints one JSON line on rank 0: ns/day, MD steps/s, per-rank owned beads.  --check compares positions after the run
with a single-domain run of the same system on rank 0's GPU (bit-exact with a fixed rebuild interval).
--equilibrated starts from the committed compacted C2 state tiled --tile times (synth.c2_equilibrated) instead of the
serpentine start, which is still compacting during a short run.
With WORLD_SIZE unset it runs the single-domain engine (the 1-GPU point of the scaling curve)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def make_comm(lib, rank, world, local):
    """NCCL communicator of the engine: rank 0 draws the unique id, torch.distributed carries it to the others."""
    import torch
    import torch.distributed as dist

    from opencafemol_b200.capi import Comm

    uid = torch.zeros(128, dtype=torch.uint8, device=f"cuda:{local}")
    if rank == 0:
        uid = torch.frombuffer(bytearray(Comm.unique_id(lib)), dtype=torch.uint8).to(f"cuda:{local}")
    dist.broadcast(uid, 0)
    return Comm.nccl(lib, bytes(uid.cpu().numpy().tobytes()), world, rank, local)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tile", type=int, nargs=2, default=[8, 8])
    ap.add_argument("--chains", type=int, default=6400)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=300)
    ap.add_argument("--axis", type=int, default=1)
    ap.add_argument("--nstlist", type=int, default=0)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--halo", choices=["push", "nccl"], default="push", help="per-step halo: peer-memory push kernel or NCCL send/recv")
    ap.add_argument("--opt", action="append", default=[], help="engine option key=value (repeatable)")
    ap.add_argument("--equilibrated", action="store_true",
                    help="start from the committed compacted C2 state tiled --tile times (needs --chains = 100 x tiles) instead of the serpentine")
    ap.add_argument("--phases", action="store_true", help="time the phases of a step with CUDA events (engine option phase_timing)")
    args = ap.parse_args()

    import torch

    from opencafemol_b200 import synth, toml_input as ti
    from opencafemol_b200.capi import product_library

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    lib = product_library()
    if args.equilibrated:
        assert args.chains == 100 * args.tile[0] * args.tile[1], "--equilibrated tiles the 100-chain state: --chains must be 100 x tiles"
        spec = synth.c2_equilibrated(tile=tuple(args.tile))
    else:
        spec = synth.hps_slab(n_chains=args.chains, tile=tuple(args.tile))
    comm = None
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
        comm = make_comm(lib, rank, world, local)
    ctx = ti.make_context(lib, spec, device=local, comm=comm, rank=rank, axis=args.axis)
    if args.nstlist:
        ctx.set_option("nstlist", args.nstlist)
    if world > 1 and args.halo == "nccl":
        ctx.set_option("halo_push", 0)
    for kv in args.opt:
        k, v = kv.split("=")
        ctx.set_option(k, float(v))
    ctx.step(args.warmup)
    if args.phases:
        ctx.set_option("phase_timing", 1)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ctx.step(args.steps)
    dev_ms = ctx.counter("last_step_ms")
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    t = torch.tensor([dev_ms * 1e-3, wall], dtype=torch.float64, device=f"cuda:{local}")
    owned = torch.tensor([ctx.counter("owned_beads")], dtype=torch.float64, device=f"cuda:{local}")
    owned_all = [owned.clone() for _ in range(world)]
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_gather(owned_all, owned)
    dev_s, wall_s = t.tolist()
    ok = None
    if args.check:
        st = ctx.get_state(positions=True, velocities=True, energy=True)
        if rank == 0:
            ref = ti.make_context(lib, spec, device=local)
            if args.nstlist:
                ref.set_option("nstlist", args.nstlist)
            ref.step(args.warmup)
            ref.step(args.steps)
            a = ref.get_state(positions=True, velocities=True, energy=True)
            ok = {"positions_equal": bool(np.array_equal(a["positions"], st["positions"])),
                  "velocities_equal": bool(np.array_equal(a["velocities"], st["velocities"])),
                  "max_abs_dx_nm": float(np.abs(a["positions"] - st["positions"]).max()),
                  "epot": [a["potential_energy"], st["potential_energy"]]}
    if rank == 0:
        print(json.dumps({
            "n_gpus": world, "beads": int(spec.mass.size), "state": spec.state or "serpentine start", "md_steps": args.steps, "device_s": dev_s, "wall_s": wall_s,
            "md_steps_per_s": args.steps / dev_s, "ns_per_day": args.steps * spec.dt * 1e-3 / dev_s * 86400.0,
            "nstlist": ctx.counter("nstlist"), "rollbacks": ctx.counter("rollbacks"), "owned": [int(o.item()) for o in owned_all],
            "halo_message_bytes": ctx.counter("halo_message_bytes"),
            "phase_ms_per_step": {k: ctx.counter("phase_ms." + k) / args.steps for k in ("rebuild", "force", "integrate", "halo")} if args.phases else None,
            "transport": "nccl" if world > 1 else "none",
            "halo": ("peer-memory push" if ctx.counter("halo_push") else "nccl send/recv") if world > 1 else None, "check": ok}))
    if world > 1:
        dist.barrier()
        ctx.close()
        comm.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
