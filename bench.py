#!/usr/bin/env python
"""bench.py — ns/day of the per-step coarse-grained MD hot path on the HPS IDR condensate slab
(BASELINE.json configs[1]: 100 x 163-bead FUS-LC chains = 16 300 beads, Ashbaugh-Hatch + Debye-Hueckel).

One bench "step" = one observer interval of the reference's run loop, i.e. ``integrator->step(save_step)``
with save_step = 1000 MD steps (src/Simulator.cpp:49-76); ns/day = MD steps/s x dt[ps] x 86.4e-3... see
`ns_per_day`.

    python bench.py --gpus 1 --steps 10 --warmup 3            # the engine (this repo)
    python bench.py --impl reference --steps 2 --warmup 1      # CPU arm: the oracle port of the reference path
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N   # N independent replicas (headline) + the
                                                                               # slab-decomposed C5 run and its bit-exact check

Both arms start from the same committed state of the workload, opencafemol_b200/data/c2_state.npz (the condensate after
40 000 steps, when its compaction has levelled off: config.state), so they time the same pair work per step.

The line printed by rank 0 follows the driver's contract (metric/value/unit/..., e2e, roofline,
cpu_baseline, clocks, gpu_launches).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MD_STEPS_PER_STEP = 1000  # save_step of the reference's FUS sample
FLOP_PER_AH_PAIR = 50.0   # algorithmic FP32 flop per in-cutoff pair, Newton-3 applied
FLOP_PER_DH_PAIR = 40.0
BYTES_PER_BEAD_INTEGRATE = 80.0


def ns_per_day(md_steps, dt_ps, seconds):
    return md_steps * dt_ps * 1e-3 / seconds * 86400.0


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.idx)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, val in zip(names, f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            return json.load(fh), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


def ncu_traffic():
    """dram bytes (read + write) per force_kernel launch from the committed `ncu --set full` capture, or None"""
    p = os.path.join(ROOT, "profiles", "force_kernel_ncu.json")
    try:
        with open(p) as fh:
            d = json.load(fh)
        return {"dram_bytes_per_launch": d["dram_bytes_read"] + d["dram_bytes_write"], "source": d.get("source", "profiles/force_kernel_ncu.json")}
    except Exception:
        return None


def cpu_port_run(spec, md_steps, threads=None, state=None):
    """Times the CPU oracle (OpenMM-equivalent CPU restatement of the reference path; OpenMM itself is
    not installable offline) on `md_steps` MD steps of the same workload; state = (positions, velocities) to start
    from (the condensate compacts over the first 10^4..10^5 steps, so the start state sets the pair count)."""
    import __graft_entry__ as g
    from opencafemol_b200 import toml_input
    from opencafemol_b200.capi import Library

    g.build_oracle()
    lib = Library(g.ORACLE_SO, "oracle_")
    ctx = toml_input.make_context(lib, spec)
    # all host cores this process may run on: torchrun exports OMP_NUM_THREADS=1 to its workers, which would time the
    # CPU arm on one thread although rank 0 is the only rank that runs it (OCM_CPU_THREADS overrides)
    threads = threads or int(os.environ.get("OCM_CPU_THREADS", "0")) or len(os.sched_getaffinity(0))
    ctx.set_option("threads", threads)
    if state is not None:
        ctx.set_positions(state[0])
        ctx.set_velocities(state[1])
    cores = int(ctx.get_option("threads"))
    ctx.step(1)
    t0 = time.perf_counter()
    ctx.step(md_steps)
    dt = time.perf_counter() - t0
    return ns_per_day(md_steps, spec.dt, dt), cores, dt


def workload():
    from opencafemol_b200 import synth

    # both arms start from the committed compacted state (tests/golden/make_c2_state.py): the serpentine start keeps compacting
    # for tens of thousands of steps, so the pair work per step would otherwise depend on the warm-up length and differ
    # between the GPU arm and the (short) CPU arm
    return synth.c2_equilibrated() if os.path.exists(synth.C2_STATE) else synth.hps_slab(n_chains=100)


def config_dict(spec, extra=None):
    c = {"workload": "C2 HPS IDR condensate slab: 100 x FUS-LC(163) = 16300 beads, Ashbaugh-Hatch (rc 3.999 nm) + "
                     "Debye-Hueckel (I=0.1 M) + harmonic bonds, Langevin 300 K, dt 0.00489 ps",
         "beads": int(spec.mass.size), "md_steps_per_step": MD_STEPS_PER_STEP, "dt_ps": spec.dt,
         "l2": "working set (< 40 MB) is L2-resident by construction in a real run; every MD step rewrites positions and "
               "forces, no flush between bench steps", "parallelism": "1 replica per GPU"}
    c["state"] = spec.state or "serpentine start at the sample's bead density (compacts during the run)"
    if extra:
        c.update(extra)
    return c


CPU_SAMPLE_STEPS = 100      # MD steps per bench step of the CPU arm (a few seconds on 16 cores)
CPU_BASELINE_STEPS = 400    # MD steps of the cpu_baseline leg of the engine arm (10-30 s of CPU work)


def run_reference(args):
    """CPU arm.  The reference's arithmetic lives in OpenMM, which cannot be installed offline; what runs here is the
    OpenMP double-precision restatement (oracle/ocm_oracle.cpp, kind "port") on all host cores.  oracle/_ref (the
    reference's own sources on a stand-in OpenMM with an expression interpreter) is the parity checker, not a fair
    performance baseline: it is single-threaded and interprets every pair."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    spec = workload()
    sample_steps = CPU_SAMPLE_STEPS
    for _ in range(args.warmup):
        cpu_port_run(spec, 2)
    vals, cores = [], 1
    t_all = 0.0
    for _ in range(args.steps):
        v, cores, dt = cpu_port_run(spec, sample_steps)
        vals.append(v); t_all += dt
    value = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": "ns/day", "value": value, "unit": "ns/day", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t_all / args.steps * 1e3, "md_steps_timed_per_step": sample_steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(spec),
        "cpu_baseline": {"value": value, "unit": "ns/day", "cores": cores, "kind": "port",
                         "sample": f"{sample_steps} MD steps of the C2 workload per bench step (ms_per_step is the measured time of "
                                   f"those {sample_steps} steps, not of {MD_STEPS_PER_STEP}); OpenMM-equivalent CPU restatement, OpenMM "
                                   "not installable offline"},
        "e2e": {"value": value, "unit": "ns/day", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def nccl_comm(lib, rank, world, local):
    """NCCL communicator of the engine: rank 0 draws the unique id, torch.distributed carries it to the others."""
    import torch
    import torch.distributed as dist

    from opencafemol_b200.capi import Comm

    uid = torch.zeros(128, dtype=torch.uint8, device=f"cuda:{local}")
    if rank == 0:
        uid = torch.frombuffer(bytearray(Comm.unique_id(lib)), dtype=torch.uint8).to(f"cuda:{local}")
    dist.broadcast(uid, 0)
    return Comm.nccl(lib, bytes(uid.cpu().numpy().tobytes()), world, rank, local)


def sharding_gather(values, local):
    from opencafemol_b200 import sharding

    return sharding.gather_concat(values, device=f"cuda:{local}")


def timed_steps(ctx, warm, steps, local, world):
    """device seconds (MAX over ranks) of `steps` MD steps after `warm` warm-up steps"""
    import torch
    import torch.distributed as dist

    from opencafemol_b200 import sharding

    ctx.step(warm)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ctx.step(steps)
    dev = ctx.counter("last_step_ms") * 1e-3
    # (world == 1: a single-rank measurement, possibly made by rank 0 alone inside a multi-rank run: no collective)
    return sharding.max_over_ranks([dev], device=f"cuda:{local}")[0] if world > 1 else dev


def single_gpu_records(lib, local):
    """N = 1: the other single-GPU configurations of BASELINE.json beside the C2 headline (each a few seconds)"""
    import sys as _sys

    from opencafemol_b200 import synth, toml_input

    out = {}
    # C5: the compacted C2 state tiled 8 x 8 = 1 043 200 beads on one GPU
    spec = synth.c2_equilibrated(tile=(8, 8))
    ctx = toml_input.make_context(lib, spec, device=local)
    dev = timed_steps(ctx, 200, 400, local, 1)
    out["c5"] = {"workload": "C5: C2 compacted state tiled 8 x 8, 1 043 200 beads, one GPU", "md_steps_per_s": 400 / dev,
                 "ns_per_day": ns_per_day(400, spec.dt, dev), "ms_per_md_step": dev / 400 * 1e3, "nstlist": ctx.counter("nstlist"),
                 "pair_kernel": "tiles" if ctx.counter("tiled.1") else "rows", "lane_pairs_evaluated_per_step": ctx.counter("pair_evals_per_step"),
                 "kernel_ms": {k: ctx.profile_kernel(k, 5) for k in ("force", "integrate", "rebuild")}}
    hbm = spec.mass.size * BYTES_PER_BEAD_INTEGRATE / (out["c5"]["kernel_ms"]["integrate"] * 1e-3) / 1e9
    peaks, how = measured_peaks()
    # FP32 line of the tile force launches (Ashbaugh-Hatch tile list + the Coulomb rows of the fused launch), algorithmic flops
    n_ah5, n_dh5 = ctx.num_neighbor_pairs(1), ctx.num_neighbor_pairs(2)
    flop5 = n_ah5 * FLOP_PER_AH_PAIR + n_dh5 * FLOP_PER_DH_PAIR
    peak32 = lib.measure_fp32_peak(local)
    out["c5"]["roofline"] = {"bound": "fp32", "kernel": "tile_force_kernel + force_kernel (all force launches of a step)",
                             "achieved": flop5 / (out["c5"]["kernel_ms"]["force"] * 1e-3) / 1e12, "peak": peak32, "unit": "TFLOP/s",
                             "frac": flop5 / (out["c5"]["kernel_ms"]["force"] * 1e-3) / 1e12 / peak32 if peak32 else None,
                             "in_cutoff_pairs": {"ashbaugh_hatch": n_ah5, "debye_huckel": n_dh5}}
    out["c5"]["roofline_hbm"] = {"bound": "hbm", "kernel": "integrate_kernel", "achieved": hbm, "peak": peaks.get("hbm_gbs"), "unit": "GB/s",
                                 "frac": hbm / peaks.get("hbm_gbs", 1.0), "peak_source": how}
    ctx.close()
    # C4: 512 SH3 replicas (8 groups x 64) in one context
    _sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import load_golden_spec

    sh3 = load_golden_spec("sh3")
    ctx = toml_input.make_context(lib, sh3, device=local, replicas=512)
    dev = timed_steps(ctx, 500, 2000, local, 1)
    out["c4"] = {"workload": "C4: 512 replicas of the SH3 AICG2+ sample (56 beads each) in one context", "replica_steps_per_s": 512 * 2000 / dev,
                 "md_steps_per_s": 2000 / dev, "kernel": "replica_step_kernel (one thread block per replica)" if ctx.get_option("replica_kernel") else "general path"}
    ctx.close()
    ctx = toml_input.make_context(lib, sh3, device=local)
    dev = timed_steps(ctx, 500, 5000, local, 1)
    out["c1"] = {"workload": "C1: sample/toml_input/sh3_AICG2+.toml (56 beads)", "md_steps_per_s": 5000 / dev, "ns_per_day": ns_per_day(5000, sh3.dt, dev),
                 "kernel": "replica_step_kernel (one thread block per replica)" if ctx.get_option("replica_kernel") else "general path"}
    # end to end with the sample's own save_step = 100: every frame downloads positions + energies (what its observers ask for),
    # the download of frame k overlapping the steps of frame k + 1
    import torch
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    req = None
    for _ in range(50):
        ctx.step(100)
        nxt = ctx.get_state_begin(positions=True, energy=True)
        if req is not None:
            ctx.get_state_end(req)
        req = nxt
    ctx.get_state_end(req)
    torch.cuda.synchronize()
    out["c1"]["e2e_md_steps_per_s_save_step_100"] = 5000 / (time.perf_counter() - t0)
    ctx.close()
    return out


def multi_gpu_records(lib, rank, world, local):
    """N > 1: the path that shards.  (1) slab_check: the C2 state decomposed into `world` slabs along z over NCCL + the
    peer-memory halo, compared with the single-domain run on rank 0 (bit for bit with a fixed rebuild interval).
    (2) c5: the 1 043 200-bead condensate decomposed along y, against the one-GPU figure measured on rank 0 in the same run."""
    import torch.distributed as dist

    from opencafemol_b200 import synth, toml_input

    comm = nccl_comm(lib, rank, world, local)
    out = {}
    # ---- slab_check ----
    spec = synth.c2_equilibrated()
    ctx = toml_input.make_context(lib, spec, device=local, comm=comm, rank=rank, axis=2)
    ctx.set_option("nstlist", 10)
    ctx.step(200)
    st = ctx.get_state(positions=True, velocities=True, energy=True)
    transport = "peer-memory push (CUDA IPC over NVLink)" if ctx.counter("halo_push") else "nccl send/recv"
    exchanges = ctx.counter("halo_exchanges")
    if rank == 0:
        ref = toml_input.make_context(lib, spec, device=local)
        ref.set_option("nstlist", 10)
        ref.step(200)
        a = ref.get_state(positions=True, velocities=True, energy=True)
        out["slab_check"] = {"workload": "C2 state, %d slabs along z, 200 MD steps, rebuild every 10" % world,
                             "positions_equal": bool(np.array_equal(a["positions"], st["positions"])),
                             "velocities_equal": bool(np.array_equal(a["velocities"], st["velocities"])),
                             "max_abs_dx_nm": float(np.abs(a["positions"] - st["positions"]).max()),
                             "epot_kj": [a["potential_energy"], st["potential_energy"]], "halo": transport, "halo_exchanges": exchanges}
        ref.close()
    dist.barrier()
    ctx.close()
    # ---- C5 ----
    spec = synth.c2_equilibrated(tile=(8, 8))
    one = None
    if rank == 0:
        c1 = toml_input.make_context(lib, spec, device=local)
        c1.step(200)
        c1.step(400)
        one = 400 / (c1.counter("last_step_ms") * 1e-3)
        c1.close()
    dist.barrier()
    ctx = toml_input.make_context(lib, spec, device=local, comm=comm, rank=rank, axis=1)
    dev = timed_steps(ctx, 300, 600, local, world)
    nst = ctx.counter("nstlist")
    ctx.set_option("phase_timing", 1)
    ctx.step(200)
    phase_run_ms = ctx.counter("last_step_ms") / 200
    phases = {k: ctx.counter("phase_ms." + k) / 200 for k in ("rebuild", "force", "integrate", "halo")}
    owned = sharding_gather([ctx.counter("owned_beads")], local)
    all_phases = sharding_gather([phases[k] for k in ("rebuild", "force", "integrate", "halo")], local).reshape(world, 4)
    evals = sharding_gather([ctx.counter("pair_evals_per_step")], local)   # (collective: every rank asks)
    if rank == 0:
        sN = 600 / dev
        out["c5"] = {"workload": "C5: C2 compacted state tiled 8 x 8, 1 043 200 beads, %d slabs along y" % world, "md_steps_per_s": sN,
                     "ns_per_day": ns_per_day(600, spec.dt, dev), "md_steps_per_s_one_gpu": one, "strong_scaling_efficiency": sN / one / world,
                     "nstlist": nst, "phase_ms_per_step_rank0": phases, "phase_run_ms_per_step_rank0": phase_run_ms,
                     "phase_note": "phases are CUDA-event intervals on rank 0's compute stream in a separate 200-step run; 'halo' is the stall at the "
                                   "join with the side stream that carries the halo positions, 'force' includes the interior launch that overlaps them", "halo_message_bytes": ctx.counter("halo_message_bytes"),
                     "halo": "peer-memory push (CUDA IPC over NVLink)" if ctx.counter("halo_push") else "nccl send/recv",
                     "pair_kernel": "tiles" if ctx.counter("tiled.1") else "rows", "owned_beads_per_rank": [int(v) for v in owned],
                     "phase_ms_per_step_by_rank": {k: [round(float(v), 4) for v in all_phases[:, i]] for i, k in enumerate(("rebuild", "force", "integrate", "halo"))},
                     "lane_pairs_evaluated_per_step_by_rank": [float(v) for v in evals]}
    dist.barrier()
    ctx.close()
    comm.close()
    # ---- C4: the fixed ensemble of 512 SH3 replicas sharded 512 / N per rank (no collective on the data path) ----
    import sys as _sys

    from opencafemol_b200 import sharding
    _sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import load_golden_spec

    sh3 = load_golden_spec("sh3")
    one = None
    if rank == 0:
        c1 = toml_input.make_context(lib, sh3, device=local, replicas=512)
        one = 512 * 2000 / timed_steps(c1, 500, 2000, local, 1)
        c1.close()
    dist.barrier()
    first, per = sharding.replica_shard(512, world, rank)
    ctx = toml_input.make_context(lib, sh3, device=local, seed=sh3.seed + first, replicas=per)
    dev = timed_steps(ctx, 500, 2000, local, world)
    if rank == 0:
        out["c4"] = {"workload": "C4: 512 SH3 replicas, %d per GPU, no collective per step" % per, "replica_steps_per_s": 512 * 2000 / dev,
                     "replica_steps_per_s_one_gpu": one, "strong_scaling_efficiency": 512 * 2000 / dev / one / world}
    dist.barrier()
    ctx.close()
    return out


def run_engine(args):
    import torch
    import torch.distributed as dist

    from opencafemol_b200 import toml_input
    from opencafemol_b200.capi import product_library

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    torch.cuda.set_device(local)

    lib = product_library()  # raises if the CUDA extension is missing: no CPU fallback
    spec = workload()
    spec.seed += rank  # independent replicas
    ctx = toml_input.make_context(lib, spec, device=local)
    n = spec.mass.size

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # warm-up: also equilibrates the serpentine start and settles the adaptive rebuild interval
    launches0 = None
    for _ in range(max(args.warmup, 3)):
        ctx.step(MD_STEPS_PER_STEP)
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    launches0 = ctx.counter("kernel_launches")
    t_dev = 0.0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ctx.step(MD_STEPS_PER_STEP)
        t_dev += ctx.counter("last_step_ms") * 1e-3
    barrier()
    t_wall = time.perf_counter() - t0
    clocks = sampler.stop()
    launches = ctx.counter("kernel_launches") - launches0
    nst = ctx.counter("nstlist")

    # e2e: the same metric through the C ABI with HOST buffers: per bench step upload positions and
    # velocities, step, download positions, velocities and the energies (what Simulator::run + observers do)
    st = ctx.get_state(positions=True, velocities=True)
    pos, vel = st["positions"], st["velocities"]
    # (warm-up of this path as well: the first calls allocate the pinned staging blocks, some 100 ms of cudaHostAlloc in a process
    # that has torch's context loaded; without it a 10-step run reads 0.81 of the resident value instead of 0.99)
    for _ in range(2):
        ctx.set_state_async(pos, vel)
        ctx.step(MD_STEPS_PER_STEP // 10)
        st = ctx.get_state_end(ctx.get_state_begin(positions=True, velocities=True, energy=True))
        pos, vel = st["positions"], st["velocities"]
    barrier()
    te0 = time.perf_counter()
    for _ in range(args.steps):
        # (pinned staging + device-side conversion on both sides; the energy kernels reuse the pair lists of the last window)
        ctx.set_state_async(pos, vel)
        ctx.step(MD_STEPS_PER_STEP)
        st = ctx.get_state_end(ctx.get_state_begin(positions=True, velocities=True, energy=True))
        pos, vel = st["positions"], st["velocities"]
    barrier()
    t_e2e = time.perf_counter() - te0
    h2d = 2 * n * 3 * 8
    d2h = 2 * n * 3 * 8 + 33 * 8

    from opencafemol_b200 import sharding

    t_dev, t_e2e, t_wall = sharding.max_over_ranks([t_dev, t_e2e, t_wall], device=f"cuda:{local}")
    # the configurations beside the headline (outside every timed region above)
    extra = {}
    if not args.no_extra:
        try:
            extra = single_gpu_records(lib, local) if world == 1 else multi_gpu_records(lib, rank, world, local)
        except Exception as e:  # noqa: BLE001  (the headline line must still be printed)
            extra = {"extra_records_error": repr(e)}

    if rank == 0:
        md = args.steps * MD_STEPS_PER_STEP
        value = world * ns_per_day(md, spec.dt, t_dev)
        e2e = world * ns_per_day(md, spec.dt, t_e2e)
        # roofline of the dominant kernel (fused force launch): algorithmic FP32 flop / measured launch time
        n_ah = len(ctx.neighbor_pairs(1))
        n_dh = len(ctx.neighbor_pairs(2))
        flop = n_ah * FLOP_PER_AH_PAIR + n_dh * FLOP_PER_DH_PAIR
        ms_force = ctx.profile_kernel("force", 200)
        ms_int = ctx.profile_kernel("integrate", 200)
        ms_reb = ctx.profile_kernel("rebuild", 50)
        peak = lib.measure_fp32_peak(local)
        peaks, how = measured_peaks()
        achieved = flop / (ms_force * 1e-3) / 1e12
        hbm_int = n * BYTES_PER_BEAD_INTEGRATE / (ms_int * 1e-3) / 1e9
        line = {
            "metric": "ns/day", "value": value, "unit": "ns/day", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": t_dev / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": config_dict(spec),
            "engine_options": {"nstlist": nst, "skin_nm": ctx.get_option("skin_nm"), "pair_kernel": "rows" if not ctx.counter("tiled.1") else "tiles"},
            "md_steps_per_s": md / t_dev, "pair_interactions_per_s": (n_ah + n_dh) * md / t_dev * world,
            "wall_ms_per_step": t_wall / args.steps * 1e3,
            "e2e": {"value": e2e, "unit": "ns/day", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "fp32", "kernel": "force_kernel (fused pair lists + bonded)", "achieved": achieved, "peak": peak,
                         "unit": "TFLOP/s", "frac": achieved / peak if peak else None, "traffic": ncu_traffic(),
                         "peak_source": "register-resident FMA microbenchmark on this GPU (SURVEY.md 8d); no tensor-core work on this path",
                         "algorithmic_flop_per_launch": flop, "ms_per_launch": ms_force,
                         "in_cutoff_pairs": {"ashbaugh_hatch": n_ah, "debye_huckel": n_dh}},
            "roofline_hbm": {"bound": "hbm", "kernel": "integrate_kernel", "achieved": hbm_int, "peak": peaks.get("hbm_gbs"), "unit": "GB/s",
                             "frac": hbm_int / peaks.get("hbm_gbs", 1.0), "peak_source": how, "ms_per_launch": ms_int,
                             "note": "16 300 beads = 1.3 MB per launch: latency-bound, L2-resident"},
            "kernel_ms": {"force": ms_force, "integrate": ms_int, "rebuild_all_lists": ms_reb},
        }
        line.update(extra)
        if world == 1 and not args.no_cpu_baseline:
            sample = CPU_BASELINE_STEPS
            v, cores, dt = cpu_port_run(spec, sample, state=(pos, vel))
            line["cpu_baseline"] = {"value": v, "unit": "ns/day", "cores": cores, "kind": "port",
                                    "sample": f"{sample} MD steps of the same C2 workload from the state the engine was timed on "
                                              f"({n_ah} in-cutoff pairs; {dt:.1f} s); OpenMM-equivalent CPU restatement (oracle), "
                                              "OpenMM not installable offline"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the C5 / C4 / slab_check records beside the C2 headline")
    args = ap.parse_args()
    import __graft_entry__ as g

    if args.impl == "reference":
        run_reference(args)
        return
    if int(os.environ.get("LOCAL_RANK", "0")) == 0:
        g.build()
    run_engine(args)


if __name__ == "__main__":
    main()
