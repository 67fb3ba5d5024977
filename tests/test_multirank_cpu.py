"""The N > 1 paths on CPU: two processes over gloo (rendezvous on 127.0.0.1), the CPU oracle standing in for the
engine.  Covers the host logic the GPU runs use under NCCL: replica sharding (config C4: no collective on the data
path, per-replica energies gathered at save frames), MAX-over-ranks timing, and the slab ownership rule (config C5)
evaluated rank by rank."""
import json
import os
import subprocess
import sys
import textwrap
import threading

import numpy as np
import pytest

from conftest import ROOT, load_golden_spec
from opencafemol_b200 import sharding, synth, toml_input as ti
from opencafemol_b200.capi import Comm


def test_replica_shards_partition():
    for total in (1, 7, 64, 512):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                first, count = sharding.replica_shard(total, world, r)
                seen += list(range(first, first + count))
            assert seen == list(range(total))
            sizes = [sharding.replica_shard(total, world, r)[1] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    assert sharding.replica_seeds(100, 8, 2, 1) == [104, 105, 106, 107]
    with pytest.raises(ValueError):
        sharding.replica_shard(8, 2, 2)


WORKER = textwrap.dedent("""
    import json, os, sys
    sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
    import numpy as np
    import torch.distributed as dist
    from conftest import load_golden_spec
    from opencafemol_b200 import sharding, toml_input as ti
    from opencafemol_b200.capi import Library
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    lib = Library(os.path.join({root!r}, "oracle", "libocm_oracle.so"), "oracle_")
    spec = load_golden_spec("sh3")
    total, steps = 6, 40
    first, count = sharding.replica_shard(total, world, rank)
    ctx = ti.make_context(lib, spec, seed=spec.seed + first, replicas=count)
    ctx.step(steps)
    e = ctx.replica_energies()["potential_energy"]
    allE = sharding.gather_concat(e)
    tmax = sharding.max_over_ranks([0.5 + rank, 2.0 - rank])
    if rank == 0:
        print(json.dumps({{"energies": allE.tolist(), "tmax": tmax, "world": world}}))
    dist.barrier()
    dist.destroy_process_group()
""")


def test_replica_ensemble_sharded_over_two_gloo_ranks(oracle, tmp_path):
    """6 SH3 replicas on 2 ranks (3 + 3) reproduce the 6-replica batch of one process replica by replica"""
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29531", str(script)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-1500:] + out.stderr[-3000:]
    got = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert got["world"] == 2 and got["tmax"] == [1.5, 2.0]
    spec = load_golden_spec("sh3")
    ctx = ti.make_context(oracle, spec, seed=spec.seed, replicas=6)
    ctx.step(40)
    want = ctx.replica_energies()["potential_energy"]
    assert np.allclose(got["energies"], want, rtol=1e-12, atol=1e-9)
    assert len(set(np.round(want, 6))) == 6  # the replicas really differ (different noise keys)


@pytest.mark.parametrize("world,axis", [(2, 1), (4, 2)])
def test_slab_ownership_rule_on_cpu(oracle, world, axis):
    """the exact integer ownership rule of the slab decomposition, evaluated independently by every rank: each bead
    has exactly one owner, a halo bead of one rank is owned by one of its two ring neighbours"""
    spec = synth.hps_slab(n_chains=12)
    comm = Comm.local(oracle, world)
    roles = [None] * world

    def body(r):
        ctx = ti.make_context(oracle, spec, comm=comm, rank=r, axis=axis)
        ctx.set_option("slab_halo_nm", 4.8)  # list radius of the Ashbaugh-Hatch term (3.999 nm) + skin, as the engine derives it
        roles[r] = ctx.roles()

    th = [threading.Thread(target=body, args=(r,)) for r in range(world)]
    [t.start() for t in th]
    [t.join(timeout=300) for t in th]
    roles = np.stack(roles)
    assert ((roles == 1).sum(0) == 1).all()
    owner = (roles == 1).argmax(0)
    assert (roles == 2).any()
    for r in range(world):
        halo = np.nonzero(roles[r] == 2)[0]  # may be empty: the slab of beads does not reach every rank's faces
        assert np.isin(owner[halo], [(r - 1) % world, (r + 1) % world]).all()
    # the owner is floor(frac(x/L) * world) on the decomposition axis
    frac = (spec.positions[:, axis] / spec.box[axis]) % 1.0
    assert np.array_equal(owner, np.minimum((frac * world).astype(int), world - 1))
