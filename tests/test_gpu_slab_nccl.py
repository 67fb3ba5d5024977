"""NCCL transport of the slab decomposition: one process per GPU under torchrun, checked bit for bit against the
single-domain run.  Needs >= 2 GPUs on the box (skipped otherwise; the in-process transport covers the logic on one)."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("halo", ["push", "nccl"])
def test_slab_nccl_two_ranks_bit_exact(engine, halo):
    """halo = push: per-step positions stored straight into the neighbours' buffers over NVLink peer memory
    (dd_push_kernel + stream wait on a sequence flag); halo = nccl: ncclSend/ncclRecv.  Both bit-exact."""
    if engine.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29517" if halo == "push" else "29518", os.path.join(ROOT, "tools", "slab_run.py"), "--tile", "1", "1", "--chains", "40",
           "--steps", "300", "--warmup", "50", "--nstlist", "10", "--check", "--halo", halo]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert line["n_gpus"] == 2 and line["transport"] == "nccl"
    assert line["halo"] == ("peer-memory push" if halo == "push" else "nccl send/recv")
    assert line["check"]["positions_equal"] and line["check"]["velocities_equal"], line
    assert sum(line["owned"]) == line["beads"]
