"""Two ranks on two GPUs: the sum over ranks done by the reduction kernel itself over peer memory (cf_comm_create / cf_comm_connect: CUDA IPC)
against one NCCL all-reduce of the per-rank results, and against the single-GPU numbers.  Needs two GPUs; skipped
on a single-GPU box (the host-side sharding logic is covered on CPU by tests/test_dist.py)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.skipif(_gpus() < 2, reason="needs two GPUs")
def test_fused_reduction_over_two_ranks():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "bench.py"), "--gpus", "2", "--steps", "3", "--warmup", "3", "--no-cpu-baseline"]
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1                                       # rank 0 prints one JSON line
    d = json.loads(lines[0])
    # bench.py itself compares the fused result with one NCCL all-reduce of the per-rank vectors before using it
    assert "inside the reduction kernel" in d["config"]["parallelism"]
    assert d["e2e"]["api"] == "dupireAADRisk (libcf_host.so)"                      # every rank calls the host API; the library shards
    assert abs(d["config"]["price"] / 0.96926107424976005 - 1) < 1e-10 and abs(d["config"]["delta"] / 0.020804057371458962 - 1) < 1e-8
    assert d["n_gpus"] == 2 and d["scaling"] == "strong"
