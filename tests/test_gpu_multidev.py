"""A single-process multi-device context (cf_init with a device list; NumericalParam::devices in the host API): every
entry point shards its paths over the devices and sums over peer memory.  Needs two GPUs (skipped on a single-GPU box;
the sharding rule itself is covered on CPU by tests/test_dist.py)."""
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
def test_entry_points_on_two_devices_match_one_device(built):
    n = min(_gpus(), 4)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "multidev_check.py"), str(n)], cwd=ROOT, capture_output=True,
                         text=True, timeout=240)
    assert out.returncode == 0, out.stderr[-3000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("MULTIDEV ")][0]
    rep = json.loads(line[len("MULTIDEV "):])
    for name, r in rep.items():
        # the shards are summed in a different order than one device sums its blocks: 1e-11 of the entry's scale;
        # per-path payoffs are the same numbers
        tol = 0.0 if name == "paths" else 1e-11
        assert r["max_rel"] <= tol, (name, r)
        assert r["bitwise_repeat"], (name, r)
