"""CPU tests of the drop-in boundary: the libraries build, load, export every symbol that
include/*.h declares, and refuse to compute without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions(header):
    text = open(header).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cf_[a-z0-9_]+)\s*\(", text)))


def test_engine_exports_every_declared_symbol(built):
    from compfinance_b200 import capi
    lib = C.CDLL(capi.LIB_PATH)
    names = _declared_functions(os.path.join(ROOT, "include", "cf_b200.h"))
    assert len(names) >= 19
    for n in names:
        assert hasattr(lib, n), f"{n} declared in cf_b200.h but not exported"
    assert sorted(capi.EXPORTED) == names


def test_host_library_exports(built):
    from compfinance_b200 import api
    lib = C.CDLL(api.HOST_LIB_PATH)
    for n in api.EXPORTED:
        assert hasattr(lib, n)


def test_engine_built_for_sm100a(built):
    from compfinance_b200 import capi
    out = subprocess.run(["cuobjdump", "-lelf", capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_direction_numbers_on_host(built):
    """The direction-number table is regenerated on the host: checkable without a GPU."""
    from compfinance_b200 import capi
    from oracle import restate as R
    lib = capi.load()
    d = R.sobol_direction_numbers()
    assert lib.cf_sobol_max_dim() == 1101
    for bit in range(32):
        for dim in range(0, 1101, 13):
            assert lib.cf_sobol_direction_number(bit, dim) == d[bit, dim]


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure mode")
def test_fails_loudly_without_gpu(built):
    from compfinance_b200 import capi
    from compfinance_b200.api import CompFinance, CfHostError
    with pytest.raises(capi.CfError, match="no CUDA device"):
        capi.Engine(device=0)
    e = capi.Engine()
    with pytest.raises(capi.CfError, match="no CUDA device"):
        e.inv_normal(np.array([0.5]))
    cf = CompFinance()
    cf.put_black_scholes(100, 0.15, False, 0, 0, "bs")
    cf.put_european(100, 1.0, 1.0, "eur")
    with pytest.raises(CfHostError, match="no CUDA device"):
        cf.value("bs", "eur", 1024)


def test_bench_reference_arm_contract(ref):
    """`bench.py --impl reference` (the reference's own CPU path on the host cores) prints ONE JSON line with the
    contract's keys; under torchrun only rank 0 works and prints."""
    import json
    import sys
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    line = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "paths/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]
    other = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=dict(os.environ, RANK="1", WORLD_SIZE="2"))
    assert other.returncode == 0 and other.stdout.strip() == ""
