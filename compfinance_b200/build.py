"""Build compfinance_b200/lib/libcf_b200.so (CUDA kernels + C ABI) for sm_100a with nvcc, in-tree."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIB_DIR, "libcf_b200.so")
OBJ_DIR = os.path.join(LIB_DIR, "obj")
COMMON = ["cf_device.cuh", "cf_comm.cuh", "cf_kernels.cuh", "cf_pick.h", os.path.join("..", "..", "include", "cf_b200.h")]
# translation unit -> headers it depends on (besides COMMON); compiled in parallel, relinked when any object changes
UNITS = {
    "cf_api.cu": ["cf_tables.h", "cf_dupire.cuh", "cf_bs.cuh", "cf_dlm.cuh", "cf_multi.cuh"],
    "cf_pick_path.cu": ["cf_multi.cuh"],
    "cf_pick_dlm.cu@4": ["cf_dlm.cuh"],      # one unit per asset-count bucket: -DCF_DLM_AMAX=<n>
    "cf_pick_dlm.cu@8": ["cf_dlm.cuh"],
    "cf_pick_dlm.cu@12": ["cf_dlm.cuh"],
    "cf_pick_dlm.cu@16": ["cf_dlm.cuh"],
    "cf_pick_dupire.cu": ["cf_dupire.cuh"],
    "cf_pick_bs.cu": ["cf_dupire.cuh", "cf_bs.cuh"],
    "cf_tables.cpp": ["cf_tables.h", "joe_kuo_init.inc"],
}
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"]
NVCC_FLAGS += os.environ.get("CF_NVCC_EXTRA", "").split()      # experiments: e.g. CF_NVCC_EXTRA="-DCF_REVS_WARPS=12"


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _src(unit):
    return unit.split("@")[0]


def _defines(unit):
    return ["-DCF_DLM_AMAX=" + unit.split("@")[1]] if "@" in unit else []


def _obj(unit):
    return os.path.join(OBJ_DIR, os.path.splitext(_src(unit))[0] + ("_" + unit.split("@")[1] if "@" in unit else "") + ".o")


def _deps(unit):
    return [os.path.join(CSRC, d) for d in [_src(unit)] + UNITS[unit] + COMMON]


def needs_build():
    return _stale(LIB, [d for u in UNITS for d in _deps(u)])


def build(force=False, verbose=False):
    """nvcc -c every stale translation unit (in parallel), then link libcf_b200.so."""
    if not force and not needs_build():
        return LIB
    from concurrent.futures import ThreadPoolExecutor
    os.makedirs(OBJ_DIR, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

    def compile_unit(unit):
        if not force and not _stale(_obj(unit), _deps(unit)):
            return
        cmd = [nvcc] + NVCC_FLAGS + _defines(unit) + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, _src(unit)), "-o", _obj(unit)]
        print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)

    with ThreadPoolExecutor(max_workers=len(UNITS)) as ex:
        list(ex.map(compile_unit, UNITS))
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static"] + [_obj(u) for u in UNITS] + ["-o", LIB]
    print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    return LIB


HOST_DIR = os.path.join(HERE, "host")
HOST_LIB = os.path.join(LIB_DIR, "libcf_host.so")
HOST_DEPS = ["cf_export.cpp", "cf_xl.h", "cf_xlcall.h", "cf_main.h", "cf_store.h", "cf_products.h", "cf_products_multi.h", "cf_models.h", "cf_models_multi.h", "cf_rng.h", "cf_base.h",
             "cf_aad.h", "cf_matrix.h", "cf_util.h", "cf_calib.h"]


def build_host(force=False):
    """Host C++17 mirror of the reference API (compfinance_b200/host) -> libcf_host.so, linked to the engine."""
    build(force=False)
    if not force and os.path.exists(HOST_LIB):
        t = os.path.getmtime(HOST_LIB)
        deps = [os.path.join(HOST_DIR, d) for d in HOST_DEPS if os.path.exists(os.path.join(HOST_DIR, d))]
        deps += [os.path.join(HERE, "..", "include", "cf_b200.h"), LIB]
        if all(os.path.getmtime(d) <= t for d in deps):
            return HOST_LIB
    cmd = ["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-Wall", os.path.join(HOST_DIR, "cf_export.cpp"),
           "-o", HOST_LIB, "-L" + LIB_DIR, "-lcf_b200", "-Wl,-rpath,$ORIGIN"]
    print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    return HOST_LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    build_host(force="--force" in sys.argv)
