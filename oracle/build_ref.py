#!/usr/bin/env python3
"""Build oracle/_ref/libcfref.so: the UNMODIFIED-ALGORITHM reference (asavine/CompFinance)
compiled with g++ for use as a test oracle and as the CPU baseline.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (compfinance_b200/, include/) may
load this library.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs use it.

The reference sources are read where they lie under /root/reference.  They are MSVC-dialect
C++17 and need six mechanical, semantics-preserving edits to compile with g++.
The edits are applied to throw-away copies in a temp directory; only the resulting .so is
written to oracle/_ref/ (git-ignored).  No reference source is copied into the repository.

Usage: python oracle/build_ref.py [--ref /root/reference] [--force]
"""
import argparse
import os
import re
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_ref")
OUT_SO = os.path.join(OUT_DIR, "libcfref.so")

HEADERS = [
    "AAD.h", "AADExpr.h", "AADNode.h", "AADNumber.h", "AADTape.h", "blocklist.h", "gaussians.h",
    "matrix.h", "interp.h", "utility.h", "choldc.h", "analytics.h", "ivs.h", "mcBase.h",
    "mcMdl.h", "mcMdlBS.h", "mcMdlDupire.h", "mcMdlMultiDisplaced.h", "mcPrd.h", "mcPrdMulti.h",
    "mrg32k3a.h", "sobol.h", "store.h", "main.h", "threadPool.h", "ConcurrentQueue.h",
]
SOURCES = ["AAD.cpp", "mcBase.cpp", "ThreadPool.cpp", "sobol.cpp"]


def patch(name: str, text: str) -> str:
    if name == "mrg32k3a.h":
        # patch 4: multi-word functional casts are MSVC-only
        text = re.sub(r"(?<![\w(])unsigned long long\s*\(", "(unsigned long long)(", text)
    if name == "AADExpr.h":
        # patch 6: dependent-name disambiguator
        text = re.sub(r"(\w|\))\.pushAdjoint<", r"\1.template pushAdjoint<", text)
    if name == "mcBase.h":
        # patch 5: in-class explicit specialisation is ill-formed outside MSVC; fold both
        # versions into one template with if constexpr.
        pat = re.compile(
            r"template<class U>\s*void putParametersOnTapeT\(\)\s*\{\s*\}\s*"
            r"//[^\n]*\n\s*template <>\s*void putParametersOnTapeT<Number>\(\)\s*\{[^}]*\}",
            re.S)
        repl = (
            "template<class U>\n    void putParametersOnTapeT()\n    {\n"
            "        if constexpr (std::is_same<U, Number>::value)\n"
            "        {\n            for (Number* param : parameters()) param->putOnTape();\n        }\n"
            "    }")
        text, n = pat.subn(repl, text)
        if n != 1:
            raise RuntimeError("mcBase.h patch 5 did not apply (reference layout changed?)")
    return text


def build(ref: str, force: bool = False, verbose: bool = True, variant: str = "") -> str:
    """variant "": the checker, -ffp-contract=off.  variant "fma": -ffp-contract=fast into libcfref_fma.so -- the same
    reference with fused multiply-adds, built only to measure the reference's own sensitivity to the compiler
    (tests/test_oracle.py: the superbucket chain moves by ~1e-4, prices and AAD risks by ~1e-13)."""
    driver = os.path.join(HERE, "ref_driver.cpp")
    OUT_SO = os.path.join(OUT_DIR, "libcfref.so" if not variant else f"libcfref_{variant}.so")
    contract = "-ffp-contract=fast" if variant == "fma" else "-ffp-contract=off"
    if not os.path.isdir(ref):
        if os.path.exists(OUT_SO):
            return OUT_SO
        raise FileNotFoundError(f"{ref} not present and no prebuilt {OUT_SO}")
    if (not force and os.path.exists(OUT_SO)
            and os.path.getmtime(OUT_SO) >= max(os.path.getmtime(driver), os.path.getmtime(__file__))):
        return OUT_SO
    os.makedirs(OUT_DIR, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="cfref_")
    try:
        for f in HEADERS + SOURCES:
            with open(os.path.join(ref, f), "r", encoding="latin-1") as fh:
                text = fh.read()
            with open(os.path.join(tmp, f), "w", encoding="latin-1") as fh:
                fh.write(patch(f, text))
        # patch 1: case-sensitive file system; includes say "ThreadPool.h"
        with open(os.path.join(tmp, "ThreadPool.h"), "w") as fh:
            fh.write('#pragma once\n#include "threadPool.h"\n')
        cmd = [
            # -ffp-contract=off: no fused multiply-adds, like the reference's own /fp:precise x64 build; with
            # contraction on, the finite differences of Dupire's formula (ivs.h:119-138) move by 1e-4 relative
            "g++", "-std=c++17", "-O3", "-march=x86-64-v3", contract, "-pthread", "-fPIC", "-shared", "-w",
            # patches 2 and 3: headers MSVC pulls in transitively
            "-include", "cstring", "-include", "functional", "-include", "algorithm",
            "-include", "vector", "-include", "string", "-include", "stdexcept",
            "-I", tmp, driver,
        ] + [os.path.join(tmp, s) for s in SOURCES] + ["-o", OUT_SO]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return OUT_SO


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--force", action="store_true")
    a = ap.parse_args()
    print(build(a.ref, a.force))
    print(build(a.ref, a.force, variant="fma"))
