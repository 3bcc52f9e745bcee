"""Build the native libraries in-tree.

    python -m lmrl_gym_b200.build            # liblmrl_b200.so (nvcc, sm_100a) + libenvcore_host.so (g++)

``liblmrl_b200.so`` is the product (C ABI of include/lmrl_b200.h).  ``libenvcore_host.so`` is a host build of
csrc/env_core.h used only by the CPU test-suite.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "liblmrl_b200.so")
HOSTLIB = os.path.join(CSRC, "libenvcore_host.so")
CU_SOURCES = ["lmrl_b200.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "--use_fast_math",
              "-Xcompiler", "-fPIC", "-shared", "-lcuda", "-Xptxas", "-v"]


def _newer(target, deps):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps)


def _deps():
    out = [os.path.join(HERE, "..", "include", "lmrl_b200.h")]
    for fn in os.listdir(CSRC):
        if fn.endswith((".cu", ".cuh", ".h", ".cpp", ".inl")):
            out.append(os.path.join(CSRC, fn))
    return out


def build_cuda(force=False, verbose=False):
    if not force and _newer(LIB, _deps()):
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cmd = [nvcc] + NVCC_FLAGS + [os.path.join(CSRC, s) for s in CU_SOURCES] + ["-o", LIB]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = res.stdout + res.stderr
    with open(os.path.join(CSRC, "build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if res.returncode != 0:
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed building liblmrl_b200.so")
    if verbose:
        print(log)
    return LIB


def build_host(force=False):
    if not force and _newer(HOSTLIB, _deps()):
        return HOSTLIB
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", os.path.join(CSRC, "envcore_host.cpp"), "-o", HOSTLIB]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("g++ failed building libenvcore_host.so")
    return HOSTLIB


if __name__ == "__main__":
    print(build_cuda(force="--force" in sys.argv, verbose=True))
    print(build_host(force="--force" in sys.argv))
