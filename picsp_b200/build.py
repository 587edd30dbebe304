"""Builds picsp_b200/libpicsp_b200.so in-tree with nvcc for sm_100a.

    python -m picsp_b200.build [--force] [--verbose]

The shared library is the product: CUDA kernels + the C ABI of include/picsp_b200.h
(+ the C++ host driver pieces).  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libpicsp_b200.so")
HOST_EXE = os.path.join(PKG, "picsp_b200_run")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function", "-ccbin", "g++",
          "-I", os.path.join(ROOT, "include")]


def _sources():
    cu = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cu")]
    host_dir = os.path.join(CSRC, "host")
    cpp = []
    if os.path.isdir(host_dir):
        cpp = [os.path.join(host_dir, f) for f in sorted(os.listdir(host_dir)) if f.endswith(".cpp") and f != "main.cpp"]
    return cu, cpp


def _deps():
    out = [os.path.join(ROOT, "include", "picsp_b200.h")]
    for d, _, files in os.walk(CSRC):
        out += [os.path.join(d, f) for f in files if f.endswith((".cu", ".cuh", ".cpp", ".h", ".hpp"))]
    return out


def _stale(target):
    if not os.path.isfile(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in _deps())


def build(force: bool = False, verbose: bool = False) -> str:
    cu, cpp = _sources()
    rebuilt = False
    if force or _stale(LIB):
        rebuilt = True
        extra = os.environ.get("PICSP_NVCC_DEFINES", "").split()     # e.g. "-DPICSP_CHUNK=4096" for tuning sweeps
        cmd = [NVCC, *ARCH, *COMMON, *extra, "-shared", "-o", LIB, *cu, *cpp, "-lcufft", "-ldl",
               "-Xlinker", "-rpath,/usr/local/cuda/lib64"]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    main_cpp = os.path.join(CSRC, "host", "main.cpp")
    if os.path.isfile(main_cpp) and (force or rebuilt or _stale(HOST_EXE)):   # never keep an executable older than the library
        cmd = ["g++", "-O2", "-std=c++17", "-Wall", "-I", os.path.join(ROOT, "include"), "-o", HOST_EXE, main_cpp,
               "-L", PKG, "-lpicsp_b200", "-Wl,-rpath,$ORIGIN", "-Wl,-rpath,/usr/local/cuda/lib64"]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(LIB)
