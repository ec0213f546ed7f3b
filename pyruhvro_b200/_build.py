"""In-tree build of the native pieces (no network, no setuptools needed).

    libruhvro_b200.so   C-ABI library: CUDA kernels (sm_100a) + host engine      [nvcc]
    _native.*.so        CPython extension: list[bytes] packing + GIL handling     [g++]

Both land next to this file so they travel with the repo snapshot to the GPU box.
"""
from __future__ import annotations

import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libruhvro_b200.so")
EXT = os.path.join(HERE, "_native" + (sysconfig.get_config_var("EXT_SUFFIX") or ".so"))

LIB_SOURCES = ["engine.cu", "kernels.cu", "schema.cpp", "plan.cpp", "result.cpp"]
LIB_HEADERS = ["arrow_c.h", "json.hpp", "kernels.cuh", "plan.hpp", "result.hpp", "schema.hpp", "walker.cuh",
               os.path.join("..", "..", "include", "ruhvro_b200.h")]

NVCC_FLAGS = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC,-Wall", "-cudart", "static"]


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def build_lib(force: bool = False, verbose: bool = False) -> str:
    srcs = [os.path.join(CSRC, s) for s in LIB_SOURCES]
    deps = srcs + [os.path.join(CSRC, h) for h in LIB_HEADERS] + [os.path.abspath(__file__)]
    if force or _stale(LIB, deps):
        cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-shared", "-o", LIB] + srcs
        subprocess.check_call(cmd, cwd=CSRC)
    return LIB


def build_ext(force: bool = False) -> str:
    src = os.path.join(CSRC, "pymod.cpp")
    if force or _stale(EXT, [src, LIB, os.path.abspath(__file__)]):
        inc = sysconfig.get_paths()["include"]
        cmd = ["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-Wall", "-pthread", "-I", inc, "-I", os.path.join(HERE, "..", "include"),
               "-o", EXT, src, "-L", HERE, "-lruhvro_b200", "-Wl,-rpath,$ORIGIN"]
        subprocess.check_call(cmd, cwd=CSRC)
    return EXT


def build_all(force: bool = False, verbose: bool = False) -> None:
    build_lib(force, verbose)
    build_ext(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
    print(EXT)
