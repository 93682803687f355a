"""Build libvhp_b200.so (CUDA kernels + C-ABI + C++ host boundary) in-tree with nvcc
for sm_100a, and the drop-in CLI.  No GPU is needed to build.

    python -m visibility_heuristic_path_planner_b200.build [--force]
"""
from __future__ import annotations

import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIBDIR = os.path.join(PKG, "lib")
LIB = os.path.join(LIBDIR, "libvhp_b200.so")
CLI = os.path.join(PKG, "bin", "visibility_heuristic_planner")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
HOST_CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-std=c++17", "-O3", "-lineinfo", "-fmad=false", "-ccbin", HOST_CXX,
          "-Xcompiler", "-fPIC,-O2,-ffp-contract=off,-Wall", "-I", os.path.join(ROOT, "include"),
          "-I", CSRC]


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC)
                  if f.endswith((".cu", ".cpp")) and not f.startswith("cli_"))


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(os.path.dirname(CLI), exist_ok=True)
    srcs = _sources()
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".hpp", ".cuh"))]
    hdrs += [os.path.join(ROOT, "include", f) for f in os.listdir(os.path.join(ROOT, "include"))]
    objs = []
    for s in srcs:
        o = os.path.join(LIBDIR, os.path.basename(s) + ".o")
        objs.append(o)
        if force or _newer(o, [s] + hdrs):
            cmd = [NVCC, *ARCH, *COMMON, "-x", "cu", "-c", s, "-o", o]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            subprocess.run(cmd, check=True)
    if force or _newer(LIB, objs):
        subprocess.run([NVCC, *ARCH, "-shared", "-ccbin", HOST_CXX, "-o", LIB, *objs,
                        "-lz", "-ldl", "-cudart", "static"], check=True)
    cli_src = os.path.join(CSRC, "cli_main.cpp")
    if os.path.exists(cli_src) and (force or _newer(CLI, [cli_src, LIB] + hdrs)):
        subprocess.run([HOST_CXX, "-std=c++20", "-O2", "-I", os.path.join(ROOT, "include"),
                        cli_src, "-o", CLI, "-L", LIBDIR, "-lvhp_b200",
                        "-Wl,-rpath,$ORIGIN/../lib"], check=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
