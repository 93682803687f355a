#!/usr/bin/env python
"""Kernel-tuning helper: build another copy of libvhp_b200.so with extra nvcc flags
(e.g. -DVHP_STEP_UNROLL=4) into visibility_heuristic_path_planner_b200/lib_<name>/ so that
several variants can be timed in one GPU session (VHP_LIB_VARIANT=<name> selects one).

    python tools/build_variant.py <name> [nvcc flags ...]
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from visibility_heuristic_path_planner_b200 import build as B  # noqa: E402


def main():
    name, flags = sys.argv[1], sys.argv[2:]
    out = os.path.join(B.PKG, "lib_" + name)
    os.makedirs(out, exist_ok=True)
    B.build()  # the default build provides the objects that do not depend on the flags
    objs = []
    for s in B._sources():
        base = os.path.basename(s)
        if base in ("kernels_sweep_tile.cu", "kernels_planner.cu"):
            o = os.path.join(out, base + ".o")
            subprocess.run([B.NVCC, *B.ARCH, *B.COMMON, *flags, "-x", "cu", "-c", s, "-o", o], check=True)
        else:
            o = os.path.join(B.LIBDIR, base + ".o")
        objs.append(o)
    lib = os.path.join(out, "libvhp_b200.so")
    subprocess.run([B.NVCC, *B.ARCH, "-shared", "-ccbin", B.HOST_CXX, "-o", lib, *objs, "-lz",
                    "-cudart", "static"], check=True)
    print(lib)


if __name__ == "__main__":
    main()
