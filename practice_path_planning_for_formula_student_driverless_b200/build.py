"""In-tree build of csrc/libraceline_b200.so for sm_100a (nvcc cross-compiles without a GPU).

    python -m practice_path_planning_for_formula_student_driverless_b200.build [--force]
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libraceline_b200.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]
SOURCES = ["raceline_inst_256.cu", "raceline_inst_cluster.cu", "raceline_geom.cu", "raceline_inst_512.cu", "raceline_inst_128.cu", "raceline_inst_64.cu", "raceline_inst_32.cu", "raceline_inst_64x4.cu", "raceline_inst_128x4.cu", "raceline_inst_256x4.cu",
           "raceline_dispatch.cu", "raceline_api.cu", "synth_tracks.cpp"]
# the geometry stage reproduces the reference's x86-64 arithmetic bit for bit: no FMA contraction in that unit
EXTRA_FLAGS = {"raceline_geom.cu": ["-fmad=false"]}
HEADERS = [os.path.join(CSRC, "raceline_device.h"), os.path.join(CSRC, "raceline_kernels.cuh"), os.path.join(CSRC, "raceline_cluster.cuh"),
           os.path.join(HERE, "..", "include", "raceline_b200.h")]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: cannot build the CUDA extension")
    return exe


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, defines=(), suffix=""):
    """Compile every source for sm_100a and link the shared library. Returns its path.

    `defines` / `suffix` build a tuning or debug VARIANT next to the product library (objects and library carry the
    suffix, e.g. libraceline_b200_dbg.so built with -DRL_DEBUG_CHECKS); the product build uses neither."""
    nvcc = _nvcc()
    objs = [os.path.join(CSRC, os.path.splitext(src)[0] + suffix + ".o") for src in SOURCES]
    lib = os.path.join(CSRC, "libraceline_b200" + suffix + ".so")
    dflags = ["-D" + d for d in defines]

    def compile_one(src, obj):
        sp = os.path.join(CSRC, src)
        if not (force or _stale(obj, [sp] + HEADERS)):
            return
        cmd = [nvcc] + NVCC_FLAGS + dflags + EXTRA_FLAGS.get(src, []) + ["-c", sp, "-o", obj]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError(f"nvcc failed on {src}")
        with open(os.path.join(CSRC, os.path.splitext(src)[0] + suffix + ".ptxas.log"), "w") as f:
            f.write(res.stdout + res.stderr)
        if verbose:
            sys.stderr.write(f"built {obj}\n")

    # the per-thread-count kernel instantiations are independent translation units: compile them in parallel
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=max(1, min(len(SOURCES), os.cpu_count() or 1))) as ex:
        for fut in [ex.submit(compile_one, s, o) for s, o in zip(SOURCES, objs)]:
            fut.result()
    if force or _stale(lib, objs):
        cmd = [nvcc, "-shared", "-o", lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError("link failed")
    return lib


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--define", "-D", action="append", default=[])
    ap.add_argument("--suffix", default="")
    a = ap.parse_args()
    print(build(force=a.force, verbose=True, defines=a.define, suffix=a.suffix))
