"""Builds ni_slam_b200/lib/libnislam.so (hand-written CUDA for sm_100a + the C ABI of include/nislam.h) in-tree with nvcc.

    python -m ni_slam_b200.build [--force]

The five translation units are compiled in parallel; the shared object links the static CUDA runtime so it has no
dependency beyond the driver.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# NIS_BUILD_TAG=<tag> (with NIS_NVCC_EXTRA="-D...") builds a compile-time variant beside the product library for A/B runs on the
# GPU box (select it with NIS_LIB=ni_slam_b200/lib/libnislam_<tag>.so)
TAG = os.environ.get("NIS_BUILD_TAG", "")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(HERE, "build" + ("_" + TAG if TAG else ""))
LIB = os.path.join(LIBDIR, "libnislam%s.so" % ("_" + TAG if TAG else ""))
UNITS = ["nis_col.cu", "nis_row.cu", "nis_misc.cu", "nis_api.cu", "nis_stitch.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
              "--threads", "4"]


def _newest_source() -> float:
    t = 0.0
    for d in (CSRC, os.path.join(HERE, "..", "include")):
        for f in os.listdir(d):
            t = max(t, os.path.getmtime(os.path.join(d, f)))
    return t


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= _newest_source():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    nvcc = os.environ.get("NVCC", "nvcc")
    extra = os.environ.get("NIS_NVCC_EXTRA", "").split()

    def compile_one(unit):
        obj = os.path.join(OBJDIR, unit.replace(".cu", ".o"))
        cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, unit), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (unit, r.stdout, r.stderr))
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=4) as ex:
        objs = list(ex.map(compile_one, UNITS))
    r = subprocess.run([nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-ldl"],
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
