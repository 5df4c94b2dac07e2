"""Builds newmsm_b200/lib/libmsmgpu.so (the C-ABI library, include/msmgpu.h) with nvcc for sm_100a.

In-tree and explicit: `nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo --fmad=false`.
--fmad=false is part of the numerical contract (csrc/geom.cuh): FP64 decisions must be made
from separately rounded products and sums, like the reference's -O2 x86-64 build.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libmsmgpu.so")
OBJDIR = os.path.join(HERE, "build")
SOURCES = ["api.cu", "octree_build.cu", "query.cu", "weights.cu", "cost.cu", "triplet.cu", "group.cu"]
HEADERS = ["common.cuh", "geom.cuh", "query.cuh", "cost.cuh", os.path.join("..", "..", "include", "msmgpu.h")]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def _host_cxx() -> str:
    # the image's $CXX wrapper lacks libgomp.spec; the distro compiler has it
    return "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else (shutil.which("g++") or "g++")


def _flags(extra=()):
    return ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "--fmad=false", "-std=c++17",
            "-ccbin", _host_cxx(), "-Xcompiler", "-fPIC,-fopenmp,-O2,-ffp-contract=off", *extra]


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build_library(force: bool = False, verbose: bool = False) -> str:
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    hdrs = [os.path.normpath(os.path.join(CSRC, h)) for h in HEADERS]
    os.makedirs(OBJDIR, exist_ok=True)
    os.makedirs(LIBDIR, exist_ok=True)
    nvcc = _nvcc()
    jobs = []
    objs = []
    for s in srcs:
        o = os.path.join(OBJDIR, os.path.basename(s)[:-3] + ".o")
        objs.append(o)
        if force or _stale(o, [s, *hdrs, __file__]):
            extra = ["-Xptxas", "-v"] if verbose else []
            jobs.append([nvcc, *_flags(extra), "-c", s, "-o", o])
    if jobs:
        with cf.ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            for cmd, res in zip(jobs, ex.map(lambda c: subprocess.run(c, capture_output=True, text=True), jobs)):
                if verbose or res.returncode:
                    sys.stderr.write(res.stdout + res.stderr)
                if res.returncode:
                    raise RuntimeError("nvcc failed: " + " ".join(cmd))
    if jobs or force or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-ccbin", _host_cxx(), "-gencode", "arch=compute_100a,code=sm_100a",
               "-Xcompiler", "-fPIC,-fopenmp", "-o", LIB, *objs, "-lgomp"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError("link failed: " + " ".join(cmd))
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
