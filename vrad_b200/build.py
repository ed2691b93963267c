"""Builds libvradcuda.so in-tree (vrad_b200/_lib/) with nvcc for sm_100a.

nvcc cross-compiles without a GPU, so this runs in the CPU-only build container as well as on
the B200 box.  fp contract: -fmad=false (no FMA contraction in device code), default
-prec-div/-prec-sqrt, no fast-math, host code with -ffp-contract=off -- see csrc/common.cuh.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "_lib")
LIB = os.path.join(OUT_DIR, "libvradcuda.so")

SOURCES = ["vrad_env.cu", "k1_trace.cu", "k1_sky.cu", "k2_transfers.cu", "k3_direct.cu", "k4_bounce.cu", "comm.cu", "kd_builder.cpp", "patch_subdivide.cpp", "light_setup.cpp",
           "bsp_file.cpp", "bsp_input.cpp", "bsp_light.cpp", "k5_finalize.cu", "kd_fast.cu", "texlights.cpp", "radial.cu", "group.cu"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "vrad_cuda.h"),
                                                               os.path.join(HERE, "..", "include", "vrad_bsp.h"), __file__]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    host_cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    objs = []
    common = [_nvcc(), "-ccbin", host_cxx, "-std=c++17", "-O3", "-lineinfo",
              "-gencode", "arch=compute_100a,code=sm_100a", "-fmad=false",
              "-Xcompiler", "-fPIC,-fopenmp,-ffp-contract=off,-fno-fast-math,-Wall",
              "-Xptxas", "-v" if verbose else "-O3"]
    procs = []
    for src in SOURCES:
        obj = os.path.join(OUT_DIR, os.path.splitext(src)[0] + ".o")
        objs.append(obj)
        cmd = common + ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"--- {src} ---\n{out}\n")
        elif verbose:
            sys.stderr.write(f"--- {src} ---\n{out}\n")
    if failed:
        raise RuntimeError("nvcc failed")
    link = [_nvcc(), "-ccbin", host_cxx, "-shared", "-o", LIB] + objs + ["-Xcompiler", "-fopenmp", "-lgomp", "-ldl"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
