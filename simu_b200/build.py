"""Build the sm_100a shared library that carries the kernels and the C ABI.

``python -m simu_b200.build`` compiles ``simu_b200/csrc/*.cu`` with nvcc into
``simu_b200/lib/libsimu_b200.so`` (in-tree, git-ignored, shipped to the GPU box
with the snapshot).  Cross-compilation needs no GPU.
"""
from __future__ import annotations

import glob
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIBDIR = os.path.join(PKG, "lib")
LIB = os.path.join(LIBDIR, "libsimu_b200.so")
HOSTDIR = os.path.join(PKG, "host")
CLI = os.path.join(PKG, "bin", "simu_b200")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-O3",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the sm_100a library cannot be built")


def _stale(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build_library(force: bool = False, verbose: bool = False) -> str:
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    deps = srcs + glob.glob(os.path.join(CSRC, "*.cuh")) + [os.path.join(PKG, "..", "include", "simu_b200.h")]
    if not force and not _stale(LIB, deps):
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    objdir = os.path.join(PKG, "build")
    os.makedirs(objdir, exist_ok=True)
    objs = []
    procs = []
    for s in srcs:
        o = os.path.join(objdir, os.path.basename(s)[:-3] + ".o")
        objs.append(o)
        if force or _stale(o, [s] + deps[len(srcs):]):
            cmd = [_nvcc(), *NVCC_FLAGS, "-c", s, "-o", o]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
                print(" ".join(cmd), flush=True)
            procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for s, p in procs:
        out, _ = p.communicate()
        if out.strip() and (verbose or p.returncode):
            print(out, file=sys.stderr)
        if p.returncode:
            failed = True
    if failed:
        raise RuntimeError("nvcc failed")
    cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB, *objs, "-cudart", "static"]
    subprocess.run(cmd, check=True)
    return LIB


def build_cli(force: bool = False) -> str | None:
    """The C++ host driver (drop-in `simu` command line) above the C ABI."""
    srcs = sorted(glob.glob(os.path.join(HOSTDIR, "*.cc")))
    if not srcs:
        return None
    deps = srcs + glob.glob(os.path.join(HOSTDIR, "*.h")) + [LIB]
    if not force and not _stale(CLI, deps):
        return CLI
    os.makedirs(os.path.dirname(CLI), exist_ok=True)
    cmd = ["g++", "-O2", "-std=c++17", "-Wall", "-I", os.path.join(PKG, "..", "include"), *srcs, "-o", CLI,
           "-L", LIBDIR, "-lsimu_b200", "-Wl,-rpath,$ORIGIN/../lib", "-ldl", "-lpthread"]
    subprocess.run(cmd, check=True)
    return CLI


if __name__ == "__main__":
    force = "--force" in sys.argv
    print(build_library(force=force, verbose="-v" in sys.argv))
    c = build_cli(force=force)
    if c:
        print(c)
