"""Build recipe for libtfs_cuda.so (in-tree, sm_100a only).

nvcc cross-compiles without a GPU, so this runs in the CPU container; the resulting .so is
git-ignored but travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libtfs_cuda.so")
E2E = os.path.join(HERE, "host", "e2e_mirror")          # frame loop through include/fluid_sim.hpp (C++ host mirror)

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",            # never contract a*b+c: Rust does not
    "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC", "-shared",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh"))] + \
           [os.path.join(os.path.dirname(HERE), "include", "tfs.h"), os.path.join(os.path.dirname(HERE), "include", "fluid_sim.hpp"),
            os.path.join(HERE, "host", "e2e_mirror.cpp")]


def needs_build() -> bool:
    if not os.path.exists(LIB) or not os.path.exists(E2E):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(s) > t for s in sources())


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    extra = os.environ.get("TFS_NVCC_EXTRA", "").split()
    cmd = [_nvcc()] + NVCC_FLAGS + extra + ["-o", LIB, os.path.join(CSRC, "tfs_api.cu")]
    env = dict(os.environ)
    # the image exports CC/CXX wrappers that are not what nvcc wants as host compiler
    host = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else None
    if host:
        cmd[1:1] = ["-ccbin", host]
    res = subprocess.run(cmd, capture_output=True, text=True, env=env)
    log = res.stdout + res.stderr
    with open(os.path.join(HERE, "build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if res.returncode != 0:
        err = [l for l in log.splitlines() if 'error' in l.lower()]
        sys.stderr.write('\n'.join(err[-20:] or log.splitlines()[-20:]) + '\n')
        raise RuntimeError("nvcc failed building libtfs_cuda.so")
    if verbose:
        print(log)
    build_host_mirror(host or "g++")
    return LIB


def build_host_mirror(cxx: str = "g++") -> str:
    """the C++ host-side mirror's frame loop (host/e2e_mirror.cpp): plain g++, links the C ABI only"""
    inc = os.path.join(os.path.dirname(HERE), "include")
    cmd = [cxx, "-O2", "-std=c++17", "-I", inc, os.path.join(HERE, "host", "e2e_mirror.cpp"), "-o", E2E, LIB,
           "-Wl,-rpath,$ORIGIN/.."]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("g++ failed building host/e2e_mirror")
    return E2E


if __name__ == "__main__":
    print(build(force=True, verbose="-v" in sys.argv))
