"""Builds the native library in-tree:  <package>/libplslam.so  (sm_100a only, nvcc cross-compiles without a GPU).

    python -m <package>.build            (or: python <package>/build.py)

Flags that matter:
  -gencode arch=compute_100a,code=sm_100a   the only target; there is no other backend and no CPU fallback
  -fmad=false                               no FMA contraction: cvRound() decisions must match the reference CPU path
  -lineinfo                                 so `ncu --import-source on` maps stalls to source lines
"""
from __future__ import annotations

import concurrent.futures
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libplslam.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false", "-std=c++17",
              "-Xcompiler", "-fPIC,-fvisibility=hidden,-O3", "--expt-relaxed-constexpr"]


def _nvcc() -> str:
    nv = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nv):
        raise RuntimeError("nvcc not found: the native sm_100a library cannot be built (there is no fallback path)")
    return nv


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps():
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    inc = os.path.join(HERE, "..", "include")
    deps += [os.path.join(inc, f) for f in os.listdir(inc)]
    hostdir = os.path.join(HERE, "host")
    if os.path.isdir(hostdir):
        deps += [os.path.join(hostdir, f) for f in os.listdir(hostdir)]
    return deps


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in _deps())


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    nv = _nvcc()
    extra = ["-Xptxas", "-v"] if verbose else []
    extra += os.environ.get("PLSLAM_NVCC_EXTRA", "").split()  # developer switches, e.g. -DPL_LSD_PROF2

    def compile_one(src):
        obj = os.path.join(OBJ, os.path.basename(src) + ".o")
        cmd = [nv, *NVCC_FLAGS, *extra, "-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        return obj, r.stderr

    with concurrent.futures.ThreadPoolExecutor(max_workers=8) as ex:
        res = list(ex.map(compile_one, sources()))
    if verbose:
        for _, err in res:
            sys.stderr.write(err)
    objs = [o for o, _ in res]
    cmd = [nv, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC",
           "-lcudart_static", "-lpthread", "-ldl", "-lrt"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed: " + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
