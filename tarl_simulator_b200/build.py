"""Builds libtarl_b200.so (sm_100a only) in-tree with nvcc. Run: python -m tarl_simulator_b200.build [--force]

The shared object stays inside the package directory so that it travels with a repo snapshot to the GPU box;
there is no JIT cache and no pip install involved.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(ROOT, "build", "obj")
LIB = os.path.join(PKG, "libtarl_b200.so")

COMMON = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC",
          "-I", os.path.join(ROOT, "include")]
# The integer/fp32 state kernels must round like ATen's scalar ops: no FMA contraction there.
PER_FILE = {
    "core_step.cu": ["-fmad=false"],
    "engine.cu": ["-fmad=false"],
    "agents.cu": ["-fmad=false"],
    "optim.cu": ["-fmad=false"],       # Adam / GAE in the operation order of the libraries they restate
}


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libtarl_b200.so cannot be built (there is no CPU fallback)")


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(ROOT, "include", f) for f in os.listdir(os.path.join(ROOT, "include"))]
    headers += [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    objs = []
    for name in sources():
        src = os.path.join(CSRC, name)
        obj = os.path.join(OBJ, name[:-3] + ".o")
        objs.append(obj)
        if force or _stale(obj, [src, __file__] + headers):
            cmd = [nvcc, *COMMON, *PER_FILE.get(name, []), "-c", src, "-o", obj]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
                print(" ".join(cmd))
            subprocess.run(cmd, check=True)
    if force or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs]
        if verbose:
            print(" ".join(cmd))
        subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv or "--verbose" in sys.argv))
