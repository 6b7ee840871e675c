"""Compile physkit_b200/libpk_collide.so for sm_100a with nvcc (in-tree, so it travels with gpurun).

-fmad=false is part of the contract, not a tuning knob: results must be bit-identical to the
reference arithmetic (x86-64 SSE2, no FMA; reference CMakeLists.txt:126-132).
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libpk_collide.so")
SOURCES = ["pk_api.cu"]
HEADERS = sorted(f for f in os.listdir(CSRC) if f.endswith(".cuh")) + ["../../include/pk_collide.h"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-fmad=false",
    "--shared", "-Xcompiler", "-fPIC",
    "-Xcompiler", "-O3",
    "-Xcompiler", "-pthread",
    "-Xptxas", "-v",
    "-cudart", "static",
]


def nvcc_path() -> str:
    for p in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if p and (os.path.sep not in p or os.path.exists(p)):
            return p
    return "nvcc"


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(os.path.normpath(d)) > t for d in deps)


def build(force: bool = False, verbose: bool = False, out: str | None = None, defines=()) -> str:
    """out/defines build an experiment variant next to the product library (select it at run time with
    PK_COLLIDE_LIB=<path>); the product itself is always physkit_b200/libpk_collide.so with no defines."""
    target = LIB if out is None else os.path.join(HERE, out)
    if out is None and not force and not needs_build():
        return LIB
    cmd = [nvcc_path()] + NVCC_FLAGS + [f"-D{d}" for d in defines] + ["-o", target] + [os.path.join(CSRC, s) for s in SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = os.path.join(HERE, "build.log")
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed ({r.returncode}); see {log}")
    return target


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if a != "--force"]
    out = None
    defines = []
    while args:
        a = args.pop(0)
        if a == "--out":
            out = args.pop(0)
        elif a == "--define":
            defines.append(args.pop(0))
    print(build(force="--force" in sys.argv, verbose=out is None, out=out, defines=defines))
