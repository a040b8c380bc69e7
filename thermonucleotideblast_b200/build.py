"""Build the sm_100a shared library in-tree (thermonucleotideblast_b200/libtntb200.so).

nvcc cross-compiles without a GPU; the .so travels to the GPU box with the gpurun snapshot.
No -use_fast_math, no FMA contraction: the dH/dS/Tm arithmetic must round like the host reference.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libtntb200.so")

SOURCES = ["engine.cu", "assemble.cpp", "thermo.cpp", "postprocess.cpp"]
HEADERS = ["kernels.cuh", "fasta.cuh", "align_core.cuh", "tnt_types.h", "thermo.h", "assemble.h",
           "santalucia_tables.inc", os.path.join("..", "..", "include", "tntb200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-fmad=false",
    "-Xcompiler", "-fPIC,-O2,-ffp-contract=off,-fno-fast-math,-Wall,-Wno-unknown-pragmas",
    "-shared",
]


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    for f in SOURCES + HEADERS + [os.path.join("..", "build.py")]:
        if os.path.getmtime(os.path.join(CSRC, f)) > t:
            return True
    return False


def build(force: bool = False, verbose: bool = False, out: str = LIB, extra: list[str] | None = None) -> str:
    """`out` / `extra` build a variant library (extra nvcc flags, e.g. -DTNT_ALIGN_THREADS=32) for A/B runs."""
    if out == LIB and not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (extra or []) + (["-Xptxas", "-v"] if verbose else []) + \
        ["-o", out] + [os.path.join(CSRC, s) for s in SOURCES] + ["-lcudart"]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd, cwd=CSRC)
    return LIB


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if a != "-v"]
    out = LIB
    if args and args[0] == "-o":
        out = os.path.abspath(args[1])
        args = args[2:]
    print(build(force=True, verbose="-v" in sys.argv, out=out, extra=args))
