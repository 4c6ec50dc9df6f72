"""Builds libpilonb200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(CSRC, "libpilonb200.so")
SOURCES = ["pb_engine.cu", "pb_output.cpp", "pb_bam.cpp"]


def deps():
    """Every source the library is compiled from: all of csrc/*.cu, csrc/*.cuh and the public header."""
    out = [f for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh", ".cpp", ".hpp")) and f != "pb_synth.cpp"]
    return out + [os.path.join("..", "..", "include", "pilon_b200.h")]



def nvcc_path() -> str:
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found")
    return p


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in deps())


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return OUT
    cmd = [nvcc_path(), "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
           "-diag-suppress=20013,20015", "-Xcompiler", "-fPIC", "-shared", "-o", OUT] + [os.path.join(CSRC, s) for s in SOURCES] + ["-lz"]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    subprocess.check_call(cmd, cwd=CSRC)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
