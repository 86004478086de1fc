"""Builds libransac_b200.so in-tree with nvcc for sm_100a.

Flags that matter:
  -gencode arch=compute_100a,code=sm_100a   B200 only, no other targets
  -fmad=false                               no FMA contraction by the compiler anywhere; code asks for
                                            fused multiply-adds explicitly (rfma / fmaf / __ffma2_rn).
                                            This is the arithmetic contract shared with the CPU checker
                                            
  -Xcompiler -ffp-contract=off              same for the host-compiled debug hooks
  -lineinfo                                 so that ncu's source page maps to these files
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libransac_b200.so")
SOURCES = ["engine.cu"]   # one translation unit; engine_*.inl are included by it


def _newer(src_dir: str, out: str) -> bool:
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    for root, _, files in os.walk(src_dir):
        for f in files:
            if os.path.getmtime(os.path.join(root, f)) > t:
                return True
    inc = os.path.join(os.path.dirname(HERE), "include")
    for root, _, files in os.walk(inc):
        for f in files:
            if os.path.getmtime(os.path.join(root, f)) > t:
                return True
    return os.path.getmtime(__file__) > t


def build(force: bool = False, verbose: bool = False) -> str:
    out = os.environ.get("RSAC_LIB_OUT", OUT)            # tuning variants: another file name, extra -D flags
    extra = os.environ.get("RSAC_EXTRA_NVCC", "").split()
    if not force and not extra and not _newer(CSRC, out):
        return out
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    cmd = [nvcc, "-shared", "-o", out, "-std=c++17", "-O3", "-lineinfo",
           "-gencode", "arch=compute_100a,code=sm_100a",
           "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
           "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math,-O3",
           "-cudart", "static"] + extra + srcs + ["-ldl"]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd))
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building libransac_b200.so")
    return out


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(OUT)
