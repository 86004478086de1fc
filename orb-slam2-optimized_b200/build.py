"""Builds libransac_b200.so in-tree with nvcc for sm_100a.

One object per translation unit (csrc/engine*.cu), compiled in parallel, then linked.

Flags that matter:
  -gencode arch=compute_100a,code=sm_100a   B200 only, no other targets
  -fmad=false                               no FMA contraction by the compiler anywhere; code asks for
                                            fused multiply-adds explicitly (rfma / fmaf / __ffma2_rn).
                                            This is the arithmetic contract shared with the CPU checker
                                            
  -Xcompiler -ffp-contract=off              same for the host-compiled debug hooks
  -lineinfo                                 so that ncu's source page maps to these files
"""
from __future__ import annotations

import concurrent.futures
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
OUT = os.path.join(HERE, "libransac_b200.so")


def _sources():
    return sorted(glob.glob(os.path.join(CSRC, "engine*.cu")))


def _headers_mtime() -> float:
    t = os.path.getmtime(__file__)
    inc = os.path.join(os.path.dirname(HERE), "include")
    for d in (CSRC, inc):
        for root, _, files in os.walk(d):
            for f in files:
                if f.endswith((".cuh", ".h", ".hpp", ".inl")):
                    t = max(t, os.path.getmtime(os.path.join(root, f)))
    return t


def build(force: bool = False, verbose: bool = False) -> str:
    out = os.environ.get("RSAC_LIB_OUT", OUT)            # tuning variants: another file name, extra -D flags
    extra = os.environ.get("RSAC_EXTRA_NVCC", "").split()
    objdir = OBJ if not extra else OBJ + "_" + str(abs(hash(tuple(extra))) % 100000)
    os.makedirs(objdir, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    hdr_t = _headers_mtime()
    flags = ["-std=c++17", "-O3", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a",
             "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
             "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math,-O3"] + extra
    jobs, objs = [], []
    for src in _sources():
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(hdr_t, os.path.getmtime(src)):
            cmd = [nvcc, "-c", "-o", obj] + flags + [src]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        return cmd, r

    failed = False
    if jobs:
        with concurrent.futures.ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            for cmd, r in ex.map(run, jobs):
                if verbose or r.returncode != 0:
                    sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
                failed = failed or r.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed building libransac_b200.so")
    if jobs or not os.path.exists(out) or any(os.path.getmtime(o) > os.path.getmtime(out) for o in objs):
        cmd = [nvcc, "-shared", "-o", out, "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"] + objs + ["-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode != 0:
            sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError("link failed for libransac_b200.so")
    return out


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(OUT)
