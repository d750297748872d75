"""Builds lpformer_b200/_C/liblpformer_b200.so in-tree with nvcc for sm_100a.

The shared library is a plain C-ABI object (include/lpformer_b200.h); it links the CUDA
runtime statically and nothing from torch.  Also builds the host-side PPR tool
(csrc/ppr_push.cpp, g++).  `python -m lpformer_b200.build` or __graft_entry__.build().
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
OUT = os.path.join(PKG, "_C")
LIB = os.path.join(OUT, "liblpformer_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]
CXX_FLAGS = ["-O3", "-std=c++17", "-fPIC", "-pthread"]


def _sources():
    cu = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))
    cpp = sorted(f for f in os.listdir(CSRC) if f.endswith(".cpp"))
    return cu, cpp


def _stamp():
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(os.path.dirname(PKG), "include")):
        for f in sorted(os.listdir(root)):
            with open(os.path.join(root, f), "rb") as fh:
                h.update(f.encode() + b"\0" + fh.read())
    h.update(" ".join(NVCC_FLAGS + CXX_FLAGS).encode())
    return h.hexdigest()


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("build failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
    return r.stdout + r.stderr


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OUT, exist_ok=True)
    stamp_file = os.path.join(OUT, "stamp")
    stamp = _stamp()
    if not force and os.path.exists(LIB) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
        return LIB
    cu, cpp = _sources()
    objs, jobs = [], []
    for f in cu:
        o = os.path.join(OUT, f[:-3] + ".o")
        objs.append(o)
        jobs.append([NVCC, *NVCC_FLAGS, "-Xptxas", "-v", "-c", os.path.join(CSRC, f), "-o", o])
    for f in cpp:
        o = os.path.join(OUT, f[:-4] + ".o")
        objs.append(o)
        jobs.append(["g++", *CXX_FLAGS, "-c", os.path.join(CSRC, f), "-o", o])
    with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
        logs = list(ex.map(_run, jobs))
    if verbose:
        print("\n".join(logs))
    with open(os.path.join(OUT, "ptxas.log"), "w") as fh:
        fh.write("\n".join(logs))
    _run([NVCC, "-shared", "-o", LIB, *objs, "-cudart", "static", "-lpthread", "-ldl", "-lrt"])
    with open(stamp_file, "w") as fh:
        fh.write(stamp)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
