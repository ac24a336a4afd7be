"""Builds libssb.so in-tree with nvcc for sm_100a (no JIT cache, no torch extension machinery)."""
from __future__ import annotations

import os
import os.path as osp
import subprocess

PKG_DIR = osp.dirname(osp.abspath(__file__))
SRC = [osp.join(PKG_DIR, "csrc", "ssb_api.cu")]
DEPS = SRC + [osp.join(PKG_DIR, "csrc", f) for f in ("ssb_sim.cuh", "ssb_types.cuh")] + [
    osp.join(osp.dirname(PKG_DIR), "include", "ssb.h")]
OUT = osp.join(PKG_DIR, "_lib", "libssb.so")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    # event times must be single IEEE f64 additions (spark_sched_sim.py:610,632): never contract
    "-fmad=false",
    "-Xcompiler", "-fPIC", "-shared",
]


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and osp.exists(OUT) and all(osp.getmtime(d) <= osp.getmtime(OUT) for d in DEPS):
        return OUT
    os.makedirs(osp.dirname(OUT), exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT] + SRC
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force=True, verbose=True))
