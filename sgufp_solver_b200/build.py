"""Builds the in-tree CUDA library for sm_100a (nvcc cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libsgufp_b200.so")
SOURCES = ["capi.cu", "capi_shard.cu", "cache.cu", "capi_dd.cu", "k1_cut.cu", "k1_cut_small.cu", "k1_lane.cu", "k2_dd.cu", "k2_build.cu", "dd_host.cpp", "model.cpp"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall", "-shared",
              # K1 list search / push variants, A/B-timed on a B200 (profiles/r02_summary.md: C4 28.15 -> 27.21 ms, C2 unchanged)
              "-DSGUFP_K1_SKIP_CONFIRM", "-DSGUFP_K1_PUSH_PAR"]


def sources():
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", h) for h in ("sgufp_b200.h", "sgufp_b200_dd.h")]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build_library(force: bool = False, verbose: bool = False) -> str:
    if force or stale():
        nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + sources()
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    import sys
    print(build_library(force=True, verbose="-v" in sys.argv))
