"""Developer tool: per-evaluation work counters of K1 (relaxation passes, label computations,
breadth-first searches) from a -DSGUFP_K1_STATS build of the library.

    python tools/k1_stats.py --build          # here (nvcc, no GPU): builds sgufp_solver_b200/libsgufp_b200_stats.so
    python tools/k1_stats.py c2 c4            # on a GPU box: runs the bench workloads once and prints the counters
"""
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sgufp_solver_b200 import build as B  # noqa: E402

STATS_LIB = os.path.join(B.HERE, "libsgufp_b200_stats.so")


def build():
    cmd = ["/usr/local/cuda/bin/nvcc"] + B.NVCC_FLAGS + ["-DSGUFP_K1_STATS", "-o", STATS_LIB] + B.sources()
    subprocess.check_call(cmd)


def main():
    if "--build" in sys.argv:
        build()
        return
    from sgufp_solver_b200 import _lib
    _lib.LIB_PATH = os.environ.get("SGUFP_STATS_LIB", STATS_LIB)
    import numpy as np
    from sgufp_solver_b200 import instances as I
    from sgufp_solver_b200.solver import GuroSolver
    L = _lib.lib()
    L.sgufp_debug_k1_stats.restype = C.c_int
    L.sgufp_debug_k1_stats.argtypes = [C.POINTER(C.c_ulonglong)]
    out = (C.c_ulonglong * 8)()
    L.sgufp_debug_k1_lane_stats.restype = C.c_int
    L.sgufp_debug_k1_lane_stats.argtypes = [C.POINTER(C.c_ulonglong)]
    lane = (C.c_ulonglong * 8)()
    L.sgufp_debug_k1_clk.restype = C.c_int
    L.sgufp_debug_k1_clk.argtypes = [C.POINTER(C.c_ulonglong)]
    clk = (C.c_ulonglong * 8)()
    import bench
    for name in [a for a in sys.argv[1:] if not a.startswith("-")]:
        inst, K = {"c2": (lambda: (I.config2(S=1000), 64)), "c4": (lambda: (I.config4(S=2000), 8))}[name]()
        solver = GuroSolver(inst)
        # the bench's DD-emitted candidates (consecutive paths of the Benders loop) unless random matchings are asked for
        paths = I.random_paths(solver, K, 31, 0.1) if "--random" in sys.argv else bench.candidate_paths(name, K, 0)[0]
        L.sgufp_debug_k1_stats(out)
        L.sgufp_debug_k1_clk(clk)
        L.sgufp_debug_k1_lane_stats(lane)
        solver.solve_paths(np.asarray(paths, dtype=np.int16))
        L.sgufp_debug_k1_stats(out)
        L.sgufp_debug_k1_lane_stats(lane)
        if lane[0]:      # the lane-per-scenario kernel ran: counters per block of 32 scenarios (warp level) and per scenario (lane level)
            print(f"{name}: lane kernel, blocks {lane[0]}  sweeps/block {lane[1] / lane[0]:.1f}  steps/block {lane[2] / lane[0]:.0f}  "
                  f"dual updates/scenario {lane[3] / max(1, lane[5]):.1f}  pushes/scenario {lane[4] / max(1, lane[5]):.1f}  ms {solver.last_kernel_ms():.3f}")
            continue
        L.sgufp_debug_k1_clk(clk)
        tot = max(1, sum(clk[:6]))
        print(f"{name}: group {os.environ.get('SGUFP_K1_GROUP', 'auto')}  warm {clk[6]} cold {clk[7]}  clocks per evaluation {tot / max(1, clk[6] + clk[7]):.0f}: "
              + "  ".join(f"{n} {100.0 * clk[i] / tot:.1f}%" for i, n in enumerate(("link", "stream", "repair", "from zero", "potentials", "lifting"))))
        ev = max(1, out[3])
        print(f"{name}: evals {out[3]}  passes/eval {out[0] / ev:.1f}  label computations/eval {out[1] / ev:.1f}  searches/eval {out[2] / ev:.1f}  ms {solver.last_kernel_ms():.3f}")
        if out[4]:
            print(f"    list searches: sweeps/search {out[4] / max(1, out[2]):.2f}  entries/sweep {out[5] / out[4]:.0f}  reached dst {100.0 * out[6] / max(1, out[2]):.0f} %  open chains/pass {out[7] / max(1, out[0]):.0f}")


if __name__ == "__main__":
    main()
