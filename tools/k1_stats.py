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
    clk = (C.c_ulonglong * 16)()
    import bench
    FLOW = ("link", "row + start", "warm_init / labels + tight list", "searches", "pushes", "dual updates + tight lists", "flow + state out")
    CUT = ("row + flow in", "SPEC-LP potentials", "chains + alphas", "per-arc lifting", "closed chains + sums")
    for name in [a for a in sys.argv[1:] if not a.startswith("-")]:
        K = bench.WORKLOADS[name]["K"]
        inst = bench.scenario_range(name, 0, 10000 if name != "c2" else 1000)
        solver = GuroSolver(inst)
        # the bench's DD-emitted candidates (consecutive paths of the Benders loop) unless random matchings are asked for; the
        # counters are those of the LAST of three calls that alternate between the batches (the bench's steady state)
        if "--random" in sys.argv:
            batches = [np.asarray(I.random_paths(solver, K, 31, 0.1), dtype=np.int16)]
        else:
            batches = bench.step_batches(bench.candidate_paths(name, K, 0, want=2 * K)[0], K)
        calls = 1 if "--first" in sys.argv else 3
        for it in range(calls):
            L.sgufp_debug_k1_stats(out)
            L.sgufp_debug_k1_clk(clk)
            L.sgufp_debug_k1_lane_stats(lane)
            solver.solve_paths(batches[it % len(batches)])
        L.sgufp_debug_k1_stats(out)
        L.sgufp_debug_k1_lane_stats(lane)
        if lane[0]:      # the lane-per-scenario kernel ran: counters per block of 32 scenarios (warp level) and per scenario (lane level)
            print(f"{name}: lane kernel, blocks {lane[0]}  sweeps/block {lane[1] / lane[0]:.1f}  steps/block {lane[2] / lane[0]:.0f}  "
                  f"dual updates/scenario {lane[3] / max(1, lane[5]):.1f}  pushes/scenario {lane[4] / max(1, lane[5]):.1f}  ms {solver.last_kernel_ms():.3f}")
            continue
        L.sgufp_debug_k1_clk(clk)
        ft, ct = max(1, sum(clk[:7])), max(1, sum(clk[8:13]))
        print(f"{name}: group {os.environ.get('SGUFP_K1_GROUP', 'auto')}  warm {clk[14]} cold {clk[15]}  kernel {solver.last_kernel_ms():.3f} ms (stats build)")
        print(f"    flow kernel: {ft / max(1, clk[7]):.0f} clocks per evaluation: " + "  ".join(f"{n} {100.0 * clk[i] / ft:.1f}%" for i, n in enumerate(FLOW)))
        print(f"    cut kernel:  {ct / max(1, clk[13]):.0f} clocks per evaluation: " + "  ".join(f"{n} {100.0 * clk[8 + i] / ct:.1f}%" for i, n in enumerate(CUT)))
        ev = max(1, out[3])
        print(f"    evals {out[3]}  passes/eval {out[0] / ev:.1f}  label computations + dual updates/eval {out[1] / ev:.1f}  searches/eval {out[2] / ev:.1f}")
        if out[4]:
            print(f"    list searches: sweeps/search {out[4] / max(1, out[2]):.2f}  entries/sweep {out[5] / out[4]:.0f}  reached dst {100.0 * out[6] / max(1, out[2]):.0f} %  open chains/pass {out[7] / max(1, out[0]):.0f}")
        solver.close()


if __name__ == "__main__":
    main()
