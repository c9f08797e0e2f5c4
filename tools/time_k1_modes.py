"""Developer tool (GPU box): K1 kernel time of the two kernels side by side on the bench workloads.

    python tools/time_k1_modes.py c2 c4 c5        # SGUFP_K1_MODE=warp / lane, best of 5 after 2 warm-ups
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from sgufp_solver_b200 import instances as I  # noqa: E402
from sgufp_solver_b200.solver import GuroSolver  # noqa: E402

WL = {"c2": lambda: (I.config2(S=1000), 64), "c4": lambda: (I.config4(S=10000), 8), "c5": lambda: (I.config4(S=100000), 8)}


def main():
    for wl in [a for a in sys.argv[1:] if a in WL]:
        inst, K = WL[wl]()
        solver = GuroSolver(inst)
        paths = np.asarray(I.random_paths(solver, K, 31, 0.1), dtype=np.int16)
        ref = None
        for mode in ("warp", "lane"):
            os.environ["SGUFP_K1_MODE"] = mode
            best = 1e9
            for it in range(5 if wl != "c5" else 3):
                res = solver.solve_paths(paths, want_obj=False, want_status=False, want_dense=True)
                if it >= 1:
                    best = min(best, solver.last_kernel_ms())
            same = "" if ref is None else ("  bit-identical to warp" if (ref.rhs == res.rhs).all() and (ref.coef_dense == res.coef_dense).all() else "  DIFFERS FROM WARP")
            if ref is None:
                ref = res
            print(f"{wl} {mode:5s}: {best:9.3f} ms  ({K * inst.S / best / 1e3:8.2f} M evals/s){same}", flush=True)


if __name__ == "__main__":
    main()
