"""Developer tool (GPU box): K1 kernel time of library variants (tools/k1_variants.py --build) in the bench's steady state —
DD-emitted candidates, the steps alternating between the batches, state between the steps.

    python tools/time_k1_variants_bench.py base arc2 st2 -- c4 c5
"""
import hashlib
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one(name, wls):
    from sgufp_solver_b200 import _lib, build as B
    _lib.LIB_PATH = B.LIB if name == "base" else os.path.join(B.HERE, f"libsgufp_b200_{name}.so")
    import numpy as np
    import bench
    from sgufp_solver_b200.solver import GuroSolver
    for wl in wls:
        w = bench.WORKLOADS[wl]
        S = bench.totals(wl, 1)
        inst = bench.scenario_range(wl, 0, S)
        paths, _ = bench.candidate_paths(wl, w["K"], 0, want=2 * w["K"])
        batches = bench.step_batches(paths, w["K"])
        solver = GuroSolver(inst)
        ms, h = [], hashlib.sha1()
        for it in range(9):
            r = solver.solve_paths(batches[it % len(batches)], want_obj=False, want_status=False, want_dense=True)
            if it >= 3:
                ms.append(solver.last_kernel_ms())
            if it < 2:
                h.update(np.ascontiguousarray(r.rhs).tobytes()); h.update(np.ascontiguousarray(r.coef_dense).tobytes())
        print(f"{name:8s} {wl} S={S} K={w['K']}: mean {np.mean(ms):8.3f} ms  min {np.min(ms):8.3f} ms  {w['K'] * S / np.mean(ms) / 1e3:7.2f} M evals/s  cuts {h.hexdigest()[:12]}", flush=True)
        solver.close()


if __name__ == "__main__":
    a = sys.argv[1:]
    if a[0] == "--one":
        one(a[1], a[2:])
    else:
        names, wls = a[:a.index("--")], a[a.index("--") + 1:]
        for n in names:
            subprocess.call([sys.executable, __file__, "--one", n] + wls)
