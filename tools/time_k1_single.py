"""Developer tool: latency of a ONE-candidate K1 call (the reference's solveSubProblem shape) on C2 (GPU box)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sgufp_solver_b200 as sg  # noqa: E402
from sgufp_solver_b200 import instances as I  # noqa: E402

for S in (1000, 10000):
    solver = sg.GuroSolver(I.config2(S=S))
    paths = np.asarray(I.random_paths(solver, 16, 31, 0.1), dtype=np.int16)
    for k in range(4):
        solver.solveSubProblem(paths[k])
    ks, ws = [], []
    for k in range(16):
        t0 = time.perf_counter()
        solver.solveSubProblem(paths[k])
        ws.append(time.perf_counter() - t0)
        ks.append(solver.last_kernel_ms())
    print(f"S={S}: solveSubProblem wall {np.median(ws) * 1e3:.3f} ms, kernel {np.median(ks):.3f} ms")
