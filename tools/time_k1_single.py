"""Developer tool: latency of a ONE-candidate K1 call (the reference's solveSubProblem shape) on C2 and C4 (GPU box), for
paths a few layers apart (the bench's DD-emitted candidates: consecutive paths of the Benders loop).  SGUFP_K1_STATE=0 turns
the state between calls off (every call from zero flow)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import sgufp_solver_b200 as sg  # noqa: E402
from sgufp_solver_b200 import instances as I  # noqa: E402

for wl, S in (("c2", 1000), ("c2", 10000), ("c4", 1000), ("c4", 10000)):
    solver = sg.GuroSolver(I.config2(S=S) if wl == "c2" else I.config4(S=S))
    paths, _ = bench.candidate_paths(wl, 16, 0)
    for k in range(4):
        solver.solveSubProblem(paths[k])
    ks, ws = [], []
    for rep in range(3):
        for k in range(16):
            t0 = time.perf_counter()
            solver.solveSubProblem(paths[k])
            ws.append(time.perf_counter() - t0)
            ks.append(solver.last_kernel_ms())
    print(f"{wl} S={S} state={os.environ.get('SGUFP_K1_STATE', '1')}: solveSubProblem wall {np.median(ws) * 1e3:.3f} ms, kernel {np.median(ks):.3f} ms", flush=True)
    solver.close()
