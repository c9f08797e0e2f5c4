"""Developer tool (GPU box): one K1 launch of a bench workload, for ncu:  python tools/run_k1_once.py c2|c4|c5 [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from sgufp_solver_b200 import instances as I  # noqa: E402
from sgufp_solver_b200.solver import GuroSolver  # noqa: E402

wl = sys.argv[1]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
import bench  # noqa: E402  (the bench's own instance blocks and DD-emitted candidate paths: what profiles/k1_traffic.json describes)
K = bench.WORKLOADS[wl]["K"]
inst = bench.scenario_range(wl, 0, bench.totals(wl, 1))
solver = GuroSolver(inst)
paths, _ = bench.candidate_paths(wl, K, 0)
for _ in range(reps):
    solver.solve_paths(paths, want_obj=False, want_status=False, want_dense=False)
print(wl, os.environ.get("SGUFP_K1_MODE", "auto"), "kernel ms", solver.last_kernel_ms())
