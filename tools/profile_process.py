"""Developer tool: where the time of the Benders loop goes (cProfile over explorer.solve on a GPU box).
    python tools/profile_process.py [c2|mid] [max_nodes]
"""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sgufp_solver_b200 as sg  # noqa: E402
from sgufp_solver_b200 import instances as I  # noqa: E402
from sgufp_solver_b200.explorer import solve  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
max_nodes = int(sys.argv[2]) if len(sys.argv) > 2 else 40
inst = I.config2(S=1000) if name == "c2" else I.make_layered([4, 5, 5, 4], 48, 12, 123, 0.7, 0.0, "mid")
solver = sg.GuroSolver(inst)
solve(solver, max_nodes=3)
t0 = time.perf_counter()
pr = cProfile.Profile()
pr.enable()
best, nodes, cuts = solve(solver, max_nodes=max_nodes)
pr.disable()
dt = time.perf_counter() - t0
print(f"{name}: {nodes} nodes, {cuts} cuts, best {best:.3f}, {dt * 1e3:.1f} ms  ({dt / max(1, nodes) * 1e3:.2f} ms per node)")
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
