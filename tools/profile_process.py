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
# the same search budget with nodes side by side: K1 time = time inside solve_paths
from sgufp_solver_b200.explorer import solve_frontier  # noqa: E402
for width in (1, 8, 32):
    s2 = sg.GuroSolver(inst)
    k1_time = [0.0]
    orig = s2.solve_paths

    def timed(*a, **k):
        t1 = time.perf_counter()
        r = orig(*a, **k)
        k1_time[0] += time.perf_counter() - t1
        return r
    s2.solve_paths = timed
    t0 = time.perf_counter()
    b, n, c, calls = solve_frontier(s2, width=width, max_nodes=100000, max_cuts=cuts)      # the same number of cuts as the run above
    dt = time.perf_counter() - t0
    print(f"frontier width {width:2d}: {n} nodes, {c} cuts in {calls} K1 calls ({c / max(1, calls):.1f} candidates per call), best {b:.3f}, "
          f"total {dt * 1e3:.1f} ms, inside solve_paths {k1_time[0] * 1e3:.1f} ms ({k1_time[0] / max(1, c) * 1e6:.0f} us per cut)")
