"""Developer tool: per-call cost of the DD entry points on a relaxed C2 diagram (GPU box)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sgufp_solver_b200 as sg  # noqa: E402
from sgufp_solver_b200 import instances as I  # noqa: E402
from sgufp_solver_b200.dd import RelaxedDDNew, random_cut  # noqa: E402

solver = sg.GuroSolver(I.config2(S=1))
rng = np.random.default_rng(1)
cuts = [random_cut(solver, rng) for _ in range(512)]
d = RelaxedDDNew(solver)
d.buildTree()
print("nodes, arcs", d.counts(), "exact", d.isTreeExact())
d.buildTree()
for opt, tag in ((-1e300, "no pruning possible"), (None, "threshold 300 below the running bound")):
    d.buildTree()
    ub = 1e300
    t0 = time.perf_counter(); ks = []
    for c in cuts[:128]:
        b = d.applyOptimalityCut(c, opt if opt is not None else ub - 300, ub)
        ks.append(d.last_stats()[0])
        if b < -1e300:
            break
        ub = min(ub, b)
    dt = time.perf_counter() - t0
    print(f"single applies ({tag}): {dt / len(ks) * 1e3:.3f} ms wall per cut, kernels {np.mean(ks):.3f} ms, {len(ks)} cuts")
    d.buildTree()
    t0 = time.perf_counter()
    bounds, n = d.applyOptimalityCuts(cuts, opt if opt is not None else -1e300)
    dt = time.perf_counter() - t0
    print(f"one sequence call ({tag if opt is not None else 'no stop'}): {n} cuts, {dt / max(1, n) * 1e3:.4f} ms wall per cut, device span {d.last_stats()[0]:.3f} ms, launches {d.last_stats()[2]}")
t0 = time.perf_counter()
for _ in range(50):
    d.getSolution()
print(f"getSolution: {(time.perf_counter() - t0) / 50 * 1e3:.3f} ms")
