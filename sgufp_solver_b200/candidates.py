"""Candidate first-stage paths as the decision-diagram master emits them ("K candidate paths ... obtained by
running RelaxedDDNew::getSolution() after applying previously generated cuts").

`dd_emitted_paths` runs the Benders loop of `explorer.solve` (NodeExplorer.cpp:915-986 over DDSolver.cpp:658-776's node
order) on a SMALL-scenario copy of the network and records every path the loop hands to `solveSubProblem`
(NodeExplorer.cpp:957) until K distinct ones are collected.  The matched fraction of such paths — which decides the size of
the contracted graph, hence the cost of an evaluation — is the DD's, not a random matching's.
"""
from __future__ import annotations

import time

import numpy as np


class _Enough(Exception):
    pass


def dd_emitted_paths(inst_small, K: int, device: int = 0, max_nodes: int = 4000, budget_s: float = 60.0):
    """Returns (paths [k, L] int16 with k <= K, info dict).  Deterministic: every rank of a partition computes the same list."""
    from .explorer import solve
    from .solver import GuroSolver
    solver = GuroSolver(inst_small, device=device)
    seen, order = set(), []
    orig = solver.solveSubProblem
    t0 = time.perf_counter()

    def spy(path):
        p = np.asarray(path, dtype=np.int16)
        key = p.tobytes()
        if key not in seen:
            seen.add(key)
            full = np.full(solver.L, -1, dtype=np.int16)
            full[: len(p)] = p
            order.append(full)
        if len(order) >= K or time.perf_counter() - t0 > budget_s:
            raise _Enough()
        return orig(path)

    solver.solveSubProblem = spy
    nodes = 0
    try:
        _, nodes, _ = solve(solver, max_nodes=max_nodes)
    except _Enough:
        pass
    finally:
        solver.solveSubProblem = orig
    paths = np.stack(order) if order else np.zeros((0, solver.L), np.int16)
    matched = float((paths >= 0).mean()) if len(paths) else 0.0
    return paths, {"emitted": int(len(paths)), "seconds": time.perf_counter() - t0, "matched_fraction": matched,
                   "scenarios_of_the_small_copy": int(inst_small.S)}
