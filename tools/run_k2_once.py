"""Developer tool (GPU box): one batched K2 launch pair on 64 DISTINCT diagrams of one B&B frontier x 64 cuts, for ncu:
    python tools/run_k2_once.py restricted 1024   |   relaxed 120"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from sgufp_solver_b200 import instances as I  # noqa: E402
from sgufp_solver_b200.dd import RelaxedDDNew, RestrictedDDNew, apply_optimality_batch, frontier_nodes, random_cut  # noqa: E402
from sgufp_solver_b200.solver import GuroSolver  # noqa: E402

kind, w = sys.argv[1], int(sys.argv[2])
solver = GuroSolver(I.config2(S=1))
rng = np.random.default_rng(5)
nodes = frontier_nodes(solver, 64)
cuts = [random_cut(solver, rng) for _ in range(64)]
dds = []
for i in range(64):
    d = RestrictedDDNew(solver, w) if kind == "restricted" else RelaxedDDNew(solver)
    d.compile(nodes[i % len(nodes)]) if kind == "restricted" else d.buildTree(nodes[i % len(nodes)])
    dds.append(d)
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
for _ in range(2):
    flush.zero_(); torch.cuda.synchronize()
    apply_optimality_batch(dds, cuts)
ms, arcs, launches = dds[0].last_stats()
cnt = [d.counts() for d in dds]
print(kind, w, "kernel ms", ms, "arcs per launch", arcs, "nodes", sum(c[0] for c in cnt), "arcs of the diagrams", sum(c[1] for c in cnt), "distinct roots", len(nodes))
