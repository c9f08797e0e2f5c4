"""Developer tool (GPU box): how long does the DD master take to emit K candidate paths on the bench networks?"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from sgufp_solver_b200 import instances as I  # noqa: E402
from sgufp_solver_b200.candidates import dd_emitted_paths  # noqa: E402

for name, inst, K in (("c2", I.config2(S=32), 64), ("c4", I.config4(S=32), 8)):
    t0 = time.perf_counter()
    paths, info = dd_emitted_paths(inst, K)
    rnd = I.random_paths(__import__("sgufp_solver_b200").GuroSolver(inst), K, 31, 0.1)
    print(name, info, "wall %.1f s" % (time.perf_counter() - t0), "random-path matched fraction %.3f" % float((rnd >= 0).mean()), flush=True)
