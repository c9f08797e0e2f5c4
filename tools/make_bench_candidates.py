"""Developer tool (GPU box): the candidate paths bench.py evaluates, as the DD master emits them
(sgufp_solver_b200/candidates.py), written to sgufp_solver_b200/data/bench_candidates.npz (and to gpurun_out/ so the
file comes back from a gpurun call).  tests/test_e2e_gpu.py::test_committed_bench_candidates_are_the_dd_emission re-emits
them and compares."""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from sgufp_solver_b200 import instances as I  # noqa: E402
from sgufp_solver_b200.candidates import dd_emitted_paths  # noqa: E402

out = {}
for net, K in (("config2", 64), ("config4", 16)):
    paths, info = dd_emitted_paths(getattr(I, net)(S=32), K, budget_s=120.0)
    assert len(paths) == K, (net, info)
    out[net] = paths
    print(net, info)
dst = os.path.join(ROOT, "sgufp_solver_b200", "data", "bench_candidates.npz")
os.makedirs(os.path.dirname(dst), exist_ok=True)
np.savez_compressed(dst, **out)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
shutil.copy(dst, os.path.join(ROOT, "gpurun_out", "bench_candidates.npz"))
