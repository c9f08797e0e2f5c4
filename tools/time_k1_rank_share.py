"""Developer tool (GPU box): K1 kernel time of ONE rank's share of the C5 strong-scaling run (scenarios [0, 100000/N) of the bench's
instance, the bench's candidates and steady state) on one GPU: what an N-GPU step costs before the exchange.

    python tools/time_k1_rank_share.py 8 4 2
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ctypes as C  # noqa: E402
import numpy as np  # noqa: E402
import torch  # noqa: E402
import bench  # noqa: E402
from sgufp_solver_b200 import _lib  # noqa: E402
from sgufp_solver_b200.solver import GuroSolver  # noqa: E402

K = bench.WORKLOADS["c5"]["K"]
paths, _ = bench.candidate_paths("c5", K, 0, want=2 * K)
batches = bench.step_batches(paths, K)
for n in [int(a) for a in sys.argv[1:]]:
    S = bench.totals("c5", n) // n
    solver = GuroSolver(bench.scenario_range("c5", 0, S), device=0, scenario_offset=0, S_total=bench.totals("c5", n))
    ms = []
    sums = torch.empty((K, solver.W), dtype=torch.int64, device="cuda:0")
    finf = torch.empty((K,), dtype=torch.int64, device="cuda:0")
    for it in range(11):
        p = batches[it % len(batches)]
        solver._check(_lib.lib().sgufp_paths_partial(solver.h, p.ctypes.data_as(_lib.i16p), K, p.shape[1], C.c_void_p(sums.data_ptr()), C.c_void_p(finf.data_ptr()), None, None, None))
        torch.cuda.synchronize()
        if it >= 3:
            ms.append(solver.last_kernel_ms())
    print(f"C5 share of one rank of {n}: S = {S}, run length {solver.run_length(K)}: K1 kernels mean {np.mean(ms):.3f} ms  min {np.min(ms):.3f} ms  "
          f"-> {K * S * n / np.mean(ms) / 1e3:.1f} M evals/s over {n} GPUs before the exchange", flush=True)
    solver.close()
