"""Developer tool (multi-GPU box): ONE process drives every visible GPU through sgufp_create_sharded (the C++ host's way:
one GuroSolver, N devices, the all-reduce inside the library) on the bench's C5 workload; wall time per solve_paths call,
and the cuts against a one-GPU handle over all scenarios.      python tools/time_single_process_partition.py [c5|c4]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import bench  # noqa: E402
from sgufp_solver_b200.solver import GuroSolver  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "c5"
n = torch.cuda.device_count()
K = bench.WORKLOADS[wl]["K"]
inst = bench.scenario_range(wl, 0, bench.totals(wl, 1))
paths, _ = bench.candidate_paths(wl, K, 0)
t0 = time.perf_counter()
one = GuroSolver(inst, device=0)
t_one = time.perf_counter() - t0
ref = one.solve_paths(paths, want_obj=False, want_status=False)
best1 = 1e9
for _ in range(4):
    t0 = time.perf_counter(); one.solve_paths(paths, want_obj=False, want_status=False); best1 = min(best1, time.perf_counter() - t0)
one.close()
t0 = time.perf_counter()
part = GuroSolver(inst, devices=list(range(n)))
t_part = time.perf_counter() - t0
res = part.solve_paths(paths, want_obj=False, want_status=False)
same = bool((res.rhs == ref.rhs).all() and (res.coef_dense == ref.coef_dense).all() and (res.nnz == ref.nnz).all())
bestn = 1e9
for _ in range(6):
    t0 = time.perf_counter(); part.solve_paths(paths, want_obj=False, want_status=False); bestn = min(bestn, time.perf_counter() - t0)
evals = K * inst.S
print(f"{wl}: one process, {n} GPUs: {bestn * 1e3:.2f} ms per call = {evals / bestn / 1e6:.2f} M evals/s; one GPU {best1 * 1e3:.2f} ms = {evals / best1 / 1e6:.2f} M; "
      f"speed-up {best1 / bestn:.2f}x; cuts {'bit-identical' if same else 'DIFFER'}; {part.comm_info()}; create {t_part:.2f} s (one GPU {t_one:.2f} s)")
