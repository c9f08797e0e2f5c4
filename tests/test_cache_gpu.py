"""The scenario-major cache file (csrc/cache.cu) and handles that share one copy of the capacities.
A handle made from the cache — the whole file or one rank's block — must give the cuts of a handle made from the arrays,
bit for bit; clones driven from 17 host threads at once (main.cpp:20 sweeps to 16 workers + the master) must give the serial
answers."""
import threading

import numpy as np
import pytest

import sgufp_solver_b200 as sg
from sgufp_solver_b200 import instances as I

pytestmark = pytest.mark.gpu


def _same(a, b):
    assert (a.cut_type == b.cut_type).all() and (a.first_infeasible == b.first_infeasible).all()
    assert (a.rhs == b.rhs).all() and (a.coef_dense == b.coef_dense).all() and (a.nnz == b.nnz).all()


@pytest.mark.parametrize("make", [lambda: I.config1(S=50, lower_prob=0.2), lambda: I.config2(S=333, lower_prob=0.02),
                                  lambda: I.make_layered([3, 4, 3], 21, 7, 77, 0.8, 0.1, "odd")], ids=["c1_lb", "c2_lb", "odd_m"])
def test_cache_round_trip(tmp_path, make):
    inst = make()
    path = str(tmp_path / "inst.sgufpc")
    I.save_cache(inst, path)
    d = I.cache_dims(path)
    assert (d["n"], d["m"], d["S"]) == (inst.n, inst.m, inst.S)
    a = sg.GuroSolver(inst)
    b = sg.GuroSolver.from_cache(path)
    assert (b.n, b.m, b.S, b.L, b.T) == (a.n, a.m, a.S, a.L, a.T) and b.layer_arc.tolist() == a.layer_arc.tolist()
    paths = I.random_paths(a, 6, 3, 0.25)
    ra, rb = a.solve_paths(paths), b.solve_paths(paths)
    _same(ra, rb)
    assert (ra.obj == rb.obj).all() and (ra.status == rb.status).all()


def test_cache_block_is_a_rank_of_a_partition(tmp_path):
    """Rows are contiguous in the file: a rank loads only its block.  Two blocks on one GPU, partial sums added on the host."""
    import ctypes as C
    import torch
    from sgufp_solver_b200 import _lib
    from sgufp_solver_b200.distributed import I64_MAX, finalize
    inst = I.config2(S=300)
    path = str(tmp_path / "c2.sgufpc")
    I.save_cache(inst, path)
    full = sg.GuroSolver(inst)
    paths = I.random_paths(full, 5, 41, 0.2)
    want = full.solve_paths(paths)
    tot = None
    for lo, n in ((0, 130), (130, 170)):
        part = sg.GuroSolver.from_cache(path, scenario_offset=lo, S_local=n)
        assert part.S == n and part.scenario_offset == lo and part.S_total == 300
        sums = torch.zeros((5, part.W), dtype=torch.int64, device="cuda")
        finf = torch.zeros((5,), dtype=torch.int64, device="cuda")
        rc = _lib.lib().sgufp_paths_partial(part.h, paths.ctypes.data_as(_lib.i16p), 5, paths.shape[1], C.c_void_p(sums.data_ptr()),
                                            C.c_void_p(finf.data_ptr()), None, None, None)
        assert rc == 0
        torch.cuda.synchronize()
        tot = sums.cpu().numpy() if tot is None else tot + sums.cpu().numpy()
        last = part
    res = finalize(last, paths, tot, np.full(5, I64_MAX, np.int64))
    assert (res.rhs == want.rhs).all() and (res.coef_dense == want.coef_dense).all()


def test_seventeen_threads_on_clones_of_one_upload():
    """N_WORKERS + 1 = 17 host threads, each with its own handle (NodeExplorer.h:115-116, main.cpp:20), one GPU, ONE copy of the
    capacities: every thread's cuts must be the serial ones, whatever the interleaving."""
    inst = I.config2(S=400, lower_prob=0.01)
    base = sg.GuroSolver(inst)
    T = 17
    clones = [base.clone() for _ in range(T)]
    batches = [I.random_paths(base, 3 + t % 4, 100 + t, 0.2) for t in range(T)]
    serial = [base.solve_paths(b) for b in batches]
    out, errs = [None] * T, []

    def work(t):
        try:
            for _ in range(6):
                out[t] = clones[t].solve_paths(batches[t])
                ctype, cut = clones[t].solveSubProblem(batches[t][0])
                assert ctype == serial[t].cut_type[0] and cut == serial[t].cut(0)
        except Exception as e:                      # noqa: BLE001 — reported below, in the main thread
            errs.append((t, repr(e)))
    ths = [threading.Thread(target=work, args=(t,)) for t in range(T)]
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    assert not errs, errs
    for t in range(T):
        _same(serial[t], out[t])
        assert (serial[t].obj == out[t].obj).all()
    base.close()                                    # the clones keep the capacities alive
    again = clones[3].solve_paths(batches[3])
    _same(serial[3], again)


def test_seventeen_benders_loops_at_once():
    """The whole caller side by side: 17 host threads, each running the Benders loop (K1 + K2, its own diagram) on its own clone."""
    from sgufp_solver_b200.explorer import solve
    inst = I.config1(S=40, lower_prob=0.1)
    base = sg.GuroSolver(inst)
    want = solve(base, max_nodes=300)
    T = 17
    clones = [base.clone() for _ in range(T)]
    got, errs = [None] * T, []

    def work(t):
        try:
            got[t] = solve(clones[t], max_nodes=300)
        except Exception as e:                      # noqa: BLE001
            errs.append((t, repr(e)))
    ths = [threading.Thread(target=work, args=(t,)) for t in range(T)]
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    assert not errs, errs
    assert all(g == want for g in got), (want, got)
