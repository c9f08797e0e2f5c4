"""Scenario partition over 2 GPUs with NCCL (one process per GPU): the cuts must be bit-identical to
the one-GPU cuts, feasibility cuts included (ray built by the owning rank, broadcast).  Skipped on a
one-GPU box; the host logic of the same path is covered on CPU by tests/test_sharded_cpu.py."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, lower_prob, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, "tests")]
    import torch
    import torch.distributed as dist
    from sgufp_solver_b200 import instances as I
    from sgufp_solver_b200.distributed import ShardedGuroSolver
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    inst = I.config2(S=301, lower_prob=lower_prob)
    sh = ShardedGuroSolver(inst, inst.S, rank, world, device=rank)
    paths = I.random_paths(sh.solver, 6, 9, 0.25)
    res = sh.solve_paths(paths)
    ctype, cut = sh.solveSubProblem(paths[0])
    if rank == 0:
        q.put((paths, res.cut_type, res.rhs, res.coef_dense, res.first_infeasible, ctype, cut.RHS, cut.keys, cut.vals))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("lower_prob", [0.0, 0.02])
def test_two_gpus_equal_one_gpu(lower_prob):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    import sgufp_solver_b200 as sg
    from sgufp_solver_b200 import instances as I
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, lower_prob, q)) for r in range(2)]
    for p in procs:
        p.start()
    paths, ct, rhs, dense, finf, ctype0, rhs0, keys0, vals0 = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    inst = I.config2(S=301, lower_prob=lower_prob)
    one = sg.GuroSolver(inst, device=0)
    ref = one.solve_paths(paths)
    assert (ct == ref.cut_type).all() and (finf == ref.first_infeasible).all()
    assert (rhs == ref.rhs).all() and (dense == ref.coef_dense).all()              # bit-identical
    t0, c0 = one.solveSubProblem(paths[0])
    assert t0 == ctype0 and c0.RHS == rhs0 and c0.keys.tolist() == keys0.tolist() and c0.vals.tolist() == vals0.tolist()
    if lower_prob > 0:
        assert (ct == 1).any(), "the case is meant to exercise the feasibility branch"
