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


def _varying_rewards(inst):
    """Rewards that differ from scenario to scenario: the production path reads column 0 only (grb.cpp:53,71,89)."""
    import dataclasses
    rew = inst.reward.copy()
    rew[:, 1:] += (np.arange(1, inst.S)[None, :] % 7) - 3
    return dataclasses.replace(inst, reward=rew)


def _worker(rank, world, port, lower_prob, q, library_exchange=None, vary=False):
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
    if vary:
        inst = _varying_rewards(inst)
    sh = ShardedGuroSolver(inst, inst.S, rank, world, device=rank, library_exchange=library_exchange)
    assert sh.library_exchange == (library_exchange is not False)
    paths = I.random_paths(sh.solver, 6, 9, 0.25)
    res = sh.solve_paths(paths)
    if sh.library_exchange:
        info = sh.solver.comm_info()
        assert info["nccl"] and info["world"] == world and info["local_ranks"] == 1
        if lower_prob == 0:
            assert info["exchanges_last_call"] == 1            # ONE collective per batch on the feasible path
    ctype, cut = sh.solveSubProblem(paths[0])
    if rank == 0:
        q.put((paths, res.cut_type, res.rhs, res.coef_dense, res.first_infeasible, ctype, cut.RHS, cut.keys, cut.vals))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("vary", [False, True], ids=["", "rewards_vary"])
@pytest.mark.parametrize("library_exchange", [True, False], ids=["nccl_in_library", "torch_distributed"])
@pytest.mark.parametrize("lower_prob", [0.0, 0.02])
def test_two_gpus_equal_one_gpu(lower_prob, library_exchange, vary):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    if vary and not library_exchange:
        pytest.skip("one exchange mode is enough for the reward rule")
    import torch.multiprocessing as mp
    import sgufp_solver_b200 as sg
    from sgufp_solver_b200 import instances as I
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, lower_prob, q, library_exchange, vary)) for r in range(2)]
    for p in procs:
        p.start()
    paths, ct, rhs, dense, finf, ctype0, rhs0, keys0, vals0 = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    inst = I.config2(S=301, lower_prob=lower_prob)
    if vary:
        inst = _varying_rewards(inst)
    one = sg.GuroSolver(inst, device=0)
    ref = one.solve_paths(paths)
    assert (ct == ref.cut_type).all() and (finf == ref.first_infeasible).all()
    assert (rhs == ref.rhs).all() and (dense == ref.coef_dense).all()              # bit-identical
    t0, c0 = one.solveSubProblem(paths[0])
    assert t0 == ctype0 and c0.RHS == rhs0 and c0.keys.tolist() == keys0.tolist() and c0.vals.tolist() == vals0.tolist()
    if lower_prob > 0:
        assert (ct == 1).any(), "the case is meant to exercise the feasibility branch"


# ---- ONE process, N scenario blocks, exchange inside the library (sgufp_create_sharded) -----------------------------------
CASES_1P = [
    ("c1", lambda I: I.config1(S=50), 6, [0, 0]),
    ("c1_lb", lambda I: I.config1(S=50, lower_prob=0.3), 8, [0, 0, 0]),
    ("c2_lb", lambda I: I.config2(S=203, lower_prob=0.03), 6, [0, 0, 0, 0]),
    ("c2_more_blocks_than_scenarios", lambda I: I.config2(S=3), 3, [0, 0, 0, 0, 0]),
    ("c4", lambda I: I.config4(S=70), 3, [0, 0]),
    ("one_block", lambda I: I.config2(S=40, lower_prob=0.05), 4, [0]),
]


@pytest.mark.parametrize("name,make,K,devices", CASES_1P, ids=[c[0] for c in CASES_1P])
def test_single_process_partition_equals_one_block(name, make, K, devices):
    """Blocks that share the one GPU of the test box: the partition logic (contiguous blocks, flags, the cold MIN / ray / row
    steps, per-block obj/status gathered into [K][S]) runs through the same entry point a multi-GPU host uses; only the
    transport differs (a reduction kernel instead of ncclAllReduce).  Everything must equal the one-block handle bit for bit."""
    import sgufp_solver_b200 as sg
    from sgufp_solver_b200 import instances as I
    inst = make(I)
    one = sg.GuroSolver(inst)
    part = sg.GuroSolver(inst, devices=devices)
    paths = I.random_paths(one, K, 17, 0.25)
    a, b = one.solve_paths(paths), part.solve_paths(paths)
    assert (a.cut_type == b.cut_type).all() and (a.first_infeasible == b.first_infeasible).all()
    assert (a.rhs == b.rhs).all() and (a.coef_dense == b.coef_dense).all() and (a.nnz == b.nnz).all()
    for k in range(K):
        assert a.cut(k).keys.tolist() == b.cut(k).keys.tolist() and a.cut(k).vals.tolist() == b.cut(k).vals.tolist()
        s = int(a.first_infeasible[k])
        upto = inst.S if s < 0 else s
        assert (a.obj[k, :upto] == b.obj[k, :upto]).all() and (a.status[k, :upto] == b.status[k, :upto]).all()
    info = part.comm_info()
    assert info["world"] == len(devices) and info["local_ranks"] == len(devices) and not info["nccl"]
    if (a.cut_type == 0).all() and len(devices) > 1:
        assert info["exchanges_last_call"] == 1
    t0, c0 = one.solveSubProblem(paths[0])
    t1, c1 = part.solveSubProblem(paths[0])
    assert t0 == t1 and c0 == c1 and c0.vals.tolist() == c1.vals.tolist()


def test_single_process_partition_over_all_gpus():
    """With more than one GPU in the box: one process, one block per device, ncclCommInitAll inside the library."""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs 2 GPUs")
    import sgufp_solver_b200 as sg
    from sgufp_solver_b200 import instances as I
    for lower_prob in (0.0, 0.03):
        inst = I.config2(S=257, lower_prob=lower_prob)
        one = sg.GuroSolver(inst)
        part = sg.GuroSolver(inst, devices=list(range(n)))
        paths = I.random_paths(one, 6, 19, 0.25)
        a, b = one.solve_paths(paths), part.solve_paths(paths)
        assert part.comm_info()["nccl"]
        assert (a.cut_type == b.cut_type).all() and (a.first_infeasible == b.first_infeasible).all()
        assert (a.rhs == b.rhs).all() and (a.coef_dense == b.coef_dense).all()
