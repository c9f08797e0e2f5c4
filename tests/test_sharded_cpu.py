"""The N>1 host logic on CPU: 2 ranks over gloo.  Each rank forms the exact integer partial sums
of its contiguous scenario block (here from Oracle B — no GPU in this test), the single exchange
step reduces them, and the product's finalize turns them into the cut.  The result must equal the
one-rank cut bit for bit, and the oracle's cut within 1e-9."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sgufp_solver_b200 import instances as I


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, lower_prob, out_q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, "tests")]
    from helpers import wlayout_partial
    from oracle.oracle import OracleNet
    from sgufp_solver_b200.distributed import I64_MAX, finalize, reduce_partials, shard_bounds
    from sgufp_solver_b200.solver import GuroSolver
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    inst = I.config1(S=21, lower_prob=lower_prob)
    net = OracleNet(inst)
    paths = I.random_paths(net, 4, 9, 0.25)
    lo, hi = shard_bounds(inst.S, world, rank)
    model = GuroSolver(inst.scenario_slice(lo, hi), device=-1, scenario_offset=lo, S_total=inst.S)
    sums = torch.zeros((len(paths), model.W), dtype=torch.int64)
    finf = torch.full((len(paths),), I64_MAX, dtype=torch.int64)
    for k, p in enumerate(paths):
        s_k, bad = wlayout_partial(net, inst, p, lo, hi)
        sums[k] = torch.from_numpy(s_k)
        if bad is not None:
            finf[k] = bad
    reduce_partials(sums, finf)
    # optimality cuts only here: the ray of a feasibility cut is a device kernel (GPU test)
    keep = (finf == I64_MAX).numpy()
    res = finalize(model, paths, sums.numpy(), finf.numpy())
    if rank == 0:
        out_q.put((paths, keep, res.rhs, res.coef_dense[:, :model.T], finf.numpy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("lower_prob", [0.0, 0.15])
def test_two_rank_reduction_equals_one_rank(lower_prob, built_lib):
    from oracle.oracle import OracleNet
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, lower_prob, q)) for r in range(2)]
    for p in procs:
        p.start()
    paths, keep, rhs, dense, finf = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    inst = I.config1(S=21, lower_prob=lower_prob)
    net = OracleNet(inst)
    for k, p in enumerate(paths):
        oc = net.solve_path(p)
        assert (oc.first_infeasible == -1) == bool(keep[k])
        if keep[k]:
            assert rhs[k] == oc.isum[0] / inst.S                      # bit-exact vs one rank
            assert (dense[k] == oc.isum[1:] / inst.S).all()
            assert np.allclose(dense[k], oc.coef_dense, rtol=1e-9, atol=1e-9)
        else:
            assert finf[k] == oc.first_infeasible


def test_shard_bounds_partition():
    from sgufp_solver_b200.distributed import owner_of, shard_bounds
    for S, w in ((10, 3), (100000, 8), (7, 8), (1000, 1)):
        cuts = [shard_bounds(S, w, r) for r in range(w)]
        assert cuts[0][0] == 0 and cuts[-1][1] == S
        assert all(cuts[r][1] == cuts[r + 1][0] for r in range(w - 1))
        for s in (0, S // 2, S - 1):
            lo, hi = cuts[owner_of(s, S, w)]
            assert lo <= s < hi


def test_a_shard_keeps_scenario_zero_rewards():
    """the rows of the dual LP use rewards[0] of the WHOLE instance for every scenario (grb.cpp:53,71,89): a scenario
    block must hand the library the full instance's column 0, not its own first column; an empty block keeps a reward column."""
    import dataclasses
    inst = I.config1(S=12)
    rew = inst.reward.copy()
    rew[:, 1:] += (np.arange(1, inst.S)[None, :] % 5) - 2          # rewards that vary with the scenario (the format allows it)
    inst = dataclasses.replace(inst, reward=rew)
    for lo, hi in ((0, 5), (5, 12), (7, 7)):
        part = inst.scenario_slice(lo, hi)
        assert part.S == hi - lo and (part.reward[:, 0] == inst.reward[:, 0]).all()
        assert part.upper.shape == (inst.m, hi - lo)
