"""BASELINE.json configs[0]: a small SGUFP instance solved end to end by the Benders loop of
NodeExplorer::process (NodeExplorer.cpp:915-986) running on K1 + K2, checked the way the
reference checks itself (main.cpp:26,43,76): |optimum - extensive-form MIP optimum| <= 1e-5."""
import numpy as np
import pytest

import sgufp_solver_b200 as sg
from ref_mip import solve_extensive_form
from sgufp_solver_b200 import instances as I
from sgufp_solver_b200.explorer import solve

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("S,lower_prob", [(50, 0.0), (20, 0.15)])
def test_c1_end_to_end(S, lower_prob):
    inst = I.config1(S=S, lower_prob=lower_prob)
    mip = solve_extensive_form(inst)
    solver = sg.GuroSolver(inst)
    if mip is None:                                        # lower bounds can make every first-stage choice infeasible
        best, nodes, cuts = solve(solver)
        assert best < -1e300
        return
    best, nodes, cuts = solve(solver)
    assert abs(best - mip) <= 1e-5, (best, mip, nodes, cuts)
    assert cuts >= 2


def test_mid_instance_with_branching():
    """A diagram that is NOT exact at the root (NodeExplorer.cpp:973-985): cut-set nodes are explored one
    by one by the minimal sequential driver.  The reference's parallel DDSolver is out of scope, so
    this only checks what must hold for ANY number of processed nodes: every incumbent is the value
    of a real first-stage solution, hence <= the MIP optimum, and it improves on the trivial bound."""
    inst = I.make_layered([4, 5, 5, 4], 48, 12, 123, 0.7, 0.0, "mid")
    mip = solve_extensive_form(inst)
    solver = sg.GuroSolver(inst)
    best, nodes, cuts = solve(solver, max_nodes=150)
    assert mip is not None and best <= mip + 1e-6, (best, mip)
    assert best > 0 and cuts >= 10 and nodes == 150


def test_batched_global_cuts_give_the_same_node_result():
    """NodeExplorer with the global optimality cuts applied in one K2 launch pair
    returns exactly what the sequential loop returns on an exact diagram."""
    from sgufp_solver_b200.dd import Node
    from sgufp_solver_b200.explorer import Container, NodeExplorer
    inst = I.config1(S=30)
    solver = sg.GuroSolver(inst)
    warm = NodeExplorer(solver)
    feas, opt = Container(), Container()
    first = warm.process(Node(ub=1e300), -1e300, feas, opt)           # fills the global cut lists
    assert len(opt) >= 2
    a = NodeExplorer(solver, batch_global_cuts=False, device_sequences=False).process(Node(ub=1e300), -1e300, Container(), _copy(opt))
    b = NodeExplorer(solver, batch_global_cuts=True, device_sequences=False).process(Node(ub=1e300), -1e300, Container(), _copy(opt))
    assert (a.lb, a.ub, a.status) == (b.lb, b.ub, b.status) == (first.lb, first.ub, first.status)


def _copy(container):
    from sgufp_solver_b200.explorer import Container
    c = Container()
    for cut in reversed(list(container)):
        c.add(cut)
    return c


def test_device_sequences_replay_the_loop_call_by_call():
    """the loops over the global cuts as ONE device call each must leave the search exactly
    where the call-by-call loops of NodeExplorer.cpp:935-944, 975-983 leave it — same incumbent, same number
    of nodes and cuts — on an instance whose root diagram is not exact (pruning, node removal, branching)."""
    inst = I.make_layered([4, 5, 5, 4], 48, 12, 123, 0.7, 0.0, "mid")
    a = solve(sg.GuroSolver(inst), max_nodes=60, device_sequences=False)
    b = solve(sg.GuroSolver(inst), max_nodes=60, device_sequences=True)
    assert a == b
    inst = I.config1(S=20, lower_prob=0.15)                     # feasibility cuts in the global list
    a = solve(sg.GuroSolver(inst), device_sequences=False)
    b = solve(sg.GuroSolver(inst), device_sequences=True)
    assert a == b


def test_committed_bench_candidates_are_the_dd_emission():
    """bench.py evaluates the candidate paths of sgufp_solver_b200/data/bench_candidates.npz: they must be what the DD master
    emits today (candidates.dd_emitted_paths on a 32-scenario copy of the network), path for path."""
    import os
    import bench
    from sgufp_solver_b200 import instances as I
    from sgufp_solver_b200.candidates import dd_emitted_paths
    if not os.path.exists(bench.CANDIDATES_FILE):
        pytest.skip("no committed emission: bench.py runs the master itself")
    z = np.load(bench.CANDIDATES_FILE)
    for net in z.files:
        fresh, info = dd_emitted_paths(getattr(I, net)(S=32), z[net].shape[0], budget_s=120.0)
        assert fresh.shape == z[net].shape and (fresh == z[net]).all(), (net, info)


@pytest.mark.parametrize("make", [lambda: I.config1(S=50), lambda: I.config1(S=50, lower_prob=0.1),
                                  lambda: I.make_layered([4, 5, 5, 4], 48, 12, 123, 0.7, 0.0, "mid")], ids=["c1", "c1_lb", "mid"])
def test_frontier_explorer_keeps_the_search_and_batches_k1(make):
    """SURVEY.md 8f-1, K1 side: nodes side by side, one K1 call per round.  Width 1 IS the one-path loop (same optimum, nodes and
    cuts as `solve`); wider frontiers reach the same optimum with several candidates per K1 call."""
    from sgufp_solver_b200.explorer import solve_frontier
    inst = make()
    best, nodes, cuts = solve(sg.GuroSolver(inst), max_nodes=2000)
    b1, n1, c1, k1 = solve_frontier(sg.GuroSolver(inst), width=1, max_nodes=2000)
    assert (b1, n1, c1) == (best, nodes, cuts) and k1 == c1
    for width in (4, 16):
        bw, nw, cw, kw = solve_frontier(sg.GuroSolver(inst), width=width, max_nodes=2000)
        if nodes < 2000 and nw < 2000:          # both searches ran to the end (not cut off by the node budget)
            assert bw == best
        assert kw <= cw
