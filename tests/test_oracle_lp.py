"""Pins Oracle B (oracle/sgufp_oracle.c) against HiGHS on the reference's own LP
(grb.cpp:41-229 restated in tests/ref_lp.py).  Solver-independent checks only:
status, optimal objective, row-by-row feasibility of the lifted dual, dual objective, validity of
the ray, RHS + coef.y == mean objective, and cut validity at other first-stage points."""
import numpy as np
import pytest

from oracle.oracle import OracleNet
from ref_lp import RefLP
from sgufp_solver_b200 import instances as I

CASES = [
    ("c1", lambda: I.config1(S=12), 5, 12, 1, 0.15),
    ("c1_lb", lambda: I.config1(S=16, lower_prob=0.3), 6, 16, 3, 0.3),
    ("c2", lambda: I.config2(S=5), 3, 5, 2, 0.1),
    ("c2_lb", lambda: I.config2(S=8, lower_prob=0.2), 3, 8, 5, 0.3),
    ("c1_allunmatched", lambda: I.config1(S=4), 2, 4, 7, 1.0),
]


@pytest.mark.parametrize("name,make,K,nscen,seed,unm", CASES, ids=[c[0] for c in CASES])
def test_oracle_matches_highs(name, make, K, nscen, seed, unm):
    inst = make()
    net = OracleNet(inst)
    lp = RefLP(inst, net.vbar)
    assert lp.T == net.T
    paths = I.random_paths(net, K, seed, unm)
    for k in range(K):
        p = paths[k]
        y = lp.ybar(p, net.layer_arc)
        cut = net.solve_path(p)
        first_bad = -1
        for s in range(min(nscen, inst.S)):
            st, obj = lp.solve(s, y)
            d = net.scenario_duals(p, s)
            assert d["status"] == st
            if st == 0:
                assert abs(d["obj"] - obj) < 1e-6
                x = lp.pack(d)
                ok, slack = lp.check_feasible(x)
                assert ok, slack.min()
                assert abs(lp.objective(s, y) @ x - obj) < 1e-6      # strong duality: the lifted dual is optimal
            else:
                if first_bad < 0:
                    first_bad = s
                x = lp.pack(net.scenario_ray(p, s))
                ok, slack = lp.check_feasible(x, homogeneous=True)
                assert ok, slack.min()
                assert lp.objective(s, y) @ x < -1e-9                # the ray proves unboundedness of the dual
        if nscen >= inst.S:
            assert cut.first_infeasible == first_bad                     # lowest-index infeasible scenario wins (grb.cpp:284-351)
            assert cut.cut_type == (1 if first_bad >= 0 else 0)
        if cut.cut_type == 0:
            val = cut.rhs + cut.coef_dense @ y
            assert abs(val - cut.obj.mean()) < 1e-9 * max(1.0, abs(val))
            assert abs(cut.rhs - cut.isum[0] / inst.S) < 1e-9 * max(1.0, abs(cut.rhs))
            assert np.allclose(cut.coef_dense, cut.isum[1:] / inst.S, rtol=1e-12, atol=1e-9)
        else:
            assert cut.rhs + cut.coef_dense @ y < 0                       # violated at y-bar


def test_optimality_cut_is_valid_elsewhere():
    """theta <= RHS + coef.y' must over-estimate mean_s Q_s(y') for every first-stage y'."""
    inst = I.config1(S=6)
    net = OracleNet(inst)
    lp = RefLP(inst, net.vbar)
    paths = I.random_paths(net, 6, 11, 0.2)
    cuts = [net.solve_path(p) for p in paths]
    for j, pj in enumerate(paths):
        yj = lp.ybar(pj, net.layer_arc)
        truth = cuts[j].obj.mean()
        for c in cuts:
            if c.cut_type == 0:
                assert c.rhs + c.coef_dense @ yj >= truth - 1e-9


def test_cut_format_matches_cutToCut():
    """keys q | i<<16 | j<<32 in (i,q,j) order with exact zeros dropped (Cut.h:342-344,406-421)."""
    inst = I.config1(S=5)
    net = OracleNet(inst)
    cut = net.solve_path(I.random_paths(net, 1, 3, 0.2)[0])
    trip = [((int(k) >> 16) & 0xFFFF, int(k) & 0xFFFF, (int(k) >> 32) & 0xFFFF) for k in cut.keys]
    assert trip == sorted(trip)
    assert np.all(cut.vals != 0)
    dense = {(int(i), int(q), int(j)): v for i, q, j, v in zip(net.slot_i, net.slot_q, net.slot_j, cut.coef_dense)}
    assert len(cut.keys) == sum(1 for v in dense.values() if v != 0)
    for t, v in zip(trip, cut.vals):
        assert dense[t] == v


def test_invalid_paths():
    inst = I.config1(S=2)
    net = OracleNet(inst)
    with pytest.raises(ValueError):
        net.solve_path(np.array([9, 9, -1, -1, -1, -1], dtype=np.int16))   # out-arc 9 claimed twice
    with pytest.raises(ValueError):
        net.solve_path(np.array([99, -1, -1, -1, -1, -1], dtype=np.int16))  # not an arc id
