"""End-to-end consistency on config C1 without a GPU: the extensive-form MIP of the reference
(StochasticModel.h, solved by HiGHS) equals the best first-stage candidate found by brute force
with Oracle B (max over every DD path of the mean subproblem objective).  This is the reference's
own acceptance check (main.cpp:26,43,76: |optimum - MIP| <= 1e-5) with the solver taken out."""
import numpy as np

from oracle.oracle import OracleNet
from ref_mip import solve_extensive_form
from sgufp_solver_b200 import instances as I


def all_paths(net, inst):
    L = len(net.layer_arc)
    out = []

    def rec(l, used, cur):
        if l == L:
            out.append(list(cur)); return
        q = int(inst.head[net.layer_arc[l]])
        if l == 0 or int(inst.head[net.layer_arc[l - 1]]) != q:
            used = set()
        for b in [-1] + [x for x in net.out_arcs(q) if x not in used]:
            rec(l + 1, used | ({b} if b >= 0 else set()), cur + [b])
    rec(0, set(), [])
    return out


def test_mip_equals_brute_force_over_dd_paths():
    inst = I.config1(S=4)
    net = OracleNet(inst)
    paths = all_paths(net, inst)
    assert len(paths) == 442                               # last layer of the exact diagram
    best = max(net.solve_path(np.array(p, np.int16)).obj.mean() for p in paths)
    mip = solve_extensive_form(inst)
    assert mip is not None and abs(best - mip) <= 1e-5
