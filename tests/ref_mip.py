"""The extensive-form MIP of the reference (`solveStochasticModel`,
/root/reference/include/StochasticModel.h:16-203), restated for HiGHS (scipy.optimize.milp).
It is the reference's own end-to-end check: |DDSolver optimum - MIP optimum| <= 1e-5
(main.cpp:26,43,76).  Gurobi is not available, so HiGHS solves the same model."""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
from scipy.optimize import Bounds, LinearConstraint, milp


def solve_extensive_form(inst):
    n, m, S = inst.n, inst.m, inst.S
    tail, head = inst.tail.astype(int), inst.head.astype(int)
    in_arcs = [[] for _ in range(n)]
    out_arcs = [[] for _ in range(n)]
    for a in range(m):
        out_arcs[tail[a]].append(a)
        in_arcs[head[a]].append(a)
    vbar = [int(v) for v in inst.vbar]
    pairs = [(q, ai, ao) for q in vbar for ai in in_arcs[q] for ao in out_arcs[q]]
    P = len(pairs)
    pid = {(ai, ao): k for k, (q, ai, ao) in enumerate(pairs)}
    nx = m * S
    X = lambda a, s: a * S + s
    Y = lambda k: nx + k
    nvar = nx + P
    c = np.zeros(nvar)
    for a in range(m):
        for s in range(S):
            c[X(a, s)] = -inst.reward[a, s] / S          # maximise (StochasticModel.h:54-62)
    rows, cols, vals, lo, hi = [], [], [], [], []
    r = 0

    def add(coefs, lb, ub):
        nonlocal r
        for j, v in coefs:
            rows.append(r); cols.append(j); vals.append(v)
        lo.append(lb); hi.append(ub); r += 1

    for q in vbar:                                         # 2a' (:67-88): a matching at every V-bar node
        for ai in in_arcs[q]:
            add([(Y(pid[(ai, ao)]), 1) for ao in out_arcs[q]], -np.inf, 1)
        for ao in out_arcs[q]:
            add([(Y(pid[(ai, ao)]), 1) for ai in in_arcs[q]], -np.inf, 1)
    for s in range(S):
        for q in range(n):                                 # 2b (:103-117): conservation at interior nodes
            if not out_arcs[q] or not in_arcs[q]:
                continue
            add([(X(a, s), 1) for a in in_arcs[q]] + [(X(a, s), -1) for a in out_arcs[q]], 0, 0)
        for q in vbar:
            for ai in in_arcs[q]:
                u_i = float(inst.upper[ai, s])
                for ao in out_arcs[q]:
                    u_o = float(inst.upper[ao, s])
                    k = pid[(ai, ao)]
                    add([(X(ai, s), 1), (X(ao, s), -1), (Y(k), u_i)], -np.inf, u_i)      # 2d (:132-146)
                    add([(X(ao, s), 1), (X(ai, s), -1), (Y(k), u_o)], -np.inf, u_o)      # 2e (:148-162)
                add([(X(ai, s), 1)] + [(Y(pid[(ai, ao)]), -u_i) for ao in out_arcs[q]], -np.inf, 0)   # 2f (:164-178)
            for ao in out_arcs[q]:
                u_o = float(inst.upper[ao, s])
                add([(X(ao, s), 1)] + [(Y(pid[(ai, ao)]), -u_o) for ai in in_arcs[q]], -np.inf, 0)    # 2g (:180-197)
    A = sp.csr_matrix((vals, (rows, cols)), shape=(r, nvar))
    lb = np.zeros(nvar); ub = np.ones(nvar)
    for a in range(m):                                     # 2c (:119-130); only arcs with a head are bounded there, all arcs have one
        for s in range(S):
            lb[X(a, s)] = inst.lower[a, s]; ub[X(a, s)] = inst.upper[a, s]
    integrality = np.zeros(nvar); integrality[nx:] = 1
    res = milp(c, constraints=LinearConstraint(A, np.array(lo), np.array(hi)), bounds=Bounds(lb, ub), integrality=integrality,
               options={"mip_rel_gap": 0.0})
    if res.status != 0:
        return None
    return -res.fun
