"""The reference's dual LP, restated row by row for HiGHS (scipy) — the solver-independent pin
of Oracle B.  Gurobi, which the reference calls, is not available.

  variables   grb.h:44-51 / grb.cpp:10-37   alpha free, everything else >= 0 (only the USED ones)
  rows        grb.cpp:49-123                one per arc in A1, A2, A3;  alpha[0]=alpha[n-1]=0 (134-135)
  objective   grb.cpp:177-229               built per scenario from u_s, l_s and y-bar
  y-bar       grb.cpp:141-150               keyed by NODE ids
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
from scipy.optimize import linprog


class RefLP:
    def __init__(self, inst, vbar_order=None):
        self.inst = inst
        n, m = inst.n, inst.m
        self.n, self.m = n, m
        tail, head = inst.tail.astype(int), inst.head.astype(int)
        self.tail, self.head = tail, head
        self.in_arcs = [[] for _ in range(n)]
        self.out_arcs = [[] for _ in range(n)]
        for a in range(m):
            self.out_arcs[tail[a]].append(a)
            self.in_arcs[head[a]].append(a)
        self.is_vbar = np.zeros(n, bool)
        self.is_vbar[inst.vbar] = True
        self.vbar = list(inst.vbar if vbar_order is None else vbar_order)
        # arc classes, Network.cpp:68-80 (A4 is cleared)
        self.A1, self.A2, self.A3 = [], [], []
        for a in range(m):
            i, j = tail[a], head[a]
            if not self.in_arcs[i]:
                if self.out_arcs[j]:
                    self.A1.append(a)
            elif not self.out_arcs[j]:
                self.A2.append(a)
            else:
                self.A3.append(a)
        # slots (i,q,j) in processingOrder x outgoingArcs order (grb.h:60-68)
        self.slot = {}
        self.slots = []
        for q in self.vbar:
            for ain in self.in_arcs[q]:
                for aout in self.out_arcs[q]:
                    self.slot[(ain, aout)] = len(self.slots)
                    self.slots.append((ain, aout))
        T = len(self.slots)
        self.T = T
        # variable layout
        self.o_alpha = 0
        self.o_beta = n
        self.o_gamma = n + m
        self.o_sigma = n + 2 * m
        self.o_phi = n + 3 * m
        self.o_lambda = n + 4 * m
        self.o_mu = n + 4 * m + T
        self.nvar = n + 4 * m + 2 * T
        rows, cols, vals, rhs = [], [], [], []

        def add(r, c, v):
            rows.append(r); cols.append(c); vals.append(v)

        r = 0
        self.row_arc = []
        for a in self.A1:  # grb.cpp:49-65
            i, q = tail[a], head[a]
            add(r, self.o_alpha + q, 1); add(r, self.o_beta + a, -1); add(r, self.o_gamma + a, 1)
            if self.is_vbar[q]:
                for b in self.out_arcs[q]:
                    add(r, self.o_lambda + self.slot[(a, b)], 1); add(r, self.o_mu + self.slot[(a, b)], -1)
                add(r, self.o_sigma + a, 1)
            rhs.append(inst.reward[a, 0]); self.row_arc.append(a); r += 1
        for a in self.A2:  # grb.cpp:67-83
            q, j = tail[a], head[a]
            add(r, self.o_alpha + q, -1); add(r, self.o_beta + a, -1); add(r, self.o_gamma + a, 1)
            if self.is_vbar[q]:
                for b in self.in_arcs[q]:
                    add(r, self.o_lambda + self.slot[(b, a)], -1); add(r, self.o_mu + self.slot[(b, a)], 1)
                add(r, self.o_phi + a, 1)
            rhs.append(inst.reward[a, 0]); self.row_arc.append(a); r += 1
        for a in self.A3:  # grb.cpp:85-123
            q, j = tail[a], head[a]
            add(r, self.o_alpha + q, -1); add(r, self.o_alpha + j, 1); add(r, self.o_beta + a, -1); add(r, self.o_gamma + a, 1)
            if self.is_vbar[q]:
                for b in self.in_arcs[q]:
                    add(r, self.o_mu + self.slot[(b, a)], 1); add(r, self.o_lambda + self.slot[(b, a)], -1)
                add(r, self.o_phi + a, 1)
            if self.is_vbar[j]:
                for b in self.out_arcs[j]:
                    add(r, self.o_lambda + self.slot[(a, b)], 1); add(r, self.o_mu + self.slot[(a, b)], -1)
                add(r, self.o_sigma + a, 1)
            rhs.append(inst.reward[a, 0]); self.row_arc.append(a); r += 1
        self.A = sp.csr_matrix((vals, (rows, cols)), shape=(r, self.nvar), dtype=float)
        self.b = np.array(rhs, dtype=float)
        self.bounds = [(None, None)] * n + [(0, None)] * (4 * m + 2 * T)
        self.bounds[0] = (0, 0)
        self.bounds[n - 1] = (0, 0)

    def ybar(self, path, layer_arc):
        """y[(in-arc, out-arc)] per grb.cpp:141-150 (node-id keyed)."""
        y = np.zeros(self.T)
        for ell, b in enumerate(path):
            if b == -1:
                continue
            a = int(layer_arc[ell])
            q, j = self.head[a], self.head[int(b)]
            for bo in self.out_arcs[q]:
                if self.head[bo] == j and (a, bo) in self.slot:
                    y[self.slot[(a, bo)]] = 1
        return y

    def objective(self, s, y):
        """grb.cpp:177-229."""
        inst = self.inst
        c = np.zeros(self.nvar)
        u, l = inst.upper[:, s].astype(float), inst.lower[:, s].astype(float)
        c[self.o_gamma:self.o_gamma + self.m] = u
        c[self.o_beta:self.o_beta + self.m] = -l
        for k, (ain, aout) in enumerate(self.slots):
            c[self.o_lambda + k] = u[ain] * (1 - y[k])
            c[self.o_mu + k] = u[aout] * (1 - y[k])
        for q in self.vbar:
            for ain in self.in_arcs[q]:
                c[self.o_sigma + ain] += u[ain] * sum(y[self.slot[(ain, b)]] for b in self.out_arcs[q])
            for aout in self.out_arcs[q]:
                c[self.o_phi + aout] += u[aout] * sum(y[self.slot[(b, aout)]] for b in self.in_arcs[q])
        return c

    def solve(self, s, y):
        """HiGHS on the reference LP.  Returns (status, objective): 0 optimal, 1 otherwise —
        the reference treats every non-OPTIMAL status as 'feasibility cut' (grb.cpp:236,284)."""
        c = self.objective(s, y)
        res = linprog(c, A_ub=-self.A, b_ub=-self.b, bounds=self.bounds, method="highs")
        if res.status == 0:
            return 0, res.fun
        return 1, None

    def pack(self, d):
        """Oracle/GPU dual arrays -> LP vector."""
        x = np.zeros(self.nvar)
        x[self.o_alpha:self.o_alpha + self.n] = d["alpha"]
        x[self.o_beta:self.o_beta + self.m] = d["beta"]
        x[self.o_gamma:self.o_gamma + self.m] = d["gamma"]
        x[self.o_sigma:self.o_sigma + self.m] = d["sigma"]
        x[self.o_phi:self.o_phi + self.m] = d["phi"]
        x[self.o_lambda:self.o_lambda + self.T] = d["lambda"]
        x[self.o_mu:self.o_mu + self.T] = d["mu"]
        return x

    def check_feasible(self, x, homogeneous=False):
        """Row-by-row feasibility of a dual point (or of a ray when homogeneous)."""
        lhs = self.A @ x
        rhs = np.zeros_like(self.b) if homogeneous else self.b
        ok_rows = np.all(lhs >= rhs - 1e-9)
        ok_sign = np.all(x[self.n:] >= -1e-12) and x[0] == 0 and x[self.n - 1] == 0
        return bool(ok_rows and ok_sign), lhs - rhs
