"""Shared test helpers (tests only)."""
from __future__ import annotations

import numpy as np

from oracle.oracle import OracleNet


def wlayout_partial(net: OracleNet, inst, path, s_lo, s_hi):
    """Exact integer partial sums of scenarios [s_lo, s_hi) in the product's W-layout
    (include/sgufp_b200.h: [0]=RHS, [1+l]=pair term of layer l, [1+L+a]=u_a*sigma_a or u_a*phi_a),
    computed from Oracle B's per-scenario duals.  Returns (sums, first_infeasible_or_None)."""
    L, m = net.L, net.m
    sums = np.zeros(1 + L + m, dtype=np.int64)
    arc_layer = {int(a): l for l, a in enumerate(net.layer_arc)}
    slot_base, t = [], 0
    for l, a in enumerate(net.layer_arc):
        slot_base.append(t)
        t += len(net.out_arcs(inst.head[a]))
    first_bad = None
    for s in range(s_lo, s_hi):
        d = net.scenario_duals(path, s)
        if d["status"] != 0:
            if first_bad is None:
                first_bad = s
            continue
        u = inst.upper[:, s].astype(np.int64); lo = inst.lower[:, s].astype(np.int64)
        sums[0] += int((u * d["gamma"]).sum() - (lo * d["beta"]).sum())
        for l, a in enumerate(net.layer_arc):
            for k, b in enumerate(net.out_arcs(inst.head[a])):
                tkn = int(u[a] * d["lambda"][slot_base[l] + k] + u[b] * d["mu"][slot_base[l] + k])
                sums[0] += tkn
                sums[1 + l] += tkn
        sums[1 + L:] += u * d["sigma"] + u * d["phi"]
    return sums, first_bad
