"""The reference's structural invariants of RestrictedDDNew (tests2.cpp:370-456, fixture width 32),
re-pointed at synthetic instances, on the product's diagrams (CPU, SGUFP_DEVICE_NONE).
INVARIANT_6 (a node's states exclude its incoming decision) and the node-by-node identity with the
older classes (INVARIANT_0/7) are covered by the bit-for-bit structure comparison with the compiled
reference in tests/test_dd_structure.py."""
import numpy as np
import pytest

from sgufp_solver_b200 import instances as I
from sgufp_solver_b200.dd import RelaxedDDNew, RestrictedDDNew
from sgufp_solver_b200.solver import GuroSolver

CASES = {"c1": lambda: I.config1(S=1), "c2": lambda: I.config2(S=1), "c4": lambda: I.config4(S=1)}


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("width", [32, 128])
def test_restricted_invariants(name, width, built_lib):
    solver = GuroSolver(CASES[name](), device=-1)
    dd = RestrictedDDNew(solver, width)
    cutset = dd.compile()
    sizes = dd.layer_sizes()
    d = dd.dump()
    nodes, arcs = dd.counts()
    assert nodes > 2 and nodes == int(sizes[:-1].sum())                       # INVARIANT_0
    assert (np.diff(d["in_ptr"])[1:] == 1).all() and d["in_ptr"][1] == 0       # INVARIANT_2/3: one parent each, none for the root
    assert len(d["terminal_weight"]) == sizes[-2]                              # INVARIANT_4: one terminal arc per last-layer node
    assert arcs == (nodes - 1) + sizes[-2]
    assert (cutset is not None and len(cutset) > 0) == (not dd.isTreeExact())  # INVARIANT_5
    assert (sizes > 0).all()                                                   # INVARIANT_7: no empty layer
    assert (sizes[1:-1] <= max(width, 1)).all() or dd.isTreeExact()            # the width cap
    assert len(sizes) == solver.L + 2                                          # one layer per processingOrder entry + root + terminal
    assert (d["terminal_weight"] == np.finfo(np.float64).max).all()            # DOUBLE_MAX before any cut (DD.cpp:3145)
    # tail positions point into the previous layer
    off = np.concatenate([[0], np.cumsum(sizes[:-1])])
    for l in range(1, len(sizes) - 1):
        tp = d["arc_tailpos"][off[l] - 1: off[l + 1] - 1]
        assert (tp >= 0).all() and (tp < sizes[l - 1]).all()


@pytest.mark.parametrize("name", sorted(CASES))
def test_relaxed_invariants(name, built_lib):
    solver = GuroSolver(CASES[name](), device=-1)
    dd = RelaxedDDNew(solver)
    dd.buildTree()
    sizes = dd.layer_sizes()
    assert len(sizes) == solver.L + 2 and (sizes > 0).all()
    collapsed = [l for l in range(1, len(sizes) - 1) if sizes[l] == 1]
    assert dd.isTreeExact() == (len(collapsed) == 0)                          # a one-node layer past the root is a collapsed layer
    # a layer only collapses before the last 5 layers (DD.cpp:3614)
    assert all(l - 1 < solver.L - 5 for l in collapsed)
    if not dd.isTreeExact():
        cs = dd.getCutset(1.0)
        first = min(l for l in collapsed if l >= 3)
        assert all(n.globalLayer == first for n in cs) and all(len(n.solutionVector) == first for n in cs)
        assert all(n.ub == 1.0 for n in cs)
