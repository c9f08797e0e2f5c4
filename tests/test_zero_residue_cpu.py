""""exact zeros are dropped where the reference's running fp64 sum would keep a ~1e-17 residue, which
changes nnz, key list and hash_val".  Pinned here on the CPU.

The reference folds `coef += (cap / S) * dual` scenario by scenario in fp64 (grb.cpp:241-278) and `cutToCut` drops
`v == 0` (Cut.h:412).  Oracle B performs that very fold (operand order included) next to the exact integer sums the product
divides once.  With integer duals (SPEC-LP) a coefficient whose exact value is 0 is a sum of terms that cancel
PAIRWISE inside a scenario (`-(u/S)*lambda + (u/S)*sigma` with lambda == sigma on the same arc): the running sum is exactly 0
too.  So key list, order and nnz of the product equal the reference-order fold's with NO filter; the values differ by fp64
rounding (one division against S additions), hence `hash_val` — a hash of the value BITS (Cut.h:247-251) — differs whenever a
value does.  That is the whole difference, and it is below the 1e-9 bar by six orders of magnitude."""
import numpy as np
import pytest

from helpers import wlayout_partial
from oracle.oracle import OracleNet
from sgufp_solver_b200 import instances as I
from sgufp_solver_b200.distributed import I64_MAX, finalize
from sgufp_solver_b200.solver import Cut, GuroSolver

CASES = [("c1", lambda: I.config1(S=50), 8, 0.15), ("c1_sparse", lambda: I.config1(S=33), 6, 0.6),
         ("c2", lambda: I.config2(S=60), 5, 0.1), ("odd", lambda: I.make_layered([3, 4, 3], 21, 24, 77, 0.8, 0.0, "odd"), 6, 0.2)]


@pytest.mark.parametrize("name,make,K,unm", CASES, ids=[c[0] for c in CASES])
def test_key_list_equals_the_reference_order_fold_without_a_filter(built_lib, name, make, K, unm):
    inst = make()
    net = OracleNet(inst)
    gs = GuroSolver(inst, device=-1)                      # host half only: sgufp_finalize_paths on a model-only handle
    paths = np.ascontiguousarray(I.random_paths(net, K, 5, unm), dtype=np.int16)
    sums = np.stack([wlayout_partial(net, inst, p, 0, inst.S)[0] for p in paths])
    ours = finalize(gs, paths, sums, np.full(K, I64_MAX, np.int64))
    worst = 0.0
    for k in range(K):
        oc = net.solve_path(paths[k])                     # reference-order running sums + cutToCut
        assert oc.cut_type == 0
        exact_zero = oc.isum[1:] == 0
        assert not (exact_zero & (oc.coef_dense != 0)).any()             # no residue: an exact 0 is 0 in the running sum
        assert not (~exact_zero & (oc.coef_dense == 0)).any()            # and no non-zero is lost to rounding
        n = int(ours.nnz[k])
        assert n == len(oc.keys) and [int(x) for x in ours._keys[k, :n]] == [int(x) for x in oc.keys]
        scale = max(1.0, np.abs(oc.vals).max())
        worst = max(worst, float(np.abs(ours._vals[k, :n] - oc.vals).max() / scale))
        same_bits = bool((ours._vals[k, :n] == oc.vals).all() and ours.rhs[k] == oc.rhs)
        a, b = Cut(ours.rhs[k], ours._keys[k, :n], ours._vals[k, :n]), Cut(oc.rhs, oc.keys, oc.vals)
        assert (a.hash_val == b.hash_val) == same_bits or not same_bits   # equal bits => equal hash; the hash is of the value bits
        if same_bits:
            assert a == b
    assert worst <= 1e-13                                 # rounding of the fold only (bar: 1e-9)
