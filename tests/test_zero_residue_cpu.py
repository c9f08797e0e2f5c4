""""exact zeros are dropped where the reference's running fp64 sum would keep a ~1e-17 residue, which
changes nnz, key list and hash_val".  Pinned here on the CPU.

The reference folds `coef += (cap / S) * dual` scenario by scenario in fp64 (grb.cpp:241-278) and `cutToCut` drops
`v == 0` (Cut.h:412).  Oracle B performs that very fold (operand order included) next to the exact integer sums the product
divides once.  A coefficient whose exact value is 0 is a sum of terms that cancel; when the cancelling terms are the same
product (`-(u/S)*lambda + (u/S)*sigma` with lambda == sigma on one arc) the running sum is exactly 0 too, when they are
different products with the same value (`u_a*x == u_b*y`) and S is not a power of two the running sum may keep a residue of a
few ulp.  What is pinned:
  * the residue slots are exactly {exact sum == 0, running sum != 0}, they are rare and below 1e-12 of the cut's scale;
  * no non-zero coefficient is ever lost;
  * the product's key list is the reference-order list MINUS the residue slots, in the same order; nnz differs by their count;
  * the values differ by fp64 rounding only (<= 1e-13 relative; bar 1e-9), so `hash_val` — a hash of the value BITS
    (Cut.h:247-251) — equals the reference-order fold's exactly when all values and the RHS have the same bits.
Reproducing the residues would take the reference's sequential fp64 additions per coefficient, i.e. every scenario's duals kept
until a serial fold: the exact sum is the better number and the difference is six orders of magnitude under the bar."""
import numpy as np
import pytest

from helpers import wlayout_partial
from oracle.oracle import OracleNet
from sgufp_solver_b200 import instances as I
from sgufp_solver_b200.distributed import I64_MAX, finalize
from sgufp_solver_b200.solver import Cut, GuroSolver

CASES = [("c1", lambda: I.config1(S=50), 8, 0.15), ("c1_sparse", lambda: I.config1(S=33), 6, 0.6),
         ("c2", lambda: I.config2(S=60), 5, 0.1), ("odd", lambda: I.make_layered([3, 4, 3], 21, 24, 77, 0.8, 0.0, "odd"), 6, 0.2),
         # S = 3 on a large network: the case where a residue does occur (one slot in ~28 000 exact zeros)
         ("large_S3", lambda: I.make_layered([100, 100, 100, 100, 100, 90], 6000, 3, 99, 0.7, 0.0, "large"), 2, 0.2)]
SEEN = {"residues": 0, "zeros": 0}


@pytest.mark.parametrize("name,make,K,unm", CASES, ids=[c[0] for c in CASES])
def test_key_list_is_the_reference_order_fold_minus_its_residues(built_lib, name, make, K, unm):
    inst = make()
    net = OracleNet(inst)
    gs = GuroSolver(inst, device=-1)                      # host half only: sgufp_finalize_paths on a model-only handle
    paths = np.ascontiguousarray(I.random_paths(net, K, 8 if name == "large_S3" else 5, unm), dtype=np.int16)
    sums = np.stack([wlayout_partial(net, inst, p, 0, inst.S)[0] for p in paths])
    ours = finalize(gs, paths, sums, np.full(K, I64_MAX, np.int64))
    lex = np.lexsort((gs.slot_j[:gs.T], gs.slot_q[:gs.T], gs.slot_i[:gs.T]))
    key_of = lambda s: int(gs.slot_q[s]) | (int(gs.slot_i[s]) << 16) | (int(gs.slot_j[s]) << 32)
    worst = 0.0
    for k in range(K):
        oc = net.solve_path(paths[k])                     # reference-order running sums + cutToCut
        assert oc.cut_type == 0
        exact_zero = oc.isum[1:] == 0
        residue = exact_zero & (oc.coef_dense != 0)
        SEEN["residues"] += int(residue.sum()); SEEN["zeros"] += int(exact_zero.sum())
        scale = max(1.0, np.abs(oc.vals).max())
        assert np.abs(oc.coef_dense[residue]).max(initial=0.0) <= 1e-12 * scale         # a few ulp
        assert residue.sum() <= max(1, exact_zero.sum() // 1000)                        # rare
        assert not (~exact_zero & (oc.coef_dense == 0)).any()                           # no non-zero is lost to rounding
        assert [int(x) for x in oc.keys] == [key_of(s) for s in lex if oc.coef_dense[s] != 0]
        n = int(ours.nnz[k])
        assert [int(x) for x in ours._keys[k, :n]] == [key_of(s) for s in lex if not exact_zero[s]]
        assert n == len(oc.keys) - int(residue.sum())
        want = np.array([oc.coef_dense[s] for s in lex if not exact_zero[s]])
        worst = max(worst, float(np.abs(ours._vals[k, :n] - want).max() / scale))
        if not residue.any():
            same_bits = bool((ours._vals[k, :n] == oc.vals).all() and ours.rhs[k] == oc.rhs)
            a, b = Cut(ours.rhs[k], ours._keys[k, :n], ours._vals[k, :n]), Cut(oc.rhs, oc.keys, oc.vals)
            assert (a.hash_val == b.hash_val) or not same_bits           # equal bits => equal hash (and then operator== holds)
            if same_bits:
                assert a == b
    assert worst <= 1e-13                                 # rounding of the fold only (bar: 1e-9)


def test_a_residue_was_actually_seen():
    """The cases above must include the phenomenon they pin (runs after them: same module, definition order)."""
    assert SEEN["zeros"] > 0
    assert SEEN["residues"] >= 1, "no residue in any case: the large S=3 instance is meant to produce one"
