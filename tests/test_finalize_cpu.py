"""Host half of K1 without a GPU: `sgufp_finalize_paths` on a model-only handle turns integer sums into
`Inavap::Cut`s (`cutToCut`, /root/reference/Cut.h:406-421): exact-zero coefficients dropped, keys packed by
`getKey` (Cut.h:342-344), (i,q,j)-lexicographic order; and it gives the same cuts whether its plans are built
for the call or taken from the first half of the operation."""
import numpy as np
import pytest

from sgufp_solver_b200 import instances as I
from sgufp_solver_b200.distributed import I64_MAX, finalize
from sgufp_solver_b200.solver import GuroSolver, getKey


@pytest.mark.parametrize("make,K", [(lambda: I.config1(S=7), 5), (lambda: I.config2(S=3), 9)], ids=["c1", "c2"])
def test_sparse_form_is_the_dense_form_in_map_order(built_lib, make, K):
    inst = make()
    gs = GuroSolver(inst, device=-1)
    rng = np.random.default_rng(5)
    paths = np.ascontiguousarray(I.random_paths(gs, K, 11, 0.3), dtype=np.int16)
    sums = rng.integers(-50, 50, size=(K, gs.W)).astype(np.int64)
    sums[rng.random((K, gs.W)) < 0.5] = 0                       # many exact zeros
    finf = np.full(K, I64_MAX, np.int64)
    finf[K // 2] = 1                                            # one feasibility cut: no 1/S on a ray
    a = finalize(gs, paths, sums, finf)
    b = finalize(gs, paths, sums, finf)                         # second call: the handle may reuse the plans
    order = sorted(range(gs.T), key=lambda s: (gs.slot_i[s], gs.slot_q[s], gs.slot_j[s]))
    for k in range(K):
        assert a.cut_type[k] == (1 if finf[k] != I64_MAX else 0)
        assert a.rhs[k] == sums[k, 0] / (1.0 if finf[k] != I64_MAX else float(inst.S))
        dense = a.coef_dense[k, :gs.T]
        want = [(getKey(int(gs.slot_q[s]), int(gs.slot_i[s]), int(gs.slot_j[s])), dense[s]) for s in order if dense[s] != 0]
        n = int(a.nnz[k])
        assert n == len(want)
        assert [int(x) for x in a._keys[k, :n]] == [w[0] for w in want]
        assert (a._vals[k, :n] == np.array([w[1] for w in want])).all()
        assert (b.coef_dense[k] == a.coef_dense[k]).all() and b.nnz[k] == n and b.rhs[k] == a.rhs[k]
        assert (b._keys[k, :n] == a._keys[k, :n]).all() and (b._vals[k, :n] == a._vals[k, :n]).all()
    # other paths on the same handle: the plans of the last call must not leak into this one
    other = np.ascontiguousarray(I.random_paths(gs, K, 12, 0.9), dtype=np.int16)
    c = finalize(gs, other, sums, finf)
    fresh = finalize(GuroSolver(inst, device=-1), other, sums, finf)
    assert (c.coef_dense == fresh.coef_dense).all() and (c.nnz == fresh.nnz).all() and (c.rhs == fresh.rhs).all()
