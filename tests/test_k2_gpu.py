"""K2 parity on a real B200: every cut application through the C ABI against Oracle A (the
unmodified reference classes), bit-exact — the recurrence is fp64 add/max/min on identical inputs
(SURVEY.md §8c).  Modelled on the reference's differential tests (tests2.cpp:470-527), re-pointed at
synthetic instances with fixed seeds."""
import numpy as np
import pytest

import sgufp_solver_b200 as sg
from oracle import ref_dd
from sgufp_solver_b200 import instances as I
from sgufp_solver_b200.dd import Node, RelaxedDDNew, RestrictedDDNew, apply_optimality_batch, random_cut

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_dd.available(), reason="oracle/_ref not available")]

INSTANCES = {
    "c1": lambda: I.config1(S=1),
    "c2": lambda: I.config2(S=1),
    "mid": lambda: I.make_layered([4, 5, 5, 4], 48, 1, 123, 0.7, 0.0, "mid"),
    "wide": lambda: I.make_layered([6, 7, 7, 7, 6], 120, 1, 321, 0.8, 0.0, "wide"),
}
LOWEST = -np.finfo(np.float64).max


def _same_state(ours, ref, relaxed=True):
    assert ours.layer_sizes().tolist() == ref.layer_sizes().tolist()
    a, b = ours.dump(), ref.dump()
    if relaxed:
        nn, na = len(a["node_layer"]), len(a["arc_tailpos"])
        assert a["in_ptr"].tolist() == b["in_ptr"][:nn + 1].tolist()
        assert a["arc_tailpos"].tolist() == b["arc_tailpos"][:na].tolist()
        assert (a["node_state"] == b["node_state"][:nn]).all()                      # bit-exact node states
        assert (a["terminal_weight"] == b["arc_weight"][na:]).all()                 # terminal arc weights: min over cuts
    else:
        assert (a["node_state"] == b["state"]).all()


@pytest.mark.parametrize("name", sorted(INSTANCES))
def test_relaxed_optimality_cuts_differential(name):
    inst = INSTANCES[name]()
    solver = sg.GuroSolver(inst)
    rn = ref_dd.RefNetwork(inst)
    ours, ref = RelaxedDDNew(solver), ref_dd.RefRelaxedDD(rn)
    ours.buildTree(); ref.build()
    rng = np.random.default_rng(7)
    ub = 1e300
    for it in range(60):
        cut = random_cut(solver, rng)
        # alternate: no pruning (optimal = -inf), then thresholds near the bound to hit DD.cpp:3987-4021
        optimal = -1e300 if it % 3 == 0 else ub - rng.uniform(0, 400)
        b_ref = ref.apply_opt(cut.RHS, cut.keys, cut.vals, optimal, ub)
        b = ours.applyOptimalityCut(cut, optimal, ub)
        assert b == b_ref, (it, b, b_ref)
        if b_ref == LOWEST:
            break                                                                     # tree pruned away (DD.cpp:4015)
        ub = min(ub, b)
        _same_state(ours, ref)
        assert ours.getSolution().tolist() == ref.solution().tolist()


@pytest.mark.parametrize("name", sorted(INSTANCES))
def test_relaxed_mixed_cuts_differential(name):
    inst = INSTANCES[name]()
    solver = sg.GuroSolver(inst)
    rn = ref_dd.RefNetwork(inst)
    ours, ref = RelaxedDDNew(solver), ref_dd.RefRelaxedDD(rn)
    ours.buildTree(); ref.build()
    rng = np.random.default_rng(11)
    for it in range(40):
        if rng.integers(0, 2):
            cut = random_cut(solver, rng, cut_type=1)
            f_ref = ref.apply_feas(cut.RHS, cut.keys, cut.vals)
            f = ours.applyFeasibilityCut(cut)
            assert f == f_ref
            if not f:
                break
        else:
            cut = random_cut(solver, rng)
            assert ours.applyOptimalityCut(cut, -1e300, 1e300) == ref.apply_opt(cut.RHS, cut.keys, cut.vals, -1e300, 1e300)
        _same_state(ours, ref)
        assert ours.getSolution().tolist() == ref.solution().tolist()


@pytest.mark.parametrize("name", ["c2", "wide"])
def test_relaxed_subtree_after_cutset(name):
    """Diagram rebuilt from a cut-set node (non-empty rootSolution => justified RHS, DD.cpp:3938-3949)."""
    inst = INSTANCES[name]()
    solver = sg.GuroSolver(inst)
    rn = ref_dd.RefNetwork(inst)
    ours, ref = RelaxedDDNew(solver), ref_dd.RefRelaxedDD(rn)
    ours.buildTree(); ref.build()
    cs = ours.getCutset(1e300)
    rng = np.random.default_rng(3)
    for nd in (cs[0], cs[len(cs) // 2], cs[-1]):
        ours.buildTree(Node(nd.states, nd.solutionVector, globalLayer=nd.globalLayer))
        ref.build(nd.states, nd.solutionVector, nd.globalLayer)
        for _ in range(10):
            cut = random_cut(solver, rng)
            assert ours.applyOptimalityCut(cut, -1e300, 1e300) == ref.apply_opt(cut.RHS, cut.keys, cut.vals, -1e300, 1e300)
            assert ours.getSolution().tolist() == ref.solution().tolist()
        _same_state(ours, ref)
        if not ours.isTreeExact():
            assert [(n.globalLayer, n.states, n.solutionVector) for n in ours.getCutset(5.0)] == ref.cutset(5.0)


@pytest.mark.parametrize("name", sorted(INSTANCES))
@pytest.mark.parametrize("width", [4, 64, 1024])
def test_restricted_differential(name, width):
    """tests2.cpp:470-527: optimality, feasibility and mixed sequences, new vs old, bit-exact."""
    inst = INSTANCES[name]()
    solver = sg.GuroSolver(inst)
    rn = ref_dd.RefNetwork(inst)
    ours, ref = RestrictedDDNew(solver, width), ref_dd.RefRestrictedDD(rn, width)
    ours.compile(); ref.compile()
    rng = np.random.default_rng(13 + width)
    for it in range(50):
        if it % 4 == 3:
            cut = random_cut(solver, rng, cut_type=1)
            f_ref = ref.apply_feas(cut.RHS, cut.keys, cut.vals)
            assert ours.applyFeasibilityCut(cut) == f_ref
            if not f_ref:
                break
        else:
            cut = random_cut(solver, rng)
            assert ours.applyOptimalityCut(cut) == ref.apply_opt(cut.RHS, cut.keys, cut.vals)
        _same_state(ours, ref, relaxed=False)
        assert ours.getMaxPath().tolist() == ref.solution().tolist()


def test_batched_equals_sequential_reference():
    """B diagrams x C cuts in one launch pair == C sequential applyOptimalityCut calls of the reference."""
    inst = INSTANCES["c2"]()
    solver = sg.GuroSolver(inst)
    rn = ref_dd.RefNetwork(inst)
    rng = np.random.default_rng(17)
    cuts = [random_cut(solver, rng) for _ in range(16)]
    widths = [8, 64, 300, 1024]
    ours = [RestrictedDDNew(solver, w) for w in widths]
    for d in ours:
        d.compile()
    bound = apply_optimality_batch(ours, cuts)
    ms, arcs, launches = ours[0].last_stats()
    assert launches == 2 and ms > 0 and arcs == 16 * sum(d.counts()[1] for d in ours)
    for k, w in enumerate(widths):
        ref = ref_dd.RefRestrictedDD(rn, w); ref.compile()
        b = None
        for c in cuts:
            b = ref.apply_opt(c.RHS, c.keys, c.vals)
        assert bound[k] == b
        assert ours[k].getMaxPath().tolist() == ref.solution().tolist()
    # exact relaxed diagram (C1): same statement
    inst1 = INSTANCES["c1"]()
    s1 = sg.GuroSolver(inst1); r1 = ref_dd.RefNetwork(inst1)
    d1 = RelaxedDDNew(s1); d1.buildTree()
    ref1 = ref_dd.RefRelaxedDD(r1); ref1.build()
    cuts1 = [random_cut(s1, rng) for _ in range(9)]
    b1 = apply_optimality_batch([d1], cuts1)[0]
    bref = None
    for c in cuts1:
        bref = ref1.apply_opt(c.RHS, c.keys, c.vals, -1e300, 1e300)
    assert b1 == bref and d1.getSolution().tolist() == ref1.solution().tolist()


def test_real_cuts_drive_the_same_paths():
    """Cuts produced by K1 applied by K2: bound and argmax path equal the reference's on every step."""
    inst = I.config1(S=50)
    solver = sg.GuroSolver(inst)
    rn = ref_dd.RefNetwork(inst)
    ours, ref = RelaxedDDNew(solver), ref_dd.RefRelaxedDD(rn)
    ours.buildTree(); ref.build()
    seen = []
    for _ in range(25):
        path = ours.getSolution()
        assert path.tolist() == ref.solution().tolist()
        if any(path.tolist() == p for p in seen):
            break
        seen.append(path.tolist())
        ctype, cut = solver.solveSubProblem(path)
        if ctype == sg.FEASIBILITY:
            assert ours.applyFeasibilityCut(cut) == ref.apply_feas(cut.RHS, cut.keys, cut.vals)
        else:
            assert ours.applyOptimalityCut(cut, -1e300, 1e300) == ref.apply_opt(cut.RHS, cut.keys, cut.vals, -1e300, 1e300)
    assert len(seen) >= 2


@pytest.fixture(params=["one_cta", "layered"])
def k2_path(request, monkeypatch):
    """the single-cut longest path has two launch shapes: one CTA (narrow diagrams) or one launch per layer"""
    monkeypatch.setenv("SGUFP_K2_LAYERED_MIN", "1" if request.param == "layered" else "1000000000")
    return request.param


@pytest.mark.parametrize("name", sorted(INSTANCES))
@pytest.mark.parametrize("seed", [5, 17])
def test_relaxed_device_resident_sequence(name, seed, k2_path):
    """a whole sequence of cuts applied on the device image — terminal weights, node
    removal with its cascade, bound-based arc pruning, getSolution — with NO structure query in
    between (so removal flags accumulate on the device); only bounds, flags and paths are compared
    per cut, the full structure once at the end, after the host mirror replays the flags."""
    inst = INSTANCES[name]()
    solver = sg.GuroSolver(inst)
    rn = ref_dd.RefNetwork(inst)
    ours, ref = RelaxedDDNew(solver), ref_dd.RefRelaxedDD(rn)
    ours.buildTree(); ref.build()
    rng = np.random.default_rng(seed)
    ub = 1e300
    arcs0 = ours.counts()[1]
    alive = True
    for it in range(40):
        if it % 3 == 2:
            cut = random_cut(solver, rng, cut_type=1)
            cut = type(cut)(cut.RHS * 0.2, cut.keys, cut.vals)     # a tighter RHS: some last-layer nodes fall below -0.01
            f_ref = ref.apply_feas(cut.RHS, cut.keys, cut.vals)
            assert ours.applyFeasibilityCut(cut) == f_ref, it
            if not f_ref:
                alive = False
                break
        else:
            cut = random_cut(solver, rng)
            optimal = -1e300 if it % 2 == 0 else ub - rng.uniform(0, 300)
            b_ref = ref.apply_opt(cut.RHS, cut.keys, cut.vals, optimal, ub)
            assert ours.applyOptimalityCut(cut, optimal, ub) == b_ref, it
            if b_ref == LOWEST:
                alive = False
                break
            ub = min(ub, b_ref)
        assert ours.getSolution().tolist() == ref.solution().tolist(), it
    _same_state(ours, ref)
    if alive:
        assert ours.getSolution().tolist() == ref.solution().tolist()
    test_relaxed_device_resident_sequence.removed = getattr(test_relaxed_device_resident_sequence, "removed", 0) + (arcs0 - ours.counts()[1])


def test_device_resident_sequences_did_remove_something():
    """the sequences above must have exercised the removal paths (runs after them in file order)"""
    assert getattr(test_relaxed_device_resident_sequence, "removed", 0) > 0


@pytest.mark.parametrize("width", [4, 64])
def test_restricted_device_resident_sequence(width, k2_path):
    inst = INSTANCES["c2"]()
    solver = sg.GuroSolver(inst)
    rn = ref_dd.RefNetwork(inst)
    ours, ref = RestrictedDDNew(solver, width), ref_dd.RefRestrictedDD(rn, width)
    ours.compile(); ref.compile()
    rng = np.random.default_rng(29 + width)
    for it in range(30):
        if it % 3 == 2:
            cut = random_cut(solver, rng, cut_type=1)
            cut = type(cut)(cut.RHS * 0.2, cut.keys, cut.vals)
            f_ref = ref.apply_feas(cut.RHS, cut.keys, cut.vals)
            assert ours.applyFeasibilityCut(cut) == f_ref
            if not f_ref:
                break
        else:
            cut = random_cut(solver, rng)
            assert ours.applyOptimalityCut(cut) == ref.apply_opt(cut.RHS, cut.keys, cut.vals)
        assert ours.getMaxPath().tolist() == ref.solution().tolist()
    _same_state(ours, ref, relaxed=False)


@pytest.mark.parametrize("name", sorted(INSTANCES))
@pytest.mark.parametrize("seed", [3, 23])
def test_sequence_call_equals_the_reference_loop(name, seed):
    """sgufp_dd_apply_sequence: a run of cuts in one device call (longest paths side by side, the
    sequential part in order, recomputation after every structural change) against the reference's
    one-by-one loop with its early return — bounds, the stopping index, the final structure."""
    inst = INSTANCES[name]()
    solver = sg.GuroSolver(inst)
    rn = ref_dd.RefNetwork(inst)
    ours, ref = RelaxedDDNew(solver), ref_dd.RefRelaxedDD(rn)
    ours.buildTree(); ref.build()
    rng = np.random.default_rng(seed)
    # 1. feasibility run (RHS tightened so that some last-layer nodes go)
    fc = [random_cut(solver, rng, cut_type=1) for _ in range(6)]
    fc = [type(c)(c.RHS * 0.25, c.keys, c.vals) for c in fc]
    want = []
    for c in fc:
        want.append(ref.apply_feas(c.RHS, c.keys, c.vals))
        if not want[-1]:
            break
    flags, n = ours.applyFeasibilityCuts(fc)
    assert n == len(want) and flags.tolist() == want
    if not want[-1]:
        return
    assert ours.getSolution().tolist() == ref.solution().tolist()
    # 2. optimality run with a threshold that prunes arcs along the way and may end the loop
    first = random_cut(solver, rng)
    ub = ref.apply_opt(first.RHS, first.keys, first.vals, -1e300, 1e300)
    assert ours.applyOptimalityCut(first, -1e300, 1e300) == ub
    oc = [random_cut(solver, rng) for _ in range(25)]
    optimal = ub - 250.0
    want = []
    for c in oc:
        want.append(ref.apply_opt(c.RHS, c.keys, c.vals, optimal, ub))
        if want[-1] <= optimal:
            break
    bounds, n = ours.applyOptimalityCuts(oc, optimal)
    assert n == len(want) and bounds.tolist() == want
    if want[-1] != LOWEST:
        _same_state(ours, ref)
        assert ours.getSolution().tolist() == ref.solution().tolist()


def test_sequence_call_on_a_restricted_tree():
    inst = INSTANCES["c2"]()
    solver = sg.GuroSolver(inst)
    rn = ref_dd.RefNetwork(inst)
    ours, ref = RestrictedDDNew(solver, 64), ref_dd.RefRestrictedDD(rn, 64)
    ours.compile(); ref.compile()
    rng = np.random.default_rng(41)
    oc = [random_cut(solver, rng) for _ in range(12)]
    want = [ref.apply_opt(c.RHS, c.keys, c.vals) for c in oc]
    bounds, n = ours.applyOptimalityCuts(oc)               # optimal = DOUBLE_MIN: never stops
    assert n == 12 and bounds.tolist() == want
    _same_state(ours, ref, relaxed=False)
    assert ours.getMaxPath().tolist() == ref.solution().tolist()


@pytest.mark.parametrize("block", range(3))
def test_fuzz_random_networks(block):
    """Random small networks: construction on the device, runs of cuts (feasibility and optimality, with thresholds
    that prune), one-cut calls, paths and cut-sets against the unmodified reference classes."""
    rng = np.random.default_rng(7000 + block)
    done = 0
    for k in range(10):
        nl = int(rng.integers(2, 5))
        layers = [int(rng.integers(2, 7)) for _ in range(nl)]
        max_m = sum(a * b for a, b in zip(layers[:-1], layers[1:])) + layers[0] + layers[-1]
        m = int(rng.integers(max(sum(layers) + 2, max_m // 2), max_m + 1))
        try:
            inst = I.make_layered(layers, m, 1, 5000 + 100 * block + k, float(rng.uniform(0.4, 0.95)), 0.0, f"dfz{k}")
            rn = ref_dd.RefNetwork(inst)
        except Exception:
            continue
        solver = sg.GuroSolver(inst)
        # relaxed
        ours, ref = RelaxedDDNew(solver), ref_dd.RefRelaxedDD(rn)
        ours.buildTree(); ref.build()
        assert ours.isTreeExact() == ref.is_exact()
        assert ours.dump_device()["built_on_device"]
        assert ours.getSolution().tolist() == ref.solution().tolist()
        alive = True
        fc = [random_cut(solver, rng, cut_type=1) for _ in range(3)]
        fc = [type(c)(c.RHS * 0.3, c.keys, c.vals) for c in fc]
        want = []
        for c in fc:
            want.append(ref.apply_feas(c.RHS, c.keys, c.vals))
            if not want[-1]:
                break
        flags, n = ours.applyFeasibilityCuts(fc)
        assert n == len(want) and flags.tolist() == want, inst.name
        alive = bool(want[-1])
        if alive:
            c0 = random_cut(solver, rng)
            ub = ref.apply_opt(c0.RHS, c0.keys, c0.vals, -1e300, 1e300)
            assert ours.applyOptimalityCut(c0, -1e300, 1e300) == ub
            assert ours.getSolution().tolist() == ref.solution().tolist()
            oc = [random_cut(solver, rng) for _ in range(12)]
            optimal = ub - float(rng.uniform(50, 400))
            want = []
            for c in oc:
                want.append(ref.apply_opt(c.RHS, c.keys, c.vals, optimal, ub))
                if want[-1] <= optimal:
                    break
            bounds, n = ours.applyOptimalityCuts(oc, optimal)
            assert n == len(want) and bounds.tolist() == want, inst.name
            if want[-1] != LOWEST:
                assert ours.getSolution().tolist() == ref.solution().tolist()
                if not ours.isTreeExact():
                    assert [(x.globalLayer, x.states, x.solutionVector) for x in ours.getCutset(1e300)] == ref.cutset(1e300)
                _same_state(ours, ref)
        # restricted
        width = int(rng.choice([1, 3, 16, 200]))
        ro, rr = RestrictedDDNew(solver, width), ref_dd.RefRestrictedDD(rn, width)
        cs = ro.compile(); rr.compile()
        assert ro.isTreeExact() == rr.is_exact()
        if not ro.isTreeExact():
            assert [(x.globalLayer, x.states, x.solutionVector) for x in cs] == rr.cutset()
        oc = [random_cut(solver, rng) for _ in range(5)]
        bounds, n = ro.applyOptimalityCuts(oc)
        assert bounds.tolist() == [rr.apply_opt(c.RHS, c.keys, c.vals) for c in oc]
        assert ro.getMaxPath().tolist() == rr.solution().tolist()
        _same_state(ro, rr, relaxed=False)
        done += 1
    assert done >= 5
