"""diagrams constructed ON THE DEVICE (k2_build.cu) against Oracle A, the unmodified
reference classes — layer sizes, in-arc order, tail positions and decisions of the CSR image that sits
in HBM must be the reference's, node for node; then cuts applied to that image must give the
reference's bounds and paths (the cut semantics are covered in depth by test_k2_gpu.py, whose
diagrams are built on the device too)."""
import numpy as np
import pytest

import sgufp_solver_b200 as sg
from oracle import ref_dd
from sgufp_solver_b200 import instances as I
from sgufp_solver_b200.dd import Node, RelaxedDDNew, RestrictedDDNew, random_cut

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_dd.available(), reason="oracle/_ref not available")]

INSTANCES = {
    "c1": lambda: I.config1(S=1),
    "c2": lambda: I.config2(S=1),
    "mid": lambda: I.make_layered([4, 5, 5, 4], 48, 1, 123, 0.7, 0.0, "mid"),
    "wide": lambda: I.make_layered([6, 7, 7, 7, 6], 120, 1, 321, 0.8, 0.0, "wide"),
}


def _relaxed_matches(ours, ref):
    dev = ours.dump_device()
    assert dev["built_on_device"]
    b = ref.dump()
    sizes = ref.layer_sizes()
    assert dev["layer_sizes"].tolist() == sizes.tolist()[:len(dev["layer_sizes"])]
    nn = int(dev["layer_sizes"].sum()); na = len(dev["arc_tailpos"])
    assert dev["in_ptr"].tolist() == b["in_ptr"][:nn + 1].tolist()
    assert dev["arc_tailpos"].tolist() == b["arc_tailpos"][:na].tolist()
    assert dev["arc_decision"].tolist() == b["arc_decision"][:na].tolist()
    host = ours.dump()                                   # builds the host mirror: it must be the same image
    assert host["in_ptr"].tolist() == dev["in_ptr"].tolist() and host["arc_tailpos"].tolist() == dev["arc_tailpos"].tolist()
    assert host["arc_decision"].tolist() == dev["arc_decision"].tolist()


@pytest.mark.parametrize("name", sorted(INSTANCES))
def test_relaxed_built_on_device_is_the_reference_structure(name):
    inst = INSTANCES[name]()
    solver = sg.GuroSolver(inst)
    rn = ref_dd.RefNetwork(inst)
    ours, ref = RelaxedDDNew(solver), ref_dd.RefRelaxedDD(rn)
    ours.buildTree(); ref.build()
    assert ours.isTreeExact() == ref.is_exact()
    _relaxed_matches(ours, ref)
    if not ours.isTreeExact():                            # sub-trees rooted at cut-set nodes (mid-V-bar roots, fixed prefixes)
        cs = ours.getCutset(1e300)
        for nd in (cs[0], cs[len(cs) // 3], cs[-1]):
            ours.buildTree(Node(nd.states, nd.solutionVector, globalLayer=nd.globalLayer))
            ref.build(nd.states, nd.solutionVector, nd.globalLayer)
            assert ours.isTreeExact() == ref.is_exact()
            _relaxed_matches(ours, ref)


@pytest.mark.parametrize("name", sorted(INSTANCES))
@pytest.mark.parametrize("width", [1, 4, 37, 1024])
def test_restricted_built_on_device_is_the_reference_structure(name, width):
    inst = INSTANCES[name]()
    solver = sg.GuroSolver(inst)
    rn = ref_dd.RefNetwork(inst)
    ours, ref = RestrictedDDNew(solver, width), ref_dd.RefRestrictedDD(rn, width)
    cs = ours.compile(); ref.compile()
    dev = ours.dump_device()
    assert dev["built_on_device"]
    b = ref.dump()
    sizes = ref.layer_sizes()
    assert dev["layer_sizes"].tolist() == sizes.tolist()[:len(dev["layer_sizes"])]
    nn = int(dev["layer_sizes"].sum())
    assert dev["arc_tailpos"].tolist() == b["parentpos"][1:nn].tolist()      # one parent per node, the root has none
    assert dev["arc_decision"].tolist() == b["decision"][1:nn].tolist()
    assert ours.isTreeExact() == ref.is_exact()
    if not ours.isTreeExact():
        assert [(n.globalLayer, n.states, n.solutionVector) for n in cs] == ref.cutset()
    else:
        assert cs is None
    rng = np.random.default_rng(width)
    for _ in range(6):                                    # and the image works: same bounds, same paths
        cut = random_cut(solver, rng)
        assert ours.applyOptimalityCut(cut) == ref.apply_opt(cut.RHS, cut.keys, cut.vals)
        assert ours.getMaxPath().tolist() == ref.solution().tolist()


@pytest.mark.parametrize("name,width", [("c2", 4), ("c2", 37), ("c2", 1024), ("c2", 4096), ("mid", 300)])
def test_cluster_builder_gives_the_reference_structure(name, width, monkeypatch):
    """k2_build_cluster (8 CTAs of a thread-block cluster build one diagram: wide layers) against the reference node for
    node — forced onto narrow widths too (SGUFP_DD_BUILD_CLUSTER=1), where slices are ragged or empty; width 4096 takes it
    by default."""
    if width < 2048:
        monkeypatch.setenv("SGUFP_DD_BUILD_CLUSTER", "1")
    inst = INSTANCES[name]()
    solver = sg.GuroSolver(inst)
    rn = ref_dd.RefNetwork(inst)
    ours, ref = RestrictedDDNew(solver, width), ref_dd.RefRestrictedDD(rn, width)
    cs = ours.compile(); ref.compile()
    dev = ours.dump_device()
    assert dev["built_on_device"]
    b = ref.dump()
    sizes = ref.layer_sizes()
    assert dev["layer_sizes"].tolist() == sizes.tolist()[:len(dev["layer_sizes"])]
    nn = int(dev["layer_sizes"].sum())
    assert dev["arc_tailpos"].tolist() == b["parentpos"][1:nn].tolist()
    assert dev["arc_decision"].tolist() == b["decision"][1:nn].tolist()
    assert ours.isTreeExact() == ref.is_exact()
    if not ours.isTreeExact():
        assert [(n.globalLayer, n.states, n.solutionVector) for n in cs] == ref.cutset()
    rng = np.random.default_rng(width)
    for _ in range(3):
        cut = random_cut(solver, rng)
        assert ours.applyOptimalityCut(cut) == ref.apply_opt(cut.RHS, cut.keys, cut.vals)
        assert ours.getMaxPath().tolist() == ref.solution().tolist()
    # the relaxed diagram through the cluster kernel as well (collapsed layers: the one node is written by rank 0)
    monkeypatch.setenv("SGUFP_DD_BUILD_CLUSTER", "1")
    ro, rr = RelaxedDDNew(solver), ref_dd.RefRelaxedDD(rn)
    ro.buildTree(); rr.build()
    assert ro.isTreeExact() == rr.is_exact()
    _relaxed_matches(ro, rr)


def test_fresh_device_tree_gives_the_reference_first_path():
    """getSolution before any cut: every state is DOUBLE_MIN, every terminal arc DOUBLE_MAX (DD.cpp:3595)"""
    inst = INSTANCES["c2"]()
    solver = sg.GuroSolver(inst)
    rn = ref_dd.RefNetwork(inst)
    ours, ref = RelaxedDDNew(solver), ref_dd.RefRelaxedDD(rn)
    ours.buildTree(); ref.build()
    assert ours.getSolution().tolist() == ref.solution().tolist()


def test_sequence_call_first_on_a_host_built_diagram(monkeypatch):
    """a host-built diagram (SGUFP_DD_BUILD=host, or the fallbacks) whose FIRST apply is a run of cuts
    (`sgufp_dd_apply_sequence`, what NodeExplorer::process does with the global cuts) must work like the one-by-one calls."""
    inst = INSTANCES["c2"]()
    solver = sg.GuroSolver(inst)
    rng = np.random.default_rng(77)
    cuts = [random_cut(solver, rng) for _ in range(9)]
    one = RelaxedDDNew(solver)                             # device-built, cut by cut
    one.buildTree()
    want = [one.applyOptimalityCut(c, -1e300, 1e300) for c in cuts]
    monkeypatch.setenv("SGUFP_DD_BUILD", "host")
    seq = RelaxedDDNew(solver)
    seq.buildTree()
    assert not seq.dump_device()["built_on_device"]
    bounds, n = seq.applyOptimalityCuts(cuts, -1e300)      # the first apply on this diagram is a sequence call
    assert n == len(cuts) and bounds.tolist() == want
    assert seq.getSolution().tolist() == one.getSolution().tolist()
    feas = RelaxedDDNew(solver)
    feas.buildTree()
    fc = [random_cut(solver, rng, cut_type=1) for _ in range(4)]
    flags, k = feas.applyFeasibilityCuts(fc)
    monkeypatch.delenv("SGUFP_DD_BUILD")
    ref = RelaxedDDNew(solver)
    ref.buildTree()
    exp = []
    for c in fc:
        exp.append(ref.applyFeasibilityCut(c))
        if not exp[-1]:
            break
    assert flags.tolist() == exp and k == len(exp)


def test_host_builder_still_reachable(monkeypatch):
    monkeypatch.setenv("SGUFP_DD_BUILD", "host")
    inst = INSTANCES["mid"]()
    solver = sg.GuroSolver(inst)
    d = RelaxedDDNew(solver)
    d.buildTree()
    assert not d.dump_device()["built_on_device"]


@pytest.mark.parametrize("threshold", [12, 500])
def test_relaxed_threshold_is_a_runtime_parameter(threshold, monkeypatch):
    """Config C3 sweeps the collapse threshold, a compile-time 120 in the reference (DD.h:732): other values cannot be
    compared with the reference binary, so the device construction is checked against the host statement of the same
    algorithm (which IS the reference's at 120, test above), structure and results."""
    inst = INSTANCES["c2"]()
    solver = sg.GuroSolver(inst)
    dev_dd = RelaxedDDNew(solver, threshold)
    dev_dd.buildTree()
    img = dev_dd.dump_device()
    assert img["built_on_device"]
    monkeypatch.setenv("SGUFP_DD_BUILD", "host")
    host_dd = RelaxedDDNew(solver, threshold)
    host_dd.buildTree()
    himg = host_dd.dump_device()
    assert not himg["built_on_device"]
    for k in ("layer_sizes", "in_ptr", "arc_tailpos", "arc_decision", "arc_slot"):
        assert img[k].tolist() == himg[k].tolist(), k
    ref120 = RelaxedDDNew(solver)                      # default threshold: a different diagram
    ref120.buildTree()
    assert ref120.layer_sizes().tolist() != dev_dd.layer_sizes().tolist()
    rng = np.random.default_rng(threshold)
    for _ in range(5):
        cut = random_cut(solver, rng)
        assert dev_dd.applyOptimalityCut(cut, -1e300, 1e300) == host_dd.applyOptimalityCut(cut, -1e300, 1e300)
        assert dev_dd.getSolution().tolist() == host_dd.getSolution().tolist()
