"""K1 parity on a real B200, through the C ABI (sgufp_solve_paths / sgufp_paths_partial).

Bar (BASELINE.json north_star): subproblem feasibility status exactly; objectives, cut
coefficients and constants within 1e-9 relative.  The kernels accumulate exact integers, so the
integer sums are compared BIT-EXACTLY with Oracle B's and only the final fp64 fold carries the
1e-9 tolerance (the reference folds `(cap / S) * dual` scenario by scenario, grb.cpp:241-278)."""
import glob
import json
import os

import numpy as np
import pytest

import sgufp_solver_b200 as sg
from oracle.oracle import OracleNet
from sgufp_solver_b200 import instances as I

pytestmark = pytest.mark.gpu
RTOL = 1e-9


def _compare(inst, gs, net, paths, res):
    for k in range(len(paths)):
        oc = net.solve_path(paths[k])
        assert oc.cut_type == res.cut_type[k]
        assert oc.first_infeasible == res.first_infeasible[k]
        div = 1.0 if oc.cut_type else float(inst.S)
        assert res.rhs[k] == oc.isum[0] / div                                   # exact integers, one division
        assert (res.coef_dense[k, :gs.T] == oc.isum[1:] / div).all()
        scale = max(1.0, np.abs(oc.coef_dense).max())
        assert abs(res.rhs[k] - oc.rhs) <= RTOL * max(1.0, abs(oc.rhs))         # vs the reference-order fp64 fold
        assert np.abs(res.coef_dense[k, :gs.T] - oc.coef_dense).max() <= RTOL * scale
        if oc.cut_type == 0:
            assert (res.status[k] == oc.status).all()
            assert (res.obj[k] == oc.obj).all()
        else:
            s = oc.first_infeasible                                              # scenarios before the first infeasible one agree
            assert (res.status[k, :s] == 0).all() and res.status[k, s] == 1
            assert (res.obj[k, :s] == oc.obj[:s]).all()
        cut = res.cut(k)
        # The oracle's (key, value) list comes from the REFERENCE-ORDER fold (running fp64 sums of `(cap / S) * dual`,
        # grb.cpp:241-278, then cutToCut's `v == 0` test, Cut.h:412).  Ours must be that list minus the RESIDUES: slots whose
        # exact integer sum is 0 but whose running fp64 sum is not (rare — tests/test_zero_residue_cpu.py — and below 1e-12).
        # No tolerance decides membership: the exact sums do.
        lex = np.lexsort((gs.slot_j[:gs.T], gs.slot_q[:gs.T], gs.slot_i[:gs.T]))
        key_of = lambda s: int(gs.slot_q[s]) | (int(gs.slot_i[s]) << 16) | (int(gs.slot_j[s]) << 32)
        assert [int(kk) for kk in oc.keys] == [key_of(s) for s in lex if oc.coef_dense[s] != 0]
        residue = (oc.isum[1:] == 0) & (oc.coef_dense != 0)
        assert np.abs(oc.coef_dense[residue]).max(initial=0.0) <= 1e-12 * scale
        assert not ((oc.isum[1:] != 0) & (oc.coef_dense == 0)).any()
        assert [int(x) for x in cut.keys] == [key_of(s) for s in lex if oc.isum[1 + s] != 0] and len(cut.keys) == res.nnz[k]
        assert np.allclose(cut.vals, [oc.coef_dense[s] for s in lex if oc.isum[1 + s] != 0], rtol=RTOL, atol=RTOL * scale)


CASES = [
    ("c1", lambda: I.config1(S=50), 8, 1, 0.15),
    ("c1_lb", lambda: I.config1(S=50, lower_prob=0.3), 8, 2, 0.3),
    ("c1_all_unmatched", lambda: I.config1(S=8), 2, 3, 1.0),
    ("c2", lambda: I.config2(S=200), 6, 3, 0.1),
    ("c2_lb", lambda: I.config2(S=200, lower_prob=0.05), 6, 4, 0.3),
    ("c4", lambda: I.config4(S=64), 3, 5, 0.1),
    ("c4_lb", lambda: I.config4(S=64, lower_prob=0.02), 3, 6, 0.3),
    ("odd_m_single_scenario", lambda: I.make_layered([3, 4, 3], 21, 1, 77, 0.8, 0.0, "odd"), 4, 7, 0.2),
    # a network whose per-scenario state (~60 KB) no longer fits 8 warps per CTA: the 2-warp launch
    ("large_m6000", lambda: I.make_layered([100, 100, 100, 100, 100, 90], 6000, 3, 99, 0.7, 0.01, "large"), 2, 8, 0.2),
]


@pytest.mark.parametrize("name,make,K,seed,unm", CASES, ids=[c[0] for c in CASES])
def test_parity_with_oracle(name, make, K, seed, unm):
    inst = make()
    net = OracleNet(inst)
    gs = sg.GuroSolver(inst)
    assert gs.layer_arc.tolist() == net.layer_arc.tolist() and gs.T == net.T
    paths = I.random_paths(net, K, seed, unm)
    res = gs.solve_paths(paths)
    launches, ms = gs.last_stats()
    assert launches >= 1 and ms > 0                                               # the CUDA path ran
    _compare(inst, gs, net, paths, res)


GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "k1_*.json")))
MAKERS = {"k1_c1": lambda: I.config1(S=50), "k1_c1_lb": lambda: I.config1(S=50, lower_prob=0.1),
          "k1_c2_small": lambda: I.config2(S=40), "k1_c2_small_lb": lambda: I.config2(S=40, lower_prob=0.08)}


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-5] for p in GOLD])
def test_golden_vectors(path):
    g = json.load(open(path))
    inst = MAKERS[g["instance"]]()
    gs = sg.GuroSolver(inst)
    paths = np.array([c["path"] for c in g["cuts"]], np.int16)
    res = gs.solve_paths(paths)
    for k, c in enumerate(g["cuts"]):
        div = 1.0 if c["cut_type"] else float(inst.S)
        assert res.cut_type[k] == c["cut_type"] and res.first_infeasible[k] == c["first_infeasible"]
        assert res.rhs[k] == c["isum"][0] / div
        assert (res.coef_dense[k, :gs.T] == np.array(c["isum"][1:]) / div).all()
        if c["obj"] is not None:
            assert res.obj[k].tolist() == c["obj"] and res.status[k].tolist() == c["status"]


def test_reference_interface_single_path():
    """solveSubProblem(path) -> (CutType, Inavap::Cut): same answer as the batched call; hash as Cut.h:243-251."""
    from oracle import ref_dd
    inst = I.config1(S=50)
    gs = sg.GuroSolver(inst)
    net = OracleNet(inst)
    paths = I.random_paths(net, 3, 21, 0.2)
    batch = gs.solve_paths(paths)
    for k in range(3):
        ctype, cut = gs.solveSubProblem(paths[k])
        assert ctype == batch.cut_type[k] and cut == batch.cut(k)
        assert cut.keys.tolist() == batch.cut(k).keys.tolist() and cut.vals.tolist() == batch.cut(k).vals.tolist()
        if ref_dd.available():
            assert cut.hash_val == ref_dd.cut_hash(cut.RHS, cut.keys, cut.vals)


def test_full_size_c2_properties():
    """configs[1] at full size (S=1000, K=64): size-independent properties."""
    inst = I.config2(S=1000)
    gs = sg.GuroSolver(inst)
    net = OracleNet(inst)
    paths = I.random_paths(net, 64, 31, 0.1)
    a = gs.solve_paths(paths)
    b = gs.solve_paths(paths)
    assert (a.rhs == b.rhs).all() and (a.coef_dense == b.coef_dense).all() and (a.obj == b.obj).all()      # run-to-run bit-identical
    assert (a.cut_type == 0).all() and (a.status == 0).all()
    # y-bar in slot order
    slot = {}
    t = 0
    for l, arc in enumerate(gs.layer_arc):
        for bo in gs.out_arcs(inst.head[arc]):
            slot[(l, bo)] = t; t += 1
    for k in range(64):
        y = np.zeros(gs.T)
        for l, bo in enumerate(paths[k]):
            if bo >= 0:
                y[slot[(l, int(bo))]] = 1
        val = a.rhs[k] + a.coef_dense[k, :gs.T] @ y
        assert abs(val - a.obj[k].mean()) <= RTOL * max(1.0, abs(val))            # RHS + coef.y == mean objective
    # every cut is valid at every other candidate (weak duality)
    truth = a.obj.mean(axis=1)
    ys = np.zeros((64, gs.T))
    for k in range(64):
        for l, bo in enumerate(paths[k]):
            if bo >= 0:
                ys[k, slot[(l, int(bo))]] = 1
    vals = a.rhs[:, None] + a.coef_dense[:, :gs.T] @ ys.T                         # [cut, point]
    assert (vals >= truth[None, :] - 1e-6).all()
    # sampled scenarios against the oracle
    for k in (0, 17, 63):
        oc = net.solve_path(paths[k])
        assert (a.obj[k] == oc.obj).all() and a.rhs[k] == oc.isum[0] / inst.S


def test_full_size_c4_properties():
    """configs[3] at full size (n=200, m=1000, S=10 000, K=8 DD-emitted candidates): size-independent properties, and a
    sample of candidates and scenarios against the oracle."""
    import bench
    inst = I.config4(S=10000)
    gs = sg.GuroSolver(inst)
    paths, _ = bench.candidate_paths("c4", 8, 0)
    a = gs.solve_paths(paths)
    b = gs.solve_paths(paths)
    assert (a.rhs == b.rhs).all() and (a.coef_dense == b.coef_dense).all() and (a.obj == b.obj).all()      # run-to-run bit-identical
    assert (a.cut_type == 0).all() and (a.status == 0).all()
    slot, t = {}, 0
    for l, arc in enumerate(gs.layer_arc):
        for bo in gs.out_arcs(inst.head[arc]):
            slot[(l, int(inst.head[bo]))] = t; t += 1                             # y-bar is keyed by node ids (grb.cpp:145-148)
    ys = np.zeros((8, gs.T))
    for k in range(8):
        for l, bo in enumerate(paths[k]):
            if bo >= 0 and (l, int(inst.head[bo])) in slot:
                ys[k, slot[(l, int(inst.head[bo]))]] = 1
    truth = a.obj.mean(axis=1)
    vals = a.rhs[:, None] + a.coef_dense[:, :gs.T] @ ys.T                         # [cut, point]
    for k in range(8):
        assert abs(vals[k, k] - truth[k]) <= RTOL * max(1.0, abs(truth[k]))       # RHS + coef.y == mean objective
    assert (vals >= truth[None, :] - 1e-6).all()                                  # every cut is valid at every other candidate
    # linearity: the sums over two halves of the scenarios add up to the sums over all of them
    half = [sg.GuroSolver(inst.scenario_slice(lo, hi), scenario_offset=lo, S_total=inst.S) for lo, hi in ((0, 4096), (4096, 10000))]
    import ctypes as C
    import torch
    from sgufp_solver_b200 import _lib
    from sgufp_solver_b200.distributed import I64_MAX, finalize
    tot = None
    for part in half:
        sums = torch.zeros((8, part.W), dtype=torch.int64, device="cuda")
        finf = torch.zeros((8,), dtype=torch.int64, device="cuda")
        assert _lib.lib().sgufp_paths_partial(part.h, paths.ctypes.data_as(_lib.i16p), 8, paths.shape[1], C.c_void_p(sums.data_ptr()),
                                              C.c_void_p(finf.data_ptr()), None, None, None) == 0
        torch.cuda.synchronize()
        tot = sums.cpu().numpy() if tot is None else tot + sums.cpu().numpy()
    res = finalize(half[1], paths, tot, np.full(8, I64_MAX, np.int64))
    assert (res.rhs == a.rhs).all() and (res.coef_dense == a.coef_dense).all()
    # the oracle on a sample: objectives of 64 scenarios of two candidates, and the exact sums of a 64-scenario block
    net = OracleNet(inst.scenario_slice(5000, 5064))
    for k in (0, 5):
        oc = net.solve_path(paths[k])
        assert (a.obj[k, 5000:5064] == oc.obj).all()
        blk = sg.GuroSolver(inst.scenario_slice(5000, 5064)).solve_paths(paths[k:k + 1])
        assert blk.rhs[0] == oc.isum[0] / 64 and (blk.coef_dense[0, :gs.T] == oc.isum[1:] / 64).all()


def test_c5_sampled_scenarios_against_the_oracle():
    """configs[4] (S = 100 000): the whole instance does not fit a test, its scenario blocks do.  Block 3 of the bench's
    instance (12 500 scenarios, the 8-GPU shard of rank 3): objectives and statuses of a sample against the oracle, exact sums of a
    sub-block against the oracle, run-to-run identity on the block."""
    import bench
    lo, hi = 37500, 50000
    inst = bench.scenario_range("c5", lo, hi)
    gs = sg.GuroSolver(inst, scenario_offset=lo, S_total=100000)
    paths, _ = bench.candidate_paths("c5", 8, 0)
    import ctypes as C
    import torch
    from sgufp_solver_b200 import _lib
    sums = torch.zeros((8, gs.W), dtype=torch.int64, device="cuda")
    finf = torch.zeros((8,), dtype=torch.int64, device="cuda")
    obj = torch.zeros((8, inst.S), dtype=torch.float64, device="cuda")
    st = torch.zeros((8, inst.S), dtype=torch.uint8, device="cuda")
    outs = []
    for _ in range(2):
        assert _lib.lib().sgufp_paths_partial(gs.h, paths.ctypes.data_as(_lib.i16p), 8, paths.shape[1], C.c_void_p(sums.data_ptr()),
                                              C.c_void_p(finf.data_ptr()), C.c_void_p(obj.data_ptr()), C.c_void_p(st.data_ptr()), None) == 0
        torch.cuda.synchronize()
        outs.append((sums.cpu().numpy().copy(), obj.cpu().numpy().copy()))
    assert (outs[0][0] == outs[1][0]).all() and (outs[0][1] == outs[1][1]).all()
    assert (st.cpu().numpy() == 0).all() and (finf.cpu().numpy() == np.iinfo(np.int64).max).all()
    import dataclasses
    sub = dataclasses.replace(inst, S=48, upper=np.ascontiguousarray(inst.upper[:, 7000:7048]), lower=np.ascontiguousarray(inst.lower[:, 7000:7048]),
                              reward=np.ascontiguousarray(np.repeat(inst.reward[:, :1], 48, axis=1)))
    net = OracleNet(sub)
    for k in (1, 6):
        oc = net.solve_path(paths[k])
        assert (outs[0][1][k, 7000:7048] == oc.obj).all()
        blk = sg.GuroSolver(sub).solve_paths(paths[k:k + 1])
        assert (blk.coef_dense[0, :gs.T] == oc.isum[1:] / 48).all() and blk.rhs[0] == oc.isum[0] / 48


def test_sharded_partials_add_up():
    """Linearity: partial sums of two scenario blocks (sgufp_paths_partial) == the one-block sums."""
    import ctypes as C
    import torch
    from sgufp_solver_b200 import _lib
    from sgufp_solver_b200.distributed import I64_MAX, finalize
    inst = I.config2(S=300, lower_prob=0.0)
    net = OracleNet(inst)
    paths = I.random_paths(net, 5, 41, 0.2)
    full = sg.GuroSolver(inst).solve_paths(paths)
    tot = None
    for lo, hi in ((0, 130), (130, 300)):
        part = sg.GuroSolver(inst.scenario_slice(lo, hi), scenario_offset=lo, S_total=inst.S)
        sums = torch.zeros((5, part.W), dtype=torch.int64, device="cuda")
        finf = torch.zeros((5,), dtype=torch.int64, device="cuda")
        st = torch.cuda.Stream()
        rc = _lib.lib().sgufp_paths_partial(part.h, paths.ctypes.data_as(_lib.i16p), 5, paths.shape[1], C.c_void_p(sums.data_ptr()),
                                            C.c_void_p(finf.data_ptr()), None, None, C.c_void_p(st.cuda_stream))
        assert rc == 0
        torch.cuda.synchronize()
        assert (finf.cpu().numpy() == I64_MAX).all()
        tot = sums.cpu().numpy() if tot is None else tot + sums.cpu().numpy()
        last = part
    res = finalize(last, paths, tot, np.full(5, I64_MAX, np.int64))
    assert (res.rhs == full.rhs).all() and (res.coef_dense == full.coef_dense).all()


def test_bad_path_is_rejected():
    inst = I.config1(S=4)
    gs = sg.GuroSolver(inst)
    with pytest.raises(sg.solver.SgufpError) as e:
        gs.solveSubProblem(np.array([9, 9, -1, -1, -1, -1], np.int16))
    assert e.value.code == -2


LANE_CASES = [c for c in CASES if c[0] in ("c1", "c1_all_unmatched", "c2", "c4", "odd_m_single_scenario")] + [
    ("c2_ragged", lambda: I.config2(S=77), 5, 23, 0.3),          # a last block of 13 scenarios: idle lanes shadow the last one
    ("c4_two_blocks", lambda: I.config4(S=40), 2, 25, 0.0),
    ("wide_caps", lambda: _scaled(I.config2(S=33), 300), 3, 27, 0.1),   # capacities up to 9000: the 16-bit state (CfgWide)
]


def _scaled(inst, f):
    import dataclasses
    return dataclasses.replace(inst, upper=inst.upper * f)


@pytest.mark.parametrize("name,make,K,seed,unm", LANE_CASES, ids=[c[0] for c in LANE_CASES])
def test_lane_kernel_gives_the_same_cuts(name, make, K, seed, unm, monkeypatch):
    """The lane-per-scenario kernel (k1_lane.cu; SGUFP_K1_MODE=lane forces it on small batches too) and the
    warp-per-scenario kernel (SGUFP_K1_MODE=warp) are two independently written K1s: same statuses, objectives and
    bit-identical cuts as each other and as Oracle B."""
    inst = make()
    net = OracleNet(inst)
    gs = sg.GuroSolver(inst)
    paths = I.random_paths(net, K, seed, unm)
    monkeypatch.setenv("SGUFP_K1_MODE", "warp")
    a = gs.solve_paths(paths)
    monkeypatch.setenv("SGUFP_K1_MODE", "lane")
    b = gs.solve_paths(paths)
    assert (a.cut_type == b.cut_type).all() and (a.first_infeasible == b.first_infeasible).all()
    assert (a.rhs == b.rhs).all() and (a.coef_dense == b.coef_dense).all()
    assert (a.status == b.status).all() and (a.obj == b.obj).all()
    _compare(inst, gs, net, paths, b)


def test_lane_mode_leaves_lower_bounds_to_the_warp_kernel(monkeypatch):
    """An instance with positive lower bounds is not for the lane kernel: forcing the mode must not change the answer."""
    inst = I.config2(S=64, lower_prob=0.05)
    net = OracleNet(inst)
    gs = sg.GuroSolver(inst)
    paths = I.random_paths(net, 4, 29, 0.2)
    monkeypatch.setenv("SGUFP_K1_MODE", "lane")
    _compare(inst, gs, net, paths, gs.solve_paths(paths))


@pytest.mark.parametrize("block", range(2))
def test_lane_kernel_fuzz(block, monkeypatch):
    """Random networks without lower bounds, 45-scenario blocks (a full warp and a ragged one), forced lane kernel."""
    monkeypatch.setenv("SGUFP_K1_MODE", "lane")
    rng = np.random.default_rng(1700 + block)
    done = 0
    for k in range(25):
        nl = int(rng.integers(2, 6))
        layers = [int(rng.integers(2, 12)) for _ in range(nl)]
        max_m = sum(a * b for a, b in zip(layers[:-1], layers[1:])) + layers[0] + layers[-1]
        m = int(rng.integers(max(sum(layers) + 2, max_m // 3), max_m + 1))
        try:
            inst = I.make_layered(layers, m, 45, 5000 + 100 * block + k, float(rng.uniform(0.1, 0.95)), 0.0, f"lfz{k}")
            net = OracleNet(inst)
        except Exception:
            continue
        gs = sg.GuroSolver(inst)
        paths = I.random_paths(net, 4, k, float(rng.choice([0.0, 0.2, 0.6])))
        _compare(inst, gs, net, paths, gs.solve_paths(paths))
        done += 1
    assert done >= 12


@pytest.mark.parametrize("block", range(3))
def test_fuzz_small_networks(block):
    """Random small networks (dense and sparse, few and many V-bar nodes, with and without lower bounds,
    33-scenario blocks so that warps of one CTA sit on different candidates): every cut through the C
    ABI against Oracle B.  The same generator drives the host emulation in test_k1_emulated_cpu.py."""
    rng = np.random.default_rng(900 + block)
    done = 0
    for k in range(25):
        nl = int(rng.integers(2, 5))
        layers = [int(rng.integers(2, 7)) for _ in range(nl)]
        max_m = sum(a * b for a, b in zip(layers[:-1], layers[1:])) + layers[0] + layers[-1]
        m = int(rng.integers(max(sum(layers) + 2, max_m // 2), max_m + 1))
        try:
            inst = I.make_layered(layers, m, 33, 3000 + 100 * block + k, float(rng.uniform(0.3, 0.95)), float(rng.choice([0.0, 0.05, 0.3])), f"gfz{k}")
            net = OracleNet(inst)
        except Exception:
            continue
        gs = sg.GuroSolver(inst)
        paths = I.random_paths(net, 5, k, float(rng.choice([0.0, 0.2, 0.6])))
        _compare(inst, gs, net, paths, gs.solve_paths(paths))
        done += 1
    assert done >= 12


# ---- runs of consecutive candidates: warm starts inside a work item (k1_cut.cu: warm_repair) --------------------------------
WARM_CASES = [
    ("c1", lambda: I.config1(S=50), 8, 1, 1, 0.15),
    ("c1_lb", lambda: I.config1(S=50, lower_prob=0.3), 8, 2, 1, 0.3),
    ("c2", lambda: I.config2(S=100), 12, 3, 3, 0.1),
    ("c2_lb", lambda: I.config2(S=100, lower_prob=0.05), 8, 5, 2, 0.3),
    ("c4", lambda: I.config4(S=24), 6, 6, 6, 0.15),
    ("c4_lb", lambda: I.config4(S=24, lower_prob=0.02), 5, 7, 4, 0.3),
    ("large_m6000", lambda: I.make_layered([100, 100, 100, 100, 100, 90], 6000, 3, 99, 0.7, 0.01, "large"), 4, 8, 5, 0.2),
]


@pytest.mark.parametrize("group", ["2", "16"])
@pytest.mark.parametrize("name,make,K,seed,changes,unm", WARM_CASES, ids=[c[0] for c in WARM_CASES])
def test_warm_started_runs_parity_with_oracle(name, make, K, seed, changes, unm, group, monkeypatch):
    """Candidates a few layers apart (the Benders loop's shape), solved in runs in which every candidate after the first
    starts from its predecessor's optimal flow and potentials: each cut must be what Oracle B gives for that candidate alone."""
    monkeypatch.setenv("SGUFP_K1_GROUP", group)
    inst = make()
    net = OracleNet(inst)
    gs = sg.GuroSolver(inst)
    paths = I.perturbed_paths(net, K, seed, changes, unm)
    res = gs.solve_paths(paths)
    _compare(inst, gs, net, paths, res)


@pytest.mark.parametrize("wl,S,K", [("c2", 1000, 64), ("c4", 1500, 16)])
def test_warm_runs_equal_cold_starts_bit_for_bit(wl, S, K, monkeypatch):
    """The bench's DD-emitted candidates: runs of 1 (every candidate from zero flow), 3, 8 and 16 give identical sums, objectives
    and cuts — the duals do not depend on which optimal flow the kernel reaches."""
    import bench
    inst = I.config2(S=S) if wl == "c2" else I.config4(S=S)
    gs = sg.GuroSolver(inst)
    paths, _ = bench.candidate_paths(wl, K, 0)
    outs = []
    monkeypatch.setenv("SGUFP_K1_ORDER", "1")          # the runs follow the nearest-neighbour chain whatever the batch's size
    for group in ("1", "3", "8", "16"):
        monkeypatch.setenv("SGUFP_K1_GROUP", group)
        r = gs.solve_paths(paths)
        outs.append(r)
        assert (r.cut_type == 0).all() and (r.status == 0).all()
    monkeypatch.setenv("SGUFP_K1_GROUP", "8")
    monkeypatch.setenv("SGUFP_K1_ORDER", "0")          # the runs take the candidates as given instead of along the nearest-neighbour chain
    outs.append(gs.solve_paths(paths))
    for r in outs[1:]:
        assert (r.rhs == outs[0].rhs).all() and (r.coef_dense == outs[0].coef_dense).all() and (r.obj == outs[0].obj).all()
        assert (r.nnz == outs[0].nnz).all()


@pytest.mark.parametrize("name,make,K,seed,changes,unm", WARM_CASES[:6], ids=[c[0] for c in WARM_CASES[:6]])
def test_one_path_per_call_keeps_state_between_calls(name, make, K, seed, changes, unm):
    """solveSubProblem(path) call after call, as NodeExplorer::process does it (NodeExplorer.cpp:949-971): the handle keeps the
    optimal flow and potentials of the last candidate per scenario and warm-starts the next call from them; every cut must be
    what a fresh handle (which starts from zero flow) and Oracle B give for that path alone."""
    inst = make()
    net = OracleNet(inst)
    gs = sg.GuroSolver(inst)
    paths = I.perturbed_paths(net, K, seed, changes, unm)
    seq = [gs.solve_paths(paths[k:k + 1]) for k in range(K)]
    fresh = sg.GuroSolver(inst)
    for k in range(K):
        _compare(inst, gs, net, paths[k:k + 1], seq[k])
        one = sg.GuroSolver(inst).solve_paths(paths[k:k + 1]) if k % 3 == 0 else None
        if one is not None:
            assert (one.rhs == seq[k].rhs).all() and (one.coef_dense == seq[k].coef_dense).all() and (one.obj == seq[k].obj).all()
    # batches after single paths and the other way round share the same state
    both = gs.solve_paths(paths)
    alone = fresh.solve_paths(paths)
    assert (both.rhs == alone.rhs).all() and (both.coef_dense == alone.coef_dense).all() and (both.obj == alone.obj).all()
