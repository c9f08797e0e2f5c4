"""The K1 kernel BODY on the CPU (tests/cpp/k1_emul.cpp compiles sgufp_solver_b200/csrc/k1_cut.cu for
the host with one lane per scenario) against Oracle B: statuses, objectives and the exact integer
partial sums of every accumulator must be identical.  This checks the kernel's algorithm (plan
decoding, level-wise shortest paths, blocking flow, SPEC-LP potentials, lifting, the feasibility
ray) where there is no GPU; the warp-parallel execution is covered by the `-m gpu` tests.
Test infrastructure only — the product never runs this build."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from helpers import wlayout_partial
from oracle.oracle import OracleNet
from sgufp_solver_b200 import instances as I

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
I64_MAX = np.iinfo(np.int64).max


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    return _build_emul(tmp_path_factory, ["SGUFP_K1_SKIP_CONFIRM", "SGUFP_K1_PUSH_PAR"])   # as the product library is built (build.py)


def _build_emul(tmp_path_factory, defines):
    so = str(tmp_path_factory.mktemp("emul") / "libk1_emul.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-w", "-fPIC", "-shared", f"-I{ROOT}/tests/cpp/emul_stub", f"-I{ROOT}/include"] +
                          [f"-D{d}" for d in defines] + [f"{ROOT}/tests/cpp/k1_emul.cpp", "-o", so])
    L = C.CDLL(so)
    ip, i16p = C.POINTER(C.c_int32), C.POINTER(C.c_int16)
    L.emul_paths.restype = C.c_int
    L.emul_paths.argtypes = [C.c_int, C.c_int, C.c_int, ip, ip, ip, ip, ip, ip, C.c_int, i16p, C.c_int, C.c_int,
                             C.POINTER(C.c_longlong), C.POINTER(C.c_longlong), C.POINTER(C.c_double), C.POINTER(C.c_uint8), C.POINTER(C.c_longlong), C.c_int]
    return L


def run_emul(L, inst, net, paths, lane_variant=0):
    ip, i16p = C.POINTER(C.c_int32), C.POINTER(C.c_int16)
    K, plen = paths.shape
    W = 1 + net.L + inst.m
    arr = lambda a: np.ascontiguousarray(a, dtype=np.int32)
    t, h, u, lo, r0, vb = arr(inst.tail), arr(inst.head), arr(inst.upper), arr(inst.lower), arr(inst.reward[:, 0]), arr(inst.vbar)
    p = np.ascontiguousarray(paths, dtype=np.int16)
    sums = np.zeros((K, W), np.int64); finf = np.zeros(K, np.int64); obj = np.zeros((K, inst.S)); st = np.zeros((K, inst.S), np.uint8)
    ray = np.zeros((K, W), np.int64)
    rc = L.emul_paths(inst.n, inst.m, inst.S, t.ctypes.data_as(ip), h.ctypes.data_as(ip), u.ctypes.data_as(ip), lo.ctypes.data_as(ip),
                      r0.ctypes.data_as(ip), vb.ctypes.data_as(ip), len(vb), p.ctypes.data_as(i16p), K, plen,
                      sums.ctypes.data_as(C.POINTER(C.c_longlong)), finf.ctypes.data_as(C.POINTER(C.c_longlong)),
                      obj.ctypes.data_as(C.POINTER(C.c_double)), st.ctypes.data_as(C.POINTER(C.c_uint8)), ray.ctypes.data_as(C.POINTER(C.c_longlong)), int(lane_variant))
    if lane_variant and rc in (98, 99):
        return None          # 99: positive lower bounds; 98: these state widths do not fit the instance — the lane kernel does not take it
    assert rc == 0
    return sums, finf, obj, st, ray


def ray_wlayout(net, inst, path, s):
    d = net.scenario_ray(path, s)
    L, m = net.L, net.m
    out = np.zeros(1 + L + m, np.int64)
    u = inst.upper[:, s].astype(np.int64); lo = inst.lower[:, s].astype(np.int64)
    out[0] = int((u * d["gamma"]).sum() - (lo * d["beta"]).sum())
    t = 0
    for l, a in enumerate(net.layer_arc):
        for b in net.out_arcs(inst.head[a]):
            v = int(u[a] * d["lambda"][t] + u[b] * d["mu"][t]); t += 1
            out[0] += v; out[1 + l] += v
    out[1 + L:] = u * d["sigma"] + u * d["phi"]
    return out


CASES = [
    ("c1", lambda: I.config1(S=30), 6, 1, 0.15),
    ("c1_lb", lambda: I.config1(S=30, lower_prob=0.3), 8, 2, 0.3),
    ("c1_all_unmatched", lambda: I.config1(S=5), 2, 3, 1.0),
    ("c2", lambda: I.config2(S=12), 4, 3, 0.1),
    ("c2_lb", lambda: I.config2(S=12, lower_prob=0.1), 4, 4, 0.3),
    ("c4_lb", lambda: I.config4(S=3, lower_prob=0.05), 2, 6, 0.3),
    ("odd_m", lambda: I.make_layered([3, 4, 3], 21, 4, 77, 0.8, 0.1, "odd"), 4, 7, 0.2),
    # > 255 contracted nodes: the searches run over the tight-chain list instead of per-node bit sets
    ("wide_lb", lambda: I.make_layered([70, 70, 70, 70, 60], 1500, 2, 98, 0.3, 0.02, "wide"), 2, 9, 0.2),
]


LANE = {"warp": 0, "lane_small": 1, "lane_mid": 2, "lane_wide": 3}    # state widths of k1_lane.cu: CfgSmall / CfgMid / CfgWide
CASES_LANE = [   # without lower bounds: what the lane-per-scenario kernel takes
    ("c1_nolb", lambda: I.config1(S=40), 6, 11, 0.15),
    ("c2_nolb", lambda: I.config2(S=24), 4, 13, 0.1),
    ("c2_nolb_sparse", lambda: I.config2(S=16), 3, 14, 0.6),
    ("c4_nolb", lambda: I.config4(S=6), 2, 16, 0.1),
    ("odd_m_nolb", lambda: I.make_layered([3, 4, 3], 21, 5, 77, 0.8, 0.0, "odd"), 4, 7, 0.2),
    ("wide_nolb", lambda: I.make_layered([70, 70, 70, 70, 60], 1500, 2, 98, 0.3, 0.0, "wide"), 2, 9, 0.2),
]


@pytest.mark.parametrize("variant", list(LANE))
@pytest.mark.parametrize("name,make,K,seed,unm", CASES + CASES_LANE, ids=[c[0] for c in CASES + CASES_LANE])
def test_kernel_body_matches_oracle(emul, name, make, K, seed, unm, variant):
    inst = make()
    net = OracleNet(inst)
    paths = I.random_paths(net, K, seed, unm)
    out = run_emul(emul, inst, net, paths, lane_variant=LANE[variant])
    if out is None:
        assert variant != "warp" and (inst.lower.max() > 0 or variant != "lane_wide")   # the wide state takes everything without lower bounds
        pytest.skip("not an instance for these lane-kernel state widths")
    sums, finf, obj, st, ray = out
    for k in range(K):
        want, first_bad = wlayout_partial(net, inst, paths[k], 0, inst.S)
        oc = net.solve_path(paths[k])
        assert (finf[k] if finf[k] != I64_MAX else -1) == oc.first_infeasible == (-1 if first_bad is None else first_bad)
        for s in range(inst.S):
            d = net.scenario_duals(paths[k], s)
            assert st[k, s] == d["status"]
            if d["status"] == 0:
                assert obj[k, s] == d["obj"]
        assert (sums[k] == want).all(), np.nonzero(sums[k] != want)
        if first_bad is not None:
            assert (ray[k] == ray_wlayout(net, inst, paths[k], first_bad)).all()


def _random_instance(rng, k, lower=None):
    nl = int(rng.integers(2, 5))
    layers = [int(rng.integers(2, 7)) for _ in range(nl)]
    n_int = sum(layers)
    max_m = sum(a * b for a, b in zip(layers[:-1], layers[1:])) + layers[0] + layers[-1]
    m = int(rng.integers(max(n_int + 2, max_m // 2), max_m + 1))
    return I.make_layered(layers, m, int(rng.integers(2, 6)), 1000 + k, float(rng.uniform(0.3, 0.95)), float(rng.choice([0.0, 0.05, 0.3])) if lower is None else lower, f"fz{k}")


@pytest.mark.parametrize("block", range(4))
def test_kernel_body_fuzz(emul, block):
    """Many small random networks (dense and sparse, few and many V-bar nodes, with and without lower
    bounds): the kernel body's sums, first infeasible scenario and ray must equal Oracle B's."""
    rng = np.random.default_rng(500 + block)
    done = 0
    for k in range(40):
        try:
            inst = _random_instance(rng, 100 * block + k)
            net = OracleNet(inst)
        except Exception:
            continue                                   # the generator refused the shape (too few arcs for the layers)
        paths = I.random_paths(net, 3, k, float(rng.choice([0.0, 0.2, 0.6])))
        sums, finf, obj, st, ray = run_emul(emul, inst, net, paths)
        for j in range(len(paths)):
            want, first_bad = wlayout_partial(net, inst, paths[j], 0, inst.S)
            assert (finf[j] if finf[j] != I64_MAX else -1) == (-1 if first_bad is None else first_bad), (inst.name, j)
            assert (sums[j] == want).all(), (inst.name, j)
            if first_bad is not None:
                assert (ray[j] == ray_wlayout(net, inst, paths[j], first_bad)).all(), (inst.name, j)
        done += 1
    assert done >= 20


@pytest.mark.parametrize("block", range(3))
def test_lane_kernel_body_fuzz(emul, block):
    """The lane-per-scenario kernel body on random networks without lower bounds (small and larger contracted graphs,
    every state width that fits): sums, objectives and statuses must equal Oracle B's."""
    rng = np.random.default_rng(1500 + block)
    done = 0
    for k in range(30):
        try:
            inst = _random_larger_instance(rng, 100 * block + k, lower=0.0) if k % 3 == 0 else _random_instance(rng, 100 * block + k, lower=0.0)
            net = OracleNet(inst)
        except Exception:
            continue
        paths = I.random_paths(net, 3, k, float(rng.choice([0.0, 0.2, 0.6])))
        for lv in (1, 2, 3):
            out = run_emul(emul, inst, net, paths, lane_variant=lv)
            if out is None:
                assert lv != 3
                continue
            sums, finf, obj, st, ray = out
            for j in range(len(paths)):
                want, first_bad = wlayout_partial(net, inst, paths[j], 0, inst.S)
                assert first_bad is None and finf[j] == I64_MAX, (inst.name, j, lv)
                assert (sums[j] == want).all(), (inst.name, j, lv)
                assert (st[j] == 0).all()
                oc = net.solve_path(paths[j])
                assert (obj[j] == oc.obj).all(), (inst.name, j, lv)
        done += 1
    assert done >= 15


def _random_larger_instance(rng, k, lower=None):
    nl = int(rng.integers(3, 7))
    layers = [int(rng.integers(7, 16)) for _ in range(nl)]
    n_int = sum(layers)
    max_m = sum(a * b for a, b in zip(layers[:-1], layers[1:])) + layers[0] + layers[-1]
    m = int(rng.integers(max(n_int + 2, max_m // 4), max(n_int + 3, max_m // 2)))
    # few V-bar nodes: little contraction, so the contracted graph keeps more than 31 nodes
    return I.make_layered(layers, m, int(rng.integers(2, 4)), 3000 + k, float(rng.uniform(0.05, 0.4)), float(rng.choice([0.0, 0.03, 0.2])) if lower is None else lower, f"fzL{k}")


@pytest.mark.parametrize("block", range(2))
def test_kernel_body_fuzz_larger_graphs(emul, block):
    """Random networks whose contracted graph has more than 31 nodes: the size class with the in-place list search,
    flag bytes and 16-bit list entries (k1_cut_eval<.., BIG = true>)."""
    rng = np.random.default_rng(900 + block)
    done = big = 0
    for k in range(12):
        try:
            inst = _random_larger_instance(rng, 100 * block + k)
            net = OracleNet(inst)
        except Exception:
            continue
        paths = I.random_paths(net, 2, k, float(rng.choice([0.0, 0.2, 0.6])))
        sums, finf, obj, st, ray = run_emul(emul, inst, net, paths)
        for j in range(len(paths)):
            want, first_bad = wlayout_partial(net, inst, paths[j], 0, inst.S)
            assert (finf[j] if finf[j] != I64_MAX else -1) == (-1 if first_bad is None else first_bad), (inst.name, j)
            assert (sums[j] == want).all(), (inst.name, j)
            if first_bad is not None:
                assert (ray[j] == ray_wlayout(net, inst, paths[j], first_bad)).all(), (inst.name, j)
        done += 1
        big += emul.emul_last_nc() > 31
    assert done >= 6 and big >= 4, (done, big)


@pytest.fixture(scope="module")
def emul_experiments(tmp_path_factory):
    return _build_emul(tmp_path_factory, [])


def test_experimental_switches_keep_parity(emul_experiments):
    """The kernel body WITHOUT the two list-search / push switches the product build turns on (build.py: SGUFP_K1_SKIP_CONFIRM,
    SGUFP_K1_PUSH_PAR) must give Oracle B's sums too: every compiled-out code path stays parity-checked."""
    emul = emul_experiments
    rng = np.random.default_rng(1234)
    done = big = 0
    for k in range(14):
        try:
            inst = _random_larger_instance(rng, 7000 + k) if k % 2 == 0 else _random_instance(rng, 7000 + k)
            net = OracleNet(inst)
        except Exception:
            continue
        paths = I.random_paths(net, 2, k, float(rng.choice([0.0, 0.2, 0.6])))
        sums, finf, obj, st, ray = run_emul(emul, inst, net, paths)
        for j in range(len(paths)):
            want, first_bad = wlayout_partial(net, inst, paths[j], 0, inst.S)
            assert (finf[j] if finf[j] != I64_MAX else -1) == (-1 if first_bad is None else first_bad), (inst.name, j)
            assert (sums[j] == want).all(), (inst.name, j)
        done += 1
        big += emul.emul_last_nc() > 31
    assert done >= 8 and big >= 3, (done, big)


# ---- runs of consecutive candidates: warm starts (k1_cut.cu: warm_repair) -----------------------------------------------
def warm_counts(L):
    out = (C.c_longlong * 2)()
    L.emul_warm_counts(out)
    return int(out[0]), int(out[1])


def check_against_oracle(inst, net, paths, out):
    sums, finf, obj, st, ray = out
    for k in range(len(paths)):
        want, first_bad = wlayout_partial(net, inst, paths[k], 0, inst.S)
        assert (finf[k] if finf[k] != I64_MAX else -1) == (-1 if first_bad is None else first_bad), (inst.name, k)
        for s in range(inst.S):
            d = net.scenario_duals(paths[k], s)
            assert st[k, s] == d["status"], (inst.name, k, s)
            if d["status"] == 0:
                assert obj[k, s] == d["obj"], (inst.name, k, s)
        assert (sums[k] == want).all(), (inst.name, k, np.nonzero(sums[k] != want))
        if first_bad is not None:
            assert (ray[k] == ray_wlayout(net, inst, paths[k], first_bad)).all(), (inst.name, k)


WARM_CASES = [
    ("c1", lambda: I.config1(S=30), 8, 1, 1, 0.15),
    ("c1_lb", lambda: I.config1(S=30, lower_prob=0.3), 8, 2, 1, 0.3),          # forced flow: those scenarios start from zero flow
    ("c2", lambda: I.config2(S=10), 8, 3, 3, 0.1),
    ("c2_sparse", lambda: I.config2(S=8), 6, 4, 4, 0.5),
    ("c2_lb", lambda: I.config2(S=10, lower_prob=0.1), 6, 5, 2, 0.3),
    ("c4", lambda: I.config4(S=3), 5, 6, 6, 0.15),
    ("c4_lb", lambda: I.config4(S=3, lower_prob=0.05), 4, 7, 4, 0.3),
    ("wide", lambda: I.make_layered([70, 70, 70, 70, 60], 1500, 2, 98, 0.3, 0.0, "wide"), 4, 9, 5, 0.2),
]


@pytest.mark.parametrize("group", [0, 2, 3])
@pytest.mark.parametrize("name,make,K,seed,changes,unm", WARM_CASES, ids=[c[0] for c in WARM_CASES])
def test_warm_started_runs_match_oracle(emul, name, make, K, seed, changes, unm, group):
    """Candidates that differ in a few layers, solved in runs of `group` (0: the whole batch): every candidate after the
    first of a run starts from its predecessor's optimal flow and potentials; sums, objectives, statuses and rays must be
    what Oracle B gives for each candidate on its own."""
    inst = make()
    net = OracleNet(inst)
    paths = I.perturbed_paths(net, K, seed, changes, unm)
    emul.emul_set_group(group)
    warm_counts(emul)
    try:
        out = run_emul(emul, inst, net, paths)
    finally:
        emul.emul_set_group(0)
    taken, given_up = warm_counts(emul)
    check_against_oracle(inst, net, paths, out)
    assert given_up == 0
    if inst.lower.max() == 0:
        runs = -(-K // (group or K))
        assert taken == (K - runs) * inst.S, (taken, given_up)        # every linked candidate was warm-started on every scenario


def test_warm_runs_on_the_bench_candidates(emul):
    """The DD-emitted paths bench.py uses (consecutive paths of the Benders loop: 2 - 8 layers apart)."""
    cand = np.load(os.path.join(ROOT, "sgufp_solver_b200", "data", "bench_candidates.npz"))
    for key, inst, K in (("config2", I.config2(S=6), 12), ("config4", I.config4(S=2), 5)):
        net = OracleNet(inst)
        paths = np.ascontiguousarray(cand[key][:K])
        warm_counts(emul)
        out = run_emul(emul, inst, net, paths)
        taken, given_up = warm_counts(emul)
        check_against_oracle(inst, net, paths, out)
        assert (taken, given_up) == ((K - 1) * inst.S, 0)


def test_the_order_of_a_run_does_not_change_the_cuts(emul):
    """Runs take the candidates along a nearest-neighbour chain (model.hpp: order_batch, from the path whose state is kept) or as
    given: the same sums either way, and Oracle B's; the chain's steps are no longer than the emission's."""
    cand = np.load(os.path.join(ROOT, "sgufp_solver_b200", "data", "bench_candidates.npz"))
    for key, inst, K in (("config2", I.config2(S=5), 16), ("config4", I.config4(S=2), 8)):
        net = OracleNet(inst)
        outs = []
        for order in (1, 0):
            emul.emul_set_order(order)
            emul.emul_state(1)
            try:
                got = []
                for b in range(2):
                    paths = np.ascontiguousarray(cand[key][b * K:(b + 1) * K]) if len(cand[key]) >= 2 * K else np.ascontiguousarray(cand[key][:K][::(1 if b == 0 else -1)])
                    out = run_emul(emul, inst, net, paths)
                    check_against_oracle(inst, net, paths, out)
                    got.append(out[0].copy())
                outs.append(got)
            finally:
                emul.emul_state(0)
                emul.emul_set_order(1)
        for a, b in zip(*outs):
            assert (a == b).all()


def test_order_batch_is_a_nearest_neighbour_chain(emul):
    """model.hpp: order_batch — a permutation; every step goes to the closest path not yet taken (ties: the order given); the chain
    starts at candidate 0, or at the path closest to `start`; batches of more than 128 candidates, and of one, keep the order given."""
    i16p, ip = C.POINTER(C.c_int16), C.POINTER(C.c_int32)
    rng = np.random.default_rng(5)

    def order(paths, start=None):
        K, L = paths.shape
        out = np.zeros(K, np.int32)
        emul.emul_order(np.ascontiguousarray(paths).ctypes.data_as(i16p), K, L,
                        None if start is None else np.ascontiguousarray(start).ctypes.data_as(i16p), out.ctypes.data_as(ip))
        return out.tolist()

    for K, L, with_start in ((1, 9, False), (1, 9, True), (2, 9, False), (2, 9, True), (7, 30, False), (16, 83, True), (64, 83, True), (128, 20, False)):
        base = rng.integers(-1, 6, size=L).astype(np.int16)
        paths = np.repeat(base[None, :], K, axis=0)
        for k in range(K):                                  # a few layers away from a common path, some rows identical
            idx = rng.integers(0, L, size=int(rng.integers(0, 5)))
            paths[k, idx] = rng.integers(-1, 6, size=len(idx))
        start = None
        if with_start:
            start = base.copy(); start[rng.integers(0, L, size=2)] = 7
        o = order(paths, start)
        assert sorted(o) == list(range(K))
        cur, left = start, set(range(K))
        if start is None and K > 2:
            assert o[0] == 0
        if K == 2 and start is None:
            assert o == [0, 1]
        if K >= 2 and not (K == 2 and start is None):
            for j, k in enumerate(o):
                if cur is not None:
                    d = {i: int((paths[i] != cur).sum()) for i in left}
                    best = min(d.values())
                    assert d[k] == best and k == min(i for i in left if d[i] == best), (K, j)
                left.discard(k); cur = paths[k]
    big = rng.integers(-1, 6, size=(129, 12)).astype(np.int16)
    assert order(big, big[5]) == list(range(129))


@pytest.mark.parametrize("block", range(3))
def test_warm_runs_fuzz(emul, block):
    """Random small and larger networks, with and without lower bounds, runs of 1 - 4 changed layers."""
    rng = np.random.default_rng(7700 + block)
    done = taken_all = 0
    for k in range(36):
        try:
            inst = _random_larger_instance(rng, 100 * block + k) if k % 3 == 0 else _random_instance(rng, 100 * block + k)
            net = OracleNet(inst)
        except Exception:
            continue
        paths = I.perturbed_paths(net, 5, k, int(rng.integers(1, 5)), float(rng.choice([0.0, 0.2, 0.6])))
        warm_counts(emul)
        out = run_emul(emul, inst, net, paths)
        taken, given_up = warm_counts(emul)
        check_against_oracle(inst, net, paths, out)
        assert given_up == 0, inst.name
        taken_all += taken
        done += 1
    assert done >= 18 and taken_all > 0


@pytest.mark.parametrize("name,make,K,seed,changes,unm", WARM_CASES, ids=[c[0] for c in WARM_CASES])
@pytest.mark.parametrize("per_call", [1, 2])
def test_state_between_calls_matches_oracle(emul, name, make, K, seed, changes, unm, per_call):
    """One path (or two) per call, as NodeExplorer::process hands them to solveSubProblem: the first candidate of a call starts
    from the flow and potentials the last candidate of the previous call left in the handle's per-scenario state."""
    inst = make()
    net = OracleNet(inst)
    paths = I.perturbed_paths(net, K, seed, changes, unm)
    emul.emul_state(1)
    warm_counts(emul)
    try:
        for k0 in range(0, K, per_call):
            sub = np.ascontiguousarray(paths[k0:k0 + per_call])
            check_against_oracle(inst, net, sub, run_emul(emul, inst, net, sub))
    finally:
        emul.emul_state(0)
    taken, given_up = warm_counts(emul)
    assert given_up == 0
    if inst.lower.max() == 0:
        assert taken == (K - 1) * inst.S, (taken, K, inst.S)       # every candidate but the very first was warm-started
