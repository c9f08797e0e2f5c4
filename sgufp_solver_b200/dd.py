"""Host-side mirror of the reference's decision-diagram interface (K2 behind the C ABI).

`RelaxedDDNew` / `RestrictedDDNew` keep the member names, argument meaning and results of
`/root/reference/DD.h:797-808` and `DD.h:713-728`.  The diagram structure is built on the host
exactly as the reference builds it; every `apply*Cut` runs the layer-wise longest path on the GPU.
`apply_optimality_batch` is the batched form: B diagrams x C cuts in one launch pair.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import cip, dp, i16p, i64p, ip, u64p, vp

DD_SIGNATURES = {
    "sgufp_dd_create": (C.c_int, [vp, C.c_int, C.c_int, C.POINTER(vp)]),
    "sgufp_dd_destroy": (None, [vp]),
    "sgufp_dd_build": (C.c_int, [vp, i16p, C.c_int, i16p, C.c_int, C.c_int, cip]),
    "sgufp_dd_is_exact": (C.c_int, [vp]),
    "sgufp_dd_num_layers": (C.c_int, [vp]),
    "sgufp_dd_layer_sizes": (C.c_int, [vp, ip]),
    "sgufp_dd_counts": (C.c_int, [vp, i64p, i64p]),
    "sgufp_dd_dump": (C.c_int, [vp, ip, dp, i64p, ip, ip, dp]),
    "sgufp_dd_dump_device": (C.c_int, [vp, ip, i64p, ip, ip, ip]),
    "sgufp_dd_apply_optimality": (C.c_int, [vp, C.c_double, u64p, dp, C.c_int, C.c_double, C.c_double, dp]),
    "sgufp_dd_apply_feasibility": (C.c_int, [vp, C.c_double, u64p, dp, C.c_int, cip]),
    "sgufp_dd_solution": (C.c_int, [vp, i16p, C.c_int]),
    "sgufp_dd_cutset": (C.c_int, [vp, C.c_double, ip, C.c_int]),
    "sgufp_dd_apply_optimality_batch": (C.c_int, [C.POINTER(vp), C.c_int, dp, u64p, dp, ip, C.c_int, dp]),
    "sgufp_dd_apply_sequence": (C.c_int, [vp, C.c_int, dp, u64p, dp, ip, C.c_int, C.c_double, dp, cip, cip]),
    "sgufp_dd_last_stats": (C.c_int, [vp, C.POINTER(C.c_float), i64p, cip]),
}
_bound = False


def _lib_dd():
    global _bound
    L = _lib.lib()
    if not _bound:
        for name, (res, args) in DD_SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _bound = True
    return L


_lib.DD_SIGNATURES = DD_SIGNATURES

DOUBLE_MIN = -np.finfo(np.float64).max   # Inavap::DOUBLE_MIN (DD.h:453)
DOUBLE_MAX = np.finfo(np.float64).max


class Node:
    """`Inavap::Node` (DD.h:456-479)."""

    def __init__(self, states=(), solutionVector=(), lb=DOUBLE_MIN, ub=DOUBLE_MIN, globalLayer=0):
        self.states = list(states)
        self.solutionVector = list(solutionVector)
        self.lb, self.ub, self.globalLayer = lb, ub, int(globalLayer)

    def __repr__(self):
        return f"Node(gl={self.globalLayer}, states={self.states}, sol={self.solutionVector})"


def _cut_arrays(cut):
    keys = np.ascontiguousarray(cut.keys, dtype=np.uint64)
    vals = np.ascontiguousarray(cut.vals, dtype=np.float64)
    return keys, vals


def _unpack_nodes(buf, k, ub=DOUBLE_MIN) -> List[Node]:
    out, i = [], 0
    while i < k:
        gl = int(buf[i]); ns = int(buf[i + 1]); st = buf[i + 2:i + 2 + ns].tolist(); i += 2 + ns
        nl = int(buf[i]); sol = buf[i + 1:i + 1 + nl].tolist(); i += 1 + nl
        out.append(Node(st, sol, DOUBLE_MIN, ub, gl))
    return out


class _DDBase:
    KIND = 0

    def __init__(self, solver, width: int = 0):
        """`solver` is the `GuroSolver` handle that owns the network model (Network.h:69-117)."""
        L = _lib_dd()
        self.solver = solver
        h = vp()
        solver._check(L.sgufp_dd_create(solver.h, self.KIND, int(width), C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            _lib_dd().sgufp_dd_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _build(self, node: Node) -> int:
        st = np.ascontiguousarray(node.states, dtype=np.int16)
        so = np.ascontiguousarray(node.solutionVector, dtype=np.int16)
        n = C.c_int(0)
        self.solver._check(_lib_dd().sgufp_dd_build(self.h, st.ctypes.data_as(i16p), len(st), so.ctypes.data_as(i16p), len(so),
                                                   int(node.globalLayer), C.byref(n)))
        return n.value

    def isTreeExact(self) -> bool:
        return bool(_lib_dd().sgufp_dd_is_exact(self.h))

    def layer_sizes(self) -> np.ndarray:
        """Sizes of tree[0..], terminal layer included (always 1) to match the reference's `tree`."""
        n = _lib_dd().sgufp_dd_num_layers(self.h)
        out = np.zeros(n, np.int32)
        _lib_dd().sgufp_dd_layer_sizes(self.h, out.ctypes.data_as(ip))
        return np.append(out, 1).astype(np.int32)

    def counts(self) -> Tuple[int, int]:
        a, b = C.c_int64(), C.c_int64()
        _lib_dd().sgufp_dd_counts(self.h, C.byref(a), C.byref(b))
        return a.value, b.value

    def dump(self):
        nn, na = self.counts()
        sizes = self.layer_sizes()
        nlast = int(sizes[-2])
        na_in = na - nlast
        node_layer = np.zeros(nn, np.int32); node_state = np.zeros(nn); inptr = np.zeros(nn + 1, np.int64)
        tailpos = np.zeros(max(1, na_in), np.int32); dec = np.zeros(max(1, na_in), np.int32); term = np.zeros(max(1, nlast))
        self.solver._check(_lib_dd().sgufp_dd_dump(self.h, node_layer.ctypes.data_as(ip), node_state.ctypes.data_as(dp), inptr.ctypes.data_as(i64p),
                                                  tailpos.ctypes.data_as(ip), dec.ctypes.data_as(ip), term.ctypes.data_as(dp)))
        return dict(node_layer=node_layer, node_state=node_state, in_ptr=inptr, arc_tailpos=tailpos[:na_in], arc_decision=dec[:na_in],
                    terminal_weight=term[:nlast])

    def _apply_sequence(self, mode: int, cuts, optimal: float):
        """A run of cuts in ONE call (sgufp_dd_apply_sequence): returns (values, applied) where values[k] is what
        the k-th applyOptimalityCut / applyFeasibilityCut of the reference's loop returns and `applied` counts the
        cuts up to and including the one that ends that loop."""
        Cn = len(cuts)
        if Cn == 0:
            return (np.zeros(0) if mode == 0 else np.zeros(0, np.int32)), 0
        rhs, keys, vals, ptr = _pack_cuts(cuts)
        bound = np.zeros(Cn); feas = np.zeros(Cn, np.int32); applied = C.c_int()
        self.solver._check(_lib_dd().sgufp_dd_apply_sequence(self.h, mode, rhs.ctypes.data_as(dp), keys.ctypes.data_as(u64p), vals.ctypes.data_as(dp),
                                                            ptr.ctypes.data_as(ip), Cn, float(optimal), bound.ctypes.data_as(dp),
                                                            feas.ctypes.data_as(cip), C.byref(applied)))
        n = applied.value
        return (bound[:n] if mode == 0 else feas[:n]), n

    def applyOptimalityCuts(self, cuts, optimal: float = DOUBLE_MIN):
        """the loop `for cut: ub = applyOptimalityCut(cut, optimal, ub); if ub <= optimal: return` in one device call"""
        return self._apply_sequence(0, cuts, optimal)

    def applyFeasibilityCuts(self, cuts):
        """the loop `for cut: if not applyFeasibilityCut(cut): return` in one device call"""
        return self._apply_sequence(1, cuts, 0.0)

    def dump_device(self):
        """The CSR image as it sits on the device (test introspection).  `built_on_device` tells whether
        k2_build made it or the host mirror was uploaded."""
        nl = _lib_dd().sgufp_dd_num_layers(self.h)
        sizes = np.zeros(nl, np.int32)
        r = _lib_dd().sgufp_dd_dump_device(self.h, sizes.ctypes.data_as(ip), None, None, None, None)
        if r < 0:
            self.solver._check(r)
        nn = int(sizes.sum()); na = nn - 1 if nn else 0
        inptr = np.zeros(nn + 1, np.int64)
        _lib_dd().sgufp_dd_dump_device(self.h, None, inptr.ctypes.data_as(i64p), None, None, None)
        na = int(inptr[nn])
        tailpos = np.zeros(max(1, na), np.int32); dec = np.zeros(max(1, na), np.int32); slot = np.zeros(max(1, na), np.int32)
        r = _lib_dd().sgufp_dd_dump_device(self.h, None, None, tailpos.ctypes.data_as(ip), dec.ctypes.data_as(ip), slot.ctypes.data_as(ip))
        return dict(layer_sizes=sizes, in_ptr=inptr, arc_tailpos=tailpos[:na], arc_decision=dec[:na], arc_slot=slot[:na], built_on_device=bool(r))

    def getSolution(self) -> np.ndarray:
        buf = np.zeros(self.solver.L + 8, np.int16)
        k = _lib_dd().sgufp_dd_solution(self.h, buf.ctypes.data_as(i16p), len(buf))
        if k < 0:
            self.solver._check(k)
        return buf[:k].copy()

    def last_stats(self):
        ms, arcs, n = C.c_float(), C.c_int64(), C.c_int()
        _lib_dd().sgufp_dd_last_stats(self.h, C.byref(ms), C.byref(arcs), C.byref(n))
        return ms.value, arcs.value, n.value

    def _cutset(self, ub) -> List[Node]:
        buf = np.zeros(1 << 22, np.int32)
        k = _lib_dd().sgufp_dd_cutset(self.h, float(ub), buf.ctypes.data_as(ip), len(buf))
        if k < 0:
            self.solver._check(k)
        return _unpack_nodes(buf, k, ub)


class RelaxedDDNew(_DDBase):
    """`Inavap::RelaxedDDNew` (DD.h:734-810).  `threshold` is the collapse threshold: the reference's compile-time
    RELAXED_MAX_WIDTH = 120 (DD.h:732) by default, a runtime parameter here for the width sweep of config C3."""
    KIND = 0

    def __init__(self, solver, threshold: int = 0):
        super().__init__(solver, threshold)

    def buildTree(self, node: Node = None) -> None:
        self._build(node or Node())

    def applyOptimalityCut(self, cut, optimal: float, upperbound: float) -> float:
        keys, vals = _cut_arrays(cut)
        b = C.c_double()
        self.solver._check(_lib_dd().sgufp_dd_apply_optimality(self.h, float(cut.RHS), keys.ctypes.data_as(u64p), vals.ctypes.data_as(dp), len(keys),
                                                              float(optimal), float(upperbound), C.byref(b)))
        return b.value

    def applyFeasibilityCut(self, cut) -> int:
        keys, vals = _cut_arrays(cut)
        f = C.c_int()
        self.solver._check(_lib_dd().sgufp_dd_apply_feasibility(self.h, float(cut.RHS), keys.ctypes.data_as(u64p), vals.ctypes.data_as(dp), len(keys), C.byref(f)))
        return f.value

    def getCutset(self, ub: float) -> List[Node]:
        return self._cutset(ub)


class RestrictedDDNew(_DDBase):
    """`Inavap::RestrictedDDNew` (DD.h:653-730); `width` is the runtime max_width (DD.h:710)."""
    KIND = 1

    def __init__(self, solver, width: int):
        super().__init__(solver, width)
        self._cs = None

    def compile(self, node: Node = None):
        """Returns the exact cut-set (list of Node) or None when the tree is exact (`nullopt`)."""
        n = self._build(node or Node())
        return None if n < 0 else self._cutset(DOUBLE_MIN)

    buildTree = compile

    def applyOptimalityCut(self, cut) -> float:
        keys, vals = _cut_arrays(cut)
        b = C.c_double()
        self.solver._check(_lib_dd().sgufp_dd_apply_optimality(self.h, float(cut.RHS), keys.ctypes.data_as(u64p), vals.ctypes.data_as(dp), len(keys),
                                                              0.0, 0.0, C.byref(b)))
        return b.value

    def applyFeasibilityCut(self, cut) -> int:
        keys, vals = _cut_arrays(cut)
        f = C.c_int()
        self.solver._check(_lib_dd().sgufp_dd_apply_feasibility(self.h, float(cut.RHS), keys.ctypes.data_as(u64p), vals.ctypes.data_as(dp), len(keys), C.byref(f)))
        return f.value

    def getMaxPath(self) -> np.ndarray:
        return self.getSolution()


def _pack_cuts(cuts):
    Cn = len(cuts)
    rhs = np.array([c.RHS for c in cuts], np.float64)
    ptr = np.zeros(Cn + 1, np.int32)
    for i, c in enumerate(cuts):
        ptr[i + 1] = ptr[i] + len(c.keys)
    keys = np.concatenate([np.asarray(c.keys, np.uint64) for c in cuts]) if ptr[-1] else np.zeros(1, np.uint64)
    vals = np.concatenate([np.asarray(c.vals, np.float64) for c in cuts]) if ptr[-1] else np.zeros(1, np.float64)
    return rhs, np.ascontiguousarray(keys), np.ascontiguousarray(vals), ptr


def apply_optimality_batch(dds: Sequence[_DDBase], cuts) -> np.ndarray:
    """B diagrams x C optimality cuts in one launch pair; returns bound[B] (DD.cpp:3975-3984)."""
    L = _lib_dd()
    B, Cn = len(dds), len(cuts)
    hs = (vp * B)(*[d.h for d in dds])
    rhs = np.array([c.RHS for c in cuts], np.float64)
    ptr = np.zeros(Cn + 1, np.int32)
    for i, c in enumerate(cuts):
        ptr[i + 1] = ptr[i] + len(c.keys)
    keys = np.concatenate([np.asarray(c.keys, np.uint64) for c in cuts]) if ptr[-1] else np.zeros(1, np.uint64)
    vals = np.concatenate([np.asarray(c.vals, np.float64) for c in cuts]) if ptr[-1] else np.zeros(1, np.float64)
    keys = np.ascontiguousarray(keys); vals = np.ascontiguousarray(vals)
    bound = np.zeros(B)
    dds[0].solver._check(L.sgufp_dd_apply_optimality_batch(hs, B, rhs.ctypes.data_as(dp), keys.ctypes.data_as(u64p), vals.ctypes.data_as(dp),
                                                           ptr.ctypes.data_as(ip), Cn, bound.ctypes.data_as(dp)))
    return bound


def random_cut(solver, rng, cut_type=0):
    """A pseudo-cut over the network's own keys, as tests2.cpp:262-291 (getActualCut) builds them."""
    from .solver import Cut, getKey
    keys, vals = [], []
    order = np.lexsort((solver.slot_j[:solver.T], solver.slot_q[:solver.T], solver.slot_i[:solver.T]))
    for s in order:
        if rng.integers(0, 11) % 2 == 0:
            keys.append(getKey(solver.slot_q[s], solver.slot_i[s], solver.slot_j[s]))
            vals.append(float(rng.uniform(-100, 100)))
    rhs = float(rng.uniform(-100, 100))
    rhs = rhs * 10 if cut_type == 0 else abs(rhs) * 7
    return Cut(rhs, keys, vals)


def bench_longest_path(device: int = 0, B: int = 64, Cn: int = 64, width: int = 1024, reps: int = 5):
    """DD arcs/s of K2 on config C3: C2 network, restricted width `width`, B diagrams x C cuts."""
    from . import instances as I
    from .solver import GuroSolver
    inst = I.config2(S=1)
    solver = GuroSolver(inst, device=device)
    rng = np.random.default_rng(5)
    dds = []
    for _ in range(B):
        d = RestrictedDDNew(solver, width)
        d.compile()
        dds.append(d)
    cuts = [random_cut(solver, rng) for _ in range(Cn)]
    apply_optimality_batch(dds, cuts)
    times = []
    for _ in range(reps):
        apply_optimality_batch(dds, cuts)
        ms, arcs, launches = dds[0].last_stats()
        times.append(ms)
    ms = float(np.median(times))
    nodes, arcs1 = dds[0].counts()
    bytes_alg = B * Cn * (arcs1 * 16 + nodes * 8) + Cn * solver.T * 8
    return {"metric": "dd_arcs_per_sec", "value": arcs / (ms / 1e3), "unit": "arcs/s", "kernel_ms": ms, "arcs_per_launch": int(arcs),
            "config": f"C3: C2 network, RestrictedDDNew width {width}, {B} diagrams x {Cn} cuts per launch",
            "achieved_GBps": bytes_alg / (ms / 1e3) / 1e9, "gpu_launches": launches}


def frontier_nodes(solver, want: int) -> List[Node]:
    """Distinct open nodes of one branch-and-bound frontier: the cut-set of the root's relaxed diagram, widened breadth-first
    through the children's own cut-sets until `want` nodes are there (DDSolver.cpp:788-791 forms the initial frontier so)."""
    dd = RelaxedDDNew(solver)
    dd.buildTree()
    frontier = dd.getCutset(DOUBLE_MAX) if not dd.isTreeExact() else []
    guard = 0
    while frontier and len(frontier) < want and guard < 64:
        node = frontier.pop(0)
        dd.buildTree(node)
        kids = dd.getCutset(DOUBLE_MAX) if not dd.isTreeExact() else []
        frontier.extend(kids if kids else [])
        if not kids:
            frontier.append(node)      # an exact sub-tree stays a frontier node itself
        guard += 1
    dd.close()
    return frontier[:want] if frontier else [Node()]


def bench_frontier(device: int = 0, widths=(64, 1024, 4096), B: int = 64, Cn: int = 64, reps: int = 5):
    """DD arcs/s of K2 on config C3 (C2 network): B DISTINCT diagrams — the restricted / relaxed sub-trees of B nodes of one
    branch-and-bound frontier — x Cn cuts per launch pair, L2 flushed before every timed launch.
    The scope table's byte model (16 B per arc + 8 B per node through HBM) does not describe this kernel: the
    node states never leave shared memory and a diagram's CSR is read once for its Cn cuts, so `model_GBps` is reported as
    what it is — a model — next to the DRAM bytes ncu measured (profiles/k2_traffic.json), and what bounds the kernel is
    stated there (issue slots and one barrier per layer), not an HBM fraction."""
    import json as _json
    import os as _os
    import torch
    from . import instances as I
    from .solver import GuroSolver
    inst = I.config2(S=1)
    solver = GuroSolver(inst, device=device)
    rng = np.random.default_rng(5)
    nodes = frontier_nodes(solver, B)
    cuts = [random_cut(solver, rng) for _ in range(Cn)]
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=f"cuda:{device}")
    tf = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "profiles", "k2_traffic.json")
    traffic = _json.load(open(tf)) if _os.path.exists(tf) else {}
    points = []
    for kind, w in [("restricted", int(x)) for x in widths] + [("relaxed", 120)]:
        dds = []
        for i in range(B):
            d = RestrictedDDNew(solver, w) if kind == "restricted" else RelaxedDDNew(solver)
            d.compile(nodes[i % len(nodes)]) if kind == "restricted" else d.buildTree(nodes[i % len(nodes)])
            dds.append(d)
        apply_optimality_batch(dds, cuts)
        times, arcs, launches = [], 0, 0
        for _ in range(reps):
            flush.zero_(); torch.cuda.synchronize()
            apply_optimality_batch(dds, cuts)
            ms, arcs, launches = dds[0].last_stats()
            times.append(ms)
        ms = float(np.median(times))
        cnt = [d.counts() for d in dds]
        nn, na = sum(c[0] for c in cnt), sum(c[1] for c in cnt)
        model = Cn * (na * 16 + nn * 8) + Cn * solver.T * 8
        points.append({"kind": kind, "width": w, "value": arcs / (ms / 1e3), "unit": "arcs/s", "kernel_ms": ms, "arcs_per_launch": int(arcs),
                       "diagrams": B, "distinct_roots": min(B, len(nodes)), "cuts": Cn, "nodes_all_diagrams": int(nn), "arcs_all_diagrams": int(na),
                       "model_GBps": model / (ms / 1e3) / 1e9, "gpu_launches": launches,
                       "ncu": traffic.get(f"{kind}_{w}")})
        for d in dds:
            d.close()
    head = next(p for p in points if p["kind"] == "restricted" and p["width"] == 1024) if any(p["width"] == 1024 for p in points) else points[0]
    del flush
    return {"metric": "dd_arcs_per_sec", "value": head["value"], "unit": "arcs/s", "kernel_ms": head["kernel_ms"], "arcs_per_launch": head["arcs_per_launch"],
            "config": f"C3: C2 network, {B} distinct diagrams of one B&B frontier x {Cn} cuts per launch pair, RestrictedDDNew widths {list(widths)} + RelaxedDDNew (threshold 120); headline = width 1024",
            "l2": "flushed (512 MiB write) before every timed launch", "points": points,
            "bound": traffic.get("bound", "node states stay in shared memory, the CSR of a diagram is read once for its cuts: not an HBM-bound kernel; see profiles/r02_k2_bound.md")}


def smoke():
    from . import instances as I
    from .solver import GuroSolver
    inst = I.config1(S=2)
    solver = GuroSolver(inst, device=0)
    dd = RelaxedDDNew(solver)
    dd.buildTree()
    from .solver import Cut, getKey
    keys, vals, v = [], [], 1.5
    for a in solver.layer_arc:
        i, q = int(inst.tail[a]), int(inst.head[a])
        for b in solver.out_arcs(q):
            keys.append(getKey(q, i, int(inst.head[b]))); vals.append(v); v += 0.75
    bound = dd.applyOptimalityCut(Cut(-3.25, keys, vals), -1e300, 1e300)
    sol = dd.getSolution().tolist()
    assert bound == 32.0 and sol == [9, 10, 11, -1, 12, 13], (bound, sol)   # SURVEY.md Appendix A.3 (reference output)
    print(f"dd smoke ok: bound={bound} solution={sol}")
