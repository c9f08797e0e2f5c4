"""Synthetic SGUFP instances in the reference's own data model.

The reference ships no instance files, so every config in BASELINE.json is
generated here.  The generator obeys the preconditions of the reference parser
(`/root/reference/Network.cpp:18-63`) and of `Network::shuffleVBarNodes`
(`Network.cpp:132-186`): layered DAG, node 0 the unique pure source, node n-1 the unique pure
sink, last interior layer = demand points with exactly one out-arc (to n-1), every V-bar node a
backward ancestor of a demand point, simple graph, no source->sink arc (the reference drops A4
rows, `Network.cpp:80`).

`Instance.write_text` emits the reference text format so the unmodified reference `Network`
(compiled as oracle/_ref) reads exactly the same data.
"""
from __future__ import annotations

import dataclasses
import os
from typing import List, Optional, Sequence

import numpy as np


@dataclasses.dataclass
class Instance:
    """Arc-major arrays exactly as `NetworkArc` stores them (`Network.h:26-41`)."""

    n: int
    m: int
    S: int
    tail: np.ndarray  # int32 [m]
    head: np.ndarray  # int32 [m]
    lower: np.ndarray  # int32 [m, S]
    upper: np.ndarray  # int32 [m, S]
    reward: np.ndarray  # int32 [m, S] (the production path only reads column 0, grb.cpp:53)
    vbar: np.ndarray  # int32 [nvbar], file order (before shuffleVBarNodes)
    name: str = "synthetic"

    def write_text(self, path: str) -> None:
        """Reference text format: `n m S`, then per arc `tail head (lb ub reward)xS`, a token,
        then the V-bar node ids (`Network.cpp:18-63`)."""
        with open(path, "w") as f:
            f.write(f"{self.n} {self.m} {self.S}\n")
            for a in range(self.m):
                trip = np.stack([self.lower[a], self.upper[a], self.reward[a]], axis=1).reshape(-1)
                f.write(f"{int(self.tail[a])} {int(self.head[a])} " + " ".join(map(str, trip.tolist())) + "\n")
            f.write("vbar\n")
            f.write(" ".join(str(int(v)) for v in self.vbar) + "\n")

    def scenario_slice(self, lo: int, hi: int) -> "Instance":
        """Contiguous scenario block [lo, hi) — the multi-GPU shard of SURVEY.md §8e.  The rows of the dual LP use
        `rewards[0]` of the FULL instance for every scenario (grb.cpp:53,71,89), so a shard keeps the full instance's
        scenario-0 rewards in its own column 0 (an empty shard keeps one reward column and no capacities)."""
        rew = np.ascontiguousarray(self.reward[:, lo:hi]) if hi > lo else np.ascontiguousarray(self.reward[:, :1])
        rew = rew.copy()
        rew[:, 0] = self.reward[:, 0]
        return dataclasses.replace(
            self,
            S=hi - lo,
            lower=np.ascontiguousarray(self.lower[:, lo:hi]),
            upper=np.ascontiguousarray(self.upper[:, lo:hi]),
            reward=rew,
        )


def _layer_arcs(rng: np.random.Generator, a_nodes: Sequence[int], b_nodes: Sequence[int], count: int):
    """`count` distinct arcs a->b, every a with >=1 out-arc and every b with >=1 in-arc."""
    na, nb = len(a_nodes), len(b_nodes)
    count = max(count, max(na, nb))
    count = min(count, na * nb)
    chosen = set()
    perm_b = rng.permutation(nb)
    for k in range(max(na, nb)):  # cover both sides
        chosen.add((k % na, int(perm_b[k % nb])))
    all_pairs = [(i, j) for i in range(na) for j in range(nb) if (i, j) not in chosen]
    extra = count - len(chosen)
    if extra > 0:
        idx = rng.choice(len(all_pairs), size=extra, replace=False)
        for t in idx:
            chosen.add(all_pairs[int(t)])
    return sorted((a_nodes[i], b_nodes[j]) for (i, j) in chosen)


def make_layered(
    layer_sizes: Sequence[int],
    m: int,
    S: int,
    seed: int,
    vbar_frac: float = 0.5,
    lower_prob: float = 0.0,
    name: str = "synthetic",
    vbar_nodes: Optional[Sequence[int]] = None,
    cap_stream: int = 0,
) -> Instance:
    """Layered DAG per SURVEY.md §8d.  `layer_sizes` are the interior layers (the last one is the
    demand layer); a source and a sink are added.  `lower_prob` is the per-(source arc, scenario)
    probability of a positive lower bound (exercises feasibility cuts); 0 gives the pure
    throughput variant."""
    rng = np.random.default_rng([seed, 0])      # topology, rewards, V-bar: independent of S
    rng_cap = np.random.default_rng([seed, 1, cap_stream])  # per-scenario capacities (one stream per scenario block)
    layers: List[List[int]] = []
    nid = 1
    for sz in layer_sizes:
        layers.append(list(range(nid, nid + sz)))
        nid += sz
    n = nid + 1
    sink = n - 1
    arcs = [(0, v) for v in layers[0]]
    fixed = len(layers[0]) + len(layers[-1])
    gaps = len(layers) - 1
    remaining = m - fixed
    if gaps:
        cap = [len(layers[g]) * len(layers[g + 1]) for g in range(gaps)]
        per = [remaining // gaps + (1 if g < remaining % gaps else 0) for g in range(gaps)]
        # push overflow of dense gaps to the others
        for g in range(gaps):
            if per[g] > cap[g]:
                over = per[g] - cap[g]
                per[g] = cap[g]
                for h in range(gaps):
                    room = cap[h] - per[h]
                    mv = min(room, over) if h != g else 0
                    per[h] += mv
                    over -= mv
        for g in range(gaps):
            arcs += _layer_arcs(rng, layers[g], layers[g + 1], per[g])
    arcs += [(v, sink) for v in layers[-1]]
    m_real = len(arcs)
    tail = np.array([a for a, _ in arcs], dtype=np.int32)
    head = np.array([b for _, b in arcs], dtype=np.int32)

    base_u = rng.integers(5, 21, size=m_real)
    noise = rng_cap.uniform(0.5, 1.5, size=(S, m_real)).T   # scenario-major draw: scenario s is the same for every S
    upper = np.maximum(1, np.rint(base_u[:, None] * noise)).astype(np.int32)
    lower = np.zeros((m_real, S), dtype=np.int32)
    if lower_prob > 0:
        src = np.nonzero(tail == 0)[0]
        mask = rng_cap.random((S, len(src))).T < lower_prob
        vals = rng_cap.integers(1, 4, size=(S, len(src))).T
        lower[src] = np.where(mask, vals, 0).astype(np.int32)
        lower = np.minimum(lower, upper)  # l <= u on every single arc; chains may still be infeasible
    r0 = np.where(head == sink, rng.integers(10, 31, size=m_real), rng.integers(-5, 1, size=m_real))
    reward = np.repeat(r0[:, None], S, axis=1).astype(np.int32)

    if vbar_nodes is None:
        cand = [v for lay in layers[1:-1] for v in lay] if len(layers) > 2 else [v for v in layers[0]]
        k = max(1, int(round(vbar_frac * len(cand))))
        vb = np.sort(rng.choice(np.array(cand), size=min(k, len(cand)), replace=False))
    else:
        vb = np.array(list(vbar_nodes))
    return Instance(n, m_real, S, tail, head, lower, upper, reward, vb.astype(np.int32), name)


# ---- the five BASELINE.json configs (shapes from SURVEY.md §8d) -------------------------------
SEED0 = 20261018


def config1(S: int = 50, lower_prob: float = 0.0) -> Instance:
    """C1: the Appendix-A topology (n=10, m=17, V-bar={4,5}) with S scenarios."""
    arcs = [(0, 1), (0, 2), (0, 3), (1, 4), (2, 4), (3, 4), (1, 5), (2, 5), (3, 5),
            (4, 6), (4, 7), (4, 8), (5, 7), (5, 8), (6, 9), (7, 9), (8, 9)]
    rng = np.random.default_rng([SEED0 + 1, 0])
    rng_cap = np.random.default_rng([SEED0 + 1, 1])
    m = len(arcs)
    tail = np.array([a for a, _ in arcs], dtype=np.int32)
    head = np.array([b for _, b in arcs], dtype=np.int32)
    base_u = rng.integers(5, 21, size=m)
    upper = np.maximum(1, np.rint(base_u[:, None] * rng_cap.uniform(0.5, 1.5, size=(S, m)).T)).astype(np.int32)
    lower = np.zeros((m, S), dtype=np.int32)
    if lower_prob > 0:
        src = np.nonzero(tail == 0)[0]
        mask = rng_cap.random((S, len(src))).T < lower_prob
        lower[src] = np.where(mask, rng_cap.integers(1, 4, size=(S, len(src))).T, 0)
        lower = np.minimum(lower, upper)
    r0 = np.where(head == 9, rng.integers(10, 31, size=m), rng.integers(-5, 1, size=m))
    reward = np.repeat(r0[:, None], S, axis=1).astype(np.int32)
    return Instance(10, m, S, tail, head, lower, upper, reward, np.array([4, 5], dtype=np.int32), "C1")


def config2(S: int = 1000, lower_prob: float = 0.0, cap_stream: int = 0) -> Instance:
    """C2: n=50 (1+8+4x8+8+1), m=200, S=1000."""
    return make_layered([8, 8, 8, 8, 8, 8], 200, S, SEED0 + 2, 0.6, lower_prob, "C2", cap_stream=cap_stream)


def config4(S: int = 10000, lower_prob: float = 0.0, cap_stream: int = 0) -> Instance:
    """C4/C5 network: n=200 (1+6x33+1), m=1000."""
    return make_layered([33, 33, 33, 33, 33, 33], 1000, S, SEED0 + 4, 0.75, lower_prob, "C4", cap_stream=cap_stream)


def random_paths(net, K: int, seed: int, unmatched_prob: float = 0.1) -> np.ndarray:
    """K first-stage candidates in the DD encoding (`DD.h:424`): one int16 per layer, the id of
    the matched out-arc or -1; each out-arc used at most once per V-bar node (the DD state rule,
    `DD.cpp:3666-3668`).  `net` is any object with `layer_arc`, `head`, `out_arcs(q)`."""
    rng = np.random.default_rng(seed)
    L = len(net.layer_arc)
    out = np.full((K, L), -1, dtype=np.int16)
    for k in range(K):
        used = set()
        for ell in range(L):
            q = int(net.head[net.layer_arc[ell]])
            avail = [b for b in net.out_arcs(q) if b not in used]
            if avail and rng.random() >= unmatched_prob:
                b = int(avail[int(rng.integers(len(avail)))])
                used.add(b)
                out[k, ell] = b
    return out


# ---- on-disk formats either side of the path -----------------------------------
def perturbed_paths(net, K: int, seed: int, changes: int = 3, unmatched_prob: float = 0.1) -> np.ndarray:
    """K candidates in which each differs from the one before in at most `changes` layers — the shape the Benders loop
    emits (`NodeExplorer.cpp:949-971`: the argmax path moves a little after every cut) and the one the warm starts of the
    K1 kernel are made for.  Same encoding and state rule as `random_paths`."""
    rng = np.random.default_rng(seed)
    out = np.repeat(random_paths(net, 1, seed, unmatched_prob), K, axis=0)
    L = out.shape[1]
    for k in range(1, K):
        out[k] = out[k - 1]
        for ell in rng.choice(L, size=min(changes, L), replace=False):
            q = int(net.head[net.layer_arc[ell]])
            used = {int(out[k, l2]) for l2 in range(L) if l2 != ell and out[k, l2] >= 0 and int(net.head[net.layer_arc[l2]]) == q}
            avail = [b for b in net.out_arcs(q) if b not in used]
            out[k, ell] = -1 if (not avail or rng.random() < unmatched_prob) else int(avail[rng.integers(len(avail))])
    return out


def read_text(path: str, name: str = "file") -> Instance:
    """Parser of the reference's instance format (`Network::Network`, Network.cpp:18-63):
    `n m S`, then per arc `tail head (lb ub reward) x S`, one separator token, then V-bar ids."""
    with open(path) as f:
        tok = f.read().split()
    n, m, S = int(tok[0]), int(tok[1]), int(tok[2])
    body = np.array(tok[3:3 + m * (2 + 3 * S)], dtype=np.int64).reshape(m, 2 + 3 * S)
    tail, head = body[:, 0].astype(np.int32), body[:, 1].astype(np.int32)
    trip = body[:, 2:].reshape(m, S, 3)
    rest = tok[3 + m * (2 + 3 * S) + 1:]           # one token is skipped (`file >> temp`, Network.cpp:53)
    vbar = np.array([int(x) for x in rest], dtype=np.int32)
    return Instance(n, m, S, tail, head, np.ascontiguousarray(trip[:, :, 0], dtype=np.int32),
                    np.ascontiguousarray(trip[:, :, 1], dtype=np.int32), np.ascontiguousarray(trip[:, :, 2], dtype=np.int32), vbar, name)


def save_cache(inst: Instance, path: str) -> None:
    """The scenario-major binary cache: header + the fp64 capacity arrays exactly as K1 reads them in HBM
    (`sgufp_cache_write`, csrc/cache.cu).  `GuroSolver.from_cache(path, ...)` streams it — or one rank's block of it — to a device."""
    import ctypes as C
    from . import _lib
    L = _lib.lib()
    arr = lambda a: np.ascontiguousarray(a, dtype=np.int32)
    t, h, u, lo, r0, vb = arr(inst.tail), arr(inst.head), arr(inst.upper), arr(inst.lower), arr(inst.reward[:, 0]), arr(inst.vbar)
    p = lambda a: a.ctypes.data_as(_lib.ip)
    rc = L.sgufp_cache_write(os.fsencode(path), inst.n, inst.m, inst.S, p(t), p(h), p(u), p(lo), p(r0), p(vb), len(vb))
    if rc:
        raise _lib.SgufpError(rc, L.sgufp_cache_last_error().decode())


def cache_dims(path: str):
    import ctypes as C
    from . import _lib
    n, m, S, nv, fb = C.c_int(), C.c_int(), C.c_int(), C.c_int(), C.c_int64()
    rc = _lib.lib().sgufp_cache_dims(os.fsencode(path), C.byref(n), C.byref(m), C.byref(S), C.byref(nv), C.byref(fb))
    if rc:
        raise _lib.SgufpError(rc, _lib.lib().sgufp_cache_last_error().decode())
    return {"n": n.value, "m": m.value, "S": S.value, "nvbar": nv.value, "file_bytes": fb.value}


def save_binary(inst: Instance, path: str) -> None:
    """Scenario-count-proof cache: the text format is O(m*S) tokens (C5: 3e8 integers); this is one
    compressed .npz with the arrays exactly as `NetworkArc` holds them."""
    np.savez_compressed(path, n=inst.n, m=inst.m, S=inst.S, tail=inst.tail, head=inst.head, lower=inst.lower, upper=inst.upper,
                        reward0=inst.reward[:, 0], reward_varies=bool((inst.reward != inst.reward[:, :1]).any()),
                        reward=inst.reward if (inst.reward != inst.reward[:, :1]).any() else np.zeros((0, 0), np.int32),
                        vbar=inst.vbar, name=inst.name)


def load_binary(path: str) -> Instance:
    z = np.load(path if path.endswith(".npz") else path + ".npz", allow_pickle=False)
    S = int(z["S"])
    reward = z["reward"] if bool(z["reward_varies"]) else np.repeat(z["reward0"][:, None], S, axis=1).astype(np.int32)
    return Instance(int(z["n"]), int(z["m"]), S, z["tail"], z["head"], z["lower"], z["upper"], reward, z["vbar"], str(z["name"]))
