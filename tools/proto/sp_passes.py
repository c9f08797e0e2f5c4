"""Prototype: relaxation passes of the SPEC-LP label computation (shortest_paths<MERGED>, the cut kernel) on the optimal flows of the
bench's candidates, with the GPU's granularity (32 chains per step, labels read before the step's atomicMin's land):
ascending chain order in every pass (as built) against alternating ascending / descending passes.

    python sp_passes.py c4 8 4      # network, scenarios, candidates
"""
import io, contextlib, os, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__)); sys.path.insert(0, HERE)
with contextlib.redirect_stdout(io.StringIO()):
    from warm_proto import G, graph, cold, I, INF, ROOT
M = 1 << 29


def passes(g, cap, x, order):
    lab = [M] * (g.nc + 1); lab[0] = 0
    n = 0
    chunks = [list(range(c0, min(g.n, c0 + 32))) for c0 in range(0, g.n, 32)]
    while True:
        n += 1
        changed = False
        seq = chunks if (order == "asc" or n % 2 == 1) else chunks[::-1]
        for ch in seq:
            snap = list(lab)
            for c in ch:
                for (t, h, d, cost) in g.arcs(c):
                    if h == g.nc or snap[t] >= M:
                        continue
                    if (cap[c] - x[c] if d == 0 else x[c]) > 0 and snap[t] + cost < lab[h]:
                        lab[h] = snap[t] + cost; changed = True
        if not changed:
            return n, lab


if __name__ == "__main__":
    name, S, K = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    inst = I.config2(S=S) if name == "c2" else I.config4(S=S)
    paths = np.load(os.path.join(ROOT, "sgufp_solver_b200", "data", "bench_candidates.npz"))["config2" if name == "c2" else "config4"]
    tot = {"asc": 0, "alt": 0}; n = 0
    for k in range(K):
        sv, ev, r, nc, ac = graph(inst, paths[k])
        g = G(sv, ev, r, nc, None)
        for s in range(S):
            cap = [INF] * g.n
            for a in range(inst.m):
                if ac[a] >= 0: cap[ac[a]] = min(cap[ac[a]], int(inst.upper[a, s]))
            x, _ = cold(g, cap, dict(searches=0, pushes=0, duals=0))
            pa, la = passes(g, cap, x, "asc"); pb, lb = passes(g, cap, x, "alt")
            assert la == lb
            tot["asc"] += pa; tot["alt"] += pb; n += 1
    print(f"{name}: {n} evaluations, passes per label computation: ascending {tot['asc'] / n:.2f}, alternating {tot['alt'] / n:.2f}")
