"""Developer tool (CPU, Oracle B): how often do scenarios of one candidate share their SPEC-LP dual?

For every candidate path: the number of DISTINCT potential vectors (alpha of grb.cpp's dual, all n nodes) and distinct
saturation patterns (which arcs carry a positive capacity multiplier) over the S scenarios, and how far the potentials
of consecutive scenarios are apart.  Decides whether a guess-and-verify / warm-start fast path can exist (VERDICT r01,
"Next round" 1a).      python tools/k1_dual_census.py c2 1000 4   |   c4 300 3
"""
import collections
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.oracle import OracleNet  # noqa: E402
from sgufp_solver_b200 import instances  # noqa: E402

cfg, S, K = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
inst = instances.config2(S) if cfg == "c2" else instances.config4(S)
net = OracleNet(inst)
paths = instances.random_paths(net, K, 7)
for k in range(K):
    duals = [net.scenario_duals(paths[k], s) for s in range(S)]
    A = np.stack([d["alpha"] for d in duals])
    pot = collections.Counter(a.tobytes() for a in A)
    sat = collections.Counter((d["gamma"] > 0).tobytes() + (d["sigma"] > 0).tobytes() + (d["phi"] > 0).tobytes() for d in duals)
    ham = (A[1:] != A[:-1]).sum(axis=1)
    near = [(A[max(0, s - 8):s] != A[s]).sum(axis=1).min() for s in range(1, S)]
    varying = int((A != A[0]).any(axis=0).sum())
    print(f"{cfg} candidate {k}: S={S}  distinct potential vectors {len(pot)} (largest class {max(pot.values())})  "
          f"distinct saturation patterns {len(sat)}  nodes whose potential varies {varying}/{A.shape[1]}  "
          f"nodes that differ between consecutive scenarios: mean {ham.mean():.1f}, nearest of the previous 8: {np.mean(near):.1f}  "
          f"potential range [{A.min()}, {A.max()}]", flush=True)
