"""Regenerates tests/golden/k1_*.json from Oracle B (oracle/sgufp_oracle.c).
The LP half has no runnable reference (Gurobi is absent): these vectors pin the
SPEC-LP cut so that neither the oracle nor the kernels drift silently.  Run from the repo root:
    python tests/golden/make_golden_k1.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.oracle import OracleNet  # noqa: E402
from sgufp_solver_b200 import instances as I  # noqa: E402

CASES = {
    "k1_c1": (lambda: I.config1(S=50), 6, 101, 0.15),
    "k1_c1_lb": (lambda: I.config1(S=50, lower_prob=0.1), 6, 102, 0.3),
    "k1_c2_small": (lambda: I.config2(S=40), 4, 103, 0.1),
    "k1_c2_small_lb": (lambda: I.config2(S=40, lower_prob=0.08), 4, 104, 0.3),
}


def main():
    for name, (make, K, seed, unm) in CASES.items():
        inst = make()
        net = OracleNet(inst)
        paths = I.random_paths(net, K, seed, unm)
        out = {"instance": name, "S": inst.S, "K": K, "seed": seed, "unmatched_prob": unm, "cuts": []}
        for p in paths:
            c = net.solve_path(p)
            out["cuts"].append({
                "path": p.tolist(), "cut_type": int(c.cut_type), "first_infeasible": int(c.first_infeasible),
                "rhs": c.rhs, "isum": c.isum.tolist(), "keys": [int(k) for k in c.keys], "vals": c.vals.tolist(),
                "obj": c.obj.tolist() if c.cut_type == 0 else None,
                "status": c.status.tolist(),
            })
        with open(os.path.join(ROOT, "tests", "golden", name + ".json"), "w") as f:
            json.dump(out, f)
        print(name, [(c["cut_type"], c["first_infeasible"]) for c in out["cuts"]])


if __name__ == "__main__":
    main()
