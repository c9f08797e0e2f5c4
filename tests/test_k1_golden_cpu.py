"""Oracle B against the committed golden vectors (tests/golden/k1_*.json)."""
import glob
import json
import os

import numpy as np
import pytest

from oracle.oracle import OracleNet
from sgufp_solver_b200 import instances as I

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "k1_*.json")))
MAKERS = {"k1_c1": lambda: I.config1(S=50), "k1_c1_lb": lambda: I.config1(S=50, lower_prob=0.1),
          "k1_c2_small": lambda: I.config2(S=40), "k1_c2_small_lb": lambda: I.config2(S=40, lower_prob=0.08)}


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-5] for p in GOLD])
def test_oracle_reproduces_golden(path):
    g = json.load(open(path))
    inst = MAKERS[g["instance"]]()
    net = OracleNet(inst)
    for c in g["cuts"]:
        oc = net.solve_path(np.array(c["path"], np.int16))
        assert oc.cut_type == c["cut_type"] and oc.first_infeasible == c["first_infeasible"]
        assert oc.rhs == c["rhs"] and oc.isum.tolist() == c["isum"]
        assert [int(k) for k in oc.keys] == c["keys"] and oc.vals.tolist() == c["vals"]
        assert oc.status.tolist() == c["status"]
        if c["obj"] is not None:
            assert oc.obj.tolist() == c["obj"]
