"""Instance formats either side of the path: the reference text format written and
read back, agreement with the reference's own parser, and the binary cache."""
import os

import numpy as np
import pytest

from oracle import ref_dd
from sgufp_solver_b200 import instances as I


def _same(a, b):
    assert (a.n, a.m, a.S) == (b.n, b.m, b.S)
    for f in ("tail", "head", "lower", "upper", "reward", "vbar"):
        assert np.array_equal(getattr(a, f), getattr(b, f)), f


@pytest.mark.parametrize("make", [lambda: I.config1(S=7, lower_prob=0.2), lambda: I.config2(S=3)], ids=["c1", "c2"])
def test_text_and_binary_round_trip(make, tmp_path):
    inst = make()
    p = str(tmp_path / "inst.txt")
    inst.write_text(p)
    _same(inst, I.read_text(p))
    b = str(tmp_path / "inst.npz")
    I.save_binary(inst, b)
    _same(inst, I.load_binary(b))
    if inst.m * inst.S > 2000:          # tiny files are dominated by the zip directory
        assert os.path.getsize(b) < os.path.getsize(p)


@pytest.mark.skipif(not ref_dd.available() and not os.path.isdir("/root/reference"), reason="oracle/_ref not available")
def test_reference_parser_reads_what_we_write(ref_available):
    inst = I.config2(S=2)
    rn = ref_dd.RefNetwork(inst)               # goes through Instance.write_text + Network::Network(file)
    assert (rn.n, rn.m) == (inst.n, inst.m)
    assert sorted(rn.vbar.tolist()) == sorted(inst.vbar.tolist())
