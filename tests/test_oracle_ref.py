"""Oracle A: the unmodified reference (oracle/_ref) against the reference's own known answers,
and Oracle B's instance model against the reference's `Network`."""
import numpy as np
import pytest

from oracle import ref_dd
from oracle.oracle import OracleNet
from sgufp_solver_b200 import instances as I

pytestmark = pytest.mark.skipif(not ref_dd.available() and not __import__("os").path.isdir("/root/reference"),
                                reason="oracle/_ref not built and /root/reference absent")


@pytest.fixture(scope="module", autouse=True)
def _built(ref_available):
    assert ref_available


def test_constant_cut_kat():
    """tests2.cpp:209-231 (ConstantCut.SHOULD_RETURN_SAME)."""
    coeff = {(2, 30, 123): 432.0, (2, 30, 124): 456.67, (1, 18, 123): 1234.56, (4, 1, 90): -1298.98, (5, 6, 7): -1298.98}
    keys, vals, h = ref_dd.cut_to_cut(1, 3012.0321, coeff)
    assert ref_dd.cut_get(3012.0321, keys, vals, ref_dd.lib().ref_get_key(30, 2, 123)) == 432.0
    assert ref_dd.cut_get(3012.0321, keys, vals, ref_dd.lib().ref_get_key(30, 2, 124)) == 456.67
    assert ref_dd.cut_get(3012.0321, keys, vals, ref_dd.lib().ref_get_key(6, 5, 7)) == -1298.98
    assert ref_dd.cut_get(3012.0321, keys, vals, ref_dd.lib().ref_get_key(6, 5, 1)) == 0.0


def test_commented_key_kats():
    """Raw 64-bit keys of the commented-out fixture (tests2.cpp:45-51,58-71,87-100)."""
    raw = [(281539401220100, 306), (281603825729543, 2664), (562997198454798, 2790), (42950066190, 2328), (281513632399378, 663),
           (281513632530450, 663), (281513632006162, 663), (563061624012827, 422), (171800133659, 1580), (281509338284057, 462),
           (281509336645657, 231), (563074509373456, 450), (51541508112, 173), (563074507931664, 450), (51540066320, 173),
           (563074509504528, 450), (51541639184, 173), (1125981511352333, 506), (38654836749, 426), (64424640525, 192),
           (47244771341, 3132), (844506536214541, 314), (38656409613, 234), (47246344205, 2940), (844506535297037, 314),
           (38655492109, 234), (47245426701, 2940)]
    keys = np.array([k for k, _ in raw], np.uint64)
    vals = np.array([v for _, v in raw], np.float64)
    gk = ref_dd.lib().ref_get_key
    for (q, i, j), want in {(4, 0, 15): 306, (7, 0, 30): 2664, (13, 2, 9): 426, (16, 31, 29): 450, (14, 6, 10): 2328,
                            (25, 28, 8): 462, (18, 17, 9): 663, (655, 342, 2): 0, (0, 0, 0): 0}.items():
        assert ref_dd.cut_get(-9821, keys, vals, gk(q, i, j)) == want


def test_appendix_a_numbers():
    """the probe instance through the reference's own classes."""
    inst = I.config1(S=2)
    rn = ref_dd.RefNetwork(inst)
    assert rn.total_layers == 6
    assert rn.layer_arc.tolist() == [3, 4, 5, 6, 7, 8]
    assert rn.has_state_changed.tolist() == [1, 0, 0, 1, 0, 0, 0]
    assert rn.state_update == {0: [-1, 9, 10, 11], 3: [-1, 12, 13]}
    assert rn.classes == (3, 3, 11, 0)
    dd = ref_dd.RefRelaxedDD(rn)
    dd.build()
    assert dd.layer_sizes().tolist() == [1, 4, 13, 34, 102, 238, 442, 1]
    assert dd.is_exact()
    keys, vals, v = [], [], 1.5
    out = {4: [9, 10, 11], 5: [12, 13]}
    for a in rn.layer_arc:
        i, q = int(inst.tail[a]), int(inst.head[a])
        for b in out[q]:
            keys.append(q | (i << 16) | (int(inst.head[b]) << 32)); vals.append(v); v += 0.75
    assert dd.apply_opt(-3.25, keys, vals, -1e300, 1e300) == 32.0
    assert dd.solution().tolist() == [9, 10, 11, -1, 12, 13]


@pytest.mark.parametrize("make", [lambda: I.config1(S=2), lambda: I.config2(S=2), lambda: I.config4(S=1)], ids=["c1", "c2", "c4"])
def test_oracle_b_model_equals_reference_network(make):
    """shuffleVBarNodes + processingOrder restated in Oracle B == the reference's (Network.cpp:94-186)."""
    inst = make()
    rn = ref_dd.RefNetwork(inst)
    on = OracleNet(inst)
    assert on.L == rn.total_layers
    assert on.vbar.tolist() == rn.vbar.tolist()
    assert on.layer_arc.tolist() == rn.layer_arc.tolist()
