"""The C-ABI library: loads, exports every symbol include/*.h declares, validates instances the
way DESIGN.md says, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes as C
import glob
import os
import re

import numpy as np
import pytest

from sgufp_solver_b200 import _lib, instances as I
from sgufp_solver_b200.solver import Cut, GuroSolver, getKey

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        txt = re.sub(r"/\*.*?\*/", "", open(h).read(), flags=re.S)
        names |= set(re.findall(r"\b(sgufp_[a-z0-9_]+)\s*\(", txt))
    return sorted(names)


def test_library_exports_every_declared_symbol(built_lib):
    L = C.CDLL(built_lib)
    syms = declared_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(L, s), s


def test_python_binding_covers_header(built_lib):
    from sgufp_solver_b200 import dd  # noqa: F401  (registers the DD signatures)
    bound = set(_lib.SIGNATURES) | set(getattr(_lib, "DD_SIGNATURES", {}))
    assert set(declared_symbols()) <= bound


def _create(inst, device):
    L = _lib.lib()
    h = C.c_void_p()
    ip = _lib.ip
    u = np.ascontiguousarray(inst.upper, np.int32); lo = np.ascontiguousarray(inst.lower, np.int32)
    r0 = np.ascontiguousarray(inst.reward[:, 0], np.int32); vb = np.ascontiguousarray(inst.vbar, np.int32)
    t = np.ascontiguousarray(inst.tail, np.int32); hd = np.ascontiguousarray(inst.head, np.int32)
    rc = L.sgufp_create(C.byref(h), inst.n, inst.m, inst.S, t.ctypes.data_as(ip), hd.ctypes.data_as(ip), u.ctypes.data_as(ip),
                        lo.ctypes.data_as(ip), r0.ctypes.data_as(ip), vb.ctypes.data_as(ip), len(vb), device, 0, inst.S)
    return rc, h, L.sgufp_last_error(None).decode()


def test_model_only_handle_matches_oracle_model(built_lib):
    from oracle.oracle import OracleNet
    for inst in (I.config1(S=2), I.config2(S=2), I.config4(S=1)):
        gs = GuroSolver(inst, device=-1)
        on = OracleNet(inst)
        assert (gs.L, gs.T) == (on.L, on.T)
        assert gs.vbar.tolist() == on.vbar.tolist()
        assert gs.layer_arc.tolist() == on.layer_arc.tolist()
        assert gs.slot_i[:gs.T].tolist() == on.slot_i.tolist() and gs.slot_j[:gs.T].tolist() == on.slot_j.tolist()
        assert gs.W == 1 + gs.L + inst.m
        # compute without a device must fail loudly
        with pytest.raises(_lib.SgufpError) as e:
            gs.solveSubProblem(np.full(gs.L, -1, np.int16))
        assert e.value.code == -6


@pytest.mark.skipif(os.path.exists("/dev/nvidia0"), reason="CPU-only check")
def test_create_on_a_device_fails_without_gpu(built_lib):
    rc, h, msg = _create(I.config1(S=2), 0)
    assert rc == -6 and "cuda" in msg.lower()


def test_instance_validation(built_lib):
    base = I.config1(S=2)
    import dataclasses
    # source->sink arc: the reference gives it no row (Network.cpp:80)
    bad = dataclasses.replace(base, m=base.m + 1, tail=np.append(base.tail, 0).astype(np.int32), head=np.append(base.head, 9).astype(np.int32),
                              upper=np.vstack([base.upper, base.upper[:1]]), lower=np.vstack([base.lower, base.lower[:1]]),
                              reward=np.vstack([base.reward, base.reward[:1]]))
    rc, _, msg = _create(bad, -1)
    assert rc == -4 and "A4" in msg
    # parallel arcs alias in the node-pair keyed variables (grb.cpp:148)
    bad = dataclasses.replace(base, m=base.m + 1, tail=np.append(base.tail, 1).astype(np.int32), head=np.append(base.head, 4).astype(np.int32),
                              upper=np.vstack([base.upper, base.upper[:1]]), lower=np.vstack([base.lower, base.lower[:1]]),
                              reward=np.vstack([base.reward, base.reward[:1]]))
    assert _create(bad, -1)[0] == -4
    # a cycle
    bad = dataclasses.replace(base, m=base.m + 1, tail=np.append(base.tail, 6).astype(np.int32), head=np.append(base.head, 1).astype(np.int32),
                              upper=np.vstack([base.upper, base.upper[:1]]), lower=np.vstack([base.lower, base.lower[:1]]),
                              reward=np.vstack([base.reward, base.reward[:1]]))
    assert _create(bad, -1)[0] == -3
    # V-bar node that no demand point descends from: shuffleVBarNodes would spin (Network.cpp:159-184)
    iso = dataclasses.replace(base, n=base.n + 2, m=base.m + 2,
                              tail=np.concatenate([base.tail, [0, 10]]).astype(np.int32), head=np.concatenate([base.head, [10, 11]]).astype(np.int32),
                              upper=np.vstack([base.upper, base.upper[:2]]), lower=np.vstack([base.lower, base.lower[:2]]),
                              reward=np.vstack([base.reward, base.reward[:2]]), vbar=np.array([4, 5, 10], np.int32))
    rc, _, msg = _create(iso, -1)
    assert rc == -4 and "shuffleVBarNodes" in msg


def test_cut_object_mirrors_inavap_cut(built_lib):
    """tests2.cpp:209-231 through the product's Cut mirror; hash as Cut.h:243-251."""
    from oracle import ref_dd
    coeff = {(1, 18, 123): 1234.56, (2, 30, 123): 432.0, (2, 30, 124): 456.67, (4, 1, 90): -1298.98, (5, 6, 7): -1298.98}
    keys = [getKey(q, i, j) for (i, q, j) in coeff]
    cut = Cut(3012.0321, keys, list(coeff.values()))
    assert cut.get(getKey(30, 2, 123)) == 432.0
    assert cut.get(getKey(30, 2, 124)) == 456.67
    assert cut.get(getKey(6, 5, 7)) == -1298.98
    assert cut.get(getKey(6, 5, 1)) == 0.0
    assert cut.get(getKey(30, 2, 123) | (7 << 48)) == 432.0       # offset bits are masked (Cut.h:278)
    if ref_dd.available():
        rk, rv, rh = ref_dd.cut_to_cut(1, 3012.0321, coeff)
        assert rk.tolist() == cut.keys.tolist() and rv.tolist() == cut.vals.tolist()
        assert rh == cut.hash_val
        assert getKey(4, 0, 15) == ref_dd.lib().ref_get_key(4, 0, 15)
