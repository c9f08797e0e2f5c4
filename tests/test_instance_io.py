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


def test_cache_file_holds_the_device_layout(tmp_path, built_lib):
    """The cache (csrc/cache.cu) on the CPU: the file is header + the scenario-major fp64 arrays K1 reads, and a model-only
    handle (no device) made from it knows the same network."""
    import ctypes as C
    import struct
    import sgufp_solver_b200 as sg
    inst = I.make_layered([3, 4, 3], 21, 9, 77, 0.8, 0.2, "odd")       # odd m: a zero pad column
    path = str(tmp_path / "odd.sgufpc")
    I.save_cache(inst, path)
    d = I.cache_dims(path)
    assert (d["n"], d["m"], d["S"], d["nvbar"]) == (inst.n, inst.m, inst.S, len(inst.vbar))
    raw = open(path, "rb").read()
    assert len(raw) == d["file_bytes"] and raw[:8] == b"SGUFPC01"
    version, n, m, S, m_pad, nvbar, max_cap, max_lower = struct.unpack_from("<8i", raw, 8)
    offs = struct.unpack_from("<7q", raw, 40)
    assert (version, n, m, S, m_pad) == (1, inst.n, inst.m, inst.S, (inst.m + 1) & ~1)
    assert max_cap == int(max(inst.upper.max(), inst.lower.max())) and max_lower == int(inst.lower.max())
    assert (np.frombuffer(raw, np.int32, m, offs[0]) == inst.tail).all() and (np.frombuffer(raw, np.int32, m, offs[1]) == inst.head).all()
    assert (np.frombuffer(raw, np.int32, m, offs[2]) == inst.reward[:, 0]).all()
    u = np.frombuffer(raw, np.float64, S * m_pad, offs[4]).reshape(S, m_pad)
    lo = np.frombuffer(raw, np.float64, S * m_pad, offs[5]).reshape(S, m_pad)
    assert (u[:, :m] == inst.upper.T).all() and (lo[:, :m] == inst.lower.T).all() and (u[:, m:] == 0).all() and (lo[:, m:] == 0).all()
    a = sg.GuroSolver(inst, device=-1)
    b = sg.GuroSolver.from_cache(path, device=-1)
    assert (b.n, b.m, b.S, b.L, b.T) == (a.n, a.m, a.S, a.L, a.T)
    assert b.layer_arc.tolist() == a.layer_arc.tolist() and b.vbar.tolist() == a.vbar.tolist() and b.tail.tolist() == a.tail.tolist()
    blk = sg.GuroSolver.from_cache(path, device=-1, scenario_offset=2, S_local=4)
    assert (blk.S, blk.scenario_offset, blk.S_total) == (4, 2, inst.S)
    with pytest.raises(sg.solver.SgufpError):
        sg.GuroSolver.from_cache(path, device=-1, scenario_offset=7, S_local=5)
    with pytest.raises(sg.solver.SgufpError):
        sg.GuroSolver.from_cache(str(tmp_path / "missing.sgufpc"), device=-1)
    c = a.clone()                                   # a model-only handle clones too (nothing to share)
    assert c.L == a.L
    # a damaged header (an offset that points outside the layout of its own sizes, a size that no longer fits the file) is refused,
    # not followed
    for where, value in ((40, 1 << 40), (40 + 4 * 8, 8), (16, inst.m + 2), (20, inst.S * 1000)):      # off_tail, off_u, m, S
        bad = bytearray(raw)
        struct.pack_into("<q" if where >= 40 else "<i", bad, where, value)
        pbad = str(tmp_path / f"bad{where}.sgufpc")
        open(pbad, "wb").write(bytes(bad))
        with pytest.raises(sg.solver.SgufpError):
            sg.GuroSolver.from_cache(pbad, device=-1)
    open(str(tmp_path / "short.sgufpc"), "wb").write(raw[: len(raw) // 2])
    with pytest.raises(sg.solver.SgufpError):
        sg.GuroSolver.from_cache(str(tmp_path / "short.sgufpc"), device=-1)
