"""include/sgufp_b200.hpp + sgufp_b200_explorer.hpp compiled against the reference's own headers and objects
(oracle/Makefile: oracle/_ref/adapter_check, built where /root/reference exists; the binary travels to the GPU box).

CPU: the adapters build with the reference's types and agree with the reference's DD on structure (SGUFP_DEVICE_NONE), and
compute fails loudly without a device.
GPU: the C++ `solveSubProblem` gives the cuts of the ctypes path (keys, values, hash_val bit for bit); the same cuts applied
to the reference's RelaxedDDNew (Oracle A, linked into the binary) and to the adapter's give identical bounds and argmax
paths; the C++ Benders loop (one node at a time and several nodes side by side) reaches the optimum of the Python loop."""
import os
import subprocess

import numpy as np
import pytest

from sgufp_solver_b200 import instances as I

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
EXE = os.path.join(ROOT, "oracle", "_ref", "adapter_check")


def _exe(built_lib=None):
    if os.path.isdir(REF):
        subprocess.check_call(["make", "-s", "-f", os.path.join(ROOT, "oracle", "Makefile"), "adapter_check"])
    if not os.path.exists(EXE):
        pytest.skip("oracle/_ref/adapter_check is not built (no /root/reference here and no prebuilt binary)")
    return EXE


def test_cpp_adapter_builds_and_agrees(built_lib, tmp_path):
    exe = _exe()
    for inst in (I.config1(S=2), I.config2(S=1)):
        f = str(tmp_path / f"{inst.name}.txt")
        inst.write_text(f)
        out = subprocess.run([exe, f], capture_output=True, text=True)
        assert out.returncode == 0, out.stdout + out.stderr
        assert "adapter ok" in out.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("make,device_count", [(lambda: I.config1(S=50), 1), (lambda: I.config1(S=50, lower_prob=0.1), 1),
                                               (lambda: I.make_layered([4, 5, 5, 4], 48, 12, 123, 0.7, 0.0, "mid"), 1)],
                         ids=["c1", "c1_lb", "mid"])
def test_cpp_compute_path_on_the_gpu(tmp_path, make, device_count):
    import sgufp_solver_b200 as sg
    from sgufp_solver_b200.explorer import solve
    exe = _exe()
    inst = make()
    f = str(tmp_path / "inst.txt")
    inst.write_text(f)
    out = subprocess.run([exe, f, "gpu", "0", str(device_count), "4"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "adapter gpu ok" in out.stdout
    gs = sg.GuroSolver(inst)
    cuts = [l for l in out.stdout.splitlines() if l.startswith("CUT ")]
    assert cuts or "mid" in inst.name, out.stdout              # the by-hand loop needs an exact root diagram ("mid" has a collapsed layer)
    for line in cuts:
        head, coeffs = line.split(" |")
        tok = head.split()
        ctype, rhs, hashv, nnz = int(tok[2]), float(tok[3]), int(tok[4]), int(tok[5])
        path = np.array([int(x) for x in tok[6:]], np.int16)
        pairs = [c.split(":") for c in coeffs.split()]
        t, cut = gs.solveSubProblem(path)                       # the ctypes path on the same candidate
        assert t == ctype and cut.RHS == rhs and len(cut.keys) == nnz == len(pairs)
        assert [int(k) for k in cut.keys] == [int(k) for k, _ in pairs]
        assert [float(v) for v in cut.vals] == [float(v) for _, v in pairs]           # printed with 17 digits: bit-exact
        assert cut.hash_val == hashv                             # Inavap::Cut's own hash (Cut.h:243-251) of the C++ object
    best, nodes, ncuts = solve(sg.GuroSolver(inst), max_nodes=400)
    ex = [l for l in out.stdout.splitlines() if l.startswith("EXPLORER")]
    assert len(ex) == 2
    o1 = float(ex[0].split("optimum ")[1].split()[0]); n1 = int(ex[0].split("nodes ")[1].split()[0]); c1 = int(ex[0].split("cuts ")[1].split()[0])
    ow = float(ex[1].split("optimum ")[1].split()[0]); kw = int(ex[1].split("k1_calls ")[1].split()[0]); cw = int(ex[1].split("cuts ")[1].split()[0])
    nw = int(ex[1].split("nodes ")[1].split()[0])
    assert o1 == best and (n1, c1) == (nodes, ncuts)             # one node at a time: the Python loop's search, node for node and cut for cut
    if nodes < 400 and nw < 400:                                 # both searches ran to the end (not cut off by the node budget):
        assert ow == best                                        # the same optimum whatever the width
    assert kw <= cw                                              # side by side: several candidates per K1 call
