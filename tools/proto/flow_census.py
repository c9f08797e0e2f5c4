"""Developer tool (CPU): census of the flow phase of K1 on the bench's candidate paths, taken on the kernel body compiled
for the host (tests/cpp/k1_emul.cpp, one lane per scenario; the counters live under SGUFP_K1_EMULATE in k1_cut.cu).
Per evaluation: searches of the repair (successful / failed), searches to the sink, chunk visits of the list search and
how many of them were made after the target was already in the reached set, push hops, dual updates, list lengths.

    python tools/proto/flow_census.py [config4|config2] [S] [group] [order: 1 = nearest-neighbour chain (as built), 0 = as emitted]
"""
import ctypes as C
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from sgufp_solver_b200 import instances as I  # noqa: E402

NAMES = ["repair searches ok", "repair searches failed", "sink searches ok", "sink searches failed", "chunk visits",
         "re-test iterations", "chunk visits after the target was reached", "sweeps", "push hops", "pushes", "dual updates",
         "tight lists built", "tight list entries", "open chains", "evaluations", "contracted nodes"]


def build(defines=("SGUFP_K1_SKIP_CONFIRM", "SGUFP_K1_PUSH_PAR")):
    so = os.path.join(tempfile.mkdtemp(), "libk1_emul.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-w", "-fPIC", "-shared", f"-I{ROOT}/tests/cpp/emul_stub", f"-I{ROOT}/include"] +
                          [f"-D{d}" for d in defines] + [f"{ROOT}/tests/cpp/k1_emul.cpp", "-o", so])
    return C.CDLL(so)


def main():
    net_name = sys.argv[1] if len(sys.argv) > 1 else "config4"
    S = int(sys.argv[2]) if len(sys.argv) > 2 else 24
    group = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    K = 8 if net_name == "config4" else 64
    inst = getattr(I, net_name)(S=S)
    paths = np.ascontiguousarray(np.load(os.path.join(ROOT, "sgufp_solver_b200", "data", "bench_candidates.npz"))[net_name], dtype=np.int16)
    L = build()
    from test_k1_emulated_cpu import run_emul
    from oracle.oracle import OracleNet
    net = OracleNet(inst)
    L.emul_set_group(group)
    L.emul_set_order(int(sys.argv[4]) if len(sys.argv) > 4 else 1)
    L.emul_state(1)
    out = (C.c_longlong * 16)()
    batches = [paths[b * K:(b + 1) * K] for b in range(max(1, len(paths) // K))]
    for step in range(4):           # the first step starts from zero flow; the others are the bench's steady state
        run_emul(L, inst, net, batches[step % len(batches)])
        L.emul_counts16(out)
        c = list(out)
        ev = max(1, c[14])
        print(f"step {step} (batch {step % len(batches)}): {c[14]} evaluations, per evaluation:")
        for i, n in enumerate(NAMES):
            if i != 14:
                print(f"    {n:45s} {c[i] / ev:9.2f}")
    L.emul_state(0)


if __name__ == "__main__":
    main()
