"""Developer tool (CPU): a longer differential campaign of the K1 kernel body (tests/cpp/k1_emul.cpp) against Oracle B than the
test suite runs — random small and larger networks, with and without lower bounds, runs of perturbed paths (warm starts), one or two
paths per call with the state kept between calls, random run lengths.  Statuses, objectives, exact integer sums, first infeasible
scenario and rays must be identical.

    python tools/fuzz_k1_emul.py [seconds] [seed]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_k1_emulated_cpu as T  # noqa: E402
from oracle.oracle import OracleNet  # noqa: E402
from sgufp_solver_b200 import instances as I  # noqa: E402


class _Tmp:
    def mktemp(self, name):
        import pathlib
        import tempfile
        return pathlib.Path(tempfile.mkdtemp(prefix=name))


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 600.0
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 20261019
    L = T._build_emul(_Tmp(), ["SGUFP_K1_SKIP_CONFIRM", "SGUFP_K1_PUSH_PAR"])
    rng = np.random.default_rng(seed)
    t0 = time.time()
    n_inst = n_eval = n_warm = n_infeas = big = 0
    k = 0
    while time.time() - t0 < budget:
        k += 1
        try:
            inst = T._random_larger_instance(rng, 50000 + k) if k % 3 == 0 else T._random_instance(rng, 50000 + k)
            net = OracleNet(inst)
        except Exception:
            continue
        K = int(rng.integers(2, 7))
        paths = I.perturbed_paths(net, K, k, int(rng.integers(1, 6)), float(rng.choice([0.0, 0.1, 0.3, 0.6])))
        mode = int(rng.integers(0, 3))
        T.warm_counts(L)
        if mode == 0:                                    # one launch, runs of a random length
            L.emul_set_group(int(rng.integers(0, K + 1)))
            try:
                out = T.run_emul(L, inst, net, paths)
            finally:
                L.emul_set_group(0)
            T.check_against_oracle(inst, net, paths, out)
            n_infeas += int((out[1] != T.I64_MAX).sum())
        else:                                            # one or two paths per call, state between the calls
            L.emul_state(1)
            try:
                for k0 in range(0, K, mode):
                    sub = np.ascontiguousarray(paths[k0:k0 + mode])
                    out = T.run_emul(L, inst, net, sub)
                    T.check_against_oracle(inst, net, sub, out)
                    n_infeas += int((out[1] != T.I64_MAX).sum())
            finally:
                L.emul_state(0)
        taken, given_up = T.warm_counts(L)
        assert given_up == 0, inst.name
        n_warm += taken
        n_inst += 1
        n_eval += K * inst.S
        big += L.emul_last_nc() > 31
    print(f"seed {seed}: {n_inst} instances ({big} with more than 31 contracted nodes), {n_eval} evaluations, {n_warm} warm-started, "
          f"{n_infeas} candidates with an infeasible scenario, 0 mismatches, {time.time() - t0:.0f} s")


if __name__ == "__main__":
    main()
