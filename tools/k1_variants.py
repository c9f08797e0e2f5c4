"""Developer tool: build K1 variants with extra nvcc flags and time them side by side.

    python tools/k1_variants.py --build mb3 "-DSGUFP_K1_MINBLOCKS=3"     # here: sgufp_solver_b200/libsgufp_b200_mb3.so
    python tools/k1_variants.py --time base mb3 -- c2 c4                   # on a GPU box ("base" = the shipped library)

Timing is the library's own CUDA-event kernel time (sgufp_last_kernel_ms), best of 5 after 2 warm-ups.
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sgufp_solver_b200 import build as B  # noqa: E402


def lib_of(name):
    return B.LIB if name == "base" else os.path.join(B.HERE, f"libsgufp_b200_{name}.so")


def time_one(name, workloads):
    from sgufp_solver_b200 import _lib
    _lib.LIB_PATH = lib_of(name)
    import numpy as np
    from sgufp_solver_b200 import instances as I
    from sgufp_solver_b200.solver import GuroSolver
    for wl in workloads:
        inst, K = {"c2": (lambda: (I.config2(S=1000), 64)), "c4": (lambda: (I.config4(S=10000), 8))}[wl]()
        solver = GuroSolver(inst)
        paths = np.asarray(I.random_paths(solver, K, 31, 0.1), dtype=np.int16)
        best = 1e9
        for it in range(7):
            solver.solve_paths(paths, want_obj=False, want_status=False, want_dense=False)
            if it >= 2:
                best = min(best, solver.last_kernel_ms())
        print(f"{name:10s} {wl}: {best:.3f} ms  ({K * inst.S / best / 1e3:.2f} M evals/s)", flush=True)


def main():
    a = sys.argv[1:]
    if a and a[0] == "--build":
        cmd = ["/usr/local/cuda/bin/nvcc"] + B.NVCC_FLAGS + a[2].split() + ["-o", lib_of(a[1])] + B.sources()
        subprocess.check_call(cmd)
    elif a and a[0] == "--time":
        names, wl = a[1:a.index("--")], a[a.index("--") + 1:]
        for n in names:
            subprocess.call([sys.executable, __file__, "--one", n] + wl)
    elif a and a[0] == "--one":
        time_one(a[1], a[2:])


if __name__ == "__main__":
    main()
