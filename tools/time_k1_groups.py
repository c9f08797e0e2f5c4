"""Developer tool: K1 kernel time on the bench workloads (DD-emitted candidates) for several run lengths of the warm start
(SGUFP_K1_GROUP; 1 = every candidate from zero flow).

    python tools/time_k1_groups.py c2 c4 c5 -- 1 2 4 8 16 0        # 0 = the library's own choice
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    a = sys.argv[1:]
    wls, groups = a[:a.index("--")], a[a.index("--") + 1:]
    import bench
    from sgufp_solver_b200.solver import GuroSolver
    for wl in wls:
        w = bench.WORKLOADS[wl]
        S = bench.totals(wl, 1)
        inst = bench.scenario_range(wl, 0, S)
        paths, _ = bench.candidate_paths(wl, w["K"], 0)
        solver = GuroSolver(inst)
        ref = None
        for g in groups:
            if g == "0":
                os.environ.pop("SGUFP_K1_GROUP", None)
            else:
                os.environ["SGUFP_K1_GROUP"] = g
            best = 1e9
            for it in range(5):
                r = solver.solve_paths(paths, want_obj=False, want_status=False, want_dense=True)
                if it >= 1:
                    best = min(best, solver.last_kernel_ms())
            same = "" if ref is None else ("  same cuts" if (r.rhs == ref.rhs).all() and (r.coef_dense == ref.coef_dense).all() else "  CUTS DIFFER")
            ref = ref or r
            print(f"{wl} S={S} K={w['K']} group={g:>2s}: {best:9.3f} ms  {w['K'] * S / best / 1e3:8.2f} M evals/s{same}", flush=True)
        solver.close()


if __name__ == "__main__":
    main()
