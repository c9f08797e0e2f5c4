#!/usr/bin/env python
"""bench_dd.py — DD arcs/s of K2 (the second half of BASELINE.json's metric) on one B200.

Config C3 (BASELINE.json configs[2]): C2 network, RestrictedDDNew width sweep 64..4096 and the
RelaxedDDNew, B diagrams x C cuts per launch pair.  Prints one JSON line per point.
  value      arcs touched per second = B*C*(in-arcs + terminal arcs) / CUDA-event time of
             k2_longest_path + k2_terminal (events on the launching stream)
  roofline   algorithmic bytes 16 B/arc + 8 B/node per (diagram, cut) + 8*T B per cut
             over the same time, against MEASURED_PEAKS.json's HBM copy bandwidth
  cpu        Oracle A (the unmodified reference class compiled in oracle/_ref) applying the same
             cuts to one diagram on one host core, if the library is present
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--widths", default="64,128,256,512,1024,2048,4096")
    ap.add_argument("--B", type=int, default=64)
    ap.add_argument("--C", type=int, default=64)
    ap.add_argument("--reps", type=int, default=7)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--relaxed-thresholds", default="120",
                    help="collapse thresholds of the RelaxedDDNew points (the reference's compile-time constant is 120)")
    ap.add_argument("--build", action="store_true",
                    help="construction time of one diagram: k2_build on the device (SURVEY 8f-3) vs the host builder vs the reference class")
    ap.add_argument("--sequential", action="store_true",
                    help="one diagram, one cut at a time + getSolution after each (the Benders inner loop, NodeExplorer.cpp:949-971): "
                         "wall time per iteration through the host API, device-side cut application (SURVEY 8f-2)")
    args = ap.parse_args()
    import torch
    from sgufp_solver_b200 import instances as I
    from sgufp_solver_b200.dd import RelaxedDDNew, RestrictedDDNew, apply_optimality_batch, random_cut
    from sgufp_solver_b200.solver import GuroSolver
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    inst = I.config2(S=1)
    solver = GuroSolver(inst, device=0)
    rng = np.random.default_rng(5)
    cuts = [random_cut(solver, rng) for _ in range(args.C)]
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    ref_net = None
    if not args.no_cpu:
        import bench                                   # the CPU-baseline leg lives in bench.py (the one place that runs oracle/)
        base = bench.DDCpuBaseline(inst)
        ref_net = base if base.available() else None
    points = [("restricted", int(w)) for w in args.widths.split(",") if w] + [("relaxed", int(t)) for t in args.relaxed_thresholds.split(",") if t]
    if args.build:
        for kind, w in points:
            line = {"metric": "dd_build_ms", "unit": "ms per buildTree/compile", "higher_is_better": False, "kind": kind, "width": w}
            for mode in ("device", "host"):
                os.environ["SGUFP_DD_BUILD"] = mode
                d = RestrictedDDNew(solver, w) if kind == "restricted" else RelaxedDDNew(solver, w)
                build = d.compile if kind == "restricted" else d.buildTree
                build(); torch.cuda.synchronize()
                ts = []
                for _ in range(5):
                    t0 = time.perf_counter(); build(); ts.append(time.perf_counter() - t0)
                line["value" if mode == "device" else "host_builder_ms"] = float(np.median(ts)) * 1e3
                if mode == "device":
                    line["built_on_device"] = bool(d.dump_device()["built_on_device"])
                    line["nodes_per_diagram"], line["arcs_per_diagram"] = d.counts()
                d.close()
            os.environ.pop("SGUFP_DD_BUILD", None)
            if ref_net is not None and ref_net.supports(kind, w):
                line["cpu_reference"] = {"value": ref_net.build_ms(kind, w), "unit": line["unit"], "cores": 1, "kind": "reference"}
            print(json.dumps(line), flush=True)
        return
    if args.sequential:
        for kind, w in points:
            d = RestrictedDDNew(solver, w) if kind == "restricted" else RelaxedDDNew(solver, w)
            d.compile() if kind == "restricted" else d.buildTree()
            nodes1, arcs1 = d.counts()
            apply = (lambda c: d.applyOptimalityCut(c)) if kind == "restricted" else (lambda c: d.applyOptimalityCut(c, -1e300, 1e300))
            sol = d.getMaxPath if kind == "restricted" else d.getSolution
            for c in cuts[:3]:
                apply(c); sol()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for c in cuts:
                apply(c); sol()
            dt = time.perf_counter() - t0
            line = {"metric": "dd_sequential_cut_ms", "value": dt / len(cuts) * 1e3, "unit": "ms per (applyOptimalityCut + getSolution)",
                    "higher_is_better": False, "kind": kind, "width": w, "arcs_per_diagram": arcs1, "nodes_per_diagram": nodes1,
                    "arcs_per_sec": arcs1 * len(cuts) / dt, "cuts": len(cuts),
                    "note": "host API wall time; longest path, terminal weights, pruning and path extraction run on the device, "
                            "only the bound and the path come back"}
            if ref_net is not None and ref_net.supports(kind, w):
                dtr = ref_net.apply_seconds(kind, w, cuts, with_solution=True)
                line["cpu_reference"] = {"value": dtr / len(cuts) * 1e3, "unit": line["unit"], "cores": 1, "kind": "reference",
                                         "sample": f"the same {len(cuts)} cuts on the unmodified reference class (oracle/_ref)"}
            print(json.dumps(line), flush=True)
            d.close()
        return
    for kind, w in points:
        dds = []
        for _ in range(args.B):
            d = RestrictedDDNew(solver, w) if kind == "restricted" else RelaxedDDNew(solver, w)
            d.compile() if kind == "restricted" else d.buildTree()
            dds.append(d)
        for _ in range(3):
            apply_optimality_batch(dds, cuts)
        times = []
        for _ in range(args.reps):
            flush.zero_(); torch.cuda.synchronize()
            apply_optimality_batch(dds, cuts)
            ms, arcs, launches = dds[0].last_stats()
            times.append(ms)
        ms = float(np.median(times))
        nodes1, arcs1 = dds[0].counts()
        bytes_alg = args.B * args.C * (arcs1 * 16 + nodes1 * 8) + args.C * solver.T * 8
        line = {"metric": "dd_arcs_per_sec", "value": arcs / (ms / 1e3), "unit": "arcs/s", "kind": kind, "width": w, "kernel_ms": ms,
                "diagrams": args.B, "cuts": args.C, "arcs_per_diagram": arcs1, "nodes_per_diagram": nodes1, "gpu_launches": launches,
                "roofline": {"bound": "hbm", "achieved": bytes_alg / (ms / 1e3) / 1e9, "peak": peak, "unit": "GB/s",
                             "frac": bytes_alg / (ms / 1e3) / 1e9 / peak, "bytes_per_launch": bytes_alg},
                "l2": "flushed (512 MiB write) before every timed launch"}
        if ref_net is not None and ref_net.supports(kind, w):
            dt = ref_net.apply_seconds(kind, w, cuts)
            line["cpu_reference"] = {"value": arcs1 * args.C / dt, "unit": "arcs/s", "cores": 1, "kind": "reference",
                                     "sample": f"1 diagram x {args.C} cuts, unmodified reference class (oracle/_ref)"}
        print(json.dumps(line), flush=True)
        for d in dds:
            d.close()


if __name__ == "__main__":
    main()
