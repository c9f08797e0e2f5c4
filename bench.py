#!/usr/bin/env python
"""bench.py — scenario-cut evaluations per second (BASELINE.json metric) on N B200s.

A STEP is one pass of the hot path over one batch of synthetic input: K candidate first-stage
paths x the rank's S scenarios = K*S exact subproblem solves folded into K cuts
(GuroSolver::solveSubProblem, /root/reference/grb.cpp:139-360).

  value   device-timed (CUDA events on the launching stream) whole-job evals/s with the capacity
          arrays resident in HBM; per step: plan upload + K1 kernel (+ the all-reduce for N>1).
  e2e     same metric through the reference-facing call (host paths in, host Inavap::Cut out:
          sgufp_solve_paths / ShardedGuroSolver.solve_paths), wall clock, copies included.
  roofline  K1's algorithmic bytes (2*m*8 + 9 per evaluation, SURVEY.md §8d, fp64 storage) over its
          own CUDA-event duration, against MEASURED_PEAKS.json's HBM copy bandwidth.
  cpu_baseline  Oracle B (oracle/sgufp_oracle.c, the CPU port of the same path — Gurobi is not
          available) on a bounded sample of the same workload.

`--impl reference` times that CPU port with every host thread on the same config.
L2 is flushed (512 MiB write) before every timed step; the flush is outside the timed events.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# The contract is ONE JSON line on stdout.  Libraries chat on file descriptor 1 (NCCL prints its version there
# when NCCL_DEBUG is set in the environment), so fd 1 is pointed at stderr for the whole run and the JSON line
# goes to a private duplicate of the real stdout.
_REAL_STDOUT = None


def isolate_stdout() -> None:
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line: dict) -> None:
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()

WORKLOADS = {
    # name: (config fn name, per-GPU scenarios, candidates per step, description)
    "c2": ("config2", 1000, 64, "C2: n=50 m=200 S=1000/GPU, K=64 candidate paths per step (BASELINE.json configs[1])"),
    "c4": ("config4", 10000, 8, "C4: n=200 m=1000 S=10000/GPU, K=8 candidate paths per step (configs[3] network)"),
    "c5": ("config4", 100000, 1, "C5: n=200 m=1000 S=100000/GPU, K=1 candidate path per step (configs[4], 1.6 GB of capacities)"),
}
METRIC = "scenario_cut_evals_per_sec"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def measured_traffic(workload):
    p = os.path.join(ROOT, "profiles", "k1_traffic.json")
    if os.path.exists(p):
        return json.load(open(p)).get(workload)
    return None


def issue_roofline(workload, kernel_ms, sm_mhz, sms):
    """Second roofline of K1 (the binding one): warp instructions per launch (ncu smsp__inst_executed.sum of
    the seeded workload, committed in profiles/k1_traffic.json) / live kernel time, against 4 issue slots per SM and clock."""
    p = os.path.join(ROOT, "profiles", "k1_traffic.json")
    if not os.path.exists(p) or not sm_mhz:
        return None
    n = json.load(open(p)).get("warp_instructions", {}).get(workload)
    if not n:
        return None
    achieved = n / (kernel_ms * 1e-3) / 1e9
    peak = sms * 4 * sm_mhz * 1e6 / 1e9
    return {"bound": "issue", "achieved": achieved, "peak": peak, "unit": "G warp-inst/s", "frac": achieved / peak,
            "warp_instructions_per_launch": n, "peak_source": "%d SMs x 4 schedulers x %.0f MHz (sampled under load)" % (sms, sm_mhz)}


class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu):
        self.gpu, self.rows, self.proc = gpu, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


def make_inputs(workload, rank, world):
    from sgufp_solver_b200 import instances as I
    fn, S, K, desc = WORKLOADS[workload]
    S_total = S * world                      # weak scaling: per-GPU scenarios fixed
    # the rank's own contiguous block, drawn independently per rank (same network: topology seed is shared)
    inst = getattr(I, fn)(S=S, cap_stream=rank)
    return inst, S, S_total, K, desc


def cpu_port_throughput(inst, paths, seconds_target, threads):
    """Oracle B on a bounded sample: `threads` host threads, each with its own handle, each
    evaluating whole scenario ranges of the sampled candidates (the GIL is released in ctypes)."""
    from oracle.oracle import OracleNet
    nets = [OracleNet(inst) for _ in range(threads)]
    # calibrate on a small slice
    t0 = time.perf_counter()
    probe = min(inst.S, 64)
    nets[0].solve_range(paths[0], 0, probe)
    per_eval = (time.perf_counter() - t0) / probe
    budget_evals = max(threads * 32, int(seconds_target / per_eval) * threads)
    n_paths = max(1, min(len(paths), budget_evals // inst.S))
    S_use = inst.S if budget_evals >= inst.S else max(threads, budget_evals)
    work = [(k, lo, min(S_use, lo + (S_use + threads - 1) // threads)) for k in range(n_paths)
            for lo in range(0, S_use, (S_use + threads - 1) // threads)]
    done = [0] * threads

    def run(t):
        for idx in range(t, len(work), threads):
            k, lo, hi = work[idx]
            nets[t].solve_range(paths[k], lo, hi)
            done[t] += hi - lo
    ts = [threading.Thread(target=run, args=(t,)) for t in range(threads)]
    t0 = time.perf_counter()
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    dt = time.perf_counter() - t0
    evals = sum(done)
    return evals / dt, evals, dt, f"{n_paths} candidate path(s) x {S_use} scenarios of the same workload ({evals} evaluations, {dt:.1f} s)"


def run_reference(args, rank, world):
    """The reference arm: the CPU implementation of the path on the box's host cores."""
    if rank != 0:
        return
    from oracle.oracle import OracleNet
    from sgufp_solver_b200 import instances as I
    inst, S, S_total, K, desc = make_inputs(args.workload, 0, 1)
    net = OracleNet(inst)
    paths = I.random_paths(net, K, 31, 0.1)
    threads = os.cpu_count() or 1
    per_step_s = max(2.0, min(20.0, 150.0 / max(1, args.steps + args.warmup)))
    vals, samples = [], ""
    for it in range(args.warmup + args.steps):
        v, evals, dt, samples = cpu_port_throughput(inst, paths, per_step_s, threads)
        if it >= args.warmup:
            vals.append((evals, dt))
    evals = sum(e for e, _ in vals); dt = sum(d for _, d in vals)
    value = evals / dt
    line = {
        "metric": METRIC, "value": value, "unit": "evals/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / max(1, args.steps), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int32 capacities, exact integer LP duals, fp64 fold", "data": "synthetic", "impl": "reference",
        "config": {"workload": desc, "note": "Gurobi (grb.cpp:231-235) is not installable here; this is Oracle B, the CPU port of the same path with the same SPEC-LP dual rule, every host thread, bounded sample per step"},
        "cpu_baseline": {"value": value, "unit": "evals/s", "cores": threads, "kind": "port", "sample": samples},
        "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    isolate_stdout()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    from sgufp_solver_b200 import _lib, instances as I
    from sgufp_solver_b200.distributed import I64_MAX, ShardedGuroSolver, finalize, reduce_partials, shard_bounds
    from sgufp_solver_b200.solver import GuroSolver

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    inst, S, S_total, K, desc = make_inputs(args.workload, rank, world)
    if world > 1:
        sh = ShardedGuroSolver(inst, S_total, rank, world, device=local, is_shard=True)
        solver = sh.solver
    else:
        sh = None
        solver = GuroSolver(inst, device=local)
    m, L, T, W = inst.m, solver.L, solver.T, solver.W
    paths = I.random_paths(solver, K, 31, 0.1)          # identical on every rank (same seed, same network)
    lib = _lib.lib()
    sums = torch.empty((K, W), dtype=torch.int64, device=dev)
    finf = torch.empty((K,), dtype=torch.int64, device=dev)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.Stream(dev)   # a real stream: the C ABI treats NULL as 'the handle's own stream'
    torch.cuda.set_stream(stream)

    def device_step():
        rc = lib.sgufp_paths_partial(solver.h, paths.ctypes.data_as(_lib.i16p), K, paths.shape[1], C.c_void_p(sums.data_ptr()),
                                     C.c_void_p(finf.data_ptr()), None, None, C.c_void_p(stream.cuda_stream))
        solver._check(rc)
        if world > 1:
            reduce_partials(sums, finf)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-timed value ----
    for _ in range(args.warmup):
        flush.zero_(); device_step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kernel_ms = []
    barrier()
    for a, b in ev:
        flush.zero_()
        a.record(stream)
        device_step()
        b.record(stream)
        kernel_ms.append(solver.last_kernel_ms())
    barrier()
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    assert int((finf.cpu() != I64_MAX).sum()) == 0, "the throughput workload has no lower bounds: every scenario must be feasible"
    # sanity: the reduced sums finalize into K cuts
    res = finalize(solver, paths, sums.cpu().numpy(), finf.cpu().numpy())
    assert np.isfinite(res.rhs).all()

    # ---- end to end through the reference-facing call ----
    api = sh if sh is not None else solver
    for _ in range(2):
        api.solve_paths(paths) if sh is not None else api.solve_paths(paths, want_obj=False, want_status=False, want_dense=False)
    barrier()
    e2e_s = 0.0
    for _ in range(args.steps):
        flush.zero_(); torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        out = api.solve_paths(paths) if sh is not None else api.solve_paths(paths, want_obj=False, want_status=False, want_dense=False)
        e2e_s += time.perf_counter() - t0
    e2e_t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_t.item())
    clocks = sampler.stop() if rank == 0 else None

    if rank == 0:
        evals_step = K * S_total
        value = evals_step * args.steps / (total_ms / 1e3)
        bytes_per_eval = 2 * m * 8 + 9
        k_ms = float(np.mean(kernel_ms))
        achieved = (K * S) * bytes_per_eval / (k_ms / 1e3) / 1e9            # this rank's kernel, its own algorithmic bytes
        peak, how = load_peaks()
        # plan words uploaded per step: measured from the C ABI's own batch (header + arrays), 4 B each
        plan_bytes = int(sum(_plan_words(solver, paths[k]) for k in range(K))) * 4 + K * 4 + K * 8
        line = {
            "metric": METRIC, "value": value, "unit": "evals/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64 capacities in HBM, exact int32/int64 LP + fold, f64 cut",
            "data": "synthetic",
            "config": {"workload": desc, "n": inst.n, "m": m, "scenarios_per_gpu": S, "scenarios_total": S_total, "candidates_per_step": K,
                       "L": L, "T": T, "l2": "flushed with a 512 MiB write before every timed step (C2's 3.2 MB of capacities are otherwise L2-resident)",
                       "timing": "sum over steps of CUDA-event pairs on the launching stream, max over ranks"},
            "e2e": {"value": evals_step * args.steps / e2e_s, "unit": "evals/s", "h2d_bytes_per_step": plan_bytes,
                    "d2h_bytes_per_step": int(K * W * 8 + K * 8), "ms_per_step": 1e3 * e2e_s / args.steps,
                    "call": "GuroSolver.solve_paths (sgufp_solve_paths): host int16 paths -> host Inavap::Cut list" if sh is None else
                            "ShardedGuroSolver.solve_paths: partial sums -> NCCL all-reduce -> host cuts"},
            "gpu_launches": args.steps * 1,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": measured_traffic(args.workload), "peak_source": how, "kernel": "k1_cut_eval", "kernel_ms": k_ms,
                         "bytes_per_eval": bytes_per_eval, "evals_per_launch": K * S,
                         "note": "K1 is instruction/latency-bound (exact LP per scenario), not HBM-bound; see DESIGN.md §6"},
            "clocks": clocks,
            "dd": None,
        }
        line["roofline"]["issue"] = issue_roofline(args.workload, k_ms, (clocks or {}).get("sm_mhz"),
                                                   torch.cuda.get_device_properties(local).multi_processor_count)
        if not args.no_cpu_baseline and world == 1:
            threads = 1
            v, evals, dt, sample = cpu_port_throughput(inst, paths, 12.0, threads)
            line["cpu_baseline"] = {"value": v, "unit": "evals/s", "cores": threads, "kind": "port", "sample": sample}
        else:
            line["cpu_baseline"] = None
        try:
            line["dd"] = dd_bench(local)
        except Exception as e:  # the DD half is reported beside the headline, never instead of it
            line["dd"] = {"error": str(e)[:200]}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


class DDCpuBaseline:
    """cpu_baseline leg of the DD benchmarks (bench_dd.py): Oracle A, the unmodified reference classes compiled into
    oracle/_ref, timed on one host core.  Like `cpu_port_throughput` this is the only place a bench executes oracle/."""

    def __init__(self, inst):
        self.net = None
        try:
            from oracle import ref_dd
            if ref_dd.available():
                self.ref_dd = ref_dd
                self.net = ref_dd.RefNetwork(inst)
        except Exception:
            self.net = None

    def available(self):
        return self.net is not None

    def _new(self, kind, width):
        return self.ref_dd.RefRestrictedDD(self.net, width) if kind == "restricted" else self.ref_dd.RefRelaxedDD(self.net)

    def supports(self, kind, width):
        return kind == "restricted" or width == 120          # the reference's relaxed threshold is a compile-time 120

    def build_ms(self, kind, width, reps=3):
        r = self._new(kind, width)
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            r.compile() if kind == "restricted" else r.build()
            ts.append(time.perf_counter() - t0)
        return float(np.median(ts)) * 1e3

    def apply_seconds(self, kind, width, cuts, with_solution=False):
        r = self._new(kind, width)
        r.compile() if kind == "restricted" else r.build()
        t0 = time.perf_counter()
        for c in cuts:
            r.apply_opt(c.RHS, c.keys, c.vals) if kind == "restricted" else r.apply_opt(c.RHS, c.keys, c.vals, -1e300, 1e300)
            if with_solution:
                r.solution()
        return time.perf_counter() - t0


def _plan_words(solver, path):
    # size of one uploaded plan: header (14 words) + 3m + 4*nch + 1 + m + ... ; upper bound used for the byte count
    return 16 + 4 * solver.m + 5 * solver.m + 2 * (solver.L + 2)


def dd_bench(device):
    """DD arcs/s of the companion longest-path kernel (K2), if built."""
    try:
        from sgufp_solver_b200 import dd as ddmod
    except Exception:
        return None
    if not hasattr(ddmod, "bench_longest_path"):
        return None
    return ddmod.bench_longest_path(device)


if __name__ == "__main__":
    main()
