#!/usr/bin/env python
"""bench.py — scenario-cut evaluations per second (BASELINE.json metric) on N B200s.

A STEP is one pass of the hot path over one batch of synthetic input: K candidate first-stage paths (emitted by the
decision-diagram master) x all scenarios = K*S exact subproblem solves folded into K cuts
(GuroSolver::solveSubProblem, /root/reference/grb.cpp:139-360).

Headline workload (the one BASELINE.md puts both targets on): **C5, strong scaling** — n=200, m=1000, S = 100 000
scenarios in total, contiguous blocks of 100 000/N per GPU, K = 8 candidates per step; N = 1 runs the same 100 000
scenarios on one GPU (1.6 GB of capacities).  Sub-records of the same run (key "sub"): C4 (S = 10 000, strong, K = 8) and
C2 (n=50, m=200, S = 1000 per GPU, K = 64; the reference's own single-GPU config, L2-resident).

  value     device-timed (CUDA events on the launching stream, max over ranks) whole-job evals/s with the capacity arrays
            resident in HBM; per step: plan upload + K1 kernel (+ the one all-reduce inside the library for N > 1)
  e2e       the same metric through the reference-facing call (host int16 paths in, host Inavap::Cut out:
            sgufp_solve_paths — a collective call on a partition), wall clock, copies included
  roofline  K1's algorithmic bytes (2*m*8 + 9 per evaluation, SURVEY.md §8d, fp64 storage) over its own CUDA-event
            duration, against MEASURED_PEAKS.json's HBM copy bandwidth; traffic = ncu dram bytes (profiles/k1_traffic.json)
  cpu_baseline  Oracle B (oracle/sgufp_oracle.c, the CPU port of the same path — Gurobi is not available)
            on a bounded sample of the same workload (N = 1 only)
  sharded_parity  N > 1: rank 0 recomputes the batch on ONE GPU over all scenarios (outside the timed region) and
            compares the reduced integer sums bit for bit

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


NBLOCKS = 8          # a strong-scaling instance is drawn in 8 blocks (one random stream each): the data do not depend on N
WORKLOADS = {
    "c5": {"net": "config4", "S_total": 100000, "K": 8, "scaling": "strong",
           "desc": "C5: n=200 m=1000, S=100000 scenarios in total, contiguous blocks of 100000/N per GPU, K=8 DD-emitted candidate paths per step (BASELINE.json configs[4]; 1.6 GB of fp64 capacities)"},
    "c4": {"net": "config4", "S_total": 10000, "K": 8, "scaling": "strong",
           "desc": "C4: n=200 m=1000, S=10000 scenarios in total, blocks of 10000/N per GPU, K=8 DD-emitted candidate paths per step (BASELINE.json configs[3])"},
    "c2": {"net": "config2", "S_per_gpu": 1000, "K": 64, "scaling": "weak",
           "desc": "C2: n=50 m=200, S=1000 scenarios per GPU, K=64 DD-emitted candidate paths per step (BASELINE.json configs[1]; 3.2 MB of capacities: L2-resident)"},
}
METRIC = "scenario_cut_evals_per_sec"
DTYPE = "f64 capacities in HBM, exact int32/int64 LP + fold, f64 cut"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def _traffic_file():
    p = os.path.join(ROOT, "profiles", "k1_traffic.json")
    return json.load(open(p)) if os.path.exists(p) else {}


def measured_traffic(workload):
    return _traffic_file().get(workload)


def issue_roofline(workload, kernel_ms, sm_mhz, sms):
    """Second roofline of K1: warp instructions per launch (ncu smsp__inst_executed.sum of the seeded workload, committed in
    profiles/k1_traffic.json) / live kernel time, against 4 issue slots per SM and clock.  It is the SM utilisation of the
    algorithm as built, not a bound of the problem."""
    n = _traffic_file().get("warp_instructions", {}).get(workload)
    if not n or not sm_mhz:
        return None
    achieved = n / (kernel_ms * 1e-3) / 1e9
    peak = sms * 4 * sm_mhz * 1e6 / 1e9
    return {"bound": "issue", "achieved": achieved, "peak": peak, "unit": "G warp-inst/s", "frac": achieved / peak,
            "warp_instructions_per_launch": n, "peak_source": "%d SMs x 4 schedulers x %.0f MHz (sampled under load)" % (sms, sm_mhz),
            "note": _traffic_file().get("_note_order")}


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


# ---- synthetic inputs -------------------------------------------------------------------------------------------------
def totals(workload, world):
    w = WORKLOADS[workload]
    return w["S_total"] if w["scaling"] == "strong" else w["S_per_gpu"] * world


def scenario_range(workload, lo, hi):
    """Scenarios [lo, hi) of the workload's instance as an `Instance` (the network does not depend on the range).  A strong
    workload is drawn in NBLOCKS blocks, a weak one in one block per GPU, each from its own random stream."""
    import dataclasses
    from sgufp_solver_b200 import instances as I
    w = WORKLOADS[workload]
    fn = getattr(I, w["net"])
    bs = w["S_total"] // NBLOCKS if w["scaling"] == "strong" else w["S_per_gpu"]
    parts = []
    for b in range(lo // bs, (max(lo, hi - 1)) // bs + 1):
        blk = fn(S=bs, cap_stream=b)
        a, z = max(lo, b * bs) - b * bs, min(hi, (b + 1) * bs) - b * bs
        parts.append((blk, a, z))
    base = parts[0][0]
    up = np.ascontiguousarray(np.concatenate([p.upper[:, a:z] for p, a, z in parts], axis=1))
    lw = np.ascontiguousarray(np.concatenate([p.lower[:, a:z] for p, a, z in parts], axis=1))
    rew = np.broadcast_to(base.reward[:, :1], (base.m, max(1, hi - lo)))      # the production path reads column 0 only (grb.cpp:53)
    return dataclasses.replace(base, S=hi - lo, upper=up, lower=lw, reward=rew)


def config_of(workload, world, inst_n, inst_m, L, T, K, paths_info):
    """The `config` object: identical for both arms (the driver compares them)."""
    w = WORKLOADS[workload]
    S_total = totals(workload, world)
    return {"workload": w["desc"], "n": inst_n, "m": inst_m, "scenarios_total": S_total,
            "scenarios_per_gpu": S_total // world if w["scaling"] == "strong" else w["S_per_gpu"],
            "candidates_per_step": K, "L": L, "T": T, "scaling": w["scaling"], "candidate_paths": paths_info,
            "l2": "flushed with a 512 MiB write before every timed step",
            "timing": "sum over steps of CUDA-event pairs on the launching stream, max over ranks"}


CANDIDATES_FILE = os.path.join(ROOT, "sgufp_solver_b200", "data", "bench_candidates.npz")


def step_batches(paths, K):
    """The batches the timed steps go through in turn: consecutive windows of K paths of the DD master's emission (two for
    C4 / C5: the committed emission holds 16 paths; one for C2).  Step i evaluates batch i % len(batches), so no step sees the
    paths of the step before it: what a step takes over from its predecessor is a STARTING flow for a different path, as
    consecutive calls of the Benders loop do (NodeExplorer.cpp:949-971), never a result."""
    nb = max(1, paths.shape[0] // K)
    return [np.ascontiguousarray(paths[b * K:(b + 1) * K]) for b in range(min(nb, 2))]


def candidate_paths(workload, K, device, want=None):
    """K candidate paths (`want` of them when the committed emission holds that many) as the DD master emits them
    (sgufp_solver_b200/candidates.py).  The committed file holds the
    emission of tools/make_bench_candidates.py (tests/test_e2e_gpu.py re-emits and compares), so both arms and every rank
    read the same list; without the file the master runs here (GPU), and random matchings fill in only if it cannot
    produce K distinct paths within its budget."""
    from sgufp_solver_b200 import instances as I
    from sgufp_solver_b200.candidates import dd_emitted_paths
    net = WORKLOADS[workload]["net"]
    if os.path.exists(CANDIDATES_FILE):
        z = np.load(CANDIDATES_FILE)
        if net in z.files and z[net].shape[0] >= K:
            n = K * max(1, min(want or K, z[net].shape[0]) // K)
            p = np.ascontiguousarray(z[net][:n], dtype=np.int16)
            how = "" if n == K else f"; the steps go through {n // K} batches of {K} consecutive emitted paths in turn"
            return p, ("DD-emitted: the paths RelaxedDDNew::getSolution hands to solveSubProblem in the Benders loop on a 32-scenario copy "
                       f"of the network (sgufp_solver_b200/data/bench_candidates.npz); matched fraction {float((p >= 0).mean()):.2f}{how}")
    fn = getattr(I, net)
    small = fn(S=32)
    paths, info = dd_emitted_paths(small, K, device=device, budget_s=45.0)
    src = "DD-emitted: the paths RelaxedDDNew::getSolution hands to solveSubProblem in the Benders loop on a 32-scenario copy of the network"
    if len(paths) < K:
        from sgufp_solver_b200.solver import GuroSolver
        extra = I.random_paths(GuroSolver(small, device=-1), K - len(paths), 31, 0.1)
        paths = np.concatenate([paths, extra], axis=0)
        src += f" ({info['emitted']} of {K}); the rest random matchings"
    return np.ascontiguousarray(paths[:K], dtype=np.int16), f"{src}; matched fraction {float((paths[:K] >= 0).mean()):.2f}"


# ---- CPU port (Oracle B): cpu_baseline and the reference arm ----------------------------------------------------------------
def cpu_port_throughput(inst, paths, seconds_target, threads):
    """Oracle B on a bounded sample: `threads` host threads, each with its own handle, each
    evaluating whole scenario ranges of the sampled candidates (the GIL is released in ctypes)."""
    from oracle.oracle import OracleNet
    nets = [OracleNet(inst) for _ in range(threads)]
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


CPU_SAMPLE_SCENARIOS = 2000     # the CPU legs hold this many scenarios of the workload per handle (a handle copies its arrays)


def run_reference(args, rank, world):
    """The reference arm: the CPU implementation of the path on the box's host cores (every host thread), on the same
    config.  Under torchrun rank 0 alone runs and prints it."""
    if rank != 0:
        return
    from sgufp_solver_b200 import instances as I
    from sgufp_solver_b200.solver import GuroSolver
    wl = args.workload
    K = WORKLOADS[wl]["K"]
    S_total = totals(wl, world)
    inst = scenario_range(wl, 0, min(S_total, CPU_SAMPLE_SCENARIOS))
    model = GuroSolver(inst, device=-1)
    if os.path.exists(CANDIDATES_FILE):
        paths, pinfo = candidate_paths(wl, K, int(os.environ.get("LOCAL_RANK", "0")), want=2 * K)     # the committed DD emission: no GPU work in this arm
        batches = step_batches(paths, K)
    else:
        paths, pinfo = np.asarray(I.random_paths(model, K, 31, 0.1), np.int16), "random matchings (no committed DD emission)"
        batches = [paths]
    threads = os.cpu_count() or 1
    per_step_s = max(2.0, min(20.0, 150.0 / max(1, args.steps + args.warmup)))
    vals, samples = [], ""
    for it in range(args.warmup + args.steps):
        v, evals, dt, samples = cpu_port_throughput(inst, batches[it % len(batches)], per_step_s, threads)   # the batches in turn, like the GPU arm
        if it >= args.warmup:
            vals.append((evals, dt))
    evals = sum(e for e, _ in vals); dt = sum(d for _, d in vals)
    value = evals / dt
    line = {
        "metric": METRIC, "value": value, "unit": "evals/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / max(1, args.steps), "higher_is_better": True, "scaling": WORKLOADS[wl]["scaling"], "vs_baseline": None,
        "dtype": DTYPE, "data": "synthetic", "impl": "reference",
        "config": config_of(wl, world, inst.n, inst.m, model.L, model.T, K, pinfo),
        "note": "Gurobi (grb.cpp:231-235) is not installable here; this is Oracle B, the CPU port of the same path with the same SPEC-LP dual rule, every host thread, bounded sample per step",
        "cpu_baseline": {"value": value, "unit": "evals/s", "cores": threads, "kind": "port", "sample": samples},
        "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ---- one workload on the GPUs ------------------------------------------------------------------------------------------------
def measure(wl, args, rank, world, local, dev, flush, steps, warmup, sample_clocks):
    import torch
    import torch.distributed as dist
    from sgufp_solver_b200 import _lib
    from sgufp_solver_b200.distributed import I64_MAX, shard_bounds
    from sgufp_solver_b200.solver import GuroSolver

    w = WORKLOADS[wl]
    K, S_total = w["K"], totals(wl, world)
    lo, hi = shard_bounds(S_total, world, rank)
    inst = scenario_range(wl, lo, hi)
    S = hi - lo
    solver = GuroSolver(inst, device=local, scenario_offset=lo, S_total=S_total)
    if world > 1:   # the partition's communicator lives in the library; its 128-byte id travels over the launcher's channel
        idt = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(GuroSolver.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, src=0)
        solver.comm_init(bytes(idt.cpu().numpy().tobytes()), rank, world)
    m, L, T, W = inst.m, solver.L, solver.T, solver.W
    all_paths, pinfo = candidate_paths(wl, K, local, want=2 * K)
    batches = step_batches(all_paths, K)
    paths = batches[0]
    turn = [0]                       # the batch the next step evaluates

    def next_batch():
        b = batches[turn[0] % len(batches)]
        turn[0] += 1
        return b
    lib = _lib.lib()
    stream = torch.cuda.ExternalStream(solver.stream_ptr(), device=dev)     # the handle's own stream: events are recorded where the kernel runs
    torch.cuda.set_stream(stream)
    sums = torch.empty((K, W), dtype=torch.int64, device=dev)
    finf = torch.empty((K,), dtype=torch.int64, device=dev)

    def device_step():
        p = next_batch()
        if world > 1:
            solver.paths_reduced(p)                 # K1 on this block + pack flags + the ONE ncclAllReduce, asynchronous
        else:
            solver._check(lib.sgufp_paths_partial(solver.h, p.ctypes.data_as(_lib.i16p), K, p.shape[1], C.c_void_p(sums.data_ptr()),
                                                  C.c_void_p(finf.data_ptr()), None, None, None))

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    last = [paths]                   # the batch of the last api_call (the sharded-parity check recomputes that one)

    def api_call():
        last[0] = next_batch()
        return solver.solve_paths(last[0], want_obj=False, want_status=False, want_dense=True)

    # ---- device-timed value ----
    for _ in range(warmup):
        flush.zero_(); device_step()
    barrier()
    sampler = ClockSampler(local) if sample_clocks and rank == 0 else None
    if sampler:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    kernel_ms, launches = [], 0
    barrier()
    for a, b in ev:
        flush.zero_()
        a.record(stream)
        device_step()
        b.record(stream)
        kernel_ms.append(solver.last_kernel_ms())
        launches += solver.last_stats()[0]
    barrier()
    total_ms = torch.tensor([sum(a.elapsed_time(b) for a, b in ev)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())

    # ---- end to end through the reference-facing call ----
    for _ in range(2):
        res = api_call()
    barrier()
    e2e_s = 0.0
    for _ in range(steps):
        flush.zero_(); torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        res = api_call()
        e2e_s += time.perf_counter() - t0
    e2e_t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_t.item())
    clocks = sampler.stop() if sampler else None

    # ---- the same step with every candidate solved from zero flow (no runs, no state between steps): what the warm starts buy ----
    cold_ms = None
    if world == 1:
        os.environ["SGUFP_K1_GROUP"], os.environ["SGUFP_K1_STATE"] = "1", "0"
        try:
            device_step(); barrier()
            cev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(3)]
            for a, b in cev:
                flush.zero_(); a.record(stream); device_step(); b.record(stream)
            barrier()
            cold_ms = sum(a.elapsed_time(b) for a, b in cev) / len(cev)
        finally:
            os.environ.pop("SGUFP_K1_GROUP", None); os.environ.pop("SGUFP_K1_STATE", None)
    assert (res.cut_type == 0).all() and np.isfinite(res.rhs).all(), "the throughput workload has no lower bounds: every cut is an optimality cut"
    info = solver.comm_info()

    # ---- N > 1: the reduced cuts against a ONE-GPU recomputation over all scenarios (rank 0, outside the timed region) ----
    parity = None
    if world > 1:
        if rank == 0:
            full = GuroSolver(scenario_range(wl, 0, S_total), device=local)
            ref = full.solve_paths(last[0], want_obj=False, want_status=False, want_dense=True)
            same = bool((ref.rhs == res.rhs).all() and (ref.coef_dense == res.coef_dense).all() and (ref.nnz == res.nnz).all())
            parity = "bit-identical" if same else "MISMATCH"
            full.close()
        dist.barrier()

    rec = None
    if rank == 0:
        evals_step = K * S_total
        bytes_per_eval = 2 * m * 8 + 9
        k_ms = float(np.mean(kernel_ms))
        achieved = (K * S) * bytes_per_eval / (k_ms / 1e3) / 1e9            # this rank's kernel, its own algorithmic bytes
        peak, how = load_peaks()
        plan_bytes = int(sum(_plan_words(solver, paths[k]) for k in range(K))) * 4 + K * 4 + K * 8
        traffic = measured_traffic(wl) if world == 1 else None
        run_len = solver.run_length(K)
        rec = {
            "value": evals_step * steps / (total_ms / 1e3), "unit": "evals/s", "ms_per_step": total_ms / steps, "scaling": w["scaling"],
            "config": config_of(wl, world, inst.n, m, L, T, K, pinfo),
            "k1": {"candidates_per_run": run_len, "state_between_steps": bool(run_len >= K),
                   "from_zero_flow": None if cold_ms is None else {"value": evals_step / (cold_ms / 1e3), "unit": "evals/s", "ms_per_step": cold_ms,
                                                                    "how": "SGUFP_K1_GROUP=1 SGUFP_K1_STATE=0, 3 steps"},
                   "batches_in_turn": len(batches),
                   "note": "a warp takes a run of consecutive candidates on one scenario and warm-starts each from the optimal flow and potentials "
                           "of the one before; when the whole batch is one run the handle keeps the last candidate's state per scenario, so the "
                           "first candidate of a step starts from the last one of the step before (consecutive paths of the Benders loop are a "
                           "few layers apart: 4 - 8 of 538 on C4 / C5, also across the batches the steps alternate between).  Every (candidate, "
                           "scenario) LP is solved to optimality in every step and no step repeats the paths of the one before it (C4 / C5); the "
                           "cuts are bit-identical to the ones from zero flow (tests/test_k1_gpu.py)"},
            "e2e": {"value": evals_step * steps / e2e_s, "unit": "evals/s", "h2d_bytes_per_step": plan_bytes,
                    "d2h_bytes_per_step": int(K * W * 8 + 2 * K * 8), "ms_per_step": 1e3 * e2e_s / steps,
                    "call": "GuroSolver.solve_paths (sgufp_solve_paths): host int16 paths -> host Inavap::Cut list" +
                            ("" if world == 1 else "; collective call on the partition: K1 -> one ncclAllReduce inside the library -> cuts on every rank")},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": how, "kernel": "k1_cut_eval + k1_cut_fold (the two kernels of one K1 launch)", "kernel_ms": k_ms,
                         "kernels_under_ncu": _traffic_file().get("kernels", {}).get(wl) if world == 1 else None,
                         "bytes_per_eval": bytes_per_eval, "evals_per_launch": K * S,
                         "note": "K1 is bound by instruction issue (flow kernel) and memory latency (cut kernel), not by HBM bandwidth: DESIGN.md §6.  The flow "
                                 "kernel reads a capacity row once per run of candidates, the cut kernel once per evaluation, and the optimal flows travel "
                                 "between the two through HBM: the measured DRAM traffic (traffic) is ~1.5 x the scope table's bytes on C4 / C5"},
            "exchange": None if world == 1 else {"collectives_per_step": info["exchanges_last_call"], "inside_library": True, "nccl": info["nccl"],
                                                 "words": K * W + K},
            "sharded_parity": parity,
            "clocks": clocks,
        }
        rec["roofline"]["issue"] = issue_roofline(wl, k_ms, (clocks or {}).get("sm_mhz"), torch.cuda.get_device_properties(local).multi_processor_count) if world == 1 else None
    torch.cuda.set_stream(torch.cuda.default_stream(dev))     # the handle's stream goes away with the handle
    solver.close()
    del sums, finf
    return rec, (inst if rank == 0 else None), paths


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c5", choices=sorted(WORKLOADS))
    ap.add_argument("--no-sub", action="store_true", help="only the headline workload (no C4 / C2 sub-records, no DD record)")
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

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    head, inst0, paths0 = measure(args.workload, args, rank, world, local, dev, flush, args.steps, args.warmup, True)
    subs = {}
    if not args.no_sub:
        for wl in ("c4", "c2"):
            if wl == args.workload:
                continue
            try:
                rec, _, _ = measure(wl, args, rank, world, local, dev, flush, args.steps, args.warmup, False)
                subs[wl] = rec
            except Exception as e:  # a sub-record is reported beside the headline, never instead of it
                subs[wl] = {"error": str(e)[:300]}
    if rank == 0:
        line = {"metric": METRIC, "value": head["value"], "unit": "evals/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": head["scaling"], "vs_baseline": None,
                "dtype": DTYPE, "data": "synthetic", "config": head["config"], "e2e": head["e2e"], "gpu_launches": head["gpu_launches"],
                "roofline": head["roofline"], "k1": head["k1"], "exchange": head["exchange"], "sharded_parity": head["sharded_parity"], "clocks": head["clocks"],
                "sub": {k: ({kk: vv for kk, vv in v.items() if kk != "clocks"} if isinstance(v, dict) else v) for k, v in subs.items()},
                "dd": None}
        if not args.no_cpu_baseline and world == 1:
            sample = scenario_range(args.workload, 0, min(totals(args.workload, 1), CPU_SAMPLE_SCENARIOS))
            v, evals, dt, desc = cpu_port_throughput(sample, paths0, 12.0, 1)
            line["cpu_baseline"] = {"value": v, "unit": "evals/s", "cores": 1, "kind": "port", "sample": desc}
        else:
            line["cpu_baseline"] = None
        if not args.no_sub:
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
    # size of one uploaded plan: header + 3m + 4*nch + 1 + m + ... ; upper bound used for the byte count
    return 24 + 4 * solver.m + 5 * solver.m + 2 * (solver.L + 2)


def dd_bench(device):
    """DD arcs/s of the companion longest-path kernel (K2): width sweep on distinct diagrams of one B&B frontier."""
    from sgufp_solver_b200 import dd as ddmod
    if hasattr(ddmod, "bench_frontier"):
        return ddmod.bench_frontier(device)
    return ddmod.bench_longest_path(device)


if __name__ == "__main__":
    main()
