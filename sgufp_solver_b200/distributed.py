"""Scenario-sharded cut evaluation: one process per GPU, contiguous scenario
blocks, ONE exchange step per batch of candidates — an all-reduce (SUM) of the exact integer
partial sums and an all-reduce (MIN) of the lowest infeasible scenario index.

`torch` is only plumbing here: device buffers, the NCCL (or gloo, for CPU tests) process group.
The sums are int64, so 1-GPU and N-GPU cuts are bit-identical whatever the reduction order.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _lib
from ._lib import cip, dp, i16p, i64p, u64p
from .solver import BatchResult, GuroSolver

I64_MAX = np.iinfo(np.int64).max


def shard_bounds(S_total: int, world: int, rank: int):
    """Contiguous block of rank `rank`: [lo, hi).  Contiguity keeps 'lowest infeasible index' a MIN."""
    base, rem = divmod(S_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def reduce_partials(sums, first_inf, group=None):
    """The single exchange step.  `sums` [K, W] int64 and `first_inf` [K] int64 torch tensors
    (CUDA for NCCL, CPU for gloo), reduced in place."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(first_inf, op=dist.ReduceOp.MIN, group=group)
    return sums, first_inf


def owner_of(scenario: int, S_total: int, world: int) -> int:
    for r in range(world):
        lo, hi = shard_bounds(S_total, world, r)
        if lo <= scenario < hi:
            return r
    raise ValueError(scenario)


class ShardedGuroSolver:
    """`GuroSolver` over a scenario partition: same `solveSubProblem(path)` contract on every rank."""

    def __init__(self, inst_full_or_shard, S_total: int, rank: int, world: int, device: int = 0, is_shard: bool = False, group=None,
                 library_exchange: Optional[bool] = None):
        """`library_exchange`: the all-reduce runs inside libsgufp_b200.so (`sgufp_comm_init`, NCCL loaded by the library) and
        `solve_paths` is one C call; default: whenever the process group runs on NCCL.  Otherwise the exchange is
        `torch.distributed` on the partial sums (what the CPU tests drive over gloo)."""
        import torch
        self.torch = torch
        self.rank, self.world, self.group = rank, world, group
        self.S_total = int(S_total)
        lo, hi = shard_bounds(self.S_total, world, rank)
        inst = inst_full_or_shard if is_shard else inst_full_or_shard.scenario_slice(lo, hi)
        assert inst.S == hi - lo
        self.lo, self.hi = lo, hi
        self.device = torch.device("cuda", device)
        self.solver = GuroSolver(inst, device=device, scenario_offset=lo, S_total=self.S_total)
        self.W, self.T, self.L = self.solver.W, self.solver.T, self.solver.L
        # a real (non-default) stream: kernel, all-reduce and read-back are ordered on it
        self.stream = torch.cuda.Stream(self.device)
        self._sums = None
        self._finf = None
        import torch.distributed as dist
        if library_exchange is None:
            library_exchange = world > 1 and dist.is_initialized() and dist.get_backend(group) == "nccl"
        self.library_exchange = bool(library_exchange)
        if self.library_exchange:
            # the 128-byte NCCL id travels over the launcher's own channel (here: the torch process group)
            idt = torch.zeros(128, dtype=torch.uint8, device=self.device)
            if rank == 0:
                idt.copy_(torch.frombuffer(bytearray(GuroSolver.comm_unique_id()), dtype=torch.uint8))
            if world > 1:
                dist.broadcast(idt, src=0, group=group)
            self.solver.comm_init(bytes(idt.cpu().numpy().tobytes()), rank, world)

    def partial(self, paths):
        """Step 1: local scenarios -> device partial sums (no host round trip)."""
        torch = self.torch
        p = np.ascontiguousarray(paths, dtype=np.int16)
        K, plen = p.shape
        if self._sums is None or self._sums.shape[0] != K:
            self._sums = torch.empty((K, self.W), dtype=torch.int64, device=self.device)
            self._finf = torch.empty((K,), dtype=torch.int64, device=self.device)
        rc = _lib.lib().sgufp_paths_partial(self.solver.h, p.ctypes.data_as(i16p), K, plen, C.c_void_p(self._sums.data_ptr()),
                                            C.c_void_p(self._finf.data_ptr()), None, None, C.c_void_p(self.stream.cuda_stream))
        self.solver._check(rc)
        return self._sums, self._finf

    def solve_paths(self, paths) -> BatchResult:
        torch = self.torch
        import torch.distributed as dist
        p = np.ascontiguousarray(paths, dtype=np.int16)
        K, plen = p.shape
        if self.library_exchange:      # one C call: K1 -> ncclAllReduce -> cuts (capi_shard.cu); obj/status cover this rank's block
            return self.solver.solve_paths(p, want_obj=False, want_status=False)
        with torch.cuda.stream(self.stream):
            sums, finf = self.partial(p)
            reduce_partials(sums, finf, self.group)
            finf_h = finf.cpu().numpy()
            if (finf_h < 0).any():   # a rank's subproblem solver hit its iteration guard (k1_cut.cu: fuel): no cut can be built
                raise _lib.SgufpError(-5, "the iteration guard of the subproblem solver was hit on some rank")
            # feasibility: the rank owning the lowest infeasible scenario builds the ray, everyone gets it
            for k in np.nonzero(finf_h != I64_MAX)[0]:
                own = owner_of(int(finf_h[k]), self.S_total, self.world)
                if own == self.rank:
                    rc = _lib.lib().sgufp_ray_partial(self.solver.h, p[k].ctypes.data_as(i16p), plen, int(finf_h[k]),
                                                      C.c_void_p(sums[k].data_ptr()), C.c_void_p(self.stream.cuda_stream))
                    self.solver._check(rc)
                if self.world > 1:
                    dist.broadcast(sums[k], src=own, group=self.group)
            sums_h = sums.cpu().numpy()
        return finalize(self.solver, p, sums_h, finf_h)

    def solveSubProblem(self, path):
        res = self.solve_paths(np.asarray(path, dtype=np.int16)[None, :])
        return int(res.cut_type[0]), res.cut(0)


def finalize(solver: GuroSolver, paths: np.ndarray, sums_h: np.ndarray, finf_h: np.ndarray) -> BatchResult:
    """Step 3: reduced integer sums -> `Inavap::Cut`s (host, O(T) per candidate)."""
    p = np.ascontiguousarray(paths, dtype=np.int16)
    K, plen = p.shape
    T1 = max(1, solver.T)
    sums_h = np.ascontiguousarray(sums_h, dtype=np.int64)
    finf_h = np.ascontiguousarray(finf_h, dtype=np.int64)
    ct = np.zeros(K, np.int32); rhs = np.zeros(K); nnz = np.zeros(K, np.int32)
    keys = np.zeros((K, T1), np.uint64); vals = np.zeros((K, T1)); dense = np.zeros((K, T1))
    rc = _lib.lib().sgufp_finalize_paths(solver.h, p.ctypes.data_as(i16p), K, plen, sums_h.ctypes.data_as(i64p), finf_h.ctypes.data_as(i64p),
                                         ct.ctypes.data_as(cip), rhs.ctypes.data_as(dp), keys.ctypes.data_as(u64p), vals.ctypes.data_as(dp),
                                         nnz.ctypes.data_as(cip), dense.ctypes.data_as(dp))
    solver._check(rc)
    fi = np.where(finf_h == I64_MAX, -1, finf_h)
    return BatchResult(ct, rhs, keys, vals, nnz, dense, None, None, fi)
