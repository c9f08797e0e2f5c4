"""Host-side mirror of the reference's cut-generation interface.

`GuroSolver.solveSubProblem(path) -> (CutType, Cut)` keeps the names, argument meaning and
result format of `/root/reference/grb.h:75` / `grb.cpp:139-159`; `Cut` mirrors `Inavap::Cut`
(`/root/reference/Cut.h:201-337`).  All arithmetic runs in the CUDA library behind the C ABI.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import SgufpError, cip, dp, i16p, i64p, ip, u64p, u8p

OPTIMALITY, FEASIBILITY = 0, 1  # enum CutType, Cut.h:22-25
IQJ_MASK = 0xFFFFFFFFFFFF      # Cut.h:197


def getKey(q: int, i: int, j: int) -> int:
    """`Inavap::getKey` (Cut.h:342-344)."""
    return int(q) | (int(i) << 16) | (int(j) << 32)


class Cut:
    """`Inavap::Cut` (Cut.h:201-337): RHS, sparse (key, value) pairs in (i,q,j) order, hash."""

    def __init__(self, RHS: float, keys, vals):
        self.RHS = float(RHS)
        self.keys = np.ascontiguousarray(keys, dtype=np.uint64)
        self.vals = np.ascontiguousarray(vals, dtype=np.float64)
        self.hash_val = _lib.lib().sgufp_cut_hash(self.keys.ctypes.data_as(u64p), self.vals.ctypes.data_as(dp), len(self.keys))

    @property
    def coeff(self) -> List[Tuple[int, float]]:
        return list(zip(self.keys.tolist(), self.vals.tolist()))

    def get(self, key: int) -> float:
        """First pair whose low 48 bits match, else 0 (Cut.h:275-282)."""
        hit = np.nonzero((self.keys & np.uint64(IQJ_MASK)) == np.uint64(int(key) & IQJ_MASK))[0]
        return float(self.vals[hit[0]]) if len(hit) else 0.0

    def getRHS(self) -> float:
        return self.RHS

    def getHash(self) -> int:
        return self.hash_val

    def __eq__(self, other) -> bool:  # Cut.h:264-267
        return self.hash_val == other.hash_val and self.RHS == other.RHS and len(self.keys) == len(other.keys)


class BatchResult:
    """Everything one device call returns for K candidate paths."""

    def __init__(self, cut_type, rhs, keys, vals, nnz, coef_dense, obj, status, first_infeasible):
        self.cut_type, self.rhs, self.nnz = cut_type, rhs, nnz
        self._keys, self._vals = keys, vals
        self.coef_dense, self.obj, self.status, self.first_infeasible = coef_dense, obj, status, first_infeasible

    def cut(self, k: int) -> Cut:
        n = int(self.nnz[k])
        return Cut(self.rhs[k], self._keys[k, :n], self._vals[k, :n])

    def __len__(self):
        return len(self.rhs)


class GuroSolver:
    """Drop-in for the reference's `GuroSolver` (grb.h:17-104) on one B200.

    Construct from an instance in the reference's data model (`Network`, Network.h:69-117);
    `solveSubProblem(path)` is grb.h:75.  `solve_paths` evaluates K candidates in one launch."""

    def __init__(self, inst, device: int = 0, scenario_offset: int = 0, S_total=None, devices=None):
        """`devices`: a list of CUDA device ids — ONE process drives them all: the scenarios are cut into len(devices)
        contiguous blocks and the exchange runs inside the library (`sgufp_create_sharded`); otherwise one device."""
        L = _lib.lib()
        self.inst = inst
        self.n, self.m, self.S = int(inst.n), int(inst.m), int(inst.S)
        self.S_total = self.S if S_total is None else int(S_total)
        self.scenario_offset = int(scenario_offset)
        self.tail = np.ascontiguousarray(inst.tail, dtype=np.int32)
        self.head = np.ascontiguousarray(inst.head, dtype=np.int32)
        u = np.ascontiguousarray(inst.upper, dtype=np.int32)
        lo = np.ascontiguousarray(inst.lower, dtype=np.int32)
        r0 = np.ascontiguousarray(inst.reward[:, 0], dtype=np.int32)
        vb = np.ascontiguousarray(inst.vbar, dtype=np.int32)
        h = C.c_void_p()
        if devices is not None:
            assert scenario_offset == 0 and self.S_total == self.S, "a single-process partition takes the whole instance"
            devs = (C.c_int * len(devices))(*[int(d) for d in devices])
            rc = L.sgufp_create_sharded(C.byref(h), self.n, self.m, self.S, self.tail.ctypes.data_as(ip), self.head.ctypes.data_as(ip),
                                        u.ctypes.data_as(ip), lo.ctypes.data_as(ip), r0.ctypes.data_as(ip), vb.ctypes.data_as(ip), len(vb),
                                        devs, len(devices))
        else:
            rc = L.sgufp_create(C.byref(h), self.n, self.m, self.S, self.tail.ctypes.data_as(ip), self.head.ctypes.data_as(ip),
                                u.ctypes.data_as(ip), lo.ctypes.data_as(ip), r0.ctypes.data_as(ip), vb.ctypes.data_as(ip), len(vb),
                                int(device), self.scenario_offset, self.S_total)
        if rc:
            raise SgufpError(rc, L.sgufp_last_error(None).decode())
        self._adopt(h)

    def _adopt(self, h):
        """Everything the wrapper knows about the network comes from the handle (also for handles made from a cache file)."""
        L = _lib.lib()
        self.h = h
        n, m, S = C.c_int(), C.c_int(), C.c_int()
        off, tot = C.c_int64(), C.c_int64()
        L.sgufp_network(h, C.byref(n), C.byref(m), C.byref(S), C.byref(off), C.byref(tot), None, None)
        self.n, self.m, self.S, self.scenario_offset, self.S_total = n.value, m.value, S.value, off.value, tot.value
        self.tail, self.head = np.zeros(self.m, np.int32), np.zeros(self.m, np.int32)
        L.sgufp_network(h, None, None, None, None, None, self.tail.ctypes.data_as(ip), self.head.ctypes.data_as(ip))
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        L.sgufp_dims(h, C.byref(a), C.byref(b), C.byref(c))
        self.L, self.T = a.value, b.value
        self.vbar = np.zeros(c.value, np.int32)
        L.sgufp_vbar_order(h, self.vbar.ctypes.data_as(ip))
        self.layer_arc = np.zeros(self.L, np.int32)
        L.sgufp_processing_order(h, self.layer_arc.ctypes.data_as(ip))
        self.slot_i, self.slot_q, self.slot_j, self.slot_rank = (np.zeros(max(1, self.T), np.int32) for _ in range(4))
        L.sgufp_slots(h, *(x.ctypes.data_as(ip) for x in (self.slot_i, self.slot_q, self.slot_j, self.slot_rank)))
        self.W = L.sgufp_partial_width(h)
        self._out = [[] for _ in range(self.n)]
        for arc in range(self.m):
            self._out[int(self.tail[arc])].append(arc)

    @classmethod
    def from_cache(cls, path: str, device: int = 0, scenario_offset: int = 0, S_local: int = -1) -> "GuroSolver":
        """A handle whose capacities stream from a cache file (`instances.save_cache`, `sgufp_create_from_cache`) straight into
        the device layout; `scenario_offset` / `S_local` pick one rank's block of a partition."""
        L = _lib.lib()
        h = C.c_void_p()
        rc = L.sgufp_create_from_cache(C.byref(h), os.fsencode(path), int(device), int(scenario_offset), int(S_local))
        if rc:
            raise SgufpError(rc, L.sgufp_cache_last_error().decode())
        self = cls.__new__(cls)
        self.inst = None
        self._adopt(h)
        return self

    def clone(self) -> "GuroSolver":
        """Another handle (own stream and buffers) on the same device-resident capacities: one per host thread."""
        L = _lib.lib()
        h = C.c_void_p()
        rc = L.sgufp_clone(self.h, C.byref(h))
        if rc:
            raise SgufpError(rc, L.sgufp_cache_last_error().decode())
        other = type(self).__new__(type(self))
        other.inst = self.inst
        other._adopt(h)
        return other

    def out_arcs(self, q):
        return self._out[int(q)]

    def close(self):
        if getattr(self, "h", None):
            _lib.lib().sgufp_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc:
            raise SgufpError(rc, _lib.lib().sgufp_last_error(self.h).decode())

    # -- reference interface ---------------------------------------------------------------
    def solveSubProblem(self, path: Sequence[int]) -> Tuple[int, Cut]:
        """`std::pair<CutType, Inavap::Cut> GuroSolver::solveSubProblem(const vector<int16_t>&)`."""
        res = self.solve_paths(np.asarray(path, dtype=np.int16)[None, :], want_obj=False, want_status=False, want_dense=False)
        return int(res.cut_type[0]), res.cut(0)

    # -- batched form ----------------------------------------------------------------------
    def solve_paths(self, paths, want_obj=True, want_status=True, want_dense=True) -> BatchResult:
        L = _lib.lib()
        p = np.ascontiguousarray(paths, dtype=np.int16)
        K, plen = p.shape
        T1 = max(1, self.T)
        ct = np.zeros(K, np.int32); rhs = np.zeros(K); nnz = np.zeros(K, np.int32); fi = np.zeros(K, np.int64)
        keys = np.zeros((K, T1), np.uint64); vals = np.zeros((K, T1))
        dense = np.zeros((K, T1)) if want_dense else None
        obj = np.zeros((K, max(1, self.S))) if want_obj else None
        st = np.zeros((K, max(1, self.S)), np.uint8) if want_status else None
        rc = L.sgufp_solve_paths(self.h, p.ctypes.data_as(i16p), K, plen, ct.ctypes.data_as(cip), rhs.ctypes.data_as(dp),
                                 keys.ctypes.data_as(u64p), vals.ctypes.data_as(dp), nnz.ctypes.data_as(cip),
                                 dense.ctypes.data_as(dp) if want_dense else None, obj.ctypes.data_as(dp) if want_obj else None,
                                 st.ctypes.data_as(u8p) if want_status else None, fi.ctypes.data_as(i64p))
        self._check(rc)
        return BatchResult(ct, rhs, keys, vals, nnz, dense, obj, st, fi)

    # -- scenario partition with the exchange inside the library (include/sgufp_b200.h) -----------
    @staticmethod
    def comm_unique_id() -> bytes:
        """Rank 0: the 128 bytes every rank hands to `comm_init` (the host's launcher carries them)."""
        buf = C.create_string_buffer(128)
        rc = _lib.lib().sgufp_comm_unique_id(buf)
        if rc:
            raise SgufpError(rc, "sgufp_comm_unique_id: NCCL is not available")
        return buf.raw

    def comm_init(self, unique_id: bytes, rank: int, world: int) -> None:
        """Joins this block (created with scenario_offset / S_total) to the partition: `solve_paths` becomes a collective call."""
        buf = C.create_string_buffer(bytes(unique_id), 128)
        self._check(_lib.lib().sgufp_comm_init(self.h, buf, int(rank), int(world)))

    def comm_info(self):
        w, l, u, e = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        _lib.lib().sgufp_comm_info(self.h, C.byref(w), C.byref(l), C.byref(u), C.byref(e))
        return {"world": w.value, "local_ranks": l.value, "nccl": bool(u.value), "exchanges_last_call": e.value}

    def paths_reduced(self, paths):
        """Device half of a partitioned call (asynchronous on `stream_ptr()`): returns the device addresses of the reduced
        sums [K*W + K flags] and of this rank's first-infeasible marks."""
        p = np.ascontiguousarray(paths, dtype=np.int16)
        a, b = C.c_void_p(), C.c_void_p()
        self._check(_lib.lib().sgufp_paths_reduced(self.h, p.ctypes.data_as(i16p), p.shape[0], p.shape[1], C.byref(a), C.byref(b)))
        return a.value, b.value

    def stream_ptr(self) -> int:
        return int(_lib.lib().sgufp_stream(self.h))

    def last_stats(self):
        a, b = C.c_int(), C.c_float()
        _lib.lib().sgufp_last_stats(self.h, C.byref(a), C.byref(b))
        return a.value, b.value

    def run_length(self, K: int) -> int:
        """Candidates per run of a batch of K (sgufp_run_length): every candidate of a run after the first is warm-started
        from its predecessor on the same scenario."""
        return int(_lib.lib().sgufp_run_length(self.h, int(K)))

    def last_kernel_ms(self) -> float:
        """Device time of the last K1 launch (CUDA events on the launching stream)."""
        t = C.c_float()
        self._check(_lib.lib().sgufp_last_kernel_ms(self.h, C.byref(t)))
        return t.value
