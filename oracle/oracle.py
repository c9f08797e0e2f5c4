"""ctypes front-end of Oracle B (oracle/sgufp_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product path must never do so.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liborc.so")
    src = os.path.join(_HERE, "sgufp_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O3", "-march=x86-64-v2", "-fPIC", "-shared", "-o", so, src])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        ip = C.POINTER(C.c_int)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_int, C.c_int, C.c_int, ip, ip, ip, ip, ip, ip, C.c_int]
        L.orc_free.argtypes = [C.c_void_p]
        for f in ("orc_L", "orc_T", "orc_nvbar"):
            getattr(L, f).argtypes = [C.c_void_p]
            getattr(L, f).restype = C.c_int
        L.orc_get_vbar.argtypes = [C.c_void_p, ip]
        L.orc_get_layers.argtypes = [C.c_void_p, ip]
        L.orc_get_slots.argtypes = [C.c_void_p, ip, ip, ip]
        L.orc_solve_path.restype = C.c_int
        L.orc_solve_path.argtypes = [C.c_void_p, C.POINTER(C.c_int16), C.c_int, ip, C.POINTER(C.c_double),
                                     C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.POINTER(C.c_double), ip,
                                     C.POINTER(C.c_double), C.POINTER(C.c_ubyte), C.POINTER(C.c_longlong), ip]
        L.orc_scenario_duals.restype = C.c_int
        L.orc_scenario_duals.argtypes = [C.c_void_p, C.POINTER(C.c_int16), C.c_int, C.c_int] + [ip] * 7 + [C.POINTER(C.c_double)]
        L.orc_scenario_ray.restype = C.c_int
        L.orc_scenario_ray.argtypes = [C.c_void_p, C.POINTER(C.c_int16), C.c_int, C.c_int] + [ip] * 7
        L.orc_solve_range.restype = C.c_int
        L.orc_solve_range.argtypes = [C.c_void_p, C.POINTER(C.c_int16), C.c_int, C.c_int, C.c_int, C.POINTER(C.c_longlong), ip]
        _LIB = L
    return _LIB


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


class OracleCut:
    def __init__(self, cut_type, rhs, coef_dense, keys, vals, obj, status, isum, first_infeasible):
        self.cut_type = cut_type  # 0 OPTIMALITY, 1 FEASIBILITY (Cut.h:22-25)
        self.rhs = rhs
        self.coef_dense = coef_dense
        self.keys = keys
        self.vals = vals
        self.obj = obj
        self.status = status
        self.isum = isum
        self.first_infeasible = first_infeasible


class OracleNet:
    """Mirror of `Network` + `GuroSolver` (Network.h:69-117, grb.h:17-104) on the CPU."""

    def __init__(self, inst):
        L = lib()
        self.inst = inst
        self.n, self.m, self.S = inst.n, inst.m, inst.S
        self.tail = np.ascontiguousarray(inst.tail, dtype=np.int32)
        self.head = np.ascontiguousarray(inst.head, dtype=np.int32)
        u = np.ascontiguousarray(inst.upper, dtype=np.int32)
        l = np.ascontiguousarray(inst.lower, dtype=np.int32)
        r0 = np.ascontiguousarray(inst.reward[:, 0], dtype=np.int32)
        vb = np.ascontiguousarray(inst.vbar, dtype=np.int32)
        self.h = L.orc_create(inst.n, inst.m, inst.S, _ip(self.tail), _ip(self.head), _ip(u), _ip(l), _ip(r0), _ip(vb), len(vb))
        if not self.h:
            raise ValueError("oracle: shuffleVBarNodes would not terminate on this instance")
        self.L = L.orc_L(self.h)
        self.T = L.orc_T(self.h)
        self.vbar = np.zeros(L.orc_nvbar(self.h), dtype=np.int32)
        L.orc_get_vbar(self.h, _ip(self.vbar))
        self.layer_arc = np.zeros(self.L, dtype=np.int32)
        L.orc_get_layers(self.h, _ip(self.layer_arc))
        self.slot_i = np.zeros(self.T, dtype=np.int32)
        self.slot_q = np.zeros(self.T, dtype=np.int32)
        self.slot_j = np.zeros(self.T, dtype=np.int32)
        L.orc_get_slots(self.h, _ip(self.slot_i), _ip(self.slot_q), _ip(self.slot_j))
        self._out = [[] for _ in range(self.n)]
        for a in range(self.m):
            self._out[int(self.tail[a])].append(a)

    def out_arcs(self, q):
        return self._out[int(q)]

    def __del__(self):
        try:
            if getattr(self, "h", None):
                lib().orc_free(self.h)
                self.h = None
        except Exception:
            pass

    def solve_path(self, path) -> OracleCut:
        L = lib()
        p = np.ascontiguousarray(path, dtype=np.int16)
        ct = C.c_int(0); rhs = C.c_double(0); nnz = C.c_int(0); fi = C.c_int(-1)
        coef = np.zeros(self.T, dtype=np.float64)
        keys = np.zeros(self.T, dtype=np.uint64)
        vals = np.zeros(self.T, dtype=np.float64)
        obj = np.zeros(self.S, dtype=np.float64)
        st = np.zeros(self.S, dtype=np.uint8)
        isum = np.zeros(self.T + 1, dtype=np.int64)
        rc = L.orc_solve_path(self.h, p.ctypes.data_as(C.POINTER(C.c_int16)), len(p), C.byref(ct), C.byref(rhs),
                              coef.ctypes.data_as(C.POINTER(C.c_double)), keys.ctypes.data_as(C.POINTER(C.c_uint64)),
                              vals.ctypes.data_as(C.POINTER(C.c_double)), C.byref(nnz),
                              obj.ctypes.data_as(C.POINTER(C.c_double)), st.ctypes.data_as(C.POINTER(C.c_ubyte)),
                              isum.ctypes.data_as(C.POINTER(C.c_longlong)), C.byref(fi))
        if rc:
            raise ValueError(f"oracle: invalid path (code {rc})")
        return OracleCut(ct.value, rhs.value, coef, keys[: nnz.value].copy(), vals[: nnz.value].copy(), obj, st, isum, fi.value)

    def _duals(self, fn, path, s, with_obj):
        p = np.ascontiguousarray(path, dtype=np.int16)
        arrs = [np.zeros(self.n, np.int32)] + [np.zeros(self.m, np.int32) for _ in range(4)] + [np.zeros(self.T, np.int32) for _ in range(2)]
        args = [self.h, p.ctypes.data_as(C.POINTER(C.c_int16)), len(p), int(s)] + [_ip(a) for a in arrs]
        obj = C.c_double(0)
        if with_obj:
            args.append(C.byref(obj))
        st = fn(*args)
        names = ["alpha", "beta", "gamma", "sigma", "phi", "lambda", "mu"]
        d = dict(zip(names, arrs))
        d["status"] = st
        d["obj"] = obj.value
        return d

    def scenario_duals(self, path, s):
        return self._duals(lib().orc_scenario_duals, path, s, True)

    def scenario_ray(self, path, s):
        return self._duals(lib().orc_scenario_ray, path, s, False)

    def solve_range(self, path, s0, s1):
        p = np.ascontiguousarray(path, dtype=np.int16)
        isum = np.zeros(self.T + 1, dtype=np.int64)
        bad = C.c_int(0)
        rc = lib().orc_solve_range(self.h, p.ctypes.data_as(C.POINTER(C.c_int16)), len(p), int(s0), int(s1),
                                   isum.ctypes.data_as(C.POINTER(C.c_longlong)), C.byref(bad))
        if rc:
            raise ValueError(f"oracle: invalid path (code {rc})")
        return isum, bad.value
