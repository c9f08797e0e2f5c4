"""ctypes front-end of Oracle A: the UNMODIFIED reference DD / Network / Cut code compiled into
oracle/_ref/libsgufp_ref.so (oracle/Makefile, oracle/ref_shim.cpp).  TEST INFRASTRUCTURE ONLY.

/root/reference does not exist on the GPU box; the prebuilt library travels with the snapshot.
`available()` says whether it can be loaded; tests fall back to tests/golden/ otherwise.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libsgufp_ref.so")
_LIB = None


def build() -> bool:
    """(Re)build from /root/reference when it is present; otherwise keep the prebuilt file."""
    if os.path.isdir("/root/reference"):
        subprocess.check_call(["make", "-s", "-C", _HERE, "ref"])
    return os.path.exists(_SO)


def available() -> bool:
    return os.path.exists(_SO)


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(_SO)
        vp, ip, dp = C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_double)
        kp, sp, lp = C.POINTER(C.c_uint64), C.POINTER(C.c_int16), C.POINTER(C.c_long)
        sig = {
            "ref_net_load": (vp, [C.c_char_p]), "ref_net_free": (None, [vp]),
            "ref_net_n": (C.c_int, [vp]), "ref_net_m": (C.c_int, [vp]), "ref_net_total_layers": (C.c_int, [vp]),
            "ref_net_nvbar": (C.c_int, [vp]), "ref_net_vbar": (None, [vp, ip]),
            "ref_net_processing_order": (None, [vp, ip, ip]), "ref_net_has_state_changed": (None, [vp, ip]),
            "ref_net_nhsc": (C.c_int, [vp]), "ref_net_state_update": (C.c_int, [vp, C.c_int, ip]),
            "ref_net_classes": (None, [vp, ip, ip, ip, ip]),
            "ref_cut_to_cut": (C.c_int, [C.c_int, C.c_double, ip, ip, ip, dp, C.c_int, kp, dp, kp]),
            "ref_cut_hash": (C.c_uint64, [C.c_double, kp, dp, C.c_int]),
            "ref_cut_get": (C.c_double, [C.c_double, kp, dp, C.c_int, C.c_uint64]),
            "ref_get_key": (C.c_uint64, [C.c_uint64, C.c_uint64, C.c_uint64]),
            "ref_rel_new": (vp, [vp]), "ref_rel_free": (None, [vp]),
            "ref_rel_build": (None, [vp, sp, C.c_int, sp, C.c_int, C.c_int]),
            "ref_rel_is_exact": (C.c_int, [vp]),
            "ref_rel_apply_opt": (C.c_double, [vp, C.c_double, kp, dp, C.c_int, C.c_double, C.c_double]),
            "ref_rel_apply_feas": (C.c_int, [vp, C.c_double, kp, dp, C.c_int]),
            "ref_rel_solution": (C.c_int, [vp, sp]), "ref_rel_nlayers": (C.c_int, [vp]),
            "ref_rel_layer_sizes": (None, [vp, ip]),
            "ref_rel_dump": (C.c_long, [vp, ip, dp, lp, ip, ip, dp]),
            "ref_rel_count_nodes": (C.c_long, [vp]), "ref_rel_count_arcs": (C.c_long, [vp]),
            "ref_rel_cutset": (C.c_int, [vp, C.c_double, ip, C.c_int]),
            "ref_res_new": (vp, [vp, C.c_int]), "ref_res_free": (None, [vp]),
            "ref_res_compile": (C.c_int, [vp, sp, C.c_int, sp, C.c_int, C.c_int]),
            "ref_res_is_exact": (C.c_int, [vp]),
            "ref_res_apply_opt": (C.c_double, [vp, C.c_double, kp, dp, C.c_int]),
            "ref_res_apply_feas": (C.c_int, [vp, C.c_double, kp, dp, C.c_int]),
            "ref_res_solution": (C.c_int, [vp, sp]), "ref_res_nlayers": (C.c_int, [vp]),
            "ref_res_layer_sizes": (None, [vp, ip]), "ref_res_dump": (C.c_long, [vp, ip, ip, dp]),
            "ref_res_cutset": (C.c_int, [vp, ip, C.c_int]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _cutargs(rhs, keys, vals):
    keys = np.ascontiguousarray(keys, dtype=np.uint64)
    vals = np.ascontiguousarray(vals, dtype=np.float64)
    return float(rhs), _p(keys, C.c_uint64), _p(vals, C.c_double), len(keys), (keys, vals)


def _unpack_nodes(buf, k):
    out, i = [], 0
    while i < k:
        gl = int(buf[i]); ns = int(buf[i + 1]); st = buf[i + 2:i + 2 + ns].tolist(); i += 2 + ns
        nl = int(buf[i]); sol = buf[i + 1:i + 1 + nl].tolist(); i += 1 + nl
        out.append((gl, st, sol))
    return out


class RefNetwork:
    """The reference `Network`, constructed by its own parser from the text form of `inst`."""

    def __init__(self, inst):
        L = lib()
        fd, path = tempfile.mkstemp(suffix=".txt")
        os.close(fd)
        try:
            inst.write_text(path)
            self.h = L.ref_net_load(path.encode())
        finally:
            os.unlink(path)
        self.n = L.ref_net_n(self.h)
        self.m = L.ref_net_m(self.h)
        self.total_layers = L.ref_net_total_layers(self.h)
        self.vbar = np.zeros(L.ref_net_nvbar(self.h), np.int32)
        L.ref_net_vbar(self.h, _p(self.vbar, C.c_int))
        lay = np.zeros(self.total_layers, np.int32)
        self.layer_arc = np.zeros(self.total_layers, np.int32)
        L.ref_net_processing_order(self.h, _p(lay, C.c_int), _p(self.layer_arc, C.c_int))
        self.has_state_changed = np.zeros(L.ref_net_nhsc(self.h), np.int32)
        L.ref_net_has_state_changed(self.h, _p(self.has_state_changed, C.c_int))
        self.state_update = {}
        buf = np.zeros(self.m + 2, np.int32)
        for k in range(self.total_layers):
            c = L.ref_net_state_update(self.h, k, _p(buf, C.c_int))
            if c >= 0:
                self.state_update[k] = buf[:c].tolist()
        a = [C.c_int(0) for _ in range(4)]
        L.ref_net_classes(self.h, *[C.byref(x) for x in a])
        self.classes = tuple(x.value for x in a)

    def __del__(self):
        try:
            if getattr(self, "h", None):
                lib().ref_net_free(self.h); self.h = None
        except Exception:
            pass


class RefRelaxedDD:
    """`Inavap::RelaxedDDNew` (DD.h:734-810)."""

    def __init__(self, net: RefNetwork):
        self.net = net
        self.h = lib().ref_rel_new(net.h)

    def __del__(self):
        try:
            if getattr(self, "h", None):
                lib().ref_rel_free(self.h); self.h = None
        except Exception:
            pass

    def build(self, states=(), solution=(), global_layer=0):
        st = np.ascontiguousarray(states, dtype=np.int16); so = np.ascontiguousarray(solution, dtype=np.int16)
        lib().ref_rel_build(self.h, _p(st, C.c_int16), len(st), _p(so, C.c_int16), len(so), int(global_layer))

    def is_exact(self):
        return bool(lib().ref_rel_is_exact(self.h))

    def apply_opt(self, rhs, keys, vals, optimal, ub):
        r, k, v, n, keep = _cutargs(rhs, keys, vals)
        return lib().ref_rel_apply_opt(self.h, r, k, v, n, float(optimal), float(ub))

    def apply_feas(self, rhs, keys, vals):
        r, k, v, n, keep = _cutargs(rhs, keys, vals)
        return lib().ref_rel_apply_feas(self.h, r, k, v, n)

    def solution(self):
        buf = np.zeros(self.net.total_layers + 4, np.int16)
        k = lib().ref_rel_solution(self.h, _p(buf, C.c_int16))
        return buf[:k].copy()

    def layer_sizes(self):
        out = np.zeros(lib().ref_rel_nlayers(self.h), np.int32)
        lib().ref_rel_layer_sizes(self.h, _p(out, C.c_int))
        return out

    def dump(self):
        L = lib()
        nn, na = L.ref_rel_count_nodes(self.h), L.ref_rel_count_arcs(self.h)
        node_layer = np.zeros(nn, np.int32); node_state = np.zeros(nn, np.float64); inptr = np.zeros(nn + 1, np.int64)
        tailpos = np.zeros(na, np.int32); dec = np.zeros(na, np.int32); w = np.zeros(na, np.float64)
        got = L.ref_rel_dump(self.h, _p(node_layer, C.c_int), _p(node_state, C.c_double), _p(inptr, C.c_long),
                             _p(tailpos, C.c_int), _p(dec, C.c_int), _p(w, C.c_double))
        assert got == na
        return dict(node_layer=node_layer, node_state=node_state, in_ptr=inptr, arc_tailpos=tailpos, arc_decision=dec, arc_weight=w)

    def cutset(self, ub):
        buf = np.zeros(1 << 22, np.int32)
        k = lib().ref_rel_cutset(self.h, float(ub), _p(buf, C.c_int), len(buf))
        assert k >= 0
        return _unpack_nodes(buf, k)


class RefRestrictedDD:
    """`Inavap::RestrictedDDNew` (DD.h:653-730)."""

    def __init__(self, net: RefNetwork, width: int):
        self.net = net
        self.h = lib().ref_res_new(net.h, int(width))

    def __del__(self):
        try:
            if getattr(self, "h", None):
                lib().ref_res_free(self.h); self.h = None
        except Exception:
            pass

    def compile(self, states=(), solution=(), global_layer=0):
        st = np.ascontiguousarray(states, dtype=np.int16); so = np.ascontiguousarray(solution, dtype=np.int16)
        return lib().ref_res_compile(self.h, _p(st, C.c_int16), len(st), _p(so, C.c_int16), len(so), int(global_layer))

    def is_exact(self):
        return bool(lib().ref_res_is_exact(self.h))

    def apply_opt(self, rhs, keys, vals):
        r, k, v, n, keep = _cutargs(rhs, keys, vals)
        return lib().ref_res_apply_opt(self.h, r, k, v, n)

    def apply_feas(self, rhs, keys, vals):
        r, k, v, n, keep = _cutargs(rhs, keys, vals)
        return lib().ref_res_apply_feas(self.h, r, k, v, n)

    def solution(self):
        buf = np.zeros(self.net.total_layers + 4, np.int16)
        k = lib().ref_res_solution(self.h, _p(buf, C.c_int16))
        return buf[:k].copy()

    def layer_sizes(self):
        out = np.zeros(lib().ref_res_nlayers(self.h), np.int32)
        lib().ref_res_layer_sizes(self.h, _p(out, C.c_int))
        return out

    def dump(self):
        sizes = self.layer_sizes()
        nn = int(sizes[:-1].sum())
        pp = np.zeros(nn, np.int32); dec = np.zeros(nn, np.int32); st = np.zeros(nn, np.float64)
        got = lib().ref_res_dump(self.h, _p(pp, C.c_int), _p(dec, C.c_int), _p(st, C.c_double))
        assert got == nn
        return dict(parentpos=pp, decision=dec, state=st)

    def cutset(self):
        buf = np.zeros(1 << 22, np.int32)
        k = lib().ref_res_cutset(self.h, _p(buf, C.c_int), len(buf))
        assert k >= 0
        return _unpack_nodes(buf, k)


def cut_to_cut(cut_type, rhs, triples):
    """`Inavap::cutToCut` (Cut.h:406-421) on a `::Cut` given as {(i,q,j): value}."""
    ci = np.array([t[0] for t in triples], np.int32); cq = np.array([t[1] for t in triples], np.int32)
    cj = np.array([t[2] for t in triples], np.int32); cv = np.array([triples[t] for t in triples], np.float64)
    keys = np.zeros(len(ci), np.uint64); vals = np.zeros(len(ci), np.float64); h = C.c_uint64(0)
    k = lib().ref_cut_to_cut(int(cut_type), float(rhs), _p(ci, C.c_int), _p(cq, C.c_int), _p(cj, C.c_int), _p(cv, C.c_double),
                             len(ci), _p(keys, C.c_uint64), _p(vals, C.c_double), C.byref(h))
    return keys[:k].copy(), vals[:k].copy(), h.value


def cut_hash(rhs, keys, vals):
    r, k, v, n, keep = _cutargs(rhs, keys, vals)
    return lib().ref_cut_hash(r, k, v, n)


def cut_get(rhs, keys, vals, key):
    r, k, v, n, keep = _cutargs(rhs, keys, vals)
    return lib().ref_cut_get(r, k, v, n, int(key))
