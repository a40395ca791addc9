"""ctypes front end of the C oracle (oracle/cone_oracle.c) -- TEST INFRASTRUCTURE.

Restates what the reference obtains from ``diffcp.solve_and_derivative_batch``
(``src/cvxpylayers/interfaces/diffcp_if.py:365``) and the adjoint closure
(``diffcp_if.py:86``).  Never imported by the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

from cvxpylayers_b200.structure import Structure

_HERE = Path(__file__).resolve().parent
_LIB = None

_i32p = C.POINTER(C.c_int32)
_f64p = C.POINTER(C.c_double)


class OrcDesc(C.Structure):
    _fields_ = [("n", C.c_int32), ("m", C.c_int32), ("nnzA", C.c_int32), ("nnzP", C.c_int32),
                ("A_indptr", _i32p), ("A_indices", _i32p), ("P_indptr", _i32p), ("P_indices", _i32p),
                ("z", C.c_int32), ("l", C.c_int32), ("nq", C.c_int32), ("ns", C.c_int32),
                ("ep", C.c_int32), ("ed", C.c_int32), ("q", _i32p), ("s", _i32p)]


class OrcSettings(C.Structure):
    _fields_ = [("eps_abs", C.c_double), ("eps_rel", C.c_double), ("eps_infeas", C.c_double),
                ("alpha", C.c_double), ("rho_x", C.c_double), ("scale", C.c_double),
                ("lsqr_atol", C.c_double), ("lsqr_btol", C.c_double), ("lsqr_conlim", C.c_double),
                ("max_iters", C.c_int32), ("normalize", C.c_int32), ("adaptive_scale", C.c_int32),
                ("check_interval", C.c_int32), ("ruiz_passes", C.c_int32), ("lsqr_iter_lim", C.c_int32),
                ("lsqr_precond", C.c_int32), ("adaptive_check", C.c_int32),
                ("acceleration_lookback", C.c_int32), ("acceleration_interval", C.c_int32)]


def build(force: bool = False) -> Path:
    so = _HERE / "libcone_oracle.so"
    src = _HERE / "cone_oracle.c"
    if force or not so.exists() or (src.exists() and so.stat().st_mtime < src.stat().st_mtime):
        subprocess.run(["make", "-C", str(_HERE), "-s", "-B"], check=True, capture_output=True)
    return so


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(str(build()))
        _LIB.orc_solve.restype = C.c_int
        _LIB.orc_vjp.restype = C.c_int
        _LIB.orc_lsqr_dense.restype = C.c_int
        _LIB.orc_max_threads.restype = C.c_int
    return _LIB


def _p(a, t=_f64p):
    return None if a is None else a.ctypes.data_as(t)


def make_settings(**kw) -> OrcSettings:
    st = OrcSettings()
    lib().orc_default_settings(C.byref(st))
    if "eps" in kw:  # diffcp maps eps -> eps_abs = eps_rel
        e = kw.pop("eps")
        st.eps_abs = e
        st.eps_rel = e
    for k, v in kw.items():
        if not hasattr(st, k):
            raise KeyError(k)
        setattr(st, k, v)
    return st


class _Desc:
    def __init__(self, st: Structure):
        self.keep = []
        d = OrcDesc()
        d.n, d.m, d.nnzA, d.nnzP = st.n, st.m, st.nnzA, st.nnzP
        d.A_indptr, d.A_indices = _p(st.A_indptr, _i32p), _p(st.A_indices, _i32p)
        if st.P_indptr is not None:
            d.P_indptr, d.P_indices = _p(st.P_indptr, _i32p), _p(st.P_indices, _i32p)
        q = np.asarray(st.cones.q, dtype=np.int32)
        s = np.asarray(st.cones.s, dtype=np.int32)
        self.keep += [q, s, st]
        d.z, d.l, d.nq, d.ns, d.ep, d.ed = st.cones.z, st.cones.l, q.size, s.size, st.cones.ep, st.cones.ed
        d.q, d.s = _p(q, _i32p), _p(s, _i32p)
        self.d = d


def _c(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def solve_batch(st: Structure, A_vals, b, c, P_vals=None, nthreads: int = 0, warm=None, **settings):
    """-> x[B,n], y[B,m], s[B,m], status[B], iters[B]   (warm = (x0, y0, s0): start from a previous solution)"""
    A_vals, b, c, P_vals = _c(A_vals), _c(b), _c(c), _c(P_vals)
    if warm is not None:
        x0, y0, s0 = map(_c, warm)
        B = A_vals.shape[0]
        x = np.empty((B, st.n)); y = np.empty((B, st.m)); s = np.empty((B, st.m))
        status = np.zeros(B, dtype=np.int32); iters = np.zeros(B, dtype=np.int32)
        D = _Desc(st); S = make_settings(**settings)
        lib().orc_solve_batch_warm(C.byref(D.d), C.c_int32(B), _p(A_vals), _p(P_vals), _p(b), _p(c), _p(x0), _p(y0), _p(s0), _p(x), _p(y), _p(s),
                                   _p(status, _i32p), _p(iters, _i32p), C.byref(S), C.c_int32(nthreads))
        return x, y, s, status, iters
    B = A_vals.shape[0]
    x = np.empty((B, st.n)); y = np.empty((B, st.m)); s = np.empty((B, st.m))
    status = np.zeros(B, dtype=np.int32); iters = np.zeros(B, dtype=np.int32)
    D = _Desc(st); S = make_settings(**settings)
    lib().orc_solve_batch(C.byref(D.d), C.c_int32(B), _p(A_vals), _p(P_vals), _p(b), _p(c), _p(x), _p(y), _p(s),
                          _p(status, _i32p), _p(iters, _i32p), C.byref(S), C.c_int32(nthreads))
    return x, y, s, status, iters


def vjp_batch(st: Structure, A_vals, b, c, x, y, s, dx, dy, P_vals=None, nthreads: int = 0, **settings):
    """-> dA[B,nnzA], dP[B,nnzP]|None, db[B,m], dc[B,n], lsqr_iters[B]"""
    A_vals, b, c, P_vals, x, y, s, dx, dy = map(_c, (A_vals, b, c, P_vals, x, y, s, dx, dy))
    B = A_vals.shape[0]
    dA = np.empty((B, st.nnzA)); db = np.empty((B, st.m)); dc = np.empty((B, st.n))
    dP = np.empty((B, st.nnzP)) if P_vals is not None else None
    its = np.zeros(B, dtype=np.int32)
    D = _Desc(st); S = make_settings(**settings)
    lib().orc_vjp_batch(C.byref(D.d), C.c_int32(B), _p(A_vals), _p(P_vals), _p(b), _p(c), _p(x), _p(y), _p(s),
                        _p(dx), _p(dy), _p(dA), _p(dP), _p(db), _p(dc), _p(its, _i32p), C.byref(S), C.c_int32(nthreads))
    return dA, dP, db, dc, its


def proj_dual_cone(st: Structure, v):
    v = np.array(v, dtype=np.float64, copy=True)
    D = _Desc(st)
    lib().orc_proj_dual_cone(C.byref(D.d), _p(v))
    return v


def dproj_dual_cone(st: Structure, v, dv):
    v, dv = _c(v), _c(dv)
    out = np.empty_like(v)
    D = _Desc(st)
    lib().orc_dproj_dual_cone(C.byref(D.d), _p(v), _p(dv), _p(out))
    return out


def lsqr_dense(M, rhs, atol=1e-8, btol=1e-8, conlim=1e8, iter_lim=-1):
    M, rhs = _c(M), _c(rhs)
    sol = np.empty(M.shape[1])
    its = lib().orc_lsqr_dense(C.c_int32(M.shape[0]), C.c_int32(M.shape[1]), _p(M), _p(rhs), _p(sol),
                               C.c_double(atol), C.c_double(btol), C.c_double(conlim), C.c_int32(iter_lim))
    return sol, its


def max_threads() -> int:
    return int(lib().orc_max_threads())
