"""ORACLE (test infrastructure, NOT product code) -- ctypes wrapper for
oracle/_ref/libfmpc_ref.so (built from oracle/fmpc_ref.c by oracle/Makefile).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
may import this."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libfmpc_ref.so")
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


class _Sys(C.Structure):
    _fields_ = [("n", C.c_int), ("m", C.c_int), ("T", C.c_int),
                ("A1", _dp), ("A2", _dp), ("B", _dp), ("Q", _dp), ("R", _dp), ("Qf", _dp),
                ("q", _dp), ("r", _dp), ("qf", _dp), ("u_min", _dp), ("u_max", _dp)]


_SOG = os.path.join(_HERE, "_ref", "libfmpc_ref_general.so")


def build(force: bool = False) -> str:
    stale = lambda so, src: not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(os.path.join(_HERE, src))
    if force or stale(_SO, "fmpc_ref.c") or stale(_SOG, "fmpc_ref_general.c"):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.fref_solve_batch.restype = C.c_int
        _lib.fref_max_threads.restype = C.c_int
    return _lib


def _f(a):  # column-major contiguous double, or None
    return None if a is None else np.asfortranarray(np.asarray(a, dtype=np.float64))


def _p(a):
    return None if a is None else a.ctypes.data_as(_dp)


def max_threads() -> int:
    return int(lib().fref_max_threads())


def solve_batch(A1, A2, B, Q, R, Qf, u_min, u_max, kappa, niters, x0, x0_pre, w, z0, nu0, xf=None,
                q=None, r=None, qf=None, ls_max=0, alpha=1e-4, beta=0.5, tol_r=1e-6, tol_p=1e-8,
                nthreads=0):
    """Per-instance arrays are 2-D with one COLUMN per instance (x0: n x nb, z0: N x nb, ...).
    Returns dict(z, nu, status, iters, halvings)."""
    A1, A2, B, Q, R, Qf = map(_f, (A1, A2, B, Q, R, Qf))
    n, m = B.shape
    z0 = _f(np.atleast_2d(np.asarray(z0, dtype=np.float64).T).T if np.ndim(z0) == 1 else z0)
    nb = z0.shape[1]
    T = z0.shape[0] // (n + m)
    keep = [A1, A2, B, Q, R, Qf]
    vec = [_f(v) for v in (q, r, qf, u_min, u_max)]
    s = _Sys(n, m, T, _p(A1), _p(A2), _p(B), _p(Q), _p(R), _p(Qf), *[_p(v) for v in vec])

    def inst(a):
        if a is None:
            return None
        a = np.asarray(a, dtype=np.float64)
        if a.ndim == 1:
            a = a[:, None]
        assert a.shape[1] == nb, "per-instance arrays need one column per instance"
        return _f(a)

    x0, x0_pre, w, xf, nu0 = map(inst, (x0, x0_pre, w, xf, nu0))
    NB = T + (1 if xf is not None else 0)
    assert nu0.shape[0] == NB * n
    z = np.zeros_like(z0, order="F")
    nu = np.zeros_like(nu0, order="F")
    status = np.zeros(nb, dtype=np.int32)
    iters = np.zeros(nb, dtype=np.int32)
    halv = np.zeros(nb, dtype=np.int32)
    rc = lib().fref_solve_batch(C.byref(s), C.c_double(kappa), C.c_int(niters), C.c_int(ls_max),
                                C.c_double(alpha), C.c_double(beta), C.c_double(tol_r), C.c_double(tol_p),
                                C.c_int(nb), _p(x0), _p(x0_pre), _p(w), _p(xf), _p(z0), _p(nu0),
                                _p(z), _p(nu), status.ctypes.data_as(_ip), iters.ctypes.data_as(_ip),
                                halv.ctypes.data_as(_ip), C.c_int(nthreads))
    if rc != 0:
        raise RuntimeError(f"fref_solve_batch failed: {rc}")
    del keep
    return dict(z=z, nu=nu, status=status, iters=iters, halvings=halv)


# ---- second structured restatement: ramp rows, literal VAR_1 C, dense Q (oracle/fmpc_ref_general.c) ----
class _SysG(C.Structure):
    _fields_ = [("n", C.c_int), ("m", C.c_int), ("T", C.c_int), ("var_order", C.c_int), ("ramp_rows", C.c_int),
                ("literal_bug", C.c_int), ("A1", _dp), ("A2", _dp), ("B", _dp), ("Q", _dp), ("R", _dp), ("Qf", _dp),
                ("q", _dp), ("r", _dp), ("qf", _dp), ("u_min", _dp), ("u_max", _dp), ("du_min", _dp), ("du_max", _dp)]


_libg = None


def solve_batch_general(A1, A2, B, Q, R, Qf, u_min, u_max, kappa, niters, x0, x0_pre, w, z0, nu0, xf=None, u_prev=None,
                        du_min=None, du_max=None, ramp_rows=False, literal_bug=False, ls_max=0, alpha=1e-4, beta=0.5,
                        tol_r=1e-6, tol_p=1e-8, nthreads=0):
    """Same conventions as solve_batch (one COLUMN per instance); VAR(1) when A2 is None."""
    global _libg
    if _libg is None:
        build()
        _libg = C.CDLL(_SOG)
        _libg.frefg_solve_batch.restype = C.c_int
    A1, A2, B, Q, R, Qf = map(_f, (A1, A2, B, Q, R, Qf))
    n, m = B.shape
    z0 = _f(z0)
    nb = z0.shape[1]
    T = z0.shape[0] // (n + m)
    vec = [_f(v) for v in (u_min, u_max, du_min, du_max)]
    s = _SysG(n, m, T, 1 if A2 is None else 2, int(bool(ramp_rows)), int(bool(literal_bug)), _p(A1), _p(A2), _p(B), _p(Q),
              _p(R), _p(Qf), None, None, None, *[_p(v) for v in vec])

    def inst(a):
        if a is None:
            return None
        a = np.asarray(a, dtype=np.float64)
        if a.ndim == 1:
            a = a[:, None]
        assert a.shape[1] == nb
        return _f(a)

    x0, x0_pre, w, xf, nu0, u_prev = map(inst, (x0, x0_pre, w, xf, nu0, u_prev))
    z = np.zeros_like(z0, order="F")
    status = np.zeros(nb, dtype=np.int32); iters = np.zeros(nb, dtype=np.int32); halv = np.zeros(nb, dtype=np.int32)
    rc = _libg.frefg_solve_batch(C.byref(s), C.c_double(kappa), C.c_int(niters), C.c_int(ls_max), C.c_double(alpha),
                                 C.c_double(beta), C.c_double(tol_r), C.c_double(tol_p), C.c_int(nb), _p(x0), _p(x0_pre),
                                 _p(u_prev), _p(w), _p(xf), _p(z0), _p(nu0), _p(z), status.ctypes.data_as(_ip),
                                 iters.ctypes.data_as(_ip), halv.ctypes.data_as(_ip), C.c_int(nthreads))
    if rc != 0:
        raise RuntimeError(f"frefg_solve_batch failed: {rc}")
    return dict(z=z, status=status, iters=iters, halvings=halv)
