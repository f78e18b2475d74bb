"""Host-side mirror of the reference's `Fast_MPC2` value class for the fastMPC hot path.

  Fast_MPC2       <-> Fast_MPC/VAR_2/Fast_MPC2.m  (23-argument ctor :28-55, solve front-ends :88-144)
  Fast_MPC2_VAR1  <-> Fast_MPC/VAR_1/Fast_MPC2.m  (21-argument ctor :26-51)
  FastMPCBatch    :   the batched C-ABI (include/fmpc.h) -- many independent instances per call

Same argument names, meaning and error strings as the reference; `None` plays MATLAB `[]`.
All compute happens in the CUDA library; nothing here falls back to the CPU.

Array conventions of the batched API (numpy, C-contiguous, instance first -- the same memory as
the C-ABI's column-major "one column per instance"):
    x0, x0_pre, xf : (nb, n)      w : (nb, T*n)      nu0 : (nb, T*n [+ n if xf])
    X0, X : (nb, T, n)            U0, U : (nb, T, m)
"""
from __future__ import annotations

import ctypes as C
import hashlib
from collections import OrderedDict

import numpy as np

from . import _lib
from ._lib import FmpcParams, FmpcSys, check, dp, load_library

FE_FIXED_LOG, FE_FIXED_NEWTON, FE_SOLVE_FULL, FE_SOLVE_CHECK = 1, 2, 3, 4

# MATLAB's global default stream (MT19937, seed 5489): shared by every Fast_MPC2 object of the
# session, one rand(length(b),1) per inf_newton_solver call (inf_newton_solver.m:2).
_matlab_stream = np.random.RandomState(5489)


def reset_matlab_stream(seed: int = 5489):
    global _matlab_stream
    _matlab_stream = np.random.RandomState(seed)


def _isempty(a) -> bool:
    return a is None or (hasattr(a, "__len__") and len(a) == 0)


def _colmajor(a):
    return np.asfortranarray(np.atleast_2d(np.asarray(a, dtype=np.float64)))


def _vec(a):
    return None if _isempty(a) else np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(-1))


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def deinterleave(z, n, m, T):
    """README.md:558-570: z -> (U (T, m), X (T, n))."""
    Z = np.asarray(z, dtype=np.float64).reshape(T, n + m)
    return Z[:, :m].copy(), Z[:, m:].copy()


def interleave(U, X):
    """(U (T, m), X (T, n)) -> z in the layout of fast_mpc_init.m:22-25."""
    return np.hstack([np.asarray(U), np.asarray(X)]).reshape(-1)


class FastMPCBatch:
    """One fmpc_handle: shared problem data + workspaces for up to `max_batch` instances."""

    def __init__(self, A1, A2, B, Q, R, Qf, u_min, u_max, T, x_min=None, x_max=None, q=None, r=None, qf=None,
                 du_min=None, du_max=None, ramp_rows=False, max_batch=1, device=0, var1_literal_bug=False):
        L = load_library()
        self._L = L
        self.var_order = 1 if _isempty(A2) else 2
        self._A1 = _colmajor(A1)
        self._A2 = None if self.var_order == 1 else _colmajor(A2)
        self._B = _colmajor(B)
        self._Q, self._R, self._Qf = _colmajor(Q), _colmajor(R), _colmajor(Qf)
        self.n, self.m = self._B.shape
        self.T = int(T)
        n, m = self.n, self.m
        # the C side cannot see array sizes: the reference's dimension checks, with its error strings
        if self._Q.shape != (n, n) or self._Qf.shape != (n, n):
            raise ValueError("State stage cost must a square matrix")                       # fast_mpc_objective.m:17-19
        if self._R.shape != (m, m):
            raise ValueError("Control stage cost must a square matrix")                     # :20-21
        if self._A1.shape != (n, n) or (self._A2 is not None and self._A2.shape != (n, n)):
            raise ValueError("The equality state dynamics matrix size does not match")      # fast_mpc_eq_const.m:27-30
        for vec, size, msg in ((q, n, "Linear state cost needs to be a vector of size n"),
                               (r, m, "Linear control cost needs to be a vector of size n"),
                               (qf, n, "State terminal linear cost needs to be a vector of size n"),
                               (x_min, n, "Check the state inequality constraints dimensions"),
                               (x_max, n, "Check the state inequality constraints dimensions"),
                               (u_min, m, "Check cotrol iequality constraint dimension"),
                               (u_max, m, "Check cotrol iequality constraint dimension"),
                               (du_min, m, "Check cotrol iequality constraint dimension"),
                               (du_max, m, "Check cotrol iequality constraint dimension")):
            if not _isempty(vec) and np.asarray(vec).size < size:
                raise ValueError(msg)
        if _isempty(u_min) or _isempty(u_max):
            raise ValueError("Check cotrol iequality constraint dimension")
        if self.T < 1:
            raise ValueError("horizon T must be >= 1")
        self._xmin = _vec(x_min) if not _isempty(x_min) else -np.ones(n)
        self._xmax = _vec(x_max) if not _isempty(x_max) else np.ones(n)
        self._umin, self._umax = _vec(u_min), _vec(u_max)
        self._q, self._r, self._qf = _vec(q), _vec(r), _vec(qf)
        self._dumin, self._dumax = _vec(du_min), _vec(du_max)
        s = FmpcSys()
        s.n, s.m, s.T, s.var_order = n, m, self.T, self.var_order
        P = lambda a: None if a is None else a.ctypes.data_as(dp)
        s.A1, s.A2, s.B = P(self._A1), P(self._A2), P(self._B)
        s.Q, s.R, s.Qf = P(self._Q), P(self._R), P(self._Qf)
        s.q, s.r, s.qf = P(self._q), P(self._r), P(self._qf)
        s.x_min, s.x_max, s.u_min, s.u_max = P(self._xmin), P(self._xmax), P(self._umin), P(self._umax)
        s.du_min, s.du_max = P(self._dumin), P(self._dumax)
        s.ramp_rows = 1 if ramp_rows else 0
        s.var1_literal_bug = 1 if (var1_literal_bug and self.var_order == 1) else 0
        self.ramp_rows = bool(ramp_rows)
        self.max_batch = int(max_batch)
        self.device = int(device)
        self._create(s)

    # the three C entry points a subclass (FastMPCMulti) swaps for their multi-device versions
    def _create(self, s):
        h = C.c_void_p()
        check(self._L.fmpc_create(C.byref(h), C.byref(s), self.max_batch, self.device))
        self._h = h
        self._step_fn, self._step_r_fn = self._L.fmpc_step, self._L.fmpc_step_r

    def close(self):
        if getattr(self, "_h", None):
            self._L.fmpc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- helpers -------------------------------------------------------------------------
    def _inst(self, a, width, name, nb):
        if a is None:
            return None
        a = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
        if a.ndim == 1:
            a = a.reshape(1, -1)
        a = a.reshape(a.shape[0], -1) if a.size else a.reshape(a.shape[0], width)
        if a.shape != (nb, width):
            raise ValueError(f"{name}: expected shape ({nb}, {width}), got {a.shape}")
        return a

    @staticmethod
    def params(kappa=0.01, niters=5, ls_max=0, **kw) -> FmpcParams:
        return FmpcParams.default(kappa=kappa, niters=niters, ls_max=ls_max, **kw)

    @property
    def launch_count(self) -> int:
        return int(self._L.fmpc_launch_count(self._h))

    @property
    def workspace_bytes(self) -> int:
        return int(self._L.fmpc_workspace_bytes(self._h))

    def last_newton_iters(self) -> int:
        return int(self._L.fmpc_last_newton_iters(self._h))

    @property
    def kernel_kind(self) -> int:
        """2 = warp-per-instance DMMA kernel (n <= 32), 1 = CTA DMMA kernel, 0 = generic kernel."""
        return int(self._L.fmpc_kernel_kind(self._h))

    def last_profile(self):
        """Phase cycle counters of the last launch (zeros unless built with -DFMPC_PROF)."""
        out = np.zeros(12, dtype=np.int64)
        check(self._L.fmpc_last_profile(self._h, out.ctypes.data_as(C.c_void_p)))
        return out

    # ---- batched solve -------------------------------------------------------------------
    def step(self, x0, x0_pre=None, w=None, xf=None, X0=None, U0=None, nu0=None, u_prev=None, params=None,
             kappa=0.01, niters=5, ls_max=0, frontend=None, k_min=None, k_max=None):
        """Batched `mpc_fixed_log_newton(niters, kappa)`; or, with `frontend`, one of the
        kappa-continuation front-ends.  Returns dict(X, U, status, iters, telapsed)."""
        n, m, T = self.n, self.m, self.T
        x0 = np.ascontiguousarray(np.asarray(x0, dtype=np.float64))
        if x0.ndim == 1:
            x0 = x0.reshape(1, -1)
        nb = x0.shape[0]
        if x0.shape[1] != n:
            raise ValueError("The equality state dynamics matrix size does not match")
        x0_pre = self._inst(x0_pre, n, "x0_pre", nb)
        if self.var_order == 2 and x0_pre is None:
            raise ValueError("The equality state dynamics matrix size does not match")
        w = self._inst(w, T * n, "w", nb)
        xf = self._inst(xf, n, "xf", nb)
        u_prev = self._inst(u_prev, m, "u_prev", nb)
        NBn = (T + (1 if xf is not None else 0)) * n
        if (X0 is None) != (U0 is None):
            raise ValueError("Initialization size mismatch (T*(n+m))")
        if X0 is not None:
            X0 = self._inst(X0, T * n, "X0", nb)
            U0 = self._inst(U0, T * m, "U0", nb)
        p = params if params is not None else self.params(kappa, niters, ls_max)
        X = np.empty((nb, T, n))
        U = np.empty((nb, T, m))
        status = np.zeros(nb, dtype=np.int32)
        iters = np.zeros(nb, dtype=np.int32)
        tel = C.c_double(0.0)
        if frontend is None:
            nu0 = self._inst(nu0, NBn, "nu0", nb)
            check(self._step_fn(self._h, C.byref(p), nb, _ptr(x0), _ptr(x0_pre), _ptr(u_prev), _ptr(w), _ptr(xf),
                                    _ptr(X0), _ptr(U0), _ptr(nu0), _ptr(X), _ptr(U), _ptr(status), _ptr(iters),
                                    C.cast(C.byref(tel), C.c_void_p)))
        else:
            nouter = self._L.fmpc_frontend_nouter(self._h, frontend)
            if nu0 is not None:
                nu0 = np.ascontiguousarray(np.asarray(nu0, dtype=np.float64))
                if nu0.shape != (nouter, nb, NBn):
                    raise ValueError(f"nu0: expected shape ({nouter}, {nb}, {NBn}), got {nu0.shape}")
            check(self._L.fmpc_frontend(self._h, frontend, C.byref(p), float(k_min or 0.0), float(k_max or 0.0), nb,
                                        _ptr(x0), _ptr(x0_pre), _ptr(u_prev), _ptr(w), _ptr(xf), _ptr(X0), _ptr(U0),
                                        _ptr(nu0), _ptr(X), _ptr(U), _ptr(status), _ptr(iters),
                                        C.cast(C.byref(tel), C.c_void_p)))
        return dict(X=X, U=U, status=status, iters=iters, telapsed=tel.value)

    def step_resident(self, x0, x0_pre=None, w=None, xf=None, nu0=None, u_prev=None, reset=False, full=False, params=None,
                      kappa=0.01, niters=5, ls_max=0):
        """One step of a closed loop whose solver state stays on the device (`fmpc_step_r`): warm start = the previous
        call's solution shifted one stage, x0_pre defaults to the previous x0, u_prev to the previous U(:,0).
        Only x0 goes in and U(:,0) comes out (README.md:589) unless `full`.  `reset=True` starts a loop (cold start).
        Returns dict(u0 (nb, m), status, iters, telapsed[, X, U])."""
        n, m, T = self.n, self.m, self.T
        x0 = np.ascontiguousarray(np.asarray(x0, dtype=np.float64))
        if x0.ndim == 1:
            x0 = x0.reshape(1, -1)
        nb = x0.shape[0]
        if x0.shape[1] != n:
            raise ValueError("The equality state dynamics matrix size does not match")
        x0_pre = self._inst(x0_pre, n, "x0_pre", nb)
        w = self._inst(w, T * n, "w", nb)
        xf = self._inst(xf, n, "xf", nb)
        u_prev = self._inst(u_prev, m, "u_prev", nb)
        nu0 = self._inst(nu0, (T + (1 if xf is not None else 0)) * n, "nu0", nb)
        p = params if params is not None else self.params(kappa, niters, ls_max)
        u0 = np.empty((nb, m))
        X = np.empty((nb, T, n)) if full else None
        U = np.empty((nb, T, m)) if full else None
        status = np.zeros(nb, dtype=np.int32)
        iters = np.zeros(nb, dtype=np.int32)
        tel = C.c_double(0.0)
        check(self._step_r_fn(self._h, C.byref(p), nb, 1 if reset else 0, _ptr(x0), _ptr(x0_pre), _ptr(u_prev), _ptr(w),
                                  _ptr(xf), _ptr(nu0), _ptr(u0), _ptr(X), _ptr(U), _ptr(status), _ptr(iters),
                                  C.cast(C.byref(tel), C.c_void_p)))
        out = dict(u0=u0, status=status, iters=iters, telapsed=tel.value)
        if full:
            out["X"], out["U"] = X, U
        return out

    def state_update(self, x, x_pre, u, w=None):
        """x+ = A1 x + A2 x- + B u (+ w) per instance (VAR_2/fast_mpc_eq_const.m:39-47 as a recurrence)."""
        x = np.ascontiguousarray(np.asarray(x, dtype=np.float64))
        if x.ndim == 1:
            x = x.reshape(1, -1)
        nb = x.shape[0]
        x_pre = self._inst(x_pre, self.n, "x_pre", nb)
        u = self._inst(u, self.m, "u", nb)
        w = self._inst(w, self.n, "w", nb)
        out = np.empty((nb, self.n))
        check(self._L.fmpc_state_update(self._h, nb, _ptr(x), _ptr(x_pre), _ptr(u), _ptr(w), _ptr(out)))
        return out

    def closed_loop(self, a, nu0=None, params=None, kappa=0.01, niters=5, ls_max=0):
        """K closed-loop steps on the device.  a: (nb, K, n) open-loop aberration sequence.
        Returns dict(U_acc (nb, K, m), X_acc (nb, K, n), iters (nb, K), telapsed)."""
        a = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
        nb, K, n = a.shape
        assert n == self.n
        if nu0 is not None:
            nu0 = np.ascontiguousarray(np.asarray(nu0, dtype=np.float64))
            if nu0.shape != (K, nb, self.T * n):
                raise ValueError(f"nu0: expected shape ({K}, {nb}, {self.T * n})")
        p = params if params is not None else self.params(kappa, niters, ls_max)
        U_acc = np.empty((nb, K, self.m))
        X_acc = np.empty((nb, K, n))
        it = np.zeros((nb, K), dtype=np.int32)
        tel = C.c_double(0.0)
        check(self._L.fmpc_closed_loop(self._h, C.byref(p), nb, K, _ptr(a), _ptr(nu0), _ptr(U_acc), _ptr(X_acc),
                                       _ptr(it), C.cast(C.byref(tel), C.c_void_p)))
        return dict(U_acc=U_acc, X_acc=X_acc, iters=it, telapsed=tel.value)


class FastMPCMulti(FastMPCBatch):
    """Every B200 of the box behind one blocking call (`fmpc_multi_*`): one handle + host thread per device in this process,
    contiguous shards of ceil(nb / G) instances, no inter-GPU traffic in the solve.  Same `step` / `step_resident` as
    FastMPCBatch; `stats()` returns the per-device records of the last step (over NCCL if `use_nccl`)."""

    def __init__(self, *args, ngpus=0, devices=None, **kw):
        self._ngpus_req = int(ngpus)
        self._devices = None if devices is None else (C.c_int * len(devices))(*devices)
        if devices is not None:
            self._ngpus_req = len(devices)
        super().__init__(*args, **kw)

    def _create(self, s):
        h = C.c_void_p()
        check(self._L.fmpc_multi_create(C.byref(h), C.byref(s), self.max_batch, self._ngpus_req,
                                        None if self._devices is None else C.cast(self._devices, C.c_void_p)))
        self._h = h
        self._step_fn, self._step_r_fn = self._L.fmpc_multi_step, self._L.fmpc_multi_step_r
        self.ngpus = int(self._L.fmpc_multi_ngpus(h))

    def close(self):
        if getattr(self, "_h", None):
            self._L.fmpc_multi_destroy(self._h)
            self._h = None

    def _dev_handles(self):
        return [C.c_void_p(self._L.fmpc_multi_handle(self._h, g)) for g in range(self.ngpus)]

    @property
    def launch_count(self) -> int:
        return sum(int(self._L.fmpc_launch_count(h)) for h in self._dev_handles())

    @property
    def workspace_bytes(self) -> int:
        return sum(int(self._L.fmpc_workspace_bytes(h)) for h in self._dev_handles())

    @property
    def kernel_kind(self) -> int:
        return int(self._L.fmpc_kernel_kind(self._dev_handles()[0]))

    def last_newton_iters(self) -> int:
        return sum(int(self._L.fmpc_last_newton_iters(h)) for h in self._dev_handles())

    def shard(self, nbatch, g):
        a, b = C.c_int(0), C.c_int(0)
        check(self._L.fmpc_multi_shard(self._h, int(nbatch), int(g), C.cast(C.byref(a), C.c_void_p), C.cast(C.byref(b), C.c_void_p)))
        return a.value, b.value

    def stats(self, use_nccl=False):
        """Per-device statistics of the last step: list of dict(device, n_solves, device_seconds, newton_iters, status_hist)."""
        rec = np.zeros((self.ngpus, 8))
        used = C.c_int(0)
        nrec = self._L.fmpc_multi_last_stats(self._h, rec.ctypes.data_as(C.c_void_p), 1 if use_nccl else 0,
                                             C.cast(C.byref(used), C.c_void_p))
        if nrec < 0:
            check(nrec)
        return [dict(device=int(r[0]), n_solves=int(r[1]), device_seconds=float(r[2]), newton_iters=int(r[3]),
                     status_hist=[int(v) for v in r[4:8]], via_nccl=bool(used.value)) for r in rec[:nrec]]

    def state_update(self, *a, **k):
        raise NotImplementedError("use a FastMPCBatch for the host-buffer state update; the multi-device handle only solves")

    def closed_loop(self, *a, **k):
        raise NotImplementedError("fmpc_closed_loop is a single-device entry point; drive the loop with step_resident")


# --------------------------------------------------------------------------------------------
# handle cache: the reference constructs a new Fast_MPC2 object every control step
# (README.md:548); the expensive part here (upload + precompute) depends only on the
# problem-constant arguments, so handles are reused across objects.
# --------------------------------------------------------------------------------------------
_handle_cache: "OrderedDict[str, FastMPCBatch]" = OrderedDict()
_HANDLE_CACHE_MAX = 8


def _cached_handle(key_arrays, ctor):
    hsh = hashlib.sha1()
    for a in key_arrays:
        if a is None:
            hsh.update(b"-")
        else:
            aa = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
            hsh.update(str(aa.shape).encode())
            hsh.update(aa.tobytes())
    key = hsh.hexdigest()
    if key in _handle_cache:
        _handle_cache.move_to_end(key)
        return _handle_cache[key]
    h = ctor()
    _handle_cache[key] = h
    while len(_handle_cache) > _HANDLE_CACHE_MAX:
        _, old = _handle_cache.popitem(last=False)
        old.close()
    return h


class Fast_MPC2:
    """VAR_2/Fast_MPC2.m -- same 23 constructor arguments, same method names; every solve is one
    call into the CUDA library with nbatch = 1."""

    var_order = 2
    device = 0

    def __init__(self, Q, R, S, Qf, q, r, qf, xmin, xmax, umin, umax, dumin, dumax, T, x0, x0_pre, u_prev,
                 A1, A2, B, w, xf, x_init):
        m2 = lambda a: None if _isempty(a) else np.atleast_2d(np.asarray(a, dtype=np.float64))
        self.Q, self.R, self.S, self.Qf = m2(Q), m2(R), S, m2(Qf)
        self.q, self.r, self.qf = _vec(q), _vec(r), _vec(qf)
        self.x_min, self.x_max, self.u_min, self.u_max = _vec(xmin), _vec(xmax), _vec(umin), _vec(umax)
        self.du_min, self.du_max = _vec(dumin), _vec(dumax)
        self.T = int(T)
        self.x0, self.x0_pre, self.u_prev = _vec(x0), _vec(x0_pre), _vec(u_prev)
        self.A1, self.A2, self.B = m2(A1), m2(A2), m2(B)
        self.w, self.x_final, self.x_init = _vec(w), _vec(xf), _vec(x_init)
        self.last = None

    # ---- the reference's error() checks (same strings) --------------------------------------
    def _validate(self):
        Q, R, Qf = self.Q, self.R, self.Qf
        if Q is None or Qf is None or Q.shape[0] != Q.shape[1] or Qf.shape[0] != Qf.shape[1]:
            raise ValueError("State stage cost must a square matrix")                       # fast_mpc_objective.m:17-19
        if R is None or R.shape[0] != R.shape[1]:
            raise ValueError("Control stage cost must a square matrix")                     # :20-21
        n, m = Q.shape[0], R.shape[0]
        if self.q is not None and self.q.shape[0] != n:
            raise ValueError("Linear state cost needs to be a vector of size n")            # :26-29
        if self.r is not None and self.r.shape[0] != m:
            raise ValueError("Linear control cost needs to be a vector of size n")          # :34-37
        if self.qf is not None and self.qf.shape[0] != n:
            raise ValueError("State terminal linear cost needs to be a vector of size n")   # :41-44
        if self.x_min is None or self.x_max is None or self.x_min.shape[0] != n or self.x_max.shape[0] != n:
            raise ValueError("Check the state inequality constraints dimensions")           # fast_mpc_ineq_const.m:4-6
        if self.u_min is None or self.u_max is None or self.u_min.shape[0] != m or self.u_max.shape[0] != m:
            raise ValueError("Check cotrol iequality constraint dimension")                 # :7-9
        if self.A1 is None or (self.var_order == 2 and self.A2 is None):
            raise ValueError("Define the state dynamics/equality constrained matrix")       # fast_mpc_eq_const.m:19-22
        if self.B is None:
            raise ValueError("Define the control dynamics/equality constrained matrix")     # :23-24
        if self.x0 is None or self.A1.shape[1] != self.x0.shape[0]:
            raise ValueError("The equality state dynamics matrix size does not match")      # :27-28
        if self.var_order == 2 and (self.x0_pre is None or self.A2.shape[1] != self.x0_pre.shape[0]):
            raise ValueError("The equality state dynamics matrix size does not match")      # :29-30
        if self.B.shape[1] != R.shape[1]:
            raise ValueError("The equality control dynamics matrix size does not match")    # :31-32
        if self.x_init is not None and self.x_init.shape[0] != self.T * (n + m):
            raise ValueError("Initialization size mismatch (T*(n+m))")                      # fast_mpc_init.m:13-14
        w = self.w
        if w is None:
            if self.T > 1:      # MATLAB: w = zeros(n,1) then w(n*i+1:...) -> index error for T > 1
                raise IndexError("Index exceeds the number of array elements (w)")
        elif w.shape[0] < self.T * n:
            raise IndexError("Index exceeds the number of array elements (w)")
        return n, m

    _ramp_rows = False
    _literal_bug = False

    def _handle(self) -> FastMPCBatch:
        keys = [self.A1, self.A2, self.B, self.Q, self.R, self.Qf, self.q, self.r, self.qf, self.x_min, self.x_max,
                self.u_min, self.u_max, np.array([self.T, self.var_order, self.device, int(self._ramp_rows),
                                                  int(self._literal_bug)], dtype=np.float64)]
        if self._ramp_rows:
            keys += [self.du_min, self.du_max]
        return _cached_handle(keys, lambda: FastMPCBatch(
            self.A1, self.A2 if self.var_order == 2 else None, self.B, self.Q, self.R, self.Qf, self.u_min, self.u_max,
            self.T, self.x_min, self.x_max, self.q, self.r, self.qf, self.du_min, self.du_max, self._ramp_rows,
            max_batch=1, device=self.device, var1_literal_bug=self._literal_bug))

    def _run(self, nw, k, nu0, frontend=None, k_min=None, k_max=None):
        n, m = self._validate()
        hb = self._handle()
        T = self.T
        NBn = (T + (0 if self.x_final is None else 1)) * n
        nouter = 1 if frontend is None else hb._L.fmpc_frontend_nouter(hb._h, frontend)
        if nu0 is None:     # rand(length(b),1) from the session stream, once per inf_newton_solver call
            nu = _matlab_stream.random_sample(nouter * NBn).reshape(nouter, 1, NBn)
        else:
            nu = np.asarray(nu0, dtype=np.float64).reshape(nouter, 1, NBn)
        X0 = U0 = None
        if self.x_init is not None:
            U0, X0 = deinterleave(self.x_init, n, m, T)
            U0, X0 = U0[None], X0[None]
        w = None if self.w is None else self.w[:T * n]
        out = hb.step(self.x0, self.x0_pre, w, self.x_final, X0, U0, nu[0] if frontend is None else nu,
                      u_prev=self.u_prev, kappa=k, niters=(1000 if nw is None else int(nw)), frontend=frontend,
                      k_min=k_min, k_max=k_max)
        self.last = out
        return interleave(out["U"][0], out["X"][0])

    # ---- the reference's solve front-ends (VAR_2/Fast_MPC2.m:88-144) ------------------------
    def mpc_fixed_log_newton(self, nw, k, nu0=None):
        """:124-130 -- fixed barrier k, nw Newton steps. Returns x_opt (interleaved z)."""
        return self._run(nw, k, nu0)

    def mpc_fixed_log(self, k, nu0=None):
        """:116-123 -- fixed barrier k, up to 1000 Newton steps."""
        return self._run(None, k, nu0, frontend=FE_FIXED_LOG)

    def mpc_fixed_newton(self, nw, nu0=None):
        """:131-144 -- k = 1, 0.1, ... while k*length(z) >= 10e-3, nw Newton steps each."""
        return self._run(nw, 1.0, nu0, frontend=FE_FIXED_NEWTON)

    def mpc_solve_full(self, nu0=None):
        """:100-115 -- same kappa schedule, up to 1000 Newton steps each."""
        return self._run(None, 1.0, nu0, frontend=FE_SOLVE_FULL)

    def mpc_solve_check(self, k_min, k_max, nu0=None):
        """:88-99 -- five linearly spaced kappa from k_max down to k_min."""
        return self._run(None, 1.0, nu0, frontend=FE_SOLVE_CHECK, k_min=k_min, k_max=k_max)

    def initialize(self):
        """fast_mpc_init.m:12-26 (host-side, no compute on the hot path)."""
        n, m = self._validate()
        if self.x_init is not None:
            return self.x_init.copy()
        z = np.zeros(self.T * (n + m))
        Z = z.reshape(self.T, n + m)
        Z[:, :m] = (self.u_min + self.u_max) / 2
        Z[:, m:] = (self.x_min + self.x_max) / 2
        return z


class Fast_MPC2_VAR1(Fast_MPC2):
    """VAR_1/Fast_MPC2.m -- 21 constructor arguments (no x0_pre, single A).

    The defaults reproduce the reference's VAR_1 exactly as written: ramp-rate rows enabled
    (VAR_1/fast_mpc_ineq_const.m:58-79) and the second block row of C written at columns n : 3n+m-1
    (VAR_1/fast_mpc_eq_const.m:34-37).  `literal_bug=False` selects the corrected
    placement m+1 : 2(n+m) (= VAR_2's code with A2 = 0); `ramp_rows=False` drops the ramp rows (then,
    with diagonal Q/R and the corrected C, the solve runs on the block-banded DMMA kernels)."""

    var_order = 1

    def __init__(self, Q, R, S, Qf, q, r, qf, xmin, xmax, umin, umax, dumin, dumax, T, x0, u_prev, A, B, w, xf,
                 x_init, ramp_rows=True, literal_bug=True):
        super().__init__(Q, R, S, Qf, q, r, qf, xmin, xmax, umin, umax, dumin, dumax, T, x0, None, u_prev, A, None,
                         B, w, xf, x_init)
        self._ramp_rows = bool(ramp_rows)
        self._literal_bug = bool(literal_bug)

    def _validate(self):
        n, m = super()._validate()
        if self._ramp_rows:
            if self.du_min is None or self.du_max is None or self.du_min.shape[0] != m or self.du_max.shape[0] != m:
                raise ValueError("Check cotrol iequality constraint dimension")
            if self.u_prev is None or self.u_prev.shape[0] != m:
                raise ValueError("Arrays have incompatible sizes for this operation (u_prev + du_max)")
        return n, m
