"""ORACLE (test infrastructure, NOT product code) -- literal dense restatement of the
reference's fastMPC solver, `Fast_MPC/VAR_2/*.m` and `Fast_MPC/VAR_1/*.m`.

PARITY UNPINNED: the reference is MATLAB, holds no golden vectors / assertions /
seeded tests for this path, and neither MATLAB nor Octave
exists in the build container.  This file follows the .m sources line by line (each
function cites the file:line range it restates); it is pinned only by the repo's own
cross-checks in tests/ (one-shot dense KKT solve, structured C restatement in
oracle/fmpc_ref.c, known-answer tests).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
may import this module.  The product path (mpc-sensorlessao_b200/) never does.

Conventions: all arrays float64; vectors are 1-D numpy arrays (MATLAB column vectors);
`None` plays the role of MATLAB `[]`.  Index arithmetic is kept 1-based inside the
builders (variables named like the .m files) and shifted by one at the slicing site so
each line can be compared with its MATLAB original.
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as sla

__all__ = [
    "MatlabRand", "Fast_MPC2", "Fast_MPC2_VAR1", "inf_newton_solver",
    "inf_newton_KKT_H", "backtracking_inf_newton", "deinterleave", "interleave",
    "dense_kkt_newton_step",
]


class MatlabRand:
    """MATLAB's default global stream: MT19937 seed 5489, 53-bit doubles
    (`rand(k,1)` == numpy RandomState(5489).random_sample(k)).
    One instance models one MATLAB session; every `inf_newton_solver` call draws
    `rand(length(b),1)` from it (inf_newton_solver.m:2)."""

    def __init__(self, seed: int = 5489):
        self._rs = np.random.RandomState(seed)

    def rand(self, k: int) -> np.ndarray:
        return self._rs.random_sample(k)


_GLOBAL_STREAM = MatlabRand()


def _isempty(a) -> bool:
    return a is None or (hasattr(a, "__len__") and len(a) == 0)


def _col(a):
    return None if _isempty(a) else np.asarray(a, dtype=np.float64).reshape(-1)


def _mat(a):
    return None if _isempty(a) else np.atleast_2d(np.asarray(a, dtype=np.float64))


# --------------------------------------------------------------------------------------
# L1: problem assembly
# --------------------------------------------------------------------------------------
def fast_mpc_init(obj) -> np.ndarray:
    """fast_mpc_init.m:1-28 (identical in VAR_1 and VAR_2)."""
    T = obj.T
    n = obj.Q.shape[0]
    m = obj.R.shape[0]
    if not _isempty(obj.x_init):                                   # :12
        if obj.x_init.shape[0] != T * (n + m):                     # :13
            raise ValueError("Initialization size mismatch (T*(n+m))")
        return obj.x_init.copy()                                   # :16
    x_init = (obj.x_min + obj.x_max) / 2                           # :19
    u_init = (obj.u_min + obj.u_max) / 2                           # :20
    z_init = np.zeros(T * (m + n))                                 # :21
    for i in range(1, T * (m + n) - (m + n) + 1 + 1, m + n):       # :22  i = 1:(m+n):T(m+n)-(m+n)+1
        z_init[i - 1:i + m - 1] = u_init                           # :23
        z_init[i + m - 1:i + m + n - 1] = x_init                   # :24
    return z_init


def fast_mpc_objective(obj):
    """fast_mpc_objective.m:1-67 (identical in VAR_1 and VAR_2).
    J = z'Hz + g'z  (no 1/2: gradient 2Hz+g)."""
    T = obj.T
    Q, R, Qf = obj.Q, obj.R, obj.Qf
    n = Q.shape[0]
    m = R.shape[0]
    if Q.shape[0] != Q.shape[1] or Qf.shape[0] != Qf.shape[1]:     # :17-19
        raise ValueError("State stage cost must a square matrix")
    if R.shape[0] != R.shape[1]:                                   # :20-21
        raise ValueError("Control stage cost must a square matrix")
    q, r, qf = obj.q, obj.r, obj.qf
    if not _isempty(q):                                            # :26-32
        if q.shape[0] != n:
            raise ValueError("Linear state cost needs to be a vector of size n")
    else:
        q = np.zeros(n)
    if not _isempty(r):                                            # :34-40
        if r.shape[0] != m:
            raise ValueError("Linear control cost needs to be a vector of size n")
    else:
        r = np.zeros(m)
    if not _isempty(qf):                                           # :41-47
        if qf.shape[0] != n:
            raise ValueError("State terminal linear cost needs to be a vector of size n")
    else:
        qf = np.zeros(n)

    N = T * (n + m)
    H = np.zeros((N, N))                                           # :50
    blk = np.block([[Q, np.zeros((n, m))], [np.zeros((m, n)), R]])
    for i in range(m + 1, N - n + 1, n + m):                       # :51  i = m+1:(n+m):N-n
        H[i - 1:i + n + m - 1, i - 1:i + n + m - 1] = blk          # :52
    H[0:m, 0:m] = R                                                # :54
    H[N - n:N, N - n:N] = Qf                                       # :55

    g = np.zeros(N)                                                # :57
    for i in range(m + 1, N + 1, n + m):                           # :58  i = m+1:(n+m):N
        if i == N - n + 1:                                         # :59
            g[N - n:N] = qf                                        # :60
        else:
            g[i - 1:i + n + m - 1] = np.concatenate([q, r])        # :62
    g[0:m] = r                                                     # :65
    return H, g


def _box_rows(obj):
    """Box part shared by both variants: VAR_2/fast_mpc_ineq_const.m:42-56 ==
    VAR_1/fast_mpc_ineq_const.m:42-56."""
    if obj.x_min.shape[0] != obj.Q.shape[0] or obj.x_max.shape[0] != obj.Q.shape[0]:   # :4-6
        raise ValueError("Check the state inequality constraints dimensions")
    if obj.u_min.shape[0] != obj.R.shape[0] or obj.u_max.shape[0] != obj.R.shape[0]:   # :7-9
        raise ValueError("Check cotrol iequality constraint dimension")
    T = obj.T
    n = obj.x_min.shape[0]
    m = obj.u_min.shape[0]
    P_box = np.zeros((2 * T * m, T * (n + m)))                     # :42
    h_box = np.zeros(2 * T * m)                                    # :43
    eye_pm = np.vstack([np.eye(m), -np.eye(m)])
    for i in range(1, 2 * T * m - 2 * m + 1 + 1, 2 * m):           # :46
        if i == 1:
            P_box[i - 1:i + 2 * m - 1, i - 1:i + m - 1] = eye_pm   # :48
        else:
            c0 = (i + 1) // 2 + n * ((i - 1) // (2 * m))           # :50 first column (1-based)
            c1 = (i + 1) // 2 + (m + n * ((i - 1) // (2 * m))) - 1
            P_box[i - 1:i + 2 * m - 1, c0 - 1:c1] = eye_pm
    for i in range(1, 2 * T * m - 2 * m + 1 + 1, 2 * m):           # :54
        h_box[i - 1:2 * m + i - 1] = np.concatenate([obj.u_max, -obj.u_min])   # :55
    return P_box, h_box, T, n, m


def fast_mpc_ineq_const_var2(obj):
    """VAR_2/fast_mpc_ineq_const.m:1-84 -- box rows on u only; the ramp section
    (:61-79) is commented out in the reference, P_ramp = h_ramp = [] (:58-59)."""
    P_box, h_box, *_ = _box_rows(obj)
    return P_box, h_box                                            # :81-82


def fast_mpc_ineq_const_var1(obj):
    """VAR_1/fast_mpc_ineq_const.m:1-82 -- box rows, then ALL ramp rows (:58-79)."""
    P_box, h_box, T, n, m = _box_rows(obj)
    P_ramp = np.zeros((2 * T * m, T * (n + m)))                    # :59
    h_ramp = np.zeros(2 * T * m)                                   # :60
    eye_pm = np.vstack([np.eye(m), -np.eye(m)])
    dblk = np.block([[-np.eye(m), np.zeros((m, n)), np.eye(m)],
                     [np.eye(m), np.zeros((m, n)), -np.eye(m)]])
    for i in range(1, 2 * T * m - 2 * m + 1 + 1, 2 * m):           # :62
        if i == 1:
            P_ramp[i - 1:i + 2 * m - 1, i - 1:i + m - 1] = eye_pm  # :64
        else:
            t = (i - 1) // (2 * m)
            c0 = (m + n) * (t - 1) + 1                             # :66
            c1 = (m + n) * t + m
            P_ramp[i - 1:i + 2 * m - 1, c0 - 1:c1] = dblk
    for i in range(1, 2 * T * m - 2 * m + 1 + 1, 2 * m):           # :70
        if i == 1:
            h_ramp[i - 1:2 * m + i - 1] = np.concatenate(
                [obj.u_prev + obj.du_max, -obj.u_prev - obj.du_min])   # :72
        else:
            h_ramp[i - 1:2 * m + i - 1] = np.concatenate([obj.du_max, -obj.du_min])  # :74
    return np.vstack([P_box, P_ramp]), np.concatenate([h_box, h_ramp])   # :78-79


def fast_mpc_eq_const_var2(obj):
    """VAR_2/fast_mpc_eq_const.m:1-72 -- two-lag dynamics written straight into C."""
    A1, A2, B = obj.A1, obj.A2, obj.B
    if _isempty(A1) or _isempty(A2):                               # :19-22
        raise ValueError("Define the state dynamics/equality constrained matrix")
    if _isempty(B):                                                # :23-24
        raise ValueError("Define the control dynamics/equality constrained matrix")
    n = A1.shape[1]
    m = B.shape[1]
    x0, x0_pre = obj.x0, obj.x0_pre
    T = obj.T
    w = obj.w
    xf = obj.x_final
    if A1.shape[1] != x0.shape[0]:                                 # :27-28
        raise ValueError("The equality state dynamics matrix size does not match")
    if A2.shape[1] != x0_pre.shape[0]:                             # :29-30
        raise ValueError("The equality state dynamics matrix size does not match")
    if B.shape[1] != obj.R.shape[1]:                               # :31-32
        raise ValueError("The equality control dynamics matrix size does not match")
    if _isempty(w):                                                # :33-34 (only valid for T == 1)
        w = np.zeros(n)
    C = np.zeros((T * n, T * (n + m)))                             # :14
    b = np.zeros(T * n)                                            # :15

    def wseg(lo, hi):   # MATLAB w(lo:hi), 1-based inclusive; out-of-range errors like MATLAB
        if hi > w.shape[0]:
            raise IndexError("Index exceeds the number of array elements (w)")
        return w[lo - 1:hi]

    C[0:n, 0:m + n] = np.hstack([-B, np.eye(n)])                   # :38
    b[0:n] = A1 @ x0 + A2 @ x0_pre + wseg(1, n)                    # :39
    for i in range(1, T):                                          # :41
        if i == 1:
            C[n * i:n * (i + 1), m:(n + m) * (i + 1)] = np.hstack([-A1, -B, np.eye(n)])       # :43
            b[n * i:n * (i + 1)] = A2 @ x0 + wseg(n * i + 1, n * (i + 1))                     # :44
        else:
            C[n * i:n * (i + 1), m + (n + m) * (i - 2):(n + m) * (i + 1)] = np.hstack(
                [-A2, np.zeros((n, m)), -A1, -B, np.eye(n)])                                  # :46
            b[n * i:n * (i + 1)] = wseg(n * i + 1, n * (i + 1))                               # :47
    if not _isempty(xf):                                           # :67-70
        b = np.concatenate([b, xf])
        C = np.vstack([C, np.zeros((n, C.shape[1]))])
    C[C.shape[0] - n:, C.shape[1] - n:] = np.eye(n)                # :71
    return C, b


def fast_mpc_eq_const_var1(obj, literal_bug: bool = True):
    """VAR_1/fast_mpc_eq_const.m:1-56.  `literal_bug=True` reproduces :34-37 exactly
    (second block row written at column n instead of m+1);
    `literal_bug=False` writes it where the else-branch formula (:40) would."""
    A, B = obj.A1, obj.B
    if _isempty(A):                                                # :17-18
        raise ValueError("Define the state dynamics/equality constrained matrix")
    if _isempty(B):                                                # :19-20
        raise ValueError("Define the control dynamics/equality constrained matrix")
    n = A.shape[1]
    m = B.shape[1]
    x = obj.x0
    T = obj.T
    w = obj.w
    xf = obj.x_final
    if A.shape[1] != x.shape[0]:                                   # :23-24
        raise ValueError("The equality state dynamics matrix size does not match")
    if B.shape[1] != obj.R.shape[1]:                               # :25-26
        raise ValueError("The equality control dynamics matrix size does not match")
    if _isempty(w):                                                # :27-28
        w = np.zeros(n)
    C = np.zeros((T * n, T * (n + m)))                             # :12
    b = np.zeros(T * n)                                            # :13
    C[0:n, 0:m + n] = np.hstack([-B, np.eye(n)])                   # :32
    blk = np.hstack([-A, -B, np.eye(n)])
    for i in range(n, T * n - n + 1 + 1, n):                       # :34  i = n:n:T*n-n+1
        if i + n > w.shape[0]:
            raise IndexError("Index exceeds the number of array elements (w)")
        if i == n and literal_bug:
            c0 = i + ((i // n) - 1) * (n + m + n)                  # :36  (= n)
            c1 = i + n + ((i // n) - 1) * (n + m + n) + m + n - 1  #      (= 3n+m-1)
            if c1 > C.shape[1]:
                # MATLAB would silently GROW C here; reproduce by padding columns.
                C = np.hstack([C, np.zeros((C.shape[0], c1 - C.shape[1]))])
            C[i:i + n, c0 - 1:c1] = blk
            b[i:i + n] = w[i:i + n]                                # :37
        else:
            c0 = ((i // n) - 1) * (n + m) + m + 1                  # :40
            c1 = ((i // n) - 1) * (n + m) + m + m + n + n
            C[i:i + n, c0 - 1:c1] = blk
            b[i:i + n] = w[i:i + n]                                # :41
    b[0:n] = A @ x + w[0:n]                                        # :48
    if not _isempty(xf):                                           # :51-54
        b = np.concatenate([b, xf])
        C = np.vstack([C, np.zeros((n, C.shape[1]))])
    C[C.shape[0] - n:, C.shape[1] - n:] = np.eye(n)                # :55
    return C, b


# --------------------------------------------------------------------------------------
# L0: Newton / KKT kernel
# --------------------------------------------------------------------------------------
def inf_newton_KKT_H(H, P, h, z, k):
    """inf_newton_KKT_H.m:1-15.  The reference builds D as a dense 2Tm x 2Tm matrix by a
    scalar loop (:5-9) and multiplies P'*D*P densely (:13); scaling the rows of P by the
    diagonal is the same arithmetic minus additions of exact zeros."""
    d_inv = h - P @ z                                              # :3
    Ddiag = (1.0 / d_inv) ** 2                                     # :8
    d = 1.0 / d_inv                                                # :12
    Hk = 2 * H + k * (P.T @ (Ddiag[:, None] * P))                  # :13
    return Hk, d


def backtracking_inf_newton(z, nu, del_z, del_nu, rp, rd, al, bt, stats=None):
    """backtracking_inf_newton.m:1-13.  The counter at :3 is never decremented and the
    error at :6-8 is unreachable; the loop ends when the test passes (at the latest when
    t underflows to 0, where both sides are equal)."""
    t = 1.0                                                        # :2
    nhalf = 0
    while (np.linalg.norm(np.concatenate([rp(z + t * del_z), rd(z + t * del_z, nu + t * del_nu)]), 2)
           > (1 - al * t) * np.linalg.norm(np.concatenate([rp(z), rd(z, nu)]), 2)):    # :4
        t = bt * t                                                 # :5
        nhalf += 1
    if stats is not None:
        stats.setdefault("halvings", []).append(nhalf)
    return z + t * del_z, nu + t * del_nu                          # :10-11


def inf_newton_solver(H, g, P, h, C, b, k, z, newton, nu0=None, stream=None, stats=None):
    """inf_newton_solver.m:1-43.

    nu0    : explicit dual start; if None, `rand(length(b),1)` is drawn from `stream`
             (default: the module-level MATLAB-session stream), :2.
    newton : None (MATLAB []) => max_iter = 1000, :4-8.
    stats  : optional dict; receives 'iters' (Newton steps taken), 'early_exit', 'halvings'.
    """
    if nu0 is None:
        nu = (stream or _GLOBAL_STREAM).rand(b.shape[0])           # :2
    else:
        nu = np.asarray(nu0, dtype=np.float64).reshape(-1).copy()
        assert nu.shape[0] == b.shape[0], "nu0 must have length(b) entries"
    max_iter = 1000 if newton is None else int(newton)             # :4-8
    tol = 1e-6                                                     # :9
    z = z.copy()
    iters = 0
    for _ in range(max_iter):                                      # :10
        KKT_H, d = inf_newton_KKT_H(H, P, h, z, k)                 # :11
        Ptd = P.T @ d                                              # d is FROZEN inside the closures

        def rd(zz, vv, Ptd=Ptd):
            return 2 * (H @ zz) + g + k * Ptd + C.T @ vv           # :12

        def rp(zz):
            return C @ zz - b                                      # :13

        tol_g = C @ z - b                                          # :15
        n_r = np.linalg.norm(np.concatenate([-rd(z, nu), -rp(z)]), 2)   # :14,16
        n_g = np.linalg.norm(tol_g)                                # :17
        if n_r <= tol and n_g <= 1e-8:                             # :19
            if stats is not None:
                stats["iters"] = iters
                stats["early_exit"] = True
            return z                                               # :20-21
        L = sla.cholesky(KKT_H, lower=True)                        # :24 (raises if not PD, like chol)
        Y = C @ sla.solve_triangular(L.T, sla.solve_triangular(L, C.T, lower=True), lower=False)   # :27
        phi_inv_rd = sla.solve_triangular(L.T, sla.solve_triangular(L, rd(z, nu), lower=True), lower=False)  # :28
        Beta = -rp(z) + C @ phi_inv_rd                             # :29
        SL = sla.cholesky(Y, lower=True)                           # :30
        int_nu = sla.solve_triangular(SL, -Beta, lower=True)       # :31
        del_nu = sla.solve_triangular(SL.T, int_nu, lower=False)   # :32
        int_z = sla.solve_triangular(L, -rd(z, nu) - C.T @ del_nu, lower=True)   # :34
        del_z = sla.solve_triangular(L.T, int_z, lower=False)      # :35
        al = 10 ** -4                                              # :36
        bt = 0.5                                                   # :37
        z, nu = backtracking_inf_newton(z, nu, del_z, del_nu, rp, rd, al, bt, stats)   # :38
        iters += 1
    if stats is not None:
        stats["iters"] = iters
        stats["early_exit"] = False
        stats["nu"] = nu
    return z                                                       # :42


# --------------------------------------------------------------------------------------
# L2: the value class
# --------------------------------------------------------------------------------------
class Fast_MPC2:
    """VAR_2/Fast_MPC2.m:1-146 -- 23-argument value class (ctor :28-55)."""

    var_order = 2

    def __init__(self, Q, R, S, Qf, q, r, qf, xmin, xmax, umin, umax, dumin, dumax, T, x0, x0_pre,
                 u_prev, A1, A2, B, w, xf, x_init):
        self.Q, self.R, self.S, self.Qf = _mat(Q), _mat(R), S, _mat(Qf)
        self.q, self.r, self.qf = _col(q), _col(r), _col(qf)
        self.x_min, self.x_max = _col(xmin), _col(xmax)
        self.u_min, self.u_max = _col(umin), _col(umax)
        self.du_min, self.du_max = _col(dumin), _col(dumax)
        self.T = int(T)
        self.x0, self.x0_pre, self.u_prev = _col(x0), _col(x0_pre), _col(u_prev)
        self.A1, self.A2, self.B = _mat(A1), _mat(A2), _mat(B)
        self.w = _col(w)
        self.x_final = _col(xf)
        self.x_init = _col(x_init)
        self.stream = None      # MatlabRand to draw nu from; None => module-level stream
        self.last_stats = None

    # :56-67
    def objective_function(self):
        return fast_mpc_objective(self)

    def inequality_const(self):
        return fast_mpc_ineq_const_var2(self)

    def equality_const(self):
        return fast_mpc_eq_const_var2(self)

    def initialize(self):
        return fast_mpc_init(self)

    def _assemble(self):
        z = self.initialize()
        H, g = self.objective_function()
        P, h = self.inequality_const()
        C, b = self.equality_const()
        return z, H, g, P, h, C, b

    def _solve(self, H, g, P, h, C, b, k, z, nw, nu0=None):
        self.last_stats = {}
        return inf_newton_solver(H, g, P, h, C, b, k, z, nw, nu0=nu0, stream=self.stream,
                                 stats=self.last_stats)

    def mpc_fixed_log_newton(self, nw, k, nu0=None):
        """:124-130 -- THE hot entry."""
        z, H, g, P, h, C, b = self._assemble()
        return self._solve(H, g, P, h, C, b, k, z, nw, nu0)

    def mpc_fixed_log(self, k, nu0=None):
        """:116-123 -- nw = [] => up to 1000 Newton steps."""
        z, H, g, P, h, C, b = self._assemble()
        return self._solve(H, g, P, h, C, b, k, z, None, nu0)

    def _kappa_continuation(self, nw, nu0_list=None):
        """:100-115 (mpc_solve_full, nw=[]) and :131-144 (mpc_fixed_newton)."""
        k = 1.0
        mu = 1 / 10
        z, H, g, P, h, C, b = self._assemble()
        x_opt = z
        j = 0
        while k * z.shape[0] >= 10e-3:                             # :108 / :138
            nu0 = None if nu0_list is None else nu0_list[j]
            x_opt = self._solve(H, g, P, h, C, b, k, z, nw, nu0)
            k = mu * k
            z = x_opt
            j += 1
        return x_opt

    def mpc_solve_full(self, nu0_list=None):
        return self._kappa_continuation(None, nu0_list)

    def mpc_fixed_newton(self, nw, nu0_list=None):
        return self._kappa_continuation(nw, nu0_list)

    def mpc_solve_check(self, k_min, k_max, nu0_list=None):
        """:88-99 -- five linearly spaced kappa from k_max down to k_min."""
        ks = np.linspace(k_max, k_min, 5)
        z, H, g, P, h, C, b = self._assemble()
        x_opt = z
        for i, k in enumerate(ks):
            nu0 = None if nu0_list is None else nu0_list[i]
            x_opt = self._solve(H, g, P, h, C, b, k, z, None, nu0)
            z = x_opt
        return x_opt

    @staticmethod
    def kappa_schedule(N):
        """The kappa values the continuation loop visits for a problem with N variables."""
        ks, k = [], 1.0
        while k * N >= 10e-3:
            ks.append(k)
            k = (1 / 10) * k
        return ks


class Fast_MPC2_VAR1(Fast_MPC2):
    """VAR_1/Fast_MPC2.m -- 21-argument ctor (:26-27): no x0_pre, single A."""

    var_order = 1

    def __init__(self, Q, R, S, Qf, q, r, qf, xmin, xmax, umin, umax, dumin, dumax, T, x0, u_prev,
                 A, B, w, xf, x_init, literal_bug: bool = True):
        super().__init__(Q, R, S, Qf, q, r, qf, xmin, xmax, umin, umax, dumin, dumax, T, x0, None,
                         u_prev, A, None, B, w, xf, x_init)
        self.literal_bug = literal_bug

    def inequality_const(self):
        return fast_mpc_ineq_const_var1(self)

    def equality_const(self):
        return fast_mpc_eq_const_var1(self, self.literal_bug)


# --------------------------------------------------------------------------------------
# helpers around the boundary
# --------------------------------------------------------------------------------------
def deinterleave(z, n, m, T):
    """README.md:558-570 -- z -> (U m x T, X n x T), columns = stages."""
    Z = np.asarray(z).reshape(T, n + m)
    return Z[:, :m].T.copy(), Z[:, m:].T.copy()


def interleave(U, X):
    """Inverse of `deinterleave` (layout of fast_mpc_init.m:22-25)."""
    return np.hstack([np.asarray(U).T, np.asarray(X).T]).reshape(-1)


def dense_kkt_newton_step(H, g, P, h, C, b, k, z, nu):
    """Independent cross-check (not in the reference): one infeasible-start Newton step
    from the full KKT system  [Phi C'; C 0][dz; dnu] = -[r_d; r_p]  solved by LU."""
    Phi, d = inf_newton_KKT_H(H, P, h, z, k)
    r_d = 2 * (H @ z) + g + k * (P.T @ d) + C.T @ nu
    r_p = C @ z - b
    N, p = H.shape[0], C.shape[0]
    K = np.block([[Phi, C.T], [C, np.zeros((p, p))]])
    sol = np.linalg.solve(K, -np.concatenate([r_d, r_p]))
    return sol[:N], sol[N:]
