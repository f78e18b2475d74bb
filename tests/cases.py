"""Seeded problem generators shared by the CPU and GPU tests."""
import numpy as np


def small_problem(seed, n, m, T, nb, umax, a2=True, xf=False, warm=False, qscale=3.0):
    rs = np.random.RandomState(seed)
    A1 = 0.5 * np.eye(n) + 0.1 * rs.randn(n, n)
    A2 = (0.2 * np.eye(n) + 0.05 * rs.randn(n, n)) if a2 else None
    B = rs.randn(n, m)
    Q = np.diag(1 + rs.rand(n)) * qscale
    R = np.diag(1 + rs.rand(m))
    Qf = Q * 2
    x0 = rs.randn(nb, n)
    x0p = rs.randn(nb, n) if a2 else None
    w = 0.1 * rs.randn(nb, T * n)
    xfv = 0.1 * rs.randn(nb, n) if xf else None
    nu0 = rs.rand(nb, (T + (1 if xf else 0)) * n)
    um = umax * np.ones(m)
    X0 = U0 = None
    if warm:
        X0 = 0.5 * rs.randn(nb, T, n)
        U0 = np.clip(0.5 * rs.randn(nb, T, m), -0.9 * umax, 0.9 * umax)
    return dict(n=n, m=m, T=T, nb=nb, A1=A1, A2=A2, B=B, Q=Q, R=R, Qf=Qf, u_min=-um, u_max=um,
                x_min=-100.0 * np.ones(n), x_max=100.0 * np.ones(n), x0=x0, x0_pre=x0p, w=w, xf=xfv, nu0=nu0,
                X0=X0, U0=U0)


def z0_of(c):
    """Interleaved start (nb, N): warm start or the midpoint cold start of fast_mpc_init.m:19-25."""
    nb, T, n, m = c["nb"], c["T"], c["n"], c["m"]
    if c["X0"] is None:
        stage = np.concatenate([(c["u_min"] + c["u_max"]) / 2, (c["x_min"] + c["x_max"]) / 2])
        return np.tile(stage, T)[None].repeat(nb, 0)
    return np.concatenate([c["U0"], c["X0"]], axis=2).reshape(nb, -1)


def dense_solve(fd, c, b, niters, kappa, var1_literal=False):
    """One instance through the literal dense oracle. Returns (z, stats)."""
    g = lambda k: None if c[k] is None else c[k][b]
    z0 = None if c["X0"] is None else z0_of(c)[b]
    if c["A2"] is not None:
        obj = fd.Fast_MPC2(c["Q"], c["R"], None, c["Qf"], None, None, None, c["x_min"], c["x_max"], c["u_min"],
                           c["u_max"], -np.ones(c["m"]), np.ones(c["m"]), c["T"], c["x0"][b], g("x0_pre"),
                           np.zeros(c["m"]), c["A1"], c["A2"], c["B"], g("w"), g("xf"), z0)
    else:
        obj = fd.Fast_MPC2_VAR1(c["Q"], c["R"], None, c["Qf"], None, None, None, c["x_min"], c["x_max"], c["u_min"],
                                c["u_max"], -np.ones(c["m"]), np.ones(c["m"]), c["T"], c["x0"][b], np.zeros(c["m"]),
                                c["A1"], c["B"], g("w"), g("xf"), z0, literal_bug=var1_literal)
        # box rows only (the GPU path's VAR(1) = VAR_2 code with A2 = 0)
        obj.inequality_const = lambda: fd.fast_mpc_ineq_const_var2(obj)
    z = obj.mpc_fixed_log_newton(niters, kappa, nu0=c["nu0"][b])
    return z, obj.last_stats


def ref_solve(fref, c, niters, kappa, **kw):
    """Whole batch through the structured C oracle. Returns dict with U (nb,T,m), X (nb,T,n)."""
    T = lambda a: None if a is None else a.T
    out = fref.solve_batch(c["A1"], c["A2"], c["B"], c["Q"], c["R"], c["Qf"], c["u_min"], c["u_max"], kappa, niters,
                           T(c["x0"]), T(c["x0_pre"]), T(c["w"]), T(z0_of(c)), T(c["nu0"]), xf=T(c["xf"]), **kw)
    Z = out["z"].T.reshape(c["nb"], c["T"], c["n"] + c["m"])
    out["U"], out["X"] = Z[:, :, :c["m"]].copy(), Z[:, :, c["m"]:].copy()
    return out


def relerr(a, b):
    """Normwise (max-norm) relative error per array, SURVEY.md 8c."""
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300))


def var1_literal_case(seed, n, m, T, nb, umax, du, xf=False, dense_q=False):
    """VAR(1) problem with ramp-rate bounds +-du, a warm start that respects them, and u_prev."""
    c = small_problem(seed, n, m, T, nb, umax, a2=False, xf=xf, warm=True)
    rs = np.random.RandomState(seed + 1000)
    u_prev = 0.3 * umax * rs.randn(nb, m).clip(-2, 2)
    U0 = u_prev[:, None, :] + np.cumsum(0.3 * du * rs.randn(nb, T, m).clip(-2, 2), axis=1)
    c["U0"] = np.clip(U0, -0.95 * umax, 0.95 * umax)
    c["u_prev"] = u_prev
    c["du_min"], c["du_max"] = -du * np.ones(m), du * np.ones(m)
    if dense_q:
        G = rs.randn(n, n)
        c["Q"] = G @ G.T / n + np.eye(n)
        G = rs.randn(n, n)
        c["Qf"] = 2 * (G @ G.T / n + np.eye(n))
    return c


def var1_literal_dense(fd, c, b, niters, kappa, ramp=True, bug=True):
    """One instance through the literal dense oracle's VAR_1 class (ramp rows / literal C placement on request)."""
    g = lambda k: None if c.get(k) is None else c[k][b]
    z0 = None if c["X0"] is None else z0_of(c)[b]
    du_min = c.get("du_min", -np.ones(c["m"]))
    du_max = c.get("du_max", np.ones(c["m"]))
    u_prev = g("u_prev") if c.get("u_prev") is not None else np.zeros(c["m"])
    if c["A2"] is None:
        obj = fd.Fast_MPC2_VAR1(c["Q"], c["R"], None, c["Qf"], None, None, None, c["x_min"], c["x_max"], c["u_min"],
                                c["u_max"], du_min, du_max, c["T"], c["x0"][b], u_prev, c["A1"], c["B"], g("w"), g("xf"), z0,
                                literal_bug=bug)
        if not ramp:
            obj.inequality_const = lambda: fd.fast_mpc_ineq_const_var2(obj)
    else:
        obj = fd.Fast_MPC2(c["Q"], c["R"], None, c["Qf"], None, None, None, c["x_min"], c["x_max"], c["u_min"],
                           c["u_max"], du_min, du_max, c["T"], c["x0"][b], g("x0_pre"), u_prev, c["A1"], c["A2"], c["B"],
                           g("w"), g("xf"), z0)
        if ramp:
            obj.inequality_const = lambda: fd.fast_mpc_ineq_const_var1(obj)
    z = obj.mpc_fixed_log_newton(niters, kappa, nu0=c["nu0"][b])
    return z, obj.last_stats
