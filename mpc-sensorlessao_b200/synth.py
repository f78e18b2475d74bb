"""Synthetic problem data for tests and bench.py -- host-side numpy, not on the hot path.

The reference's MPC matrices are not shipped (A1/A2 are identified from OOMAO phase screens,
B needs Zs.mat), so every matrix is synthesised here, deterministically:
  * A1, A2 : a stable "true" VAR(2) is simulated and then IDENTIFIED by least squares exactly as the
             reference does (README.md:116-130);
  * B      : Gaussian influence functions of a 12 x 12 actuator grid, coupling 0.1 (README.md:196-232),
             least-squares projected onto the Zernike basis on the pupil grid (README.md:271);
  * costs / bounds : README.md:344-356;
  * aberration sequences : the true VAR(2) driven by white noise with a Kolmogorov-like modal
             spectrum, scaled by mag_conv_10 (README.md:279).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

MAG_CONV_10 = 1.781797436291855          # README.md:279


def zernike_modes(N: int):
    """(n, m) index vectors in the reference's order (zernmodfit.m:195-198): n = 0..N, m = -n:2:n."""
    n = np.concatenate([[k] * (k + 1) for k in range(N + 1)]).astype(int)
    m = np.concatenate([np.arange(-k, k + 1, 2) for k in range(N + 1)]).astype(int)
    return n, m


def zernike_basis(N: int, r: np.ndarray, theta: np.ndarray) -> np.ndarray:
    """Un-normalised Zernike functions (Fricker's zernfun convention) -- data generation only."""
    n, m = zernike_modes(N)
    Z = np.zeros((r.shape[0], n.shape[0]))
    for j, (nj, mj) in enumerate(zip(n, m)):
        ma = abs(int(mj))
        rad = np.zeros_like(r)
        for s in range((nj - ma) // 2 + 1):
            c = ((-1) ** s * math.factorial(nj - s)
                 / (math.factorial(s) * math.factorial((nj + ma) // 2 - s) * math.factorial((nj - ma) // 2 - s)))
            rad = rad + c * r ** (nj - 2 * s)
        Z[:, j] = rad * (np.cos(ma * theta) if mj > 0 else np.sin(ma * theta) if mj < 0 else 1.0)
    return Z


def pupil(nL: int):
    """README.md:78-84 grid; returns (r_in, theta_in, X_in, Y_in) in column-major pixel order."""
    x = np.arange(-(nL - 1), nL, 2, dtype=np.float64) / (nL - 1)
    X, Y = np.meshgrid(x, x)
    r = np.hypot(X, Y)
    sel = (r <= 1.0).T.reshape(-1)
    f = lambda A: A.T.reshape(-1)[sel]
    return f(r), f(np.arctan2(Y, X)), f(X), f(Y)


def influence_matrix(N: int, m1: int = 12, coupling: float = 0.1, nL: int = 128, drop_piston: bool = False):
    """B (nmodes x m1^2): Gaussian actuator influence functions projected on the Zernike basis."""
    r, th, X, Y = pupil(nL)
    Z = zernike_basis(N, r, th)
    pos = np.linspace(-1.0, 1.0, m1)
    d = pos[1] - pos[0]
    G = np.zeros((r.shape[0], m1 * m1))
    a = 0
    for i in range(m1):            # README.md:220-230: row index i -> y, column index j -> x
        for j in range(m1):
            G[:, a] = np.exp(math.log(coupling) * ((X - pos[j]) ** 2 + (Y + pos[i]) ** 2) / d ** 2)
            a += 1
    B = np.linalg.lstsq(Z, G, rcond=None)[0]
    return B[1:] if drop_piston else B


@dataclass
class Problem:
    n: int
    m: int
    T: int
    A1: np.ndarray
    A2: np.ndarray
    B: np.ndarray
    Q: np.ndarray
    R: np.ndarray
    Qf: np.ndarray
    u_min: np.ndarray
    u_max: np.ndarray
    x_min: np.ndarray
    x_max: np.ndarray
    du_min: np.ndarray
    du_max: np.ndarray
    A1_true: np.ndarray
    A2_true: np.ndarray
    sigma: np.ndarray


def make_problem(N: int = 6, T: int = 20, var_order: int = 2, drop_piston: bool = False, m1: int = 12,
                 u_bound: float = 28.0, seed: int = 5489, nL: int = 128, n_train: int = 1000) -> Problem:
    """The benchmark problem family: N = 6 -> 28 modes (27 without piston), N = 10 -> 66 modes."""
    nm, _ = zernike_modes(N)
    if drop_piston:
        nm = nm[1:]
    n = nm.shape[0]
    rs = np.random.RandomState(seed)
    G1, G2 = rs.randn(n, n) / math.sqrt(n), rs.randn(n, n) / math.sqrt(n)
    A1t = 0.9 * np.eye(n) + 0.02 * G1
    A2t = (0.05 * np.eye(n) + 0.01 * G2) if var_order == 2 else np.zeros((n, n))
    comp = np.block([[A1t, A2t], [np.eye(n), np.zeros((n, n))]])
    rho = np.max(np.abs(np.linalg.eigvals(comp)))
    if rho > 0.98:                                   # rescale so the companion radius is 0.98
        s = 0.98 / rho
        A1t, A2t = A1t * s, A2t * s * s
    sigma = (nm + 1.0) ** (-11.0 / 6.0)
    sigma = sigma / sigma.max()
    # training sequence + least-squares identification, README.md:116-130
    a = np.zeros((n_train, n))
    for k in range(2, n_train):
        a[k] = A1t @ a[k - 1] + A2t @ a[k - 2] + sigma * rs.randn(n)
    PN = var_order
    AA = np.hstack([a[PN - j - 1:n_train - j - 1] for j in range(PN)])
    BB = a[PN:]
    PARA = np.linalg.solve(AA.T @ AA, AA.T @ BB)
    A1 = PARA[:n].T.copy()
    A2 = PARA[n:2 * n].T.copy() if var_order == 2 else None
    B = influence_matrix(N, m1, 0.1, nL, drop_piston)
    m = B.shape[1]
    Q = 1.5e4 * np.eye(n)                             # README.md:344-346
    return Problem(n, m, T, A1, A2, B, Q, np.eye(m), Q.copy(),
                   -u_bound * np.ones(m), u_bound * np.ones(m), -100.0 * np.ones(n), 100.0 * np.ones(n),
                   -0.2121 * np.ones(m), 0.2121 * np.ones(m), A1t, A2t, sigma)


def aberrations(p: Problem, nbatch: int, K: int, seed: int = 1, amp: float = 2.0) -> np.ndarray:
    """(nbatch, K, n) open-loop modal aberration sequences a[k] = A1 a[k-1] + A2 a[k-2] + w_k,
    burn-in 200 steps, scaled so the largest mode has rms ~ amp/3 rad, times mag_conv_10."""
    rs = np.random.RandomState(seed)
    n = p.n
    burn = 200
    a1 = np.zeros((nbatch, n))
    a2 = np.zeros((nbatch, n))
    out = np.empty((nbatch, K, n))
    for k in range(burn + K):
        a0 = a1 @ p.A1_true.T + a2 @ p.A2_true.T + p.sigma * rs.randn(nbatch, n)
        a2, a1 = a1, a0
        if k >= burn:
            out[:, k - burn] = a0
    scale = (amp / 3.0) / max(out[..., 0 if n > 1 else 0].std(), 1e-12)
    return out * scale * MAG_CONV_10


def warm_inputs(p: Problem, nbatch: int, seed: int = 2, amp: float = 2.0):
    """One batch of per-step solver inputs in the closed-loop regime without running a loop:
    x0, x0_pre from consecutive aberration samples (plus a small actuator residual), a warm start
    near the unconstrained optimum (so Hessians differ per instance), nu0 uniform."""
    rs = np.random.RandomState(seed)
    a = aberrations(p, nbatch, 2, seed=seed + 17, amp=amp)
    x0_pre, x0 = a[:, 0].copy(), a[:, 1].copy()
    T, n, m = p.T, p.n, p.m
    Bp = np.linalg.pinv(p.B)
    X0 = np.zeros((nbatch, T, n))
    U0 = np.zeros((nbatch, T, m))
    xm1, x = x0_pre, x0
    for t in range(T):                 # roll the model forward with the minimum-norm correcting input
        A2 = p.A2 if p.A2 is not None else np.zeros((n, n))
        free = x @ p.A1.T + xm1 @ A2.T
        u = -(free @ Bp.T) * (0.9 + 0.1 * rs.rand(nbatch, 1)) + 0.05 * rs.randn(nbatch, m)
        u = np.clip(u, 0.95 * p.u_min, 0.95 * p.u_max)
        xn = free + u @ p.B.T
        U0[:, t], X0[:, t] = u, xn + 1e-3 * rs.randn(nbatch, n)
        xm1, x = x, xn
    nu0 = rs.rand(nbatch, T * n)
    return dict(x0=x0, x0_pre=x0_pre, X0=X0, U0=U0, nu0=nu0)
