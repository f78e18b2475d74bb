"""ORACLE (test infrastructure, NOT product code) -- restatement of the reference's
`zernfun.m` (Fricker 2012) and `zernmodfit.m`, plus the driver's fixed pupil grid
(README.md:78-93).

PARITY UNPINNED: the reference has no golden vectors for this path and MATLAB/Octave
are absent.  Pins are the repo's own known-answer tests (a frame
synthesised from known coefficients returns them; QR == normal equations == pinv).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as sla

__all__ = ["mode_indices", "zernfun", "zernmodfit", "pupil_grid", "fit_frames_literal"]


def _prod2(k: int) -> float:
    """MATLAB prod(2:k): empty product = 1 for k < 2 (zernfun.m:167-170)."""
    p = 1.0
    for j in range(2, int(k) + 1):
        p *= j
    return p


def mode_indices(N: int):
    """zernmodfit.m:195-198 -- n = 0..N repeated n+1 times; within each n,
    m = [(-n:2:-1) fliplr(n:-2:0)] i.e. ascending -n:2:n."""
    ns, ms = [], []
    for x in range(0, N + 1):
        ns += [x] * (x + 1)
        neg = list(range(-x, 0, 2))                 # (-x:2:-1)
        pos = list(range(x, -1, -2))[::-1]          # fliplr(x:-2:0)
        ms += neg + pos
    return np.array(ns, dtype=np.int64), np.array(ms, dtype=np.int64)


def zernfun(n, m, r, theta, norm: bool = False) -> np.ndarray:
    """zernfun.m:140-192 -- Z_n^m(r,theta), un-normalised unless norm=True.
    Returns length(r) x length(n)."""
    n = np.asarray(n, dtype=np.int64).reshape(-1)
    m = np.asarray(m, dtype=np.int64).reshape(-1)
    r = np.asarray(r, dtype=np.float64).reshape(-1)
    theta = np.asarray(theta, dtype=np.float64).reshape(-1)
    if np.any((n - m) % 2 != 0):
        raise ValueError("All N and M must differ by multiples of 2 (including 0).")
    if np.any(m > n):
        raise ValueError("Each M must be less than or equal to its corresponding N.")
    if np.any((r > 1) | (r < 0)):
        raise ValueError("All R must be between 0 and 1.")
    length_r = r.shape[0]
    m_abs = np.abs(m)                                             # :140
    rpowers = np.unique(np.concatenate([np.arange(ma, nn + 1, 2) for ma, nn in zip(m_abs, n)]))  # :141-145
    # :150-157 -- r.^p columns; the p == 0 column is ones
    rpowern = {int(p): (np.ones(length_r) if p == 0 else r ** int(p)) for p in rpowers}
    z = np.zeros((length_r, n.shape[0]))                          # :161
    for j in range(n.shape[0]):                                   # :162
        nj, mj = int(n[j]), int(m_abs[j])
        s = list(range(0, (nj - mj) // 2 + 1))                    # :163
        pows = list(range(nj, mj - 1, -2))                        # :164
        for k in range(len(s) - 1, -1, -1):                       # :165
            p = ((1 - 2 * (s[k] % 2)) * _prod2(nj - s[k]) / _prod2(s[k])
                 / _prod2((nj - mj) // 2 - s[k]) / _prod2((nj + mj) // 2 - s[k]))   # :166-170
            z[:, j] = z[:, j] + p * rpowern[pows[k]]              # :171-172
        if norm:
            z[:, j] *= np.sqrt((1 + (m[j] != 0)) * (nj + 1) / np.pi)   # :175-177
    idx_pos = m > 0                                               # :184-185
    idx_neg = m < 0
    if np.any(idx_pos):
        z[:, idx_pos] *= np.cos(np.outer(theta, m_abs[idx_pos]))  # :187-189
    if np.any(idx_neg):
        z[:, idx_neg] *= np.sin(np.outer(theta, m_abs[idx_neg]))  # :190-192
    return z


def zernmodfit(r, theta, data, N: int):
    """zernmodfit.m:154-213.  Returns (ad, nm): ad is nmodes x 2 with column 2 == 0
    (the 'modified' rotation part is commented out, :216-244), nm = [n m]."""
    r = np.asarray(r, dtype=np.float64).reshape(-1)
    theta = np.asarray(theta, dtype=np.float64).reshape(-1)
    data = np.asarray(data, dtype=np.float64).reshape(-1)
    if not (r.shape[0] == theta.shape[0] == data.shape[0]):       # :161-166
        raise ValueError("The inputs R, THETA, and DATA must all have the same number of elements.")
    if N < 0 or N != round(N):                                    # :173-176
        raise ValueError("N must be a positive integer or zero.")
    if np.any((r > 1) | (r < 0)):                                 # :182-184
        raise ValueError("All R must be between 0 and 1.")
    n, m = mode_indices(int(N))                                   # :195-198
    z = zernfun(n, m, r, theta)                                   # :205
    # :209  c = z\data : MATLAB mldivide on a tall matrix = QR with column pivoting
    Qm, Rm, piv = sla.qr(z, mode="economic", pivoting=True)
    y = sla.solve_triangular(Rm, Qm.T @ data, lower=False)
    c = np.empty_like(y)
    c[piv] = y
    ad = np.column_stack([c, np.zeros_like(c)])                   # :213
    return ad, np.column_stack([n, m])


def pupil_grid(nL: int):
    """README.md:78-84 -- fixed grid of the driver.  Returns (r_in, theta_in, is_in) with
    is_in an nL x nL boolean mask indexed [row, col] like MATLAB's meshgrid output;
    r_in/theta_in are taken in MATLAB COLUMN-MAJOR order (`r(is_in)`)."""
    x = np.arange(-(nL - 1), (nL - 1) + 1, 2, dtype=np.float64) / (nL - 1)
    X, Y = np.meshgrid(x, x)                                      # X varies along columns
    theta = np.arctan2(Y, X)                                      # cart2pol
    r = np.hypot(X, Y)
    is_in = r <= np.max(np.abs(x))
    sel = is_in.T.reshape(-1)                                     # column-major linear order
    return r.T.reshape(-1)[sel], theta.T.reshape(-1)[sel], is_in


def fit_frames_literal(frames: np.ndarray, N: int) -> np.ndarray:
    """The driver loop README.md:88-93: per frame `zernmodfit(r,theta,z(is_in),N)`,
    keeping column 1.  frames: (nf, nL, nL) with frames[j] indexed [row, col].
    Returns nf x nmodes (rows like `ad_acc`)."""
    nf, nL, _ = frames.shape
    r, theta, is_in = pupil_grid(nL)
    sel = is_in.T.reshape(-1)
    out = []
    for j in range(nf):
        zcol = frames[j].T.reshape(-1)[sel]                       # z(is_in), column-major
        ad, _ = zernmodfit(r, theta, zcol, N)
        out.append(ad[:, 0])
    return np.array(out)
