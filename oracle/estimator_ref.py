"""ORACLE / TEST INFRASTRUCTURE -- CPU restatement of the estimator step of the reference's closed loop,
README.md:478:   ad_est = lsqminnorm((A_s'*A_s), ((A_s)'*(Y_M - b_s)));
with A_s (npix x nmodes, piston column already removed, README.md:289-290) and b_s from model_approx.mat.
lsqminnorm is MATLAB's minimum-norm least-squares solve (complete orthogonal decomposition); for the
full-rank 27 x 27 Gram matrix of the reference model (cond(A_s) = 9.29) it is the unique solution, restated
here literally (Gram matrix, then numpy's minimum-norm lstsq) and, as a cross-check, through pinv(A_s).
PARITY UNPINNED by the reference (MATLAB, no golden vectors, cannot run here): pinned on the reference's own DATA
(model_approx.mat, checksums of SURVEY.md 8c) and by the agreement of the two routes below (tests/test_oracle_estimator.py).
Only tests/, bench.py's baseline leg and __graft_entry__.smoke() may import this module."""
import numpy as np


def estimate(A_s: np.ndarray, b_s: np.ndarray, y: np.ndarray) -> np.ndarray:
    """y: (nb, npix) measurements -> (nb, nmodes) estimates, one lsqminnorm per measurement like the reference."""
    G = A_s.T @ A_s
    out = np.empty((y.shape[0], A_s.shape[1]))
    for i in range(y.shape[0]):
        rhs = A_s.T @ (y[i] - b_s)
        # minimum-norm least squares; lsqminnorm's default rank tolerance max(size(G)) * eps(norm(G)) = numpy's rcond = n * eps
        out[i] = np.linalg.lstsq(G, rhs, rcond=G.shape[0] * np.finfo(float).eps)[0]
    return out


def estimate_pinv(A_s: np.ndarray, b_s: np.ndarray, y: np.ndarray) -> np.ndarray:
    return (y - b_s[None, :]) @ np.linalg.pinv(A_s).T
