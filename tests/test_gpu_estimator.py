"""GPU parity of the batched estimator step (README.md:478) against the lsqminnorm oracle on the reference's
model_approx.mat (committed fixture).  Tolerance 1e-10 normwise, like zernmodfit."""
import os

import numpy as np
import pytest

from cases import relerr
from oracle import estimator_ref as er

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "model_approx.npz")
TOL = 1e-10


def test_estimator_golden(pk):
    g = np.load(GOLD)
    est = pk.Estimator(g["A_s"][:, 1:], g["b_s"], max_batch=8)
    xh, tel = est.estimate(g["y"])
    assert relerr(xh, g["x_hat"]) < TOL and tel > 0 and est.launch_count >= 1
    x1, _ = est.estimate(g["y"][0])                       # a single measurement vector, like the reference's loop
    assert relerr(x1[0], g["x_hat"][0]) < TOL
    est.close()


@pytest.mark.parametrize("nb", [1, 9, 300, 5000])
def test_estimator_batches_vs_oracle(pk, nb):
    g = np.load(GOLD)
    A, b = g["A_s"][:, 1:], g["b_s"]
    rs = np.random.RandomState(nb)
    x = 0.3 * rs.randn(nb, 27)
    y = b[None] + x @ A.T + 1e-3 * rs.randn(nb, A.shape[0])
    est = pk.Estimator(A, b, max_batch=nb)
    xh, _ = est.estimate(y)
    assert relerr(xh, er.estimate_pinv(A, b, y)) < TOL
    if nb <= 9:
        assert relerr(xh, er.estimate(A, b, y)) < TOL
    est.close()


def test_estimator_other_shapes_and_errors(pk):
    rs = np.random.RandomState(4)
    for npix, nm in ((64, 5), (1001, 66), (40, 40), (333, 80)):       # aligned rows, odd rows, square, > 72 modes (scalar kernel)
        A = rs.randn(npix, nm)
        b = rs.randn(npix)
        y = rs.randn(7, npix)
        est = pk.Estimator(A, b, max_batch=7)
        xh, _ = est.estimate(y)
        assert relerr(xh, er.estimate_pinv(A, b, y)) < 1e-9
        est.close()
    A0 = rs.randn(50, 3)
    est = pk.Estimator(A0, None, max_batch=2)                          # b_s = []
    y = rs.randn(2, 50)
    assert relerr(est.estimate(y)[0], er.estimate_pinv(A0, np.zeros(50), y)) < 1e-9
    est.close()
    A = rs.randn(30, 4)
    A[:, 3] = A[:, 0]                                                  # rank deficient
    with pytest.raises(pk.FmpcError) as e:
        pk.Estimator(A, None)
    assert e.value.code == -13
    with pytest.raises(pk.FmpcError):
        pk.Estimator(rs.randn(3, 5), None)                             # more modes than pixels
