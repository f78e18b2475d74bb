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
    # rank-deficient A_s: lsqminnorm's minimum-norm solution (README.md:478), tolerance 1e-10 like zernmodfit
    for npix, nm, dup in ((30, 4, [(3, 0)]), (200, 27, [(5, 2), (20, 19)]), (64, 9, [(8, 0), (7, 0)])):
        A = rs.randn(npix, nm)
        for dst, src in dup:
            A[:, dst] = A[:, src]                                      # duplicated columns
        if nm == 27:
            A[:, 11] = A[:, 3] - 2.0 * A[:, 4]                         # and a linear combination
        b = rs.randn(npix)
        y = rs.randn(5, npix)
        est = pk.Estimator(A, b, max_batch=5)
        xh, _ = est.estimate(y)
        ref = er.estimate(A, b, y)                                     # Gram matrix + minimum-norm lstsq, one per measurement
        assert relerr(xh, ref) < 1e-10
        assert relerr(xh, er.estimate_pinv(A, b, y)) < 1e-9
        for dst, src in dup:                                           # minimum norm: duplicated columns share the weight
            assert np.abs(xh[:, dst] - xh[:, src]).max() < 1e-10 * np.abs(xh).max()
        est.close()
    with pytest.raises(pk.FmpcError):
        pk.Estimator(rs.randn(3, 5), None)                             # more modes than pixels


def test_c1_closed_loop_estimator_then_var1_solve(pk):
    """BASELINE.json configs[0] in miniature: the reference's closed-loop step (README.md:444-570) with the estimator
    of model_approx.mat in front of the VAR(1) fastMPC solve exactly as Fast_MPC/VAR_1 assembles it (ramp rows,
    literal C), n = 27, m = 144, T = 10 -- GPU estimator + GPU solve against the lsqminnorm / dense oracles."""
    from mpc_sensorlessao_b200 import synth
    from oracle import fastmpc_dense as fd
    g = np.load(GOLD)
    A_s, b_s = g["A_s"][:, 1:], g["b_s"]
    p = synth.make_problem(6, 10, var_order=1, drop_piston=True, u_bound=28.0)
    K = 2
    a = synth.aberrations(p, 1, K, seed=11, amp=0.3)[0]
    rs = np.random.RandomState(12)
    noise = 1e-3 * rs.randn(K, A_s.shape[0])
    nu0 = rs.rand(K, p.T * p.n)
    est = pk.Estimator(A_s, b_s, max_batch=1)
    hb = pk.FastMPCBatch(p.A1, None, p.B, p.Q, p.R, p.Qf, p.u_min, p.u_max, p.T, p.x_min, p.x_max, du_min=p.du_min,
                         du_max=p.du_max, ramp_rows=True, var1_literal_bug=True, max_batch=1)
    u_prev_g = np.zeros(p.m); u_prev_o = np.zeros(p.m)
    zg = zo = None
    for k in range(K):
        # measurement of the residual aberration through the first-order PSF model, y = b_s + A_s x + noise
        yg = b_s + A_s @ (a[k] + p.B @ u_prev_g) + noise[k]
        yo = b_s + A_s @ (a[k] + p.B @ u_prev_o) + noise[k]
        x0g = est.estimate(yg)[0][0]
        x0o = er.estimate(A_s, b_s, yo[None])[0]
        assert relerr(x0g, x0o) < 1e-9
        X0 = U0 = None
        if zg is not None:
            Zs = np.vstack([zg.reshape(p.T, -1)[1:], zg.reshape(p.T, -1)[-1:]])
            U0, X0 = Zs[None, :, :p.m], Zs[None, :, p.m:]
            zo = np.vstack([zo.reshape(p.T, -1)[1:], zo.reshape(p.T, -1)[-1:]]).reshape(-1)
        out = hb.step(x0g, None, None, None, X0, U0, nu0[k], u_prev=u_prev_g, kappa=0.01, niters=3)
        zg = np.hstack([out["U"][0], out["X"][0]]).reshape(-1)
        o = fd.Fast_MPC2_VAR1(p.Q, p.R, None, p.Qf, None, None, None, p.x_min, p.x_max, p.u_min, p.u_max, p.du_min, p.du_max,
                              p.T, x0o, u_prev_o, p.A1, p.B, np.zeros(p.T * p.n), None, zo)
        zo = o.mpc_fixed_log_newton(3, 0.01, nu0=nu0[k])
        assert relerr(zg[:p.m], zo[:p.m]) < 1e-9, k
        u_prev_g, u_prev_o = zg[:p.m].copy(), zo[:p.m].copy()
    est.close(); hb.close()
