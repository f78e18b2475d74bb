"""CPU tests pinning the zernfun / zernmodfit restatement."""
import numpy as np
import pytest

from oracle import zernike_ref as zr


def test_mode_order_is_osa_ansi():
    n, m = zr.mode_indices(3)
    assert n.tolist() == [0, 1, 1, 2, 2, 2, 3, 3, 3, 3]
    assert m.tolist() == [0, -1, 1, -2, 0, 2, -3, -1, 1, 3]
    assert zr.mode_indices(6)[0].shape[0] == 28 and zr.mode_indices(10)[0].shape[0] == 66


def test_known_polynomials():
    r = np.linspace(0, 1, 11)
    th = np.linspace(-3, 3, 11)
    Z = zr.zernfun([0, 1, 1, 2, 2, 4, 3], [0, -1, 1, 0, 2, 0, -1], r, th)
    assert np.allclose(Z[:, 0], 1.0)
    assert np.allclose(Z[:, 1], r * np.sin(th))
    assert np.allclose(Z[:, 2], r * np.cos(th))
    assert np.allclose(Z[:, 3], 2 * r ** 2 - 1)
    assert np.allclose(Z[:, 4], r ** 2 * np.cos(2 * th))
    assert np.allclose(Z[:, 5], 6 * r ** 4 - 6 * r ** 2 + 1)
    assert np.allclose(Z[:, 6], (3 * r ** 3 - 2 * r) * np.sin(th))
    Zn = zr.zernfun([2], [0], r, th, norm=True)
    assert np.allclose(Zn[:, 0], np.sqrt(3 / np.pi) * (2 * r ** 2 - 1))


def test_pupil_grid_matches_survey_counts():
    r, th, is_in = zr.pupil_grid(128)
    assert r.shape[0] == 12644 and is_in.sum() == 12644 and r.max() <= 1.0
    n, m = zr.mode_indices(6)
    assert abs(np.linalg.cond(zr.zernfun(n, m, r, th)) - 3.82) < 0.01
    n, m = zr.mode_indices(10)
    assert abs(np.linalg.cond(zr.zernfun(n, m, r, th)) - 5.0) < 0.1


@pytest.mark.parametrize("N", [0, 3, 6, 10])
def test_fit_recovers_known_coefficients_and_ignores_nan_outside(N):
    nL = 64
    r, th, is_in = zr.pupil_grid(nL)
    n, m = zr.mode_indices(N)
    Z = zr.zernfun(n, m, r, th)
    c = np.random.RandomState(N).randn(n.shape[0])
    fr = np.full(nL * nL, np.nan)                       # NaN outside the pupil (zernmodfit.m:30)
    fr[is_in.T.reshape(-1)] = Z @ c
    out = zr.fit_frames_literal(fr.reshape(nL, nL).T[None], N)
    assert np.abs(out[0] - c).max() < 1e-12


def test_qr_normal_equations_and_pinv_agree():
    r, th, _ = zr.pupil_grid(128)
    n, m = zr.mode_indices(6)
    Z = zr.zernfun(n, m, r, th)
    d = np.random.RandomState(1).randn(Z.shape[0])
    ad, nm = zr.zernmodfit(r, th, d, 6)
    assert ad.shape == (28, 2) and np.all(ad[:, 1] == 0) and nm.shape == (28, 2)
    c_ne = np.linalg.solve(Z.T @ Z, Z.T @ d)
    c_pi = np.linalg.pinv(Z) @ d
    assert np.abs(ad[:, 0] - c_ne).max() < 1e-12 and np.abs(ad[:, 0] - c_pi).max() < 1e-12
    assert np.abs(np.linalg.pinv(Z) @ Z - np.eye(28)).max() < 1e-12


def test_input_checks():
    with pytest.raises(ValueError, match="same number of elements"):
        zr.zernmodfit([0.1, 0.2], [0.0], [1.0, 2.0], 2)
    with pytest.raises(ValueError, match="between 0 and 1"):
        zr.zernmodfit([0.1, 1.2], [0.0, 0.1], [1.0, 2.0], 0)
    with pytest.raises(ValueError, match="positive integer"):
        zr.zernmodfit([0.1, 0.2], [0.0, 0.1], [1.0, 2.0], -1)
