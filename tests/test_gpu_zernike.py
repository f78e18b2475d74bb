"""GPU parity tests of the batched zernmodfit kernel (tolerance 1e-10 normwise, north-star)."""
import os

import numpy as np
import pytest

from cases import relerr
from oracle import zernike_ref as zr

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-10


@pytest.mark.parametrize("N", [6, 10])
def test_golden(pk, N):
    g = np.load(os.path.join(GOLD, f"zernmodfit_N{N}.npz"))
    zf = pk.ZernikeFitter(128, N, max_frames=16)
    coef, _ = zf.fit(g["frames"].astype(np.float64))
    assert relerr(coef, g["coef"]) < TOL
    zf.close()


@pytest.mark.parametrize("nL,N", [(128, 6), (64, 3), (33, 0), (16, 2)])
def test_basis_mask_and_fit_vs_oracle(pk, nL, N):
    zf = pk.ZernikeFitter(nL, N, max_frames=64)
    r, th, is_in = zr.pupil_grid(nL)
    n_, m_ = zr.mode_indices(N)
    Z = zr.zernfun(n_, m_, r, th)
    assert (zf.mask() == is_in).all() and zf.npix_in == r.shape[0] and zf.nmodes == n_.shape[0]
    assert np.abs(zf.basis() - Z).max() < 1e-12
    rs = np.random.RandomState(nL)
    nf = 37                                         # ragged: not a multiple of the kernel's frame tile
    frames = np.full((nf, nL * nL), np.nan)         # NaN outside the pupil must not leak (zernmodfit.m:30)
    frames[:, is_in.T.reshape(-1)] = rs.randn(nf, Z.shape[0])
    frames = frames.reshape(nf, nL, nL).transpose(0, 2, 1)
    coef, _ = zf.fit(frames)
    assert np.isfinite(coef).all()
    assert relerr(coef, zr.fit_frames_literal(frames, N)) < TOL
    zf.close()


def test_known_coefficients_roundtrip_2000_frames(pk):
    """BASELINE config 3 size: 2000 frames 128 x 128, N = 6; frames synthesised from known coefficients."""
    zf = pk.ZernikeFitter(128, 6, max_frames=2000)
    Z = zf.basis()
    is_in = zf.mask()
    rs = np.random.RandomState(3)
    c = rs.randn(2000, 28)
    frames = np.zeros((2000, 128 * 128))
    frames[:, is_in.T.reshape(-1)] = c @ Z.T
    frames = frames.reshape(2000, 128, 128).transpose(0, 2, 1)
    coef, tel = zf.fit(frames)
    assert relerr(coef, c) < TOL and tel > 0
    # linearity
    coef2, _ = zf.fit(2.5 * frames[:64] + frames[64:128])
    assert relerr(coef2, 2.5 * c[:64] + c[64:128]) < TOL
    zf.close()


def test_function_mirror(pk):
    r, th, is_in = zr.pupil_grid(64)
    d = np.random.RandomState(0).randn(r.shape[0])
    ad, nm = pk.zernmodfit(r, th, d, 4)
    ad_ref, nm_ref = zr.zernmodfit(r, th, d, 4)
    assert relerr(ad, ad_ref) < TOL and np.array_equal(nm, nm_ref)
    with pytest.raises(ValueError, match="between 0 and 1"):
        pk.zernmodfit(r * 2, th, d, 4)


@pytest.mark.parametrize("npts,N,nf", [(500, 4, 1), (3001, 6, 40), (12644, 6, 3), (257, 10, 9), (90, 11, 5)])
def test_arbitrary_sample_sets(pk, npts, N, nf):
    """zernmodfit.m:154-213 takes ANY sample vectors (its docstring example :21-90 uses a Cartesian patch of `peaks`):
    random points in the unit disk, GPU fit (zmf_create_samples) against the oracle's zernmodfit per data vector."""
    rs = np.random.RandomState(1000 + npts)
    r = np.sqrt(rs.rand(npts))
    th = rs.uniform(-np.pi, np.pi, npts)
    r[0], r[1] = 0.0, 1.0                                  # both ends of the allowed range
    data = rs.randn(nf, npts)
    sf = pk.SampleFitter(r, th, N, max_frames=nf)
    coef, _ = sf.fit(data)
    sf.close()
    for j in range(nf):
        ad_ref, _ = zr.zernmodfit(r, th, data[j], N)
        assert relerr(coef[j], ad_ref[:, 0]) < TOL, j
    # the function mirror takes the same route for anything that is not the driver's pupil grid
    ad, nm = pk.zernmodfit(r, th, data[0], N)
    ad_ref, nm_ref = zr.zernmodfit(r, th, data[0], N)
    assert relerr(ad, ad_ref) < TOL and np.array_equal(nm, nm_ref) and np.all(ad[:, 1] == 0)


def test_arbitrary_sample_sets_errors(pk):
    rs = np.random.RandomState(3)
    with pytest.raises(ValueError, match="between 0 and 1"):
        pk.SampleFitter(np.array([0.1, 1.2, 0.3]), np.zeros(3), 0)
    with pytest.raises(pk.FmpcError):                      # fewer samples than modes
        pk.SampleFitter(rs.rand(5), rs.rand(5), 3)
    with pytest.raises(pk.FmpcError) as e:                 # all samples on one ray: the azimuthal modes are not determined
        pk.SampleFitter(rs.rand(200), np.zeros(200), 4)
    assert e.value.code == -13


@pytest.mark.parametrize("nL,N,nf", [(128, 6, 1), (128, 6, 9), (128, 10, 70), (64, 2, 130), (128, 6, 4500), (20, 11, 33), (50, 4, 17)])
def test_dmma_path_shapes(pk, nL, N, nf):
    """Every tile-count / frame-tile / split-K configuration of the DMMA kernel (and, for nL % 4 != 0 or > 72 modes,
    the scalar kernel) against W = pinv(Z) applied in numpy (the oracle's QR route == pinv route, test_oracle_zernike)."""
    zf = pk.ZernikeFitter(nL, N, max_frames=nf)
    r, th, is_in = zr.pupil_grid(nL)
    n_, m_ = zr.mode_indices(N)
    Z = zr.zernfun(n_, m_, r, th)
    rs = np.random.RandomState(nL + N + nf)
    vals = rs.randn(nf, Z.shape[0])
    frames = np.full((nf, nL * nL), np.nan)
    frames[:, is_in.T.reshape(-1)] = vals
    frames = frames.reshape(nf, nL, nL).transpose(0, 2, 1)
    coef, _ = zf.fit(frames)
    ref = np.linalg.lstsq(Z, vals.T, rcond=None)[0].T
    assert np.isfinite(coef).all() and relerr(coef, ref) < TOL
    if nf <= 70:
        assert relerr(coef, zr.fit_frames_literal(frames, N)) < TOL
    zf.close()


@pytest.mark.parametrize("nL,N,nf", [(128, 6, 5), (128, 6, 300), (128, 10, 66), (64, 3, 4200), (50, 4, 9), (16, 12, 3)])
def test_synthesis_matches_basis(pk, nL, N, nf):
    """zmf_synth (README.md:592-598, phase_cor = sum_j ad_cor(j) Z_j): frames = Z coef inside the pupil, 0 outside; and
    fit(synth(c)) == c (the two kernels are each other's pseudo-inverse on the pupil)."""
    zf = pk.ZernikeFitter(nL, N, max_frames=nf)
    r, th, is_in = zr.pupil_grid(nL)
    n_, m_ = zr.mode_indices(N)
    Z = zr.zernfun(n_, m_, r, th)
    c = np.random.RandomState(nf).randn(nf, Z.shape[1])
    frames, tel = zf.synth(c)
    ref = np.zeros((nf, nL * nL))
    ref[:, is_in.T.reshape(-1)] = c @ Z.T
    ref = ref.reshape(nf, nL, nL).transpose(0, 2, 1)
    assert relerr(frames, ref) < 1e-13 and tel > 0
    assert (frames[:, ~is_in] == 0).all()
    back, _ = zf.fit(frames)
    assert relerr(back, c) < TOL
    zf.close()


def test_dm_influence_matrix_is_a_zernike_fit(pk):
    """README.md:196-271: B = pinv(Z'Z) Z' I -- the influence matrix is zernmodfit applied to the 144 actuator
    influence functions, i.e. one zmf_fit call with the influence functions as frames."""
    from mpc_sensorlessao_b200 import synth
    nL, N, m1 = 128, 6, 12
    B_ref = synth.influence_matrix(N, m1, 0.1, nL, False)                # host construction used by the benchmark problems
    x = np.arange(-(nL - 1), nL, 2) / (nL - 1)
    X, Y = np.meshgrid(x, x)
    ax = np.linspace(-1.0, 1.0, m1)
    d = ax[1] - ax[0]
    infl = np.array([np.exp(np.log(0.1) * ((X - ax[j]) ** 2 + (Y + ax[i]) ** 2) / d ** 2) for i in range(m1) for j in range(m1)])
    zf = pk.ZernikeFitter(nL, N, max_frames=m1 * m1)
    B, _ = zf.fit(infl)
    zf.close()
    assert B.T.shape == B_ref.shape and relerr(B.T, B_ref) < TOL
