"""Host-side mirror of the reference's `zernmodfit.m` for the batched GPU fit.

  zernmodfit(r, theta, data, N)  <->  zernmodfit.m:1  (same arguments; returns (ad, nm))
  ZernikeFitter(nL, N)           :   the driver loop README.md:78-93 batched over frames

  SampleFitter(r, theta, N)      :   the same fit for an ARBITRARY sample set (zernmodfit.m:154-213 takes any vectors)

`zernmodfit` uses the frame kernel when (r, theta) is the driver's fixed pupil grid (README.md:78-84, every call the
reference makes) and a `SampleFitter` otherwise; all arithmetic runs on the GPU, there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import check, load_library


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class ZernikeFitter:
    def __init__(self, nL: int, N: int, max_frames: int = 2048, device: int = 0):
        self._L = load_library()
        self.nL, self.N = int(nL), int(N)
        h = C.c_void_p()
        check(self._L.zmf_create(C.byref(h), self.nL, self.N, int(max_frames), int(device)))
        self._h = h
        self.nmodes = int(self._L.zmf_nmodes(h))
        self.npix_in = int(self._L.zmf_npix_in(h))
        self.max_frames = int(max_frames)

    def close(self):
        if getattr(self, "_h", None):
            self._L.zmf_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launch_count(self) -> int:
        return int(self._L.zmf_launch_count(self._h))

    def mask(self) -> np.ndarray:
        """is_in as an (nL, nL) boolean array indexed [row, col] (MATLAB orientation)."""
        mk = np.zeros(self.nL * self.nL, dtype=np.uint8)
        check(self._L.zmf_get_mask(self._h, _ptr(mk)))
        return mk.reshape(self.nL, self.nL).T.astype(bool)        # stored column-major

    def basis(self) -> np.ndarray:
        """Z (npix_in, nmodes): zernfun(n, m, r(is_in), theta(is_in)) in column-major pixel order."""
        Z = np.zeros((self.nmodes, self.npix_in))
        check(self._L.zmf_get_basis(self._h, _ptr(Z)))
        return Z.T.copy()

    def fit(self, frames: np.ndarray):
        """frames: (nf, nL, nL) with frames[j][row, col] = phase(row, col, j).
        Returns (coef (nf, nmodes) -- rows like `ad_acc` README.md:92 --, telapsed seconds)."""
        frames = np.asarray(frames, dtype=np.float64)
        if frames.ndim == 2:
            frames = frames[None]
        nf = frames.shape[0]
        if frames.shape[1:] != (self.nL, self.nL):
            raise ValueError("frames must be (nf, nL, nL)")
        # C-ABI wants MATLAB column-major frames: element (row, col) at row + nL*col
        fcm = np.ascontiguousarray(np.transpose(frames, (0, 2, 1)))
        coef = np.empty((nf, self.nmodes))
        tel = C.c_double(0.0)
        check(self._L.zmf_fit(self._h, nf, _ptr(fcm), _ptr(coef), C.cast(C.byref(tel), C.c_void_p)))
        return coef, tel.value


    def synth(self, coef: np.ndarray):
        """coef: (nf, nmodes) -> (frames (nf, nL, nL) with frames[j][row, col] = sum_k coef[j, k] Z_k(row, col) inside the
        pupil and 0 outside (README.md:592-598), telapsed seconds)."""
        coef = np.ascontiguousarray(np.asarray(coef, dtype=np.float64))
        if coef.ndim == 1:
            coef = coef.reshape(1, -1)
        if coef.shape[1] != self.nmodes:
            raise ValueError("coef must be (nf, nmodes)")
        nf = coef.shape[0]
        out = np.empty((nf, self.nL, self.nL))
        tel = C.c_double(0.0)
        check(self._L.zmf_synth(self._h, nf, _ptr(coef), _ptr(out), C.cast(C.byref(tel), C.c_void_p)))
        return np.ascontiguousarray(np.transpose(out, (0, 2, 1))), tel.value


class SampleFitter:
    """zernmodfit for an arbitrary sample set (r, theta): the basis and its least-squares operator are built once
    (`zmf_create_samples`), `fit(data)` projects any number of data vectors sampled at those points."""

    def __init__(self, r, theta, N: int, max_frames: int = 64, device: int = 0):
        self._L = load_library()
        self._r = np.ascontiguousarray(np.asarray(r, dtype=np.float64).reshape(-1))
        self._th = np.ascontiguousarray(np.asarray(theta, dtype=np.float64).reshape(-1))
        if self._r.shape != self._th.shape:
            raise ValueError("The inputs R, THETA, and DATA must all have the same number of elements.")
        if np.any((self._r > 1) | (self._r < 0)):
            raise ValueError("All R must be between 0 and 1.")                                           # zernmodfit.m:182-184
        self.N, self.npts = int(N), int(self._r.shape[0])
        h = C.c_void_p()
        check(self._L.zmf_create_samples(C.byref(h), self.npts, _ptr(self._r), _ptr(self._th), self.N, int(max_frames), int(device)))
        self._h = h
        self.nmodes = int(self._L.zmf_nmodes(h))

    def close(self):
        if getattr(self, "_h", None):
            self._L.zmf_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def fit(self, data: np.ndarray):
        """data: (nf, npts) or (npts,) -> (coef (nf, nmodes), telapsed seconds)."""
        data = np.ascontiguousarray(np.asarray(data, dtype=np.float64))
        if data.ndim == 1:
            data = data.reshape(1, -1)
        if data.shape[1] != self.npts:
            raise ValueError("The inputs R, THETA, and DATA must all have the same number of elements.")
        coef = np.empty((data.shape[0], self.nmodes))
        tel = C.c_double(0.0)
        check(self._L.zmf_fit(self._h, data.shape[0], _ptr(data), _ptr(coef), C.cast(C.byref(tel), C.c_void_p)))
        return coef, tel.value


_fitters = {}
_sample_fitters = {}


def _mode_table(N):
    n = np.concatenate([[k] * (k + 1) for k in range(N + 1)]).astype(int)
    m = np.concatenate([np.arange(-k, k + 1, 2) for k in range(N + 1)]).astype(int)
    return np.column_stack([n, m])


def _fit_samples(r, theta, data, N, device):
    import hashlib
    key = (hashlib.sha1(r.tobytes() + theta.tobytes()).hexdigest(), int(N), int(device))
    f = _sample_fitters.get(key)
    if f is None:
        if len(_sample_fitters) >= 4:
            _sample_fitters.pop(next(iter(_sample_fitters))).close()
        f = _sample_fitters[key] = SampleFitter(r, theta, int(N), max_frames=64, device=device)
    coef, _ = f.fit(data[None])
    return np.column_stack([coef[0], np.zeros(f.nmodes)]), _mode_table(int(N))


def zernmodfit(r, theta, data, N, device: int = 0):
    """zernmodfit.m:1 -- [ad, nm] = zernmodfit(r, theta, data, N) on the driver's pupil grid.
    ad is nmodes x 2 with column 2 == 0 (:213), nm = [n m] (:214)."""
    r = np.asarray(r, dtype=np.float64).reshape(-1)
    theta = np.asarray(theta, dtype=np.float64).reshape(-1)
    data = np.asarray(data, dtype=np.float64).reshape(-1)
    if not (r.shape[0] == theta.shape[0] == data.shape[0]):
        raise ValueError("The inputs R, THETA, and DATA must all have the same number of elements.")     # :161-166
    if N < 0 or N != round(N):
        raise ValueError("N must be a positive integer or zero.")                                        # :173-176
    if np.any((r > 1) | (r < 0)):
        raise ValueError("All R must be between 0 and 1.")                                               # :182-184
    # recover nL from the number of in-pupil samples of the README grid
    key = None
    for (nL, NN), f in _fitters.items():
        if NN == N and f.npix_in == r.shape[0]:
            key = (nL, NN)
    if key is None:
        for nL in range(2, 2049):
            x = np.arange(-(nL - 1), nL, 2) / (nL - 1)
            cnt = int((np.hypot(*np.meshgrid(x, x)) <= 1.0).sum())
            if cnt == r.shape[0]:
                key = (nL, int(N))
                _fitters[key] = ZernikeFitter(nL, int(N), max_frames=64, device=device)
                break
            if cnt > r.shape[0]:
                break
    if key is None:
        return _fit_samples(r, theta, data, N, device)          # any other sample set: zernmodfit.m:154-213 as it stands
    f = _fitters[key]
    nL = key[0]
    x = np.arange(-(nL - 1), nL, 2) / (nL - 1)
    X, Y = np.meshgrid(x, x)
    is_in = f.mask()
    sel = is_in.T.reshape(-1)
    r_ref = np.hypot(X, Y).T.reshape(-1)[sel]
    th_ref = np.arctan2(Y, X).T.reshape(-1)[sel]
    if np.max(np.abs(r_ref - r)) > 1e-12 or np.max(np.abs(np.angle(np.exp(1j * (th_ref - theta))))) > 1e-12:
        return _fit_samples(r, theta, data, N, device)          # as many samples as a pupil grid, but not that grid
    frame_cm = np.zeros(nL * nL)
    frame_cm[sel] = data
    frame = frame_cm.reshape(nL, nL).T
    coef, _ = f.fit(frame[None])
    n = np.concatenate([[k] * (k + 1) for k in range(N + 1)]).astype(int)
    m = np.concatenate([np.arange(-k, k + 1, 2) for k in range(N + 1)]).astype(int)
    return np.column_stack([coef[0], np.zeros(f.nmodes)]), np.column_stack([n, m])
