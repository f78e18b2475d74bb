"""Host-side mirror of the estimator step of the reference's closed loop (README.md:478):

    ad_est = lsqminnorm((A_s'*A_s), ((A_s)'*(Y_M - b_s)));

`Estimator(A_s, b_s)` holds the model of model_approx.mat (piston column already removed, README.md:289-290);
`estimate(Y)` runs the batched GPU kernel for one measurement vector per row of Y.  No CPU fallback."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import check, load_library


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Estimator:
    def __init__(self, A_s, b_s=None, max_batch: int = 1, device: int = 0):
        self._L = load_library()
        A = np.asfortranarray(np.asarray(A_s, dtype=np.float64))             # npix x nmodes, column-major like MATLAB
        if A.ndim != 2:
            raise ValueError("A_s must be a matrix (npix x nmodes)")
        self.npix, self.nmodes = A.shape
        b = None if b_s is None else np.ascontiguousarray(np.asarray(b_s, dtype=np.float64).reshape(-1))
        if b is not None and b.shape[0] != self.npix:
            raise ValueError("Arrays have incompatible sizes for this operation (Y_M - b_s)")
        h = C.c_void_p()
        check(self._L.est_create(C.byref(h), self.npix, self.nmodes, _ptr(A), _ptr(b), int(max_batch), int(device)))
        self._h = h
        self.max_batch = int(max_batch)

    def close(self):
        if getattr(self, "_h", None):
            self._L.est_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launch_count(self) -> int:
        return int(self._L.est_launch_count(self._h))

    def estimate(self, Y):
        """Y: (nb, npix) or (npix,) measurements -> (x_hat (nb, nmodes), telapsed seconds)."""
        Y = np.ascontiguousarray(np.asarray(Y, dtype=np.float64))
        if Y.ndim == 1:
            Y = Y.reshape(1, -1)
        if Y.shape[1] != self.npix:
            raise ValueError("Arrays have incompatible sizes for this operation (Y_M - b_s)")
        nb = Y.shape[0]
        out = np.empty((nb, self.nmodes))
        tel = C.c_double(0.0)
        check(self._L.est_apply(self._h, nb, _ptr(Y), _ptr(out), C.cast(C.byref(tel), C.c_void_p)))
        return out, tel.value


def identify_var(ad_acc, order: int = 2, device: int = 0):
    """README.md:116-130 on the GPU.  ad_acc: (K, n) training series (rows = time, like ad_acc(1:num_train,:)) or a batch
    (nseq, K, n).  Returns (A (order, n, n) or (nseq, order, n, n) with A[j-1] = A_j, telapsed seconds)."""
    L = load_library()
    a = np.asarray(ad_acc, dtype=np.float64)
    single = a.ndim == 2
    if single:
        a = a[None]
    nseq, K, n = a.shape
    acm = np.ascontiguousarray(np.transpose(a, (0, 2, 1)))              # per sequence column-major K x n
    out = np.empty((nseq, order, n, n))
    info = np.zeros(nseq, dtype=np.int32)
    tel = C.c_double(0.0)
    check(L.var_identify(nseq, K, n, int(order), _ptr(acm), _ptr(out), _ptr(info), int(device), C.cast(C.byref(tel), C.c_void_p)))
    if info.any():
        raise np.linalg.LinAlgError(f"AA'*AA is not positive definite for sequences {np.nonzero(info)[0].tolist()}")
    A = np.ascontiguousarray(np.transpose(out, (0, 1, 3, 2)))           # stored column-major n x n
    return (A[0] if single else A), tel.value
