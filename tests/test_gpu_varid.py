"""GPU parity of the batched VAR identification (README.md:116-130) against its literal oracle.  The normal
equations square the conditioning of the lagged data matrix (cond(AA'AA) ~ 1e4..1e5 here), the kernel accumulates
them in double-double; tolerance 1e-9 relative per matrix, like U and X."""
import numpy as np
import pytest

from cases import relerr
from oracle import varid_ref

pytestmark = pytest.mark.gpu
TOL = 1e-9


def ar_series(seed, n, K, order=2, scale=None):
    rs = np.random.RandomState(seed)
    A1 = 0.9 * np.eye(n) + 0.02 * rs.randn(n, n) / np.sqrt(n)
    A2 = (0.05 * np.eye(n) + 0.01 * rs.randn(n, n) / np.sqrt(n)) if order >= 2 else np.zeros((n, n))
    sig = np.ones(n) if scale is None else scale
    a = np.zeros((K + 200, n))
    for k in range(2, K + 200):
        a[k] = A1 @ a[k - 1] + A2 @ a[k - 2] + sig * rs.randn(n)
    return a[200:]


def test_readme_size(pk):
    """n = 27 modes (piston removed), num_train = 1000, PN = 2, Kolmogorov-like mode variances (README.md:116-130)."""
    nm = np.concatenate([[k] * (k + 1) for k in range(7)])[1:]
    sig = (nm + 1.0) ** (-11.0 / 6.0)
    a = ar_series(1, 27, 1000, scale=sig / sig.max())
    A, tel = pk.identify_var(a, 2)
    ref = varid_ref.identify(a, 2)
    assert relerr(A[0], ref[0]) < TOL and relerr(A[1], ref[1]) < TOL and tel > 0


@pytest.mark.parametrize("n,K,order,nseq", [(6, 100, 2, 5), (27, 1000, 1, 2), (9, 333, 3, 3), (1, 50, 2, 2), (30, 600, 2, 1)])
def test_batches_and_orders(pk, n, K, order, nseq):
    a = np.stack([ar_series(10 * s + n, n, K, order=min(order, 2)) for s in range(nseq)])
    A, _ = pk.identify_var(a, order)
    for s in range(nseq):
        ref = varid_ref.identify(a[s], order)
        for j in range(order):
            assert relerr(A[s, j], ref[j]) < TOL, (s, j)


def test_errors(pk):
    with pytest.raises(pk.FmpcError) as e:       # (order n)(order + 1) n accumulators exceed the kernel's register budget
        pk.identify_var(np.random.RandomState(0).randn(600, 45), 2)
    assert e.value.code == -2
    with pytest.raises(pk.FmpcError):            # fewer equations than unknowns
        pk.identify_var(np.random.RandomState(0).randn(20, 15), 2)
    a = np.zeros((100, 4)); a[:, 0] = np.random.RandomState(1).randn(100)
    with pytest.raises(np.linalg.LinAlgError):   # singular Gram matrix (three identically zero series)
        pk.identify_var(a, 2)
