"""CPU: the VAR-identification oracle recovers a known VAR(2) model and agrees with the benchmark generator."""
import numpy as np

from cases import relerr
from oracle import varid_ref


def test_identify_recovers_known_model():
    rs = np.random.RandomState(0)
    n = 5
    A1 = 0.5 * np.eye(n) + 0.05 * rs.randn(n, n)
    A2 = 0.2 * np.eye(n) + 0.05 * rs.randn(n, n)
    a = np.zeros((4000, n))
    for k in range(2, 4000):
        a[k] = A1 @ a[k - 1] + A2 @ a[k - 2] + rs.randn(n)
    A = varid_ref.identify(a, 2)
    assert np.abs(A[0] - A1).max() < 0.06 and np.abs(A[1] - A2).max() < 0.06
    # the one-step predictions of the fitted model reproduce the data up to the driving noise (unit variance)
    res = a[2:] - a[1:-1] @ A[0].T - a[:-2] @ A[1].T
    assert abs(res.std() - 1.0) < 0.05


def test_identify_matches_benchmark_generator(pk):
    """synth.make_problem identifies A1 / A2 from its training sequence with the same normal equations."""
    from mpc_sensorlessao_b200 import synth
    import math
    p = synth.make_problem(2, 4, m1=3)
    n = p.n
    rs = np.random.RandomState(5489)
    rs.randn(n, n); rs.randn(n, n)               # G1, G2 draws of make_problem
    a = np.zeros((1000, n))
    for k in range(2, 1000):
        a[k] = p.A1_true @ a[k - 1] + p.A2_true @ a[k - 2] + p.sigma * rs.randn(n)
    A = varid_ref.identify(a, 2)
    assert relerr(A[0], p.A1) < 1e-9 and relerr(A[1], p.A2) < 1e-9
