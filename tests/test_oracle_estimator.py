"""CPU: the hand-written MATLAB-v7.3 reader against the reference's model_approx.mat (when /root/reference is
mounted) and against the committed fixture; the estimator oracle's two routes; known-answer recovery."""
import os

import numpy as np
import pytest

from cases import relerr
from oracle import estimator_ref as er

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "model_approx.npz")
REF = "/root/reference/model_approx.mat"


def test_fixture_pins():
    """shapes and checksums of A_s, b_s."""
    g = np.load(GOLD)
    A, b = g["A_s"], g["b_s"]
    assert A.shape == (2883, 28) and b.shape == (2883,)
    assert np.abs(A).max() == 8.534834222687513
    assert abs(b.sum() - 283.1102428882918) < 1e-10 and b.max() == 74.48320434840755
    assert abs(np.linalg.norm(A[:, 0]) - 0.03047259966674046) < 1e-12          # piston column carries no signal
    sv = np.linalg.svd(A[:, 1:], compute_uv=False)
    assert abs(sv[0] - 16.831944172860897) < 1e-9 and abs(sv[-1] - 1.8126513854648034) < 1e-9


@pytest.mark.skipif(not os.path.exists(REF), reason="the reference tree is not mounted (GPU box)")
def test_reader_matches_reference_file():
    from oracle.mat73 import Mat73, load_model_approx
    f = Mat73(REF)
    assert sorted(f.group_entries(f.root_btree, f.root_heap)) == ["A_s", "b_s"]
    A, b = load_model_approx(REF)
    g = np.load(GOLD)
    assert np.array_equal(A, g["A_s"]) and np.array_equal(b, g["b_s"])


def test_estimator_routes_agree_and_recover():
    g = np.load(GOLD)
    A, b = g["A_s"][:, 1:], g["b_s"]
    x1 = er.estimate(A, b, g["y"])
    x2 = er.estimate_pinv(A, b, g["y"])
    assert relerr(x1, x2) < 1e-12 and relerr(x1, g["x_hat"]) < 1e-13
    # noise-free measurements return the coefficients that made them
    x = np.random.RandomState(0).randn(3, 27)
    assert relerr(er.estimate(A, b, b[None] + x @ A.T), x) < 1e-12
