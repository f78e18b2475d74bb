"""Extracts A_s (2883 x 28) and b_s (2883) from the reference's model_approx.mat (MATLAB v7.3 / HDF5, read by the
hand-written parser oracle/mat73.py) into tests/golden/model_approx.npz, plus estimator golden vectors
x_hat = lsqminnorm(A_s'A_s, A_s'(y - b_s)) (README.md:478, piston column removed as README.md:289-290 does) for
synthetic measurements y = b_s + A_s [0; x] + noise.  /root/reference does not exist on the GPU box, hence the fixture.

    python tests/golden/make_model_approx.py [/root/reference/model_approx.mat]
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle.mat73 import load_model_approx  # noqa: E402
from oracle.estimator_ref import estimate  # noqa: E402


def main():
    path = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/model_approx.mat"
    A_s, b_s = load_model_approx(path)
    rs = np.random.RandomState(2883)
    nb = 6
    x = 0.3 * rs.randn(nb, 27)
    y = b_s[None, :] + x @ A_s[:, 1:].T + 1e-3 * rs.randn(nb, A_s.shape[0])
    xhat = estimate(A_s[:, 1:], b_s, y)
    np.savez_compressed(os.path.join(HERE, "model_approx.npz"), A_s=A_s, b_s=b_s, y=y, x_true=x, x_hat=xhat)
    print("A_s", A_s.shape, "max|A_s|", np.abs(A_s).max(), "sum(b_s)", b_s.sum(), "est err", np.abs(xhat - x).max())


if __name__ == "__main__":
    main()
