"""CPU: the structured C oracle and the zernike oracle against the committed golden fixtures
(generated from the literal dense oracle by tests/golden/make_golden.py)."""
import glob
import os

import numpy as np
import pytest

from cases import ref_solve, relerr
from oracle import zernike_ref as zr

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FMPC_FIXTURES = sorted(glob.glob(os.path.join(GOLD, "fmpc_*.npz")))


def load_case(path):
    g = np.load(path)
    c = {k: g[k] for k in g.files}
    for k in ("n", "m", "T", "nb", "niters"):
        c[k] = int(c[k])
    c["kappa"] = float(c["kappa"])
    for k in ("A2", "x0_pre", "xf", "X0", "U0", "w"):
        c.setdefault(k, None)
    return c


def test_fixture_inventory():
    assert len(FMPC_FIXTURES) == 7 and len(glob.glob(os.path.join(GOLD, "zernmodfit_*.npz"))) == 2


@pytest.mark.parametrize("path", FMPC_FIXTURES, ids=lambda p: os.path.basename(p)[5:-4])
def test_structured_oracle_matches_golden(path, fref):
    c = load_case(path)
    ref = ref_solve(fref, c, c["niters"], c["kappa"])
    for b in range(c["nb"]):
        assert relerr(ref["U"][b], c["U"][b]) < 1e-9
        assert relerr(ref["X"][b], c["X"][b]) < 1e-9
    assert np.array_equal(ref["iters"], c["iters"])
    assert np.array_equal(ref["status"] == 1, c["early_exit"])


@pytest.mark.parametrize("N", [6, 10])
def test_zernike_oracle_matches_golden(N):
    g = np.load(os.path.join(GOLD, f"zernmodfit_N{N}.npz"))
    coef = zr.fit_frames_literal(g["frames"].astype(np.float64), N)
    assert relerr(coef, g["coef"]) < 1e-12
