"""CPU: the second structured C restatement (oracle/fmpc_ref_general.c: ramp rows, literal VAR_1 C, dense Q -- Thomas
solves per actuator, dense Schur complement) against the literal dense oracle and the committed var1lit_* fixtures.
Two independent CPU routes that agree pin the reference semantics the CUDA general-structure kernel is tested against."""
import glob
import os

import numpy as np
import pytest

from cases import relerr, small_problem, var1_literal_case, var1_literal_dense, z0_of
from oracle import fastmpc_dense as fd

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def c_general(fref, c, niters, kappa, ramp, bug):
    T_ = lambda a: None if a is None else np.asarray(a).T
    out = fref.solve_batch_general(c["A1"], c["A2"], c["B"], c["Q"], c["R"], c["Qf"], c["u_min"], c["u_max"], kappa, niters,
                                   T_(c["x0"]), T_(c["x0_pre"]), T_(c["w"]), T_(z0_of(c)), T_(c["nu0"]), xf=T_(c["xf"]),
                                   u_prev=T_(c.get("u_prev")), du_min=c.get("du_min"), du_max=c.get("du_max"), ramp_rows=ramp,
                                   literal_bug=bug)
    Z = out["z"].T.reshape(c["nb"], c["T"], c["n"] + c["m"])
    return Z[:, :, :c["m"]], Z[:, :, c["m"]:], out


@pytest.mark.parametrize("seed,n,m,T,xf,ramp,bug,dq", [
    (301, 6, 9, 6, False, True, True, False), (302, 6, 9, 6, False, True, False, False), (303, 6, 9, 6, False, False, True, False),
    (304, 7, 5, 8, True, True, False, False), (306, 9, 4, 5, False, True, True, False), (307, 6, 9, 6, False, False, False, True),
    (308, 8, 7, 7, True, True, True, True)])
def test_general_c_oracle_matches_dense(fref, seed, n, m, T, xf, ramp, bug, dq):
    c = var1_literal_case(seed, n, m, T, 2, 0.6 if m > 5 else 0.5, 0.15 if m > 5 else 0.1, xf=xf, dense_q=dq)
    U, X, out = c_general(fref, c, 5, 0.01, ramp, bug)
    for b in range(c["nb"]):
        z, st = var1_literal_dense(fd, c, b, 5, 0.01, ramp=ramp, bug=bug)
        Ud, Xd = fd.deinterleave(z, n, m, T)
        assert relerr(U[b], Ud.T) < 1e-9 and relerr(X[b], Xd.T) < 1e-9
        assert out["iters"][b] == st["iters"]


def test_general_c_oracle_var2_with_ramp(fref):
    c = small_problem(311, 6, 5, 6, 2, 0.8)
    c["du_min"], c["du_max"] = -0.5 * np.ones(5), 0.5 * np.ones(5)
    c["u_prev"] = 0.1 * np.random.RandomState(5).randn(2, 5)
    U, X, _ = c_general(fref, c, 4, 0.01, True, False)
    for b in range(2):
        z, _ = var1_literal_dense(fd, c, b, 4, 0.01, ramp=True, bug=False)
        Ud, Xd = fd.deinterleave(z, 6, 5, 6)
        assert relerr(U[b], Ud.T) < 1e-9 and relerr(X[b], Xd.T) < 1e-9


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "var1lit_*.npz"))), ids=lambda p: os.path.basename(p)[8:-4])
def test_general_c_oracle_matches_golden(fref, path):
    g = np.load(path)
    c = {k: g[k] for k in g.files}
    for k in ("n", "m", "T", "nb", "niters"):
        c[k] = int(c[k])
    for k in ("A2", "x0_pre", "xf"):
        c.setdefault(k, None)
    U, X, out = c_general(fref, c, c["niters"], float(c["kappa"]), bool(c["ramp"]), bool(c["bug"]))
    for b in range(c["nb"]):
        assert relerr(U[b], c["U"][b]) < 1e-9 and relerr(X[b], c["X"][b]) < 1e-9
    assert np.array_equal(out["iters"], c["iters"])
