"""CPU tests pinning the oracles for the fastMPC path (the reference has no golden vectors,
SURVEY.md 4 / 8c, so the pins are created here):
  1. literal dense restatement == structured C restatement == one-shot dense KKT solve;
  2. known answers (equality-constrained QP optimum, terminal xf, r_p = 0 after a full step,
     nu-independence of a full step, VAR(2) with A2 = 0 == corrected VAR(1));
  3. the MATLAB random stream; the reference's error() strings; the VAR_1 literal bug switch.
"""
import numpy as np
import pytest

from cases import dense_solve, ref_solve, relerr, small_problem, z0_of
from oracle import fastmpc_dense as fd

TOL = 1e-9      # north-star tolerance on U and X (relative, per array)

CASES = [
    dict(seed=1, n=6, m=4, T=5, nb=2, umax=2.0),
    dict(seed=2, n=6, m=4, T=5, nb=2, umax=0.3, xf=True),
    dict(seed=3, n=6, m=4, T=5, nb=2, umax=0.3, a2=False),
    dict(seed=4, n=8, m=5, T=10, nb=2, umax=0.2, xf=True, warm=True),
    dict(seed=5, n=8, m=5, T=10, nb=2, umax=0.1, warm=True),
    dict(seed=6, n=5, m=7, T=1, nb=2, umax=0.5),
    dict(seed=7, n=5, m=7, T=2, nb=2, umax=0.5, xf=True),
    dict(seed=8, n=3, m=9, T=3, nb=2, umax=0.5, a2=False, xf=True, warm=True),
]


@pytest.mark.parametrize("kw", CASES, ids=lambda k: f"s{k['seed']}")
def test_dense_equals_structured(kw, fref):
    c = small_problem(**kw)
    ref = ref_solve(fref, c, 6, 0.01)
    for b in range(c["nb"]):
        z, st = dense_solve(fd, c, b, 6, 0.01)
        U, X = fd.deinterleave(z, c["n"], c["m"], c["T"])
        assert relerr(ref["U"][b], U.T) < TOL
        assert relerr(ref["X"][b], X.T) < TOL
        assert ref["iters"][b] == st["iters"]
        assert (ref["status"][b] == 1) == st["early_exit"]
        assert ref["halvings"][b] == sum(st.get("halvings", []))


def test_dense_step_equals_one_shot_kkt():
    c = small_problem(11, 7, 5, 6, 1, 0.4, warm=True)
    z1, _ = dense_solve(fd, c, 0, 1, 0.01)
    obj = fd.Fast_MPC2(c["Q"], c["R"], None, c["Qf"], None, None, None, c["x_min"], c["x_max"], c["u_min"], c["u_max"],
                       None, None, c["T"], c["x0"][0], c["x0_pre"][0], None, c["A1"], c["A2"], c["B"], c["w"][0], None,
                       z0_of(c)[0])
    z0, H, g, P, h, C, b = obj._assemble()
    dz, dnu = fd.dense_kkt_newton_step(H, g, P, h, C, b, 0.01, z0, c["nu0"][0])
    assert obj.last_stats is None
    _, st = dense_solve(fd, c, 0, 1, 0.01)
    assert st["halvings"] == [0]
    assert relerr(z1, z0 + dz) < 1e-12
    # (iii) r_p = 0 after a full Newton step
    assert np.abs(C @ z1 - b).max() < 1e-12


def test_full_step_independent_of_nu():
    c = small_problem(12, 6, 4, 5, 1, 1.0)
    z_a, _ = dense_solve(fd, c, 0, 1, 0.01)
    c["nu0"] = np.random.RandomState(99).rand(*c["nu0"].shape) * 5
    z_b, _ = dense_solve(fd, c, 0, 1, 0.01)
    assert relerr(z_a, z_b) < 1e-12


def test_inactive_bounds_reach_equality_qp_optimum(fref):
    """(i) bounds far away + tiny kappa + many iterations => KKT point of the equality-constrained QP."""
    c = small_problem(13, 6, 4, 6, 1, 1e6)
    ref = ref_solve(fref, c, 30, 1e-12)
    obj = fd.Fast_MPC2(c["Q"], c["R"], None, c["Qf"], None, None, None, c["x_min"], c["x_max"], c["u_min"], c["u_max"],
                       None, None, c["T"], c["x0"][0], c["x0_pre"][0], None, c["A1"], c["A2"], c["B"], c["w"][0], None, None)
    _, H, g, P, h, C, b = obj._assemble()
    N, p = H.shape[0], C.shape[0]
    K = np.block([[2 * H, C.T], [C, np.zeros((p, p))]])
    sol = np.linalg.solve(K, np.concatenate([-g, b]))
    z = np.concatenate([ref["U"][0], ref["X"][0]], axis=1).reshape(-1)
    assert relerr(z, sol[:N]) < 1e-9


def test_terminal_state_is_met(fref):
    """(ii) xf given => x_T = xf."""
    c = small_problem(14, 6, 8, 7, 2, 5.0, xf=True)
    ref = ref_solve(fref, c, 8, 0.01)
    assert np.abs(ref["X"][:, -1, :] - c["xf"]).max() < 1e-12


def test_var2_with_zero_A2_equals_corrected_var1(fref):
    """(v)"""
    c1 = small_problem(15, 6, 4, 5, 2, 0.4, a2=False, warm=True)
    c2 = dict(c1)
    c2["A2"] = np.zeros((6, 6))
    c2["x0_pre"] = np.random.RandomState(3).randn(2, 6)
    r1, r2 = ref_solve(fref, c1, 5, 0.01), ref_solve(fref, c2, 5, 0.01)
    assert relerr(r1["U"], r2["U"]) < 1e-13 and relerr(r1["X"], r2["X"]) < 1e-13
    z, _ = dense_solve(fd, c1, 0, 5, 0.01)                      # corrected VAR_1 dense == structured
    U, X = fd.deinterleave(z, 6, 4, 5)
    assert relerr(r1["U"][0], U.T) < TOL


def test_var1_literal_bug_changes_C():
    """VAR_1/fast_mpc_eq_const.m:34-37 writes block row 2 at column n, not m+1."""
    n, m, T = 3, 6, 3
    rs = np.random.RandomState(0)
    args = (np.eye(n), np.eye(m), None, np.eye(n), None, None, None, -np.ones(n), np.ones(n), -np.ones(m), np.ones(m),
            -np.ones(m), np.ones(m), T, rs.randn(n), np.zeros(m), 0.5 * np.eye(n) + 0.1 * rs.randn(n, n), rs.randn(n, m),
            rs.randn(T * n), None, None)
    Cb, _ = fd.Fast_MPC2_VAR1(*args, literal_bug=True).equality_const()
    Cg, _ = fd.Fast_MPC2_VAR1(*args, literal_bug=False).equality_const()
    A = args[16]
    assert np.array_equal(Cg[n:2 * n, m:m + n], -A)
    assert np.array_equal(Cb[n:2 * n, n - 1:2 * n - 1], -A)        # 1-based column n
    assert not np.array_equal(Cb, Cg)
    assert np.array_equal(Cb[2 * n:], Cg[2 * n:])                   # rows >= 3 are correct (:39-41)


def test_var1_ramp_rows_shape_and_content():
    n, m, T = 3, 2, 3
    up = np.array([0.1, -0.2])
    obj = fd.Fast_MPC2_VAR1(np.eye(n), np.eye(m), None, np.eye(n), None, None, None, -np.ones(n), np.ones(n),
                            -2 * np.ones(m), 2 * np.ones(m), -0.5 * np.ones(m), 0.5 * np.ones(m), T, np.zeros(n), up,
                            np.eye(n), np.ones((n, m)), np.zeros(T * n), None, None)
    P, h = obj.inequality_const()
    assert P.shape == (4 * T * m, T * (n + m))
    Pr, hr = P[2 * T * m:], h[2 * T * m:]
    assert np.array_equal(hr[:2 * m], np.concatenate([up + 0.5, -up + 0.5]))      # :72
    z = np.arange(T * (n + m), dtype=float)
    u = z.reshape(T, n + m)[:, :m]
    assert np.allclose(Pr[2 * m:3 * m] @ z, u[1] - u[0])                          # :66
    assert np.allclose(Pr[3 * m:4 * m] @ z, u[0] - u[1])


def test_matlab_default_stream():
    """rand after start-up = 0.8147 0.9058 0.1270 0.9134 0.6324."""
    v = fd.MatlabRand().rand(5)
    assert np.allclose(v, [0.8147, 0.9058, 0.1270, 0.9134, 0.6324], atol=5e-5)
    assert abs(v[0] - 0.8147236863931789) < 1e-15


def test_reference_error_strings():
    n, m, T = 3, 2, 2
    base = dict(Q=np.eye(n), R=np.eye(m), S=None, Qf=np.eye(n), q=None, r=None, qf=None, xmin=-np.ones(n),
                xmax=np.ones(n), umin=-np.ones(m), umax=np.ones(m), dumin=None, dumax=None, T=T, x0=np.zeros(n),
                x0_pre=np.zeros(n), u_prev=None, A1=np.eye(n), A2=np.eye(n), B=np.ones((n, m)), w=np.zeros(T * n),
                xf=None, x_init=None)

    def mk(**kw):
        d = dict(base)
        d.update(kw)
        return fd.Fast_MPC2(*d.values())

    with pytest.raises(ValueError, match="Initialization size mismatch"):
        mk(x_init=np.zeros(3)).mpc_fixed_log_newton(1, 0.01)
    with pytest.raises(ValueError, match="State stage cost must a square matrix"):
        mk(Q=np.ones((n, n + 1))).mpc_fixed_log_newton(1, 0.01)
    with pytest.raises(ValueError, match="Check cotrol iequality"):     # reachable only with a warm start
        mk(umin=-np.ones(m + 1), x_init=np.zeros(T * (n + m))).mpc_fixed_log_newton(1, 0.01)
    with pytest.raises(ValueError, match="Define the state dynamics"):
        mk(A2=None).mpc_fixed_log_newton(1, 0.01)
    with pytest.raises(ValueError, match="equality state dynamics matrix size"):
        mk(x0=np.zeros(n + 1)).mpc_fixed_log_newton(1, 0.01)
    with pytest.raises(IndexError):
        mk(w=None).mpc_fixed_log_newton(1, 0.01)            # w = [] only valid for T == 1


def test_line_search_halvings_are_exercised(fref):
    """Tight bounds + a warm start hugging them: backtracking must actually halve.  Both regimes of
    SURVEY.md F6 occur (a few genuine halvings, or ~50 per step when t collapses to roundoff and the
    iterate stays put); the two restatements must agree on z in both (and on the halving count in the first)."""
    genuine = saturated = 0
    for seed in range(40, 70):
        c = small_problem(seed, 6, 5, 6, 1, 0.05, warm=True)
        c["U0"] = np.clip(c["U0"] * 10, -0.0499, 0.0499)
        ref = ref_solve(fref, c, 4, 0.01)
        z, st = dense_solve(fd, c, 0, 4, 0.01)
        U, X = fd.deinterleave(z, c["n"], c["m"], c["T"])
        if max(st["halvings"]) < 40:        # in the saturated regime the count itself is roundoff-level
            assert ref["halvings"][0] == sum(st["halvings"])
        assert relerr(ref["U"][0], U.T) < TOL and relerr(ref["X"][0], X.T) < TOL
        genuine += any(1 <= h <= 12 for h in st["halvings"])
        saturated += any(h >= 40 for h in st["halvings"])
    assert genuine >= 3 and saturated >= 3


def test_frontends_dense_kappa_schedule():
    assert fd.Fast_MPC2.kappa_schedule(3440) == pytest.approx([1.0, 0.1, 0.01, 1e-3, 1e-4, 1e-5])
    assert len(fd.Fast_MPC2.kappa_schedule(45)) == 4
