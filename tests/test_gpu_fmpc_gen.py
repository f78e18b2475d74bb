"""GPU parity tests of the general-structure kernel (VAR_1 ramp-rate rows, the literal VAR_1 placement of the
second block row of C, dense Q / Qf) through the C-ABI, against the literal dense oracle (live, at sizes it
finishes in seconds) and the committed var1lit_* golden fixtures.  Tolerance 1e-9 relative per array on U and X."""
import glob
import os

import numpy as np
import pytest

from cases import relerr, small_problem, var1_literal_case, var1_literal_dense, z0_of

pytestmark = pytest.mark.gpu
TOL = 1e-9
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
VAR1_FIXTURES = sorted(glob.glob(os.path.join(GOLD, "var1lit_*.npz")))


def gen_handle(pk, c, ramp, bug, max_batch=None):
    return pk.FastMPCBatch(c["A1"], c["A2"], c["B"], c["Q"], c["R"], c["Qf"], c["u_min"], c["u_max"], c["T"],
                           c["x_min"], c["x_max"], du_min=c.get("du_min"), du_max=c.get("du_max"), ramp_rows=ramp,
                           var1_literal_bug=bug, max_batch=max_batch or c["nb"])


def gen_solve(pk, c, niters, kappa, ramp, bug, **kw):
    hb = gen_handle(pk, c, ramp, bug)
    assert hb.kernel_kind == 3
    out = hb.step(c["x0"], c["x0_pre"], c["w"], c["xf"], c["X0"], c["U0"], c["nu0"], u_prev=c.get("u_prev"), kappa=kappa,
                  niters=niters, **kw)
    hb.close()
    return out


def check_against_dense(pk, c, niters, kappa, ramp, bug):
    from oracle import fastmpc_dense as fd
    out = gen_solve(pk, c, niters, kappa, ramp, bug)
    for b in range(c["nb"]):
        z, st = var1_literal_dense(fd, c, b, niters, kappa, ramp=ramp, bug=bug)
        U, X = fd.deinterleave(z, c["n"], c["m"], c["T"])
        assert relerr(out["U"][b], U.T) < TOL, f"U instance {b}"
        assert relerr(out["X"][b], X.T) < TOL, f"X instance {b}"
        assert out["iters"][b] == st["iters"]
        assert (out["status"][b] == 1) == bool(st["early_exit"])


GEN_CASES = [
    # seed n  m  T  nb umax  du   xf     ramp  bug   denseQ
    (301, 6, 9, 6, 3, 0.6, 0.15, False, True, True, False),     # the reference's VAR_1 as written
    (302, 6, 9, 6, 3, 0.6, 0.15, False, True, False, False),    # ramp rows, corrected C
    (303, 6, 9, 6, 3, 0.6, 0.15, False, False, True, False),    # literal C, box rows only
    (304, 7, 5, 8, 3, 0.5, 0.10, True, True, False, False),     # ramp rows + terminal row x_T = xf
    (305, 5, 12, 4, 2, 0.5, 0.10, True, True, True, False),     # literal C + xf (n < m + 1)
    (306, 9, 4, 5, 2, 0.5, 0.10, False, True, True, False),     # n > m + 1: the mis-placed row reaches into x_2
    (307, 6, 9, 6, 3, 0.6, 0.15, False, False, False, True),    # dense Q / Qf only
    (308, 8, 7, 7, 2, 0.6, 0.15, True, True, True, True),       # everything at once
    (309, 33, 20, 6, 2, 0.5, 0.10, False, True, True, False),   # n > 32
    (310, 3, 2, 3, 2, 0.5, 0.10, False, True, True, False),     # smallest horizon the literal C supports
]


@pytest.mark.parametrize("block", ["wide", "narrow"])
@pytest.mark.parametrize("cfg", GEN_CASES, ids=lambda g: f"s{g[0]}_n{g[1]}m{g[2]}T{g[3]}")
def test_gen_kernel_matches_dense_oracle(pk, cfg, block, monkeypatch):
    """Both block sizes of the kernel: batches smaller than half the resident CTAs run one 512-thread CTA per instance,
    larger ones 256-thread CTAs, two per SM (FMPC_GEN_NARROW=1 forces the latter)."""
    if block == "narrow":
        monkeypatch.setenv("FMPC_GEN_NARROW", "1")
    seed, n, m, T, nb, umax, du, xf, ramp, bug, dq = cfg
    c = var1_literal_case(seed, n, m, T, nb, umax, du, xf=xf, dense_q=dq)
    check_against_dense(pk, c, 5, 0.01, ramp, bug)


def test_gen_kernel_feasible_regime_converges(pk):
    """Small states: iterates stay inside box and ramp bounds, the early-exit test fires like the oracle's."""
    c = var1_literal_case(203, 6, 9, 6, 3, 0.6, 0.15)
    c["x0"] *= 0.2; c["w"] *= 0.2; c["X0"] *= 0.2
    c["U0"] = 0.2 * c["U0"]; c["u_prev"] = 0.2 * c["u_prev"]
    check_against_dense(pk, c, 8, 0.01, True, True)
    out = gen_solve(pk, c, 8, 0.01, True, True)
    d = np.diff(np.concatenate([c["u_prev"][:, None, :], out["U"]], axis=1), axis=1)
    assert np.abs(d).max() < 0.15 and np.abs(out["U"]).max() < 0.6
    assert (out["status"] == 1).all()


def test_gen_var2_with_ramp_rows_and_cold_start(pk):
    """ramp_rows on the two-lag model (the rows VAR_2/fast_mpc_ineq_const.m:58-82 has commented out), cold start."""
    from oracle import fastmpc_dense as fd
    c = small_problem(311, 6, 5, 6, 2, 0.8)
    c["du_min"], c["du_max"] = -0.5 * np.ones(5), 0.5 * np.ones(5)
    c["u_prev"] = 0.1 * np.random.RandomState(5).randn(2, 5)
    out = gen_solve(pk, c, 4, 0.01, True, False)
    for b in range(2):
        z, st = var1_literal_dense(fd, c, b, 4, 0.01, ramp=True, bug=False)
        U, X = fd.deinterleave(z, 6, 5, 6)
        assert relerr(out["U"][b], U.T) < TOL and relerr(out["X"][b], X.T) < TOL


@pytest.mark.parametrize("path", VAR1_FIXTURES, ids=lambda p: os.path.basename(p)[8:-4])
def test_gen_kernel_matches_golden(pk, path):
    g = np.load(path)
    c = {k: g[k] for k in g.files}
    for k in ("n", "m", "T", "nb", "niters"):
        c[k] = int(c[k])
    for k in ("A2", "x0_pre", "xf"):
        c.setdefault(k, None)
    out = gen_solve(pk, c, c["niters"], float(c["kappa"]), bool(c["ramp"]), bool(c["bug"]))
    for b in range(c["nb"]):
        assert relerr(out["U"][b], c["U"][b]) < TOL and relerr(out["X"][b], c["X"][b]) < TOL
    assert np.array_equal(out["iters"], c["iters"])


@pytest.mark.parametrize("block", ["wide", "narrow"])
def test_gen_batch_at_c1_size_matches_golden(pk, block, monkeypatch):
    """The C1-size fixture (n = 27, m = 144, T = 10, ramp rows + literal C) replicated to a batch larger than the resident CTAs:
    persistent 256-thread CTAs (two per SM, dynamic instance counter) on a full machine, and -- with a batch below half the
    resident CTAs -- the 512-thread CTA per instance.  Every copy has to reproduce the golden solution."""
    path = os.path.join(GOLD, "var1lit_readme_c1_n27_T10.npz")
    g = np.load(path)
    c = {k: g[k] for k in g.files}
    for k in ("n", "m", "T", "nb", "niters"):
        c[k] = int(c[k])
    for k in ("A2", "x0_pre", "xf"):
        c.setdefault(k, None)
    reps = 700 // c["nb"] if block == "narrow" else max(1, 8 // c["nb"])
    if block == "narrow" and reps * c["nb"] < 600:
        pytest.skip("fixture too small to fill the machine")
    big = dict(c)
    for k in ("x0", "x0_pre", "w", "xf", "X0", "U0", "nu0", "u_prev"):
        if c.get(k) is not None:
            big[k] = np.ascontiguousarray(np.concatenate([c[k]] * reps, axis=0))
    big["nb"] = reps * c["nb"]
    out = gen_solve(pk, big, c["niters"], float(c["kappa"]), bool(c["ramp"]), bool(c["bug"]))
    for b in range(big["nb"]):
        o = b % c["nb"]
        assert relerr(out["U"][b], c["U"][o]) < TOL and relerr(out["X"][b], c["X"][o]) < TOL, f"copy {b}"
    assert np.array_equal(out["iters"], np.tile(c["iters"], reps))


def test_var1_class_is_the_reference_by_default(pk):
    """Fast_MPC2_VAR1 with the reference's 21 ctor arguments = ramp rows + literal C (VAR_1/Fast_MPC2.m:26-51)."""
    from oracle import fastmpc_dense as fd
    c = var1_literal_case(312, 6, 9, 6, 1, 0.6, 0.15)
    args = (c["Q"], c["R"], None, c["Qf"], None, None, None, c["x_min"], c["x_max"], c["u_min"], c["u_max"],
            c["du_min"], c["du_max"], c["T"], c["x0"][0], c["u_prev"][0], c["A1"], c["B"], c["w"][0], None, z0_of(c)[0])
    z = pk.Fast_MPC2_VAR1(*args).mpc_fixed_log_newton(5, 0.01, nu0=c["nu0"][0])
    zr = fd.Fast_MPC2_VAR1(*args).mpc_fixed_log_newton(5, 0.01, nu0=c["nu0"][0])
    assert relerr(z, zr) < TOL
    # kappa-continuation front-end on the same path
    nouter = len(fd.Fast_MPC2.kappa_schedule(c["T"] * (c["n"] + c["m"])))
    nus = np.random.RandomState(1).rand(nouter, c["T"] * c["n"])
    z2 = pk.Fast_MPC2_VAR1(*args).mpc_fixed_newton(2, nu0=nus)
    zr2 = fd.Fast_MPC2_VAR1(*args).mpc_fixed_newton(2, nu0_list=list(nus))
    assert relerr(z2, zr2) < TOL


def test_gen_closed_loop_ramp(pk):
    """fmpc_closed_loop on the ramp-row path: u_prev of step k is U(:,0) of step k-1 (kept on the device)."""
    from mpc_sensorlessao_b200 import synth
    from oracle import fastmpc_dense as fd
    p = synth.make_problem(2, 4, var_order=1, m1=3, u_bound=5.0)
    nb, K = 2, 3
    a = synth.aberrations(p, nb, K, seed=3, amp=0.2)
    du = 0.5
    hb = pk.FastMPCBatch(p.A1, None, p.B, p.Q, p.R, p.Qf, p.u_min, p.u_max, p.T, p.x_min, p.x_max,
                         du_min=-du * np.ones(p.m), du_max=du * np.ones(p.m), ramp_rows=True, var1_literal_bug=True, max_batch=nb)
    nu0 = np.random.RandomState(9).rand(K, nb, p.T * p.n)
    out = hb.closed_loop(a, nu0=nu0, kappa=0.01, niters=3)
    hb.close()
    for b in range(nb):
        u_prev = np.zeros(p.m); x0 = np.zeros(p.n); z = None
        for k in range(K):
            x0 = a[b, k] + p.B @ u_prev
            if z is not None:     # shift the warm start one stage, last stage repeated
                Z = z.reshape(p.T, p.n + p.m)
                z = np.vstack([Z[1:], Z[-1:]]).reshape(-1)
            o = fd.Fast_MPC2_VAR1(p.Q, p.R, None, p.Qf, None, None, None, p.x_min, p.x_max, p.u_min, p.u_max,
                                  -du * np.ones(p.m), du * np.ones(p.m), p.T, x0, u_prev, p.A1, p.B, np.zeros(p.T * p.n), None, z)
            z = o.mpc_fixed_log_newton(3, 0.01, nu0=nu0[k, b])
            u_prev = z[:p.m].copy()
            assert relerr(out["U_acc"][b, k], u_prev) < TOL, (b, k)
            assert relerr(out["X_acc"][b, k], x0) < 1e-12


def test_gen_argument_errors(pk):
    c = var1_literal_case(313, 6, 9, 6, 2, 0.6, 0.15)
    hb = gen_handle(pk, c, True, True)
    with pytest.raises(pk.FmpcError) as e:       # ramp rows need u_prev
        hb.step(c["x0"], None, c["w"], None, c["X0"], c["U0"], c["nu0"], u_prev=None)
    assert e.value.code == -7
    hb.close()
    Rd = c["R"].copy(); Rd[0, 1] = Rd[1, 0] = 0.1
    with pytest.raises(pk.FmpcError) as e:       # non-diagonal R together with ramp rows is the one combination not covered
        pk.FastMPCBatch(c["A1"], None, c["B"], c["Q"], Rd, c["Qf"], c["u_min"], c["u_max"], c["T"], c["x_min"], c["x_max"],
                        du_min=c["du_min"], du_max=c["du_max"], ramp_rows=True)
    assert e.value.code == -14
    c2 = var1_literal_case(314, 6, 9, 2, 1, 0.6, 0.15)
    with pytest.raises(pk.FmpcError) as e:       # literal C with T < 3: fast_mpc_eq_const.m:55 rewrites the row itself
        gen_handle(pk, c2, False, True)
    assert e.value.code == -14


def dense_spd(rs, m, scale=1.0):
    G = rs.randn(m, m)
    return scale * (np.eye(m) + 0.3 * G @ G.T / m)


@pytest.mark.parametrize("cfg", [(401, 6, 5, 4, 3, 0.6, True, False, False), (402, 8, 12, 6, 2, 0.4, True, True, False),
                                 (403, 7, 9, 5, 2, 0.5, False, False, True), (404, 12, 33, 7, 2, 0.5, True, False, True),
                                 (405, 28, 144, 3, 1, 3.0, True, False, False)],
                         ids=lambda g: f"s{g[0]}_n{g[1]}m{g[2]}T{g[3]}")
def test_dense_R_matches_dense_oracle(pk, cfg):
    """fast_mpc_objective.m:20-21 accepts any square R: a dense SPD R makes Phi_uu of every stage a dense m x m matrix
    (general-structure kernel: packed Cholesky + explicit inverse per stage in shared memory).  VAR(2) and VAR(1), with and
    without the terminal row, with dense Q as well, and at the README's input dimension m = 144."""
    from oracle import fastmpc_dense as fd
    from cases import dense_solve
    seed, n, m, T, nb, umax, a2, xf, denseq = cfg
    c = small_problem(seed, n, m, T, nb, umax, a2=a2, xf=xf, warm=True)
    rs = np.random.RandomState(seed)
    c["R"] = dense_spd(rs, m)
    if denseq:
        c["Q"] = dense_spd(rs, n, 3.0)
        c["Qf"] = 2.0 * c["Q"]
    niters = 4 if m < 100 else 2
    hb = gen_handle(pk, c, False, False)
    assert hb.kernel_kind == 3
    out = hb.step(c["x0"], c["x0_pre"], c["w"], c["xf"], c["X0"], c["U0"], c["nu0"], kappa=0.01, niters=niters)
    hb.close()
    for b in range(nb):
        z, st = dense_solve(fd, c, b, niters, 0.01)
        U, X = fd.deinterleave(z, n, m, T)
        assert relerr(out["U"][b], U.T) < TOL, f"U instance {b}"
        assert relerr(out["X"][b], X.T) < TOL, f"X instance {b}"
        assert out["iters"][b] == st["iters"]


def test_dense_R_errors(pk):
    c = small_problem(410, 5, 4, 3, 1, 0.5)
    c["R"] = np.array([[1.0, 2.0, 0, 0], [2.0, 1.0, 0, 0], [0, 0, 1.0, 0], [0, 0, 0, 1.0]])       # symmetric, indefinite
    with pytest.raises(pk.FmpcError) as e:
        gen_handle(pk, c, False, False)
    assert e.value.code == -13
    c["R"] = dense_spd(np.random.RandomState(0), 4)
    c["du_min"], c["du_max"] = -np.ones(4), np.ones(4)
    with pytest.raises(pk.FmpcError) as e:                     # dense R together with ramp rows is not covered
        gen_handle(pk, c, True, False)
    assert e.value.code == -14
