"""GPU parity tests of the fastMPC CUDA path, through the C-ABI (ctypes), against
  * the structured C oracle on seeded inputs (sizes the oracle finishes in seconds),
  * the committed golden fixtures (literal dense oracle),
  * size-independent properties at BASELINE.json's full size (4096 instances, n=28, m=144, T=20).
Tolerance (north-star): 1e-9 relative per array on U and X."""
import numpy as np
import pytest

from cases import ref_solve, relerr, small_problem
from test_golden_cpu import FMPC_FIXTURES, load_case

pytestmark = pytest.mark.gpu
TOL = 1e-9


def make_handle(pk, c, max_batch=None):
    return pk.FastMPCBatch(c["A1"], c["A2"], c["B"], c["Q"], c["R"], c["Qf"], c["u_min"], c["u_max"], c["T"],
                           c["x_min"], c["x_max"], max_batch=max_batch or c["nb"])


def gpu_solve(pk, c, niters, kappa, hb=None, **kw):
    own = hb is None
    hb = hb or make_handle(pk, c)
    out = hb.step(c["x0"], c["x0_pre"], c["w"], c["xf"], c["X0"], c["U0"], c["nu0"], kappa=kappa, niters=niters, **kw)
    if own:
        hb.close()
    return out


def assert_parity(out, ref, nb):
    for b in range(nb):
        assert relerr(out["U"][b], ref["U"][b]) < TOL, f"U instance {b}"
        assert relerr(out["X"][b], ref["X"][b]) < TOL, f"X instance {b}"


SMALL = [
    dict(seed=1, n=6, m=4, T=5, nb=3, umax=2.0),
    dict(seed=2, n=6, m=4, T=5, nb=3, umax=0.3, xf=True),
    dict(seed=3, n=6, m=4, T=5, nb=3, umax=0.3, a2=False),
    dict(seed=4, n=8, m=5, T=10, nb=5, umax=0.2, xf=True, warm=True),
    dict(seed=5, n=8, m=5, T=10, nb=5, umax=0.1, warm=True),
    dict(seed=6, n=5, m=7, T=1, nb=2, umax=0.5),                      # T = 1: single block row
    dict(seed=7, n=5, m=7, T=2, nb=2, umax=0.5, xf=True),
    dict(seed=8, n=33, m=20, T=6, nb=4, umax=0.5, warm=True),         # n > 32: multi-row-per-lane paths
    dict(seed=9, n=1, m=1, T=3, nb=2, umax=1.0),                      # degenerate sizes
    dict(seed=10, n=27, m=144, T=10, nb=4, umax=3.0, a2=False, warm=True),   # C1 shape, VAR(1)
    dict(seed=11, n=28, m=144, T=20, nb=6, umax=0.5, warm=True),      # C2 shape, active barrier
    dict(seed=12, n=66, m=144, T=30, nb=2, umax=1.0, warm=True),      # C5 shape
    # every instantiation of the warp-per-instance kernel (block size classes 8/16/24/28/32, with and without a
    # spare padding row for the forward-substitution row), odd m, T across 8-stage tile boundaries
    dict(seed=13, n=12, m=9, T=7, nb=3, umax=0.4, warm=True),
    dict(seed=14, n=16, m=11, T=9, nb=3, umax=0.4, warm=True, xf=True),
    dict(seed=15, n=20, m=17, T=26, nb=3, umax=0.4, warm=True),
    dict(seed=16, n=24, m=8, T=4, nb=2, umax=0.4, warm=True),
    dict(seed=17, n=30, m=33, T=5, nb=2, umax=0.4, warm=True),
    dict(seed=18, n=32, m=16, T=5, nb=2, umax=0.4, warm=True, xf=True),
    dict(seed=19, n=28, m=144, T=20, nb=40, umax=0.5, warm=True, xf=True),   # more instances than one CTA holds
    dict(seed=20, n=8, m=6, T=17, nb=9, umax=0.3, a2=False, xf=True, warm=True),
    # 32 < n <= 72: the 256-thread CTA DMMA kernel (whole-CTA potrf + inverse)
    dict(seed=21, n=72, m=10, T=4, nb=2, umax=0.4, warm=True, xf=True),
    dict(seed=22, n=40, m=50, T=9, nb=3, umax=0.4, a2=False, warm=True),
    dict(seed=23, n=65, m=30, T=3, nb=2, umax=0.5),
    dict(seed=25, n=66, m=144, T=30, nb=150, umax=1.0, warm=True),   # C5 shape, more instances than SMs
    # NPOT = 28 (n = 25..28): the last tile row of B diag(w) B' is shared by two consecutive stages -- odd and even horizons,
    # every n of the class, with and without the terminal row, VAR(1)
    dict(seed=26, n=25, m=33, T=7, nb=3, umax=0.4, warm=True),
    dict(seed=27, n=26, m=20, T=4, nb=3, umax=0.4, warm=True, xf=True),
    dict(seed=28, n=28, m=12, T=1, nb=2, umax=0.5),
    dict(seed=29, n=27, m=9, T=3, nb=9, umax=0.3, a2=False, warm=True, xf=True),
    dict(seed=30, n=28, m=144, T=21, nb=8, umax=0.5, warm=True),
]


@pytest.mark.parametrize("kw", SMALL, ids=lambda k: f"s{k['seed']}_n{k['n']}m{k['m']}T{k['T']}")
def test_parity_vs_structured_oracle(pk, fref, kw):
    c = small_problem(**kw)
    out = gpu_solve(pk, c, 6, 0.01)
    ref = ref_solve(fref, c, 6, 0.01)
    assert_parity(out, ref, c["nb"])
    assert np.array_equal(out["iters"], ref["iters"])
    assert np.array_equal(out["status"], ref["status"])


@pytest.mark.parametrize("path", FMPC_FIXTURES, ids=lambda p: p.split("fmpc_")[-1][:-4])
def test_parity_vs_golden(pk, path):
    c = load_case(path)
    out = gpu_solve(pk, c, c["niters"], c["kappa"])
    assert_parity(out, c, c["nb"])
    assert np.array_equal(out["iters"], c["iters"])
    assert np.array_equal(out["status"] == 1, c["early_exit"])


def test_kernel_selection(pk):
    """n <= 32 runs on the warp-per-instance DMMA kernel, 32 < n <= 72 on the CTA-per-instance DMMA kernel (never a
    CPU path); larger blocks do not fit the shared-memory block chain and go to the general-structure kernel."""
    for n, kind in ((6, 2), (28, 2), (32, 2), (33, 1), (66, 1), (72, 1)):
        c = small_problem(1, n, 4, 3, 1, 1.0)
        hb = make_handle(pk, c)
        assert hb.kernel_kind == kind
        hb.close()
    hb = make_handle(pk, small_problem(1, 73, 4, 3, 1, 1.0))          # n > 72: the general-structure kernel takes over
    assert hb.kernel_kind == 3
    hb.close()


def test_large_state_dimension_runs_on_the_general_kernel(pk, fref):
    """The reference has no size limit (fast_mpc_eq_const.m:14); n > 72 does not fit the stage blocks of the block-banded
    kernels and is solved by the general-structure kernel (dense Schur complement)."""
    for kw in (dict(seed=71, n=80, m=10, T=4, nb=2, umax=0.5, warm=True), dict(seed=72, n=100, m=12, T=3, nb=2, umax=0.4, xf=True)):
        c = small_problem(**kw)
        out = gpu_solve(pk, c, 4, 0.01)
        ref = ref_solve(fref, c, 4, 0.01)
        assert_parity(out, ref, c["nb"])
        assert np.array_equal(out["iters"], ref["iters"])


def test_line_search_both_regimes(pk, fref):
    """Tight bounds: genuine halvings and the FP-saturated regime."""
    for seed in (59, 61, 63, 67, 40, 41):
        c = small_problem(seed, 6, 5, 6, 1, 0.05, warm=True)
        c["U0"] = np.clip(c["U0"] * 10, -0.0499, 0.0499)
        out = gpu_solve(pk, c, 4, 0.01)
        ref = ref_solve(fref, c, 4, 0.01)
        assert ref["halvings"][0] > 0
        assert_parity(out, ref, 1)


def test_ls_max_is_reported(pk):
    c = small_problem(40, 6, 5, 6, 1, 0.05, warm=True)
    c["U0"] = np.clip(c["U0"] * 10, -0.0499, 0.0499)
    out = gpu_solve(pk, c, 2, 0.01, ls_max=5)
    assert out["status"][0] == 3


def test_not_pd_is_reported_per_instance_not_fatal(pk):
    """Iterate outside the box far enough that a barrier weight goes hugely negative -> chol(Schur) fails
    for that instance only (the reference would throw, inf_newton_solver.m:30)."""
    c = small_problem(21, 6, 4, 5, 3, 0.5, warm=True)
    out_ok = gpu_solve(pk, c, 3, 0.01)
    c["U0"] = c["U0"].copy()
    c["U0"][1] = np.nan
    out = gpu_solve(pk, c, 3, 0.01)
    assert out["status"][1] in (2, 4)
    assert relerr(out["U"][0], out_ok["U"][0]) == 0.0 and relerr(out["U"][2], out_ok["U"][2]) == 0.0


def test_empty_batch_and_batch_limit(pk):
    c = small_problem(22, 4, 3, 3, 2, 1.0)
    hb = make_handle(pk, c)
    e = hb.step(np.zeros((0, 4)), np.zeros((0, 4)), None, None, None, None, np.zeros((0, 12)))
    assert e["U"].shape == (0, 3, 3)
    c3 = small_problem(22, 4, 3, 3, 3, 1.0)
    with pytest.raises(pk.FmpcError) as ei:
        gpu_solve(pk, c3, 2, 0.01, hb=hb)
    assert ei.value.code == -15
    hb.close()


def test_matlab_stream_default_nu(pk, fref):
    """nu0 = NULL draws rand(length(b),1) per instance from MT19937(5489), instance after instance."""
    c = small_problem(23, 6, 4, 5, 3, 0.3, warm=True)
    nu = np.random.RandomState(5489).random_sample(c["nu0"].size * 2).reshape(2, *c["nu0"].shape)
    hb = make_handle(pk, c)
    for call in range(2):
        c["nu0"] = None
        out = gpu_solve(pk, c, 1, 0.01, hb=hb)        # 1 step: z independent of nu, but iters/status are not
        c["nu0"] = nu[call]
        ref = ref_solve(fref, c, 1, 0.01)
        assert_parity(out, ref, 3)
    hb.close()
    # a 2-step solve depends on nu through the accept/exit tests; same stream => same result
    hb = make_handle(pk, c)
    c["nu0"] = None
    out = gpu_solve(pk, c, 4, 0.01, hb=hb)
    c["nu0"] = nu[0]
    assert_parity(out, ref_solve(fref, c, 4, 0.01), 3)
    hb.close()


def test_interleaved_entry_point_and_class_mirror(pk, fref):
    """Fast_MPC2(...).mpc_fixed_log_newton(nw, k) -- the reference's own call shape (README.md:548-555)."""
    from oracle import fastmpc_dense as fd
    from mpc_sensorlessao_b200 import fast_mpc2
    c = small_problem(24, 7, 5, 6, 1, 0.4, warm=False)
    args = (c["Q"], c["R"], None, c["Qf"], None, None, None, c["x_min"], c["x_max"], c["u_min"], c["u_max"],
            -np.ones(5), np.ones(5), c["T"], c["x0"][0], c["x0_pre"][0], np.zeros(5), c["A1"], c["A2"], c["B"], c["w"][0],
            None, None)
    fast_mpc2.reset_matlab_stream()
    z_gpu = pk.Fast_MPC2(*args).mpc_fixed_log_newton(5, 0.01)
    z_gpu2 = pk.Fast_MPC2(*args).mpc_fixed_log_newton(5, 0.01)      # second object: next draws of the session stream
    o = fd.Fast_MPC2(*args)
    o.stream = fd.MatlabRand()
    z_ref = o.mpc_fixed_log_newton(5, 0.01)
    z_ref2 = o.mpc_fixed_log_newton(5, 0.01)
    assert relerr(z_gpu, z_ref) < TOL and relerr(z_gpu2, z_ref2) < TOL
    # warm start through x_init
    z_w = pk.Fast_MPC2(*args[:-1], z_ref).mpc_fixed_log_newton(2, 0.01, nu0=np.full(c["T"] * 7, 0.5))
    o2 = fd.Fast_MPC2(*args[:-1], z_ref)
    assert relerr(z_w, o2.mpc_fixed_log_newton(2, 0.01, nu0=np.full(c["T"] * 7, 0.5))) < TOL


def test_var1_class_mirror(pk):
    from oracle import fastmpc_dense as fd
    c = small_problem(25, 6, 4, 5, 1, 0.4, a2=False)
    args = (c["Q"], c["R"], None, c["Qf"], None, None, None, c["x_min"], c["x_max"], c["u_min"], c["u_max"],
            -np.ones(4), np.ones(4), c["T"], c["x0"][0], np.zeros(4), c["A1"], c["B"], c["w"][0], None, None)
    z = pk.Fast_MPC2_VAR1(*args, ramp_rows=False, literal_bug=False).mpc_fixed_log_newton(5, 0.01, nu0=c["nu0"][0])
    o = fd.Fast_MPC2_VAR1(*args, literal_bug=False)
    o.inequality_const = lambda: fd.fast_mpc_ineq_const_var2(o)
    assert relerr(z, o.mpc_fixed_log_newton(5, 0.01, nu0=c["nu0"][0])) < TOL


def test_frontends_kappa_continuation(pk):
    """mpc_fixed_newton / mpc_solve_full / mpc_solve_check / mpc_fixed_log (VAR_2/Fast_MPC2.m:88-144)."""
    from oracle import fastmpc_dense as fd
    c = small_problem(26, 6, 4, 5, 1, 0.6)
    args = (c["Q"], c["R"], None, c["Qf"], None, None, None, c["x_min"], c["x_max"], c["u_min"], c["u_max"],
            -np.ones(4), np.ones(4), c["T"], c["x0"][0], c["x0_pre"][0], np.zeros(4), c["A1"], c["A2"], c["B"], c["w"][0],
            None, None)
    nouter = len(fd.Fast_MPC2.kappa_schedule(c["T"] * 10))
    nus = np.random.RandomState(7).rand(max(nouter, 5), 30)
    g, o = pk.Fast_MPC2(*args), fd.Fast_MPC2(*args)
    assert relerr(g.mpc_fixed_newton(3, nu0=nus[:nouter]), o.mpc_fixed_newton(3, nu0_list=list(nus[:nouter]))) < TOL
    assert relerr(g.mpc_solve_full(nu0=nus[:nouter]), o.mpc_solve_full(nu0_list=list(nus[:nouter]))) < TOL
    assert relerr(g.mpc_solve_check(0.01, 1.0, nu0=nus[:5]), o.mpc_solve_check(0.01, 1.0, nu0_list=list(nus[:5]))) < TOL
    assert relerr(g.mpc_fixed_log(0.01, nu0=nus[0]), o.mpc_fixed_log(0.01, nu0=nus[0])) < TOL


def test_state_update(pk):
    c = small_problem(27, 9, 6, 3, 17, 1.0)
    hb = make_handle(pk, c)
    rs = np.random.RandomState(0)
    x, xp, u, w = rs.randn(17, 9), rs.randn(17, 9), rs.randn(17, 6), rs.randn(17, 9)
    out = hb.state_update(x, xp, u, w)
    ref = x @ c["A1"].T + xp @ c["A2"].T + u @ c["B"].T + w
    assert relerr(out, ref) < 1e-13
    assert relerr(hb.state_update(x, xp, u), ref - w) < 1e-13
    hb.close()


def test_closed_loop_matches_host_driven_loop(pk, fref):
    """fmpc_closed_loop == the same loop driven from the host through the oracle."""
    from mpc_sensorlessao_b200 import synth
    p = synth.make_problem(3, 6, m1=4)            # n = 10, m = 16
    nb, K = 3, 5
    a = synth.aberrations(p, nb, K, seed=5)
    nu = np.random.RandomState(1).rand(K, nb, p.T * p.n)
    hb = pk.FastMPCBatch(p.A1, p.A2, p.B, p.Q, p.R, p.Qf, p.u_min, p.u_max, p.T, p.x_min, p.x_max, max_batch=nb)
    out = hb.closed_loop(a, nu0=nu, niters=3)
    hb.close()
    u_prev = np.zeros((nb, p.m)); x0 = np.zeros((nb, p.n)); U = X = None
    for k in range(K):
        x0_pre = x0 if k else np.zeros((nb, p.n))
        x0 = a[:, k] + u_prev @ p.B.T
        c = dict(n=p.n, m=p.m, T=p.T, nb=nb, A1=p.A1, A2=p.A2, B=p.B, Q=p.Q, R=p.R, Qf=p.Qf, u_min=p.u_min, u_max=p.u_max,
                 x_min=p.x_min, x_max=p.x_max, x0=x0, x0_pre=x0_pre, w=None, xf=None, nu0=nu[k],
                 X0=None if k == 0 else np.concatenate([X[:, 1:], X[:, -1:]], axis=1),
                 U0=None if k == 0 else np.concatenate([U[:, 1:], U[:, -1:]], axis=1))
        ref = ref_solve(fref, c, 3, 0.01)
        U, X = ref["U"], ref["X"]
        u_prev = U[:, 0]
        assert relerr(out["U_acc"][:, k], u_prev) < 1e-8, k
        assert relerr(out["X_acc"][:, k], x0) < 1e-8, k
        assert np.array_equal(out["iters"][:, k], ref["iters"])


def test_resident_steps_match_oracle_driven_loop(pk, fref):
    """K `fmpc_step_r` calls (state on the device: only x0 in, U(:,0) out) == the loop driven from the host through the
    oracle with the warm start shifted by hand (README.md:444-626 around the solve)."""
    from mpc_sensorlessao_b200 import synth
    p = synth.make_problem(3, 6, m1=4)            # n = 10, m = 16
    nb, K = 5, 6
    a = synth.aberrations(p, nb, K, seed=7)
    nu = np.random.RandomState(2).rand(K, nb, p.T * p.n)
    hb = pk.FastMPCBatch(p.A1, p.A2, p.B, p.Q, p.R, p.Qf, p.u_min, p.u_max, p.T, p.x_min, p.x_max, max_batch=nb)
    with pytest.raises(pk.FmpcError) as ei:       # nothing resident yet: the first call must reset
        hb.step_resident(a[:, 0], nu0=nu[0], niters=3)
    assert ei.value.code == -12
    u_prev = np.zeros((nb, p.m)); x0 = np.zeros((nb, p.n)); U = X = None
    for k in range(K):
        x0_pre = x0 if k else np.zeros((nb, p.n))
        x0 = a[:, k] + u_prev @ p.B.T
        out = hb.step_resident(x0, nu0=nu[k], reset=(k == 0), full=(k % 2 == 1), niters=3)
        c = dict(n=p.n, m=p.m, T=p.T, nb=nb, A1=p.A1, A2=p.A2, B=p.B, Q=p.Q, R=p.R, Qf=p.Qf, u_min=p.u_min, u_max=p.u_max,
                 x_min=p.x_min, x_max=p.x_max, x0=x0, x0_pre=x0_pre, w=None, xf=None, nu0=nu[k],
                 X0=None if k == 0 else np.concatenate([X[:, 1:], X[:, -1:]], axis=1),
                 U0=None if k == 0 else np.concatenate([U[:, 1:], U[:, -1:]], axis=1))
        ref = ref_solve(fref, c, 3, 0.01)
        U, X = ref["U"], ref["X"]
        u_prev = U[:, 0]
        assert relerr(out["u0"], u_prev) < 1e-8, k
        assert np.array_equal(out["iters"], ref["iters"]) and np.array_equal(out["status"], ref["status"])
        if "U" in out:
            assert relerr(out["U"], U) < 1e-8 and relerr(out["X"], X) < 1e-8
            assert np.array_equal(out["U"][:, 0], out["u0"])
    hb.close()


@pytest.mark.parametrize("shape", [(28, 144, 20, 300), (40, 12, 5, 4)], ids=["warp_kernel", "cta_kernel"])
def test_resident_step_equals_full_surface_step(pk, shape):
    """Bit-for-bit: fmpc_step_r == fmpc_step fed with the previous outputs shifted one stage and the previous x0
    (both the fused warp-kernel variant and the separate shift / extract kernels of the other solve kernels)."""
    n, m, T, nb = shape
    c = small_problem(31, n, m, T, nb, 0.5, warm=False)
    c["w"] = None
    rs = np.random.RandomState(3)
    hb = make_handle(pk, c)
    hr = make_handle(pk, c)
    X = U = None
    x0_pre = np.zeros((nb, n))
    for k in range(3):
        x0 = rs.randn(nb, n) * 0.3
        nu = rs.rand(nb, T * n)
        ref = hb.step(x0, x0_pre, None, None, None if k == 0 else np.concatenate([X[:, 1:], X[:, -1:]], axis=1),
                      None if k == 0 else np.concatenate([U[:, 1:], U[:, -1:]], axis=1), nu, niters=4)
        out = hr.step_resident(x0, nu0=nu, reset=(k == 0), full=True, niters=4)
        X, U = ref["X"], ref["U"]
        assert np.array_equal(out["U"], U) and np.array_equal(out["X"], X) and np.array_equal(out["u0"], U[:, 0])
        assert np.array_equal(out["iters"], ref["iters"])
        x0_pre = x0
    hb.close(); hr.close()


def test_pinned_and_pageable_host_buffers_agree(pk):
    """fmpc_step stages pageable caller buffers through its own pinned ring (copy threads); pinned buffers go to the DMA
    engines directly.  Same bits either way, for a batch cut into several chunks."""
    import ctypes as C
    import torch
    c = small_problem(51, 28, 144, 20, 1300, 0.5, warm=True)
    hb = make_handle(pk, c)
    ref = gpu_solve(pk, c, 3, 0.01, hb=hb)                      # numpy arrays: pageable
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    t = {k: pin(c[k]) for k in ("x0", "x0_pre", "w", "X0", "U0", "nu0")}
    X, U = pin(np.zeros_like(ref["X"])), pin(np.zeros_like(ref["U"]))
    it = torch.zeros(c["nb"], dtype=torch.int32).pin_memory()
    p = hb.params(0.01, 3, 0)
    vp = lambda a: C.c_void_p(a.data_ptr())
    rc = hb._L.fmpc_step(hb._h, C.byref(p), c["nb"], vp(t["x0"]), vp(t["x0_pre"]), None, vp(t["w"]), None, vp(t["X0"]), vp(t["U0"]),
                         vp(t["nu0"]), vp(X), vp(U), None, vp(it), None)
    assert rc == 0
    assert np.array_equal(X.numpy(), ref["X"]) and np.array_equal(U.numpy(), ref["U"]) and np.array_equal(it.numpy(), ref["iters"])
    hb.close()


def test_matlab_stream_on_device_across_batch_sizes(pk):
    """nu0 = NULL consumes ONE MT19937(5489) stream in order whatever the sequence of calls and batch sizes
    (device generator running one call ahead, left-overs carried over): every call equals the explicit-nu0 call."""
    c = small_problem(41, 6, 4, 5, 9, 0.3, warm=True)
    T, n = c["T"], c["n"]
    stream = np.random.RandomState(5489).random_sample(64 * T * n)
    hb = make_handle(pk, c)
    hx = make_handle(pk, c)
    pos = 0
    for nbk in (9, 9, 4, 9, 1, 7, 7):
        sub = {k: (v[:nbk] if isinstance(v, np.ndarray) and v.ndim >= 1 and v.shape[0] == c["nb"] and k in
                   ("x0", "x0_pre", "w", "X0", "U0") else v) for k, v in c.items()}
        sub["nb"] = nbk
        sub["nu0"] = None
        out = gpu_solve(pk, sub, 4, 0.01, hb=hb)
        sub["nu0"] = stream[pos:pos + nbk * T * n].reshape(nbk, T * n)
        pos += nbk * T * n
        ref = gpu_solve(pk, sub, 4, 0.01, hb=hx)
        assert np.array_equal(out["U"], ref["U"]) and np.array_equal(out["iters"], ref["iters"])
    # the closed loop draws from the same stream: K steps of nb instances each
    from mpc_sensorlessao_b200 import synth
    p = synth.make_problem(3, 6, m1=4)
    a = synth.aberrations(p, 3, 4, seed=5)
    h1 = pk.FastMPCBatch(p.A1, p.A2, p.B, p.Q, p.R, p.Qf, p.u_min, p.u_max, p.T, p.x_min, p.x_max, max_batch=3)
    h2 = pk.FastMPCBatch(p.A1, p.A2, p.B, p.Q, p.R, p.Qf, p.u_min, p.u_max, p.T, p.x_min, p.x_max, max_batch=3)
    o1 = h1.closed_loop(a, nu0=None, niters=3)
    o2 = h2.closed_loop(a, nu0=np.random.RandomState(5489).random_sample(4 * 3 * p.T * p.n).reshape(4, 3, -1), niters=3)
    assert np.array_equal(o1["U_acc"], o2["U_acc"]) and np.array_equal(o1["iters"], o2["iters"])
    h1.close(); h2.close(); hb.close(); hx.close()


# ---- BASELINE.json full size: properties that need no oracle ---------------------------------
@pytest.fixture(scope="module")
def c2(pk):
    from mpc_sensorlessao_b200 import synth
    p = synth.make_problem(6, 20)
    nb = 4096
    wi = synth.warm_inputs(p, nb)
    c = dict(n=p.n, m=p.m, T=p.T, nb=nb, A1=p.A1, A2=p.A2, B=p.B, Q=p.Q, R=p.R, Qf=p.Qf, u_min=p.u_min, u_max=p.u_max,
             x_min=p.x_min, x_max=p.x_max, x0=wi["x0"], x0_pre=wi["x0_pre"], w=None, xf=None, nu0=wi["nu0"],
             X0=wi["X0"], U0=wi["U0"])
    hb = make_handle(pk, c)
    out = gpu_solve(pk, c, 5, 0.01, hb=hb)
    yield c, hb, out
    hb.close()


def test_c2_equality_constraints_hold(c2):
    """r_p = 0 after a full Newton step: x_{i+1} = A1 x_i + A2 x_{i-1} + B u_i for every instance and stage."""
    c, _, out = c2
    X, U = out["X"], out["U"]
    xs = np.concatenate([c["x0_pre"][:, None], c["x0"][:, None], X], axis=1)      # x_{-1}, x_0, x_1..x_T
    pred = xs[:, 1:-1] @ c["A1"].T + xs[:, :-2] @ c["A2"].T + U @ c["B"].T
    assert np.abs(pred - X).max() < 1e-9 * max(1.0, np.abs(X).max())
    assert (np.abs(U) < 28.0).all() and np.isin(out["status"], (0, 1)).all() and (out["iters"] >= 1).all()


def test_c2_sample_against_oracle(c2, fref):
    c, _, out = c2
    idx = np.random.RandomState(0).choice(c["nb"], 48, replace=False)
    sub = dict(c)
    for k in ("x0", "x0_pre", "nu0", "X0", "U0"):
        sub[k] = c[k][idx]
    sub["nb"] = len(idx)
    ref = ref_solve(fref, sub, 5, 0.01)
    for j, b in enumerate(idx):
        assert relerr(out["U"][b], ref["U"][j]) < TOL and relerr(out["X"][b], ref["X"][j]) < TOL
    assert np.array_equal(out["iters"][idx], ref["iters"])


def test_c2_permutation_and_batch_size_invariance(pk, c2):
    """An instance's result does not depend on its position in the batch or on the batch size (bit-exact)."""
    c, hb, out = c2
    perm = np.random.RandomState(1).permutation(c["nb"])
    sub = dict(c)
    for k in ("x0", "x0_pre", "nu0", "X0", "U0"):
        sub[k] = c[k][perm]
    out_p = gpu_solve(pk, sub, 5, 0.01, hb=hb)
    assert np.array_equal(out_p["U"], out["U"][perm]) and np.array_equal(out_p["X"], out["X"][perm])
    for k in ("x0", "x0_pre", "nu0", "X0", "U0"):
        sub[k] = c[k][:7]
    sub["nb"] = 7
    out_s = gpu_solve(pk, sub, 5, 0.01, hb=hb)
    assert np.array_equal(out_s["U"], out["U"][:7])


def test_c2_full_step_is_independent_of_nu(pk, c2):
    c, hb, _ = c2
    sub = dict(c)
    for k in ("x0", "x0_pre", "nu0", "X0", "U0"):
        sub[k] = c[k][:64]
    sub["nb"] = 64
    a = gpu_solve(pk, sub, 1, 0.01, hb=hb)
    sub["nu0"] = sub["nu0"][::-1].copy() * 3.0
    b = gpu_solve(pk, sub, 1, 0.01, hb=hb)
    assert relerr(a["U"], b["U"]) < 1e-11 and relerr(a["X"], b["X"]) < 1e-9
