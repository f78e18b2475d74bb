"""CPU tests of the host-side mirror (argument checks with the reference's error strings,
layout helpers, synthetic generators, import hygiene)."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _mk(pk, **kw):
    n, m, T = 3, 2, 2
    base = dict(Q=np.eye(n), R=np.eye(m), S=None, Qf=np.eye(n), q=None, r=None, qf=None, xmin=-np.ones(n),
                xmax=np.ones(n), umin=-np.ones(m), umax=np.ones(m), dumin=None, dumax=None, T=T, x0=np.zeros(n),
                x0_pre=np.zeros(n), u_prev=None, A1=np.eye(n), A2=np.eye(n), B=np.ones((n, m)), w=np.zeros(T * n),
                xf=None, x_init=None)
    base.update(kw)
    return pk.Fast_MPC2(*base.values())


@pytest.mark.parametrize("kw,exc,msg", [
    (dict(x_init=np.zeros(3)), ValueError, "Initialization size mismatch"),
    (dict(Q=np.ones((3, 4))), ValueError, "State stage cost must a square matrix"),
    (dict(R=np.ones((2, 3))), ValueError, "Control stage cost must a square matrix"),
    (dict(q=np.zeros(5)), ValueError, "Linear state cost"),
    (dict(xmin=-np.ones(4)), ValueError, "state inequality constraints"),
    (dict(umin=-np.ones(3)), ValueError, "Check cotrol iequality"),
    (dict(A2=None), ValueError, "Define the state dynamics"),
    (dict(B=None), ValueError, "Define the control dynamics"),
    (dict(x0=np.zeros(4)), ValueError, "equality state dynamics matrix size"),
    (dict(x0_pre=None), ValueError, "equality state dynamics matrix size"),
    (dict(w=None), IndexError, "w"),
])
def test_mirror_raises_reference_errors_before_touching_the_gpu(pk, kw, exc, msg):
    with pytest.raises(exc, match=msg):
        _mk(pk, **kw).mpc_fixed_log_newton(1, 0.01)


def test_interleave_roundtrip(pk):
    T, n, m = 4, 3, 2
    z = np.arange(T * (n + m), dtype=float)
    U, X = pk.deinterleave(z, n, m, T)
    assert U.shape == (T, m) and X.shape == (T, n)
    assert np.array_equal(U[1], z[(n + m):(n + m) + m]) and np.array_equal(X[0], z[m:m + n])
    assert np.array_equal(pk.interleave(U, X), z)


def test_cold_start_midpoint(pk):
    o = _mk(pk, umin=np.array([-1.0, 0.0]), umax=np.array([3.0, 2.0]), xmin=-2 * np.ones(3), xmax=4 * np.ones(3))
    z = o.initialize().reshape(2, 5)
    assert np.array_equal(z[:, :2], [[1, 1], [1, 1]]) and np.all(z[:, 2:] == 1.0)


def test_synth_problem_is_stable_and_deterministic(pk):
    from mpc_sensorlessao_b200 import synth
    p = synth.make_problem(6, 20)
    q = synth.make_problem(6, 20)
    assert (p.n, p.m, p.T) == (28, 144, 20) and np.array_equal(p.A1, q.A1) and np.array_equal(p.B, q.B)
    comp = np.block([[p.A1, p.A2], [np.eye(28), np.zeros((28, 28))]])
    assert np.abs(np.linalg.eigvals(comp)).max() < 1.0
    p27 = synth.make_problem(6, 10, var_order=1, drop_piston=True)
    assert (p27.n, p27.A2) == (27, None)
    assert synth.make_problem(10, 30).n == 66
    a = synth.aberrations(p, 3, 7)
    assert a.shape == (3, 7, 28) and np.isfinite(a).all()


def test_product_path_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under the package may reference it."""
    pkg = os.path.join(ROOT, "mpc-sensorlessao_b200")
    for dp_, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".c", ".m")):
                src = open(os.path.join(dp_, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f
                assert "fmpc_ref" not in src and "fref_" not in src, f


def test_var1_mirror_is_the_reference_problem_by_default(pk):
    """Fast_MPC2_VAR1 = VAR_1/Fast_MPC2.m: 21 ctor arguments, ramp rows + literal C unless switched off; the ramp
    rows need u_prev and du bounds of size m (VAR_1/fast_mpc_ineq_const.m:72-74)."""
    n, m, T = 3, 2, 4
    args = [np.eye(n), np.eye(m), None, np.eye(n), None, None, None, -np.ones(n), np.ones(n), -np.ones(m), np.ones(m),
            -0.1 * np.ones(m), 0.1 * np.ones(m), T, np.zeros(n), np.zeros(m), 0.5 * np.eye(n), np.ones((n, m)), np.zeros(T * n), None, None]
    o = pk.Fast_MPC2_VAR1(*args)
    assert o._ramp_rows and o._literal_bug and o.var_order == 1
    assert o._validate() == (n, m)
    bad = list(args); bad[15] = None                      # u_prev = []
    with pytest.raises(ValueError, match="incompatible sizes"):
        pk.Fast_MPC2_VAR1(*bad)._validate()
    bad = list(args); bad[11] = np.ones(m + 1)            # du_min of the wrong size
    with pytest.raises(ValueError, match="cotrol iequality"):
        pk.Fast_MPC2_VAR1(*bad)._validate()
    assert pk.Fast_MPC2_VAR1(*bad, ramp_rows=False, literal_bug=False)._validate() == (n, m)


def test_side_entry_points_validate_before_cuda(pk):
    """Argument errors of the estimator / identification entry points carry their own codes even without a GPU."""
    with pytest.raises(pk.FmpcError) as e:
        pk.Estimator(np.zeros((3, 5)))                    # more modes than pixels
    assert e.value.code == -2
    with pytest.raises(ValueError, match="incompatible sizes"):
        pk.Estimator(np.ones((6, 2)), np.ones(5))
    with pytest.raises(pk.FmpcError) as e:
        pk.identify_var(np.zeros((10, 8)), 2)             # fewer equations than unknowns
    assert e.value.code == -2
    with pytest.raises(pk.FmpcError) as e:
        pk.identify_var(np.zeros((600, 45)), 2)           # accumulators exceed the kernel's budget
    assert e.value.code == -2


def test_library_mt19937_is_matlabs_default_stream(tmp_path):
    """nu0 = NULL => rand(length(b),1) (inf_newton_solver.m:2) from MATLAB's default stream MT19937(5489).
    The host seeds the state (struct MT19937 of csrc/fmpc_api.cu, compiled here); the stream itself is generated on the
    device, every word of the next 624-word block expressed by words of the current one
    (fmpc_mt_twist_kernel in csrc/fmpc_kernels.cu) -- restated here in numpy with the kernel's index ranges and compared with numpy's MT19937; the kernel itself is checked by the -m gpu tests."""
    import subprocess
    src = open(os.path.join(ROOT, "mpc-sensorlessao_b200", "csrc", "fmpc_api.cu")).read()
    a = src.index("struct MT19937 {")
    b = src.index("#define CU_OK", a)
    code = ("#include <cstdint>\n#include <cstdio>\n" + src[a:b] +
            "\nint main(){ MT19937 g(5489u); for (int i = 0; i < 624; ++i) printf(\"%u\\n\", g.mt[i]); return 0; }\n")
    cpp, exe = tmp_path / "mt.cpp", tmp_path / "mt"
    cpp.write_text(code)
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([cxx, "-O2", "-o", str(exe), str(cpp)])
    seed_state = np.array([int(x) for x in subprocess.check_output([str(exe)]).split()], dtype=np.uint32)
    assert np.array_equal(seed_state, np.random.RandomState(5489).get_state()[1])

    def twist(a, b):
        y = (a & np.uint32(0x80000000)) | (b & np.uint32(0x7fffffff))
        return (y >> np.uint32(1)) ^ (np.where(y & np.uint32(1), np.uint32(0x9908b0df), np.uint32(0)))

    def temper(y):
        y = y ^ (y >> np.uint32(11)); y = y ^ ((y << np.uint32(7)) & np.uint32(0x9d2c5680))
        y = y ^ ((y << np.uint32(15)) & np.uint32(0xefc60000)); return y ^ (y >> np.uint32(18))

    o, out = seed_state.copy(), []
    for _ in range(7):                                  # 7 blocks = 2184 doubles
        w = np.zeros(624, dtype=np.uint32)              # every word of the next block from the current one
        e = np.arange(0, 227)
        w[e] = o[e + 397] ^ twist(o[e], o[e + 1])
        e = np.arange(227, 454)
        w[e] = o[e + 170] ^ twist(o[e - 227], o[e - 226]) ^ twist(o[e], o[e + 1])
        e = np.arange(454, 623)
        w[e] = o[e - 57] ^ twist(o[e - 454], o[e - 453]) ^ twist(o[e - 227], o[e - 226]) ^ twist(o[e], o[e + 1])
        n0 = o[397:398] ^ twist(o[0:1], o[1:2])
        n396 = o[566:567] ^ twist(o[169:170], o[170:171]) ^ twist(o[396:397], o[397:398])
        w[623] = (n396 ^ twist(o[623:624], n0))[0]
        y = temper(w)
        out.append(((y[0::2] >> np.uint32(5)).astype(np.float64) * 67108864.0 + (y[1::2] >> np.uint32(6)).astype(np.float64))
                   * (1.0 / 9007199254740992.0))
        o = w
    out = np.concatenate(out)
    ref = np.random.RandomState(5489).random_sample(out.size)
    assert np.array_equal(out, ref)
    assert abs(out[0] - 0.8147236863931789) < 1e-16 and abs(out[1] - 0.9057919370756192) < 1e-16       # MATLAB: rand after start-up


def test_batched_handle_checks_shapes_before_touching_the_gpu(pk):
    """FastMPCBatch hands raw pointers to fmpc_create, which cannot see array sizes: the reference's dimension checks
    (same error strings) run on the host first."""
    import pytest
    n, m, T = 4, 3, 5
    A, B, Q, R = np.eye(n), np.ones((n, m)), np.eye(n), np.eye(m)
    um = np.ones(m)
    ok = dict(A1=A, A2=A, B=B, Q=Q, R=R, Qf=Q, u_min=-um, u_max=um, T=T)
    def make(**kw):
        a = dict(ok); a.update(kw)
        return pk.FastMPCBatch(a["A1"], a["A2"], a["B"], a["Q"], a["R"], a["Qf"], a["u_min"], a["u_max"], a["T"],
                               q=a.get("q"), r=a.get("r"), x_min=a.get("x_min"), x_max=a.get("x_max"))
    with pytest.raises(ValueError, match="State stage cost"):
        make(Q=np.eye(n + 1))
    with pytest.raises(ValueError, match="Control stage cost"):
        make(R=np.eye(m + 1))
    with pytest.raises(ValueError, match="equality state dynamics"):
        make(A2=np.eye(n + 1))
    with pytest.raises(ValueError, match="cotrol iequality"):
        make(u_max=np.ones(m - 1))
    with pytest.raises(ValueError, match="Linear state cost"):
        make(q=np.ones(n - 1))
    with pytest.raises(ValueError, match="state inequality"):
        make(x_min=-np.ones(n - 2), x_max=np.ones(n))
