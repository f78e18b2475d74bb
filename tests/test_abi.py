"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/fmpc.h declares, and fails LOUDLY (no CPU fallback) when no B200 is present."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "fmpc.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b((?:fmpc|zmf|est|var)_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_list_agree(pk):
    from mpc_sensorlessao_b200 import _lib
    assert declared_symbols() == sorted(_lib.ABI_SYMBOLS)


def test_library_exports_every_declared_symbol(pk):
    so = pk.lib_path()
    if not os.path.exists(so):
        pk.build_library()
    L = ctypes.CDLL(so)
    for sym in declared_symbols():
        assert hasattr(L, sym), f"{sym} declared in include/fmpc.h but not exported"
    out = subprocess.run(["nm", "-D", "--defined-only", so], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T ((?:fmpc|zmf|est|var)_\w+)", out))
    assert exported == set(declared_symbols()), "exported C-ABI symbols differ from the header"


def test_header_compiles_as_plain_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "fmpc.h"\nint main(void){ fmpc_params p; fmpc_sys s; (void)p; (void)s; return FMPC_VERSION > 0 ? 0 : 1; }\n')
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    subprocess.check_call([cc, "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(src),
                           "-o", str(tmp_path / "t.o")])


def test_mex_shim_compiles_against_the_header():
    """The MATLAB gateway cannot be linked here (no MATLAB): compile-check it against a stub mex.h so it stays in
    sync with include/fmpc.h."""
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    subprocess.check_call([cc, "-std=c99", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "tests", "stubs"),
                           "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "mpc-sensorlessao_b200", "matlab", "fmpc_mex.c")])


def test_sass_is_sm100_only(pk):
    out = subprocess.run(["cuobjdump", "-lelf", pk.lib_path()], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_strerror_and_defaults(pk):
    assert "square" in pk.strerror(-3)
    assert "no CPU fallback" in pk.strerror(-16)
    p = pk.FmpcParams.default()
    assert (p.kappa, p.niters, p.alpha, p.beta, p.tol_r, p.tol_p) == (0.01, 5, 1e-4, 0.5, 1e-6, 1e-8)


def test_no_cpu_fallback_without_gpu(pk):
    import numpy as np
    if pk.device_count() > 0:
        pytest.skip("a B200 is present")
    n, m, T = 3, 2, 4
    with pytest.raises(pk.FmpcError) as e:
        pk.FastMPCBatch(np.eye(n), np.eye(n), np.ones((n, m)), np.eye(n), np.eye(m), np.eye(n), -np.ones(m), np.ones(m), T)
    assert e.value.code == -16
    with pytest.raises(pk.FmpcError):
        pk.ZernikeFitter(32, 3)
    with pytest.raises(pk.FmpcError):
        pk.fp64_peak(0, 0, 10)


def test_create_argument_errors_come_before_cuda(pk):
    """Reference error() conditions are reported as their own codes even without a GPU."""
    import numpy as np
    from mpc_sensorlessao_b200._lib import FmpcSys, load_library, dp
    L = load_library()
    n, m, T = 3, 2, 4
    keep = [np.asfortranarray(np.eye(n)), np.asfortranarray(np.ones((n, m))), np.asfortranarray(np.eye(m)), np.ones(n), np.ones(m)]
    P = lambda a: a.ctypes.data_as(dp)

    def sys_(**kw):
        s = FmpcSys()
        s.n, s.m, s.T, s.var_order = n, m, T, 2
        s.A1 = s.A2 = s.Q = s.Qf = P(keep[0]); s.B = P(keep[1]); s.R = P(keep[2])
        s.x_min = s.x_max = P(keep[3]); s.u_min = s.u_max = P(keep[4])
        for k, v in kw.items():
            setattr(s, k, v)
        return s

    h = ctypes.c_void_p()
    null = ctypes.cast(None, dp)
    assert L.fmpc_create(ctypes.byref(h), ctypes.byref(sys_(A2=null)), 1, 0) == -8       # NO_A
    assert L.fmpc_create(ctypes.byref(h), ctypes.byref(sys_(B=null)), 1, 0) == -9        # NO_B
    assert L.fmpc_create(ctypes.byref(h), ctypes.byref(sys_(u_min=null)), 1, 0) == -7    # U_BOUND_SIZE
    assert L.fmpc_create(ctypes.byref(h), ctypes.byref(sys_(n=0)), 1, 0) == -2           # DIM
    dense = np.asfortranarray(np.eye(m) + 0.1)
    assert L.fmpc_create(ctypes.byref(h), ctypes.byref(sys_(R=P(dense))), 1, 0) == -16   # dense R is covered: fails only for lack of a GPU
    denseq = np.asfortranarray(np.eye(n) + 0.1)
    assert L.fmpc_create(ctypes.byref(h), ctypes.byref(sys_(Q=P(denseq))), 1, 0) == -16  # dense Q is covered: fails only for lack of a GPU
    assert L.fmpc_create(ctypes.byref(h), ctypes.byref(sys_(ramp_rows=1)), 1, 0) == -7   # ramp rows without du bounds
    neg = np.asfortranarray(-np.eye(m))
    assert L.fmpc_create(ctypes.byref(h), ctypes.byref(sys_(R=P(neg))), 1, 0) == -13     # NOT_PD
