"""ctypes binding of include/fmpc.h.  Loading fails loudly (no fallback) when the CUDA
library has not been built; compute calls fail with FMPC_ERR_CUDA when there is no B200."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
# FMPC_B200_LIB selects an alternative BUILD of the same library (e.g. the -DFMPC_PROF profiling variant); never a fallback
_SO = os.environ.get("FMPC_B200_LIB") or os.path.join(_HERE, "lib", "libfmpc_b200.so")
dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int)


class FmpcError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"fmpc error {code}: {msg}")
        self.code = code


class FmpcSys(C.Structure):
    _fields_ = [("n", C.c_int), ("m", C.c_int), ("T", C.c_int), ("var_order", C.c_int),
                ("A1", dp), ("A2", dp), ("B", dp), ("Q", dp), ("R", dp), ("Qf", dp),
                ("q", dp), ("r", dp), ("qf", dp), ("x_min", dp), ("x_max", dp), ("u_min", dp), ("u_max", dp),
                ("du_min", dp), ("du_max", dp), ("ramp_rows", C.c_int), ("var1_literal_bug", C.c_int)]


class FmpcParams(C.Structure):
    _fields_ = [("kappa", C.c_double), ("niters", C.c_int), ("ls_max", C.c_int), ("alpha", C.c_double),
                ("beta", C.c_double), ("tol_r", C.c_double), ("tol_p", C.c_double)]

    @classmethod
    def default(cls, **kw) -> "FmpcParams":
        p = cls()
        load_library().fmpc_default_params(C.byref(p))
        for k, v in kw.items():
            setattr(p, k, v)
        return p


# every symbol include/fmpc.h declares (tests/test_abi.py checks the .so exports all of them)
ABI_SYMBOLS = [
    "fmpc_default_params", "fmpc_device_count", "fmpc_create", "fmpc_destroy", "fmpc_step", "fmpc_step_d",
    "fmpc_step_r", "fmpc_step_r_d", "fmpc_step_z", "fmpc_frontend", "fmpc_frontend_nouter", "fmpc_state_update", "fmpc_state_update_d",
    "fmpc_closed_loop", "fmpc_get_dims", "fmpc_workspace_bytes", "fmpc_launch_count", "fmpc_last_newton_iters", "fmpc_kernel_kind", "fmpc_last_profile", "fmpc_strerror",
    "fmpc_fp64_peak", "fmpc_seed_stream", "fmpc_multi_create", "fmpc_multi_destroy", "fmpc_multi_ngpus", "fmpc_multi_shard",
    "fmpc_multi_handle", "fmpc_multi_step", "fmpc_multi_step_r", "fmpc_multi_last_stats", "zmf_create", "zmf_create_samples", "zmf_destroy", "zmf_nmodes", "zmf_npix_in", "zmf_fit", "zmf_fit_d", "zmf_synth", "zmf_synth_d",
    "zmf_get_basis", "zmf_get_mask", "zmf_launch_count",
    "est_create", "est_destroy", "est_apply", "est_apply_d", "est_launch_count", "var_identify",
]


def lib_path() -> str:
    return _SO


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a with nvcc (cross-compiles without a GPU)."""
    cmd = ["make", "-C", os.path.join(_HERE, "csrc")] + (["-B"] if force else [])
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or out.returncode != 0:
        print(out.stdout)
    if out.returncode != 0:
        raise RuntimeError("building libfmpc_b200.so failed")
    return _SO


_lib = None


def load_library():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_SO):
        raise ImportError(f"{_SO} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(or make -C mpc-sensorlessao_b200/csrc). There is no CPU fallback.")
    L = C.CDLL(_SO)
    vp = C.c_void_p
    L.fmpc_default_params.argtypes = [C.POINTER(FmpcParams)]
    L.fmpc_default_params.restype = None
    L.fmpc_device_count.restype = C.c_int
    L.fmpc_create.argtypes = [C.POINTER(vp), C.POINTER(FmpcSys), C.c_int, C.c_int]
    L.fmpc_destroy.argtypes = [vp]
    L.fmpc_destroy.restype = None
    step_args = [vp, C.POINTER(FmpcParams), C.c_int] + [vp] * 10 + [vp, vp]
    L.fmpc_step.argtypes = step_args + [vp]                # ..., status, iters, telapsed
    L.fmpc_step_d.argtypes = step_args + [vp]              # ..., status, iters, stream
    L.fmpc_step_r.argtypes = [vp, C.POINTER(FmpcParams), C.c_int, C.c_int] + [vp] * 6 + [vp] * 6
    L.fmpc_step_r_d.argtypes = [vp, C.POINTER(FmpcParams), C.c_int, C.c_int] + [vp] * 6 + [vp] * 6
    L.fmpc_seed_stream.argtypes = [vp, C.c_uint]
    L.fmpc_multi_create.argtypes = [C.POINTER(vp), C.POINTER(FmpcSys), C.c_int, C.c_int, vp]
    L.fmpc_multi_destroy.argtypes = [vp]
    L.fmpc_multi_destroy.restype = None
    L.fmpc_multi_ngpus.argtypes = [vp]
    L.fmpc_multi_shard.argtypes = [vp, C.c_int, C.c_int, vp, vp]
    L.fmpc_multi_handle.argtypes = [vp, C.c_int]
    L.fmpc_multi_handle.restype = vp
    L.fmpc_multi_step.argtypes = step_args + [vp]
    L.fmpc_multi_step_r.argtypes = [vp, C.POINTER(FmpcParams), C.c_int, C.c_int] + [vp] * 6 + [vp] * 6
    L.fmpc_multi_last_stats.argtypes = [vp, vp, C.c_int, vp]
    L.fmpc_step_z.argtypes = [vp, C.POINTER(FmpcParams), C.c_int] + [vp] * 8 + [vp, vp, vp]
    L.fmpc_frontend.argtypes = [vp, C.c_int, C.POINTER(FmpcParams), C.c_double, C.c_double, C.c_int] + [vp] * 10 + [vp, vp, vp]
    L.fmpc_frontend_nouter.argtypes = [vp, C.c_int]
    L.fmpc_state_update.argtypes = [vp, C.c_int] + [vp] * 5
    L.fmpc_state_update_d.argtypes = [vp, C.c_int] + [vp] * 6
    L.fmpc_closed_loop.argtypes = [vp, C.POINTER(FmpcParams), C.c_int, C.c_int] + [vp] * 6
    L.var_identify.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp, C.c_int, vp]
    L.est_create.argtypes = [C.POINTER(vp), C.c_int, C.c_int, vp, vp, C.c_int, C.c_int]
    L.est_destroy.argtypes = [vp]
    L.est_destroy.restype = None
    L.est_apply.argtypes = [vp, C.c_int, vp, vp, vp]
    L.est_apply_d.argtypes = [vp, C.c_int, vp, vp, vp]
    for f in ("fmpc_workspace_bytes", "fmpc_launch_count", "fmpc_last_newton_iters", "zmf_launch_count", "est_launch_count"):
        getattr(L, f).argtypes = [vp]
        getattr(L, f).restype = C.c_longlong
    L.fmpc_kernel_kind.argtypes = [vp]
    L.fmpc_last_profile.argtypes = [vp, vp]
    L.fmpc_strerror.argtypes = [C.c_int]
    L.fmpc_strerror.restype = C.c_char_p
    L.fmpc_fp64_peak.argtypes = [C.c_int, C.c_int, C.c_int]
    L.fmpc_fp64_peak.restype = C.c_double
    L.zmf_create.argtypes = [C.POINTER(vp), C.c_int, C.c_int, C.c_int, C.c_int]
    L.zmf_create_samples.argtypes = [C.POINTER(vp), C.c_int, vp, vp, C.c_int, C.c_int, C.c_int]
    L.zmf_destroy.argtypes = [vp]
    L.zmf_destroy.restype = None
    L.zmf_nmodes.argtypes = [vp]
    L.zmf_npix_in.argtypes = [vp]
    L.zmf_fit.argtypes = [vp, C.c_int, vp, vp, vp]
    L.zmf_fit_d.argtypes = [vp, C.c_int, vp, vp, vp]
    L.zmf_synth.argtypes = [vp, C.c_int, vp, vp, vp]
    L.zmf_synth_d.argtypes = [vp, C.c_int, vp, vp, vp]
    L.zmf_get_basis.argtypes = [vp, vp]
    L.zmf_get_mask.argtypes = [vp, vp]
    _lib = L
    return L


def strerror(code: int) -> str:
    return load_library().fmpc_strerror(int(code)).decode()


def check(code: int):
    if code != 0:
        raise FmpcError(code, strerror(code))


def device_count() -> int:
    return int(load_library().fmpc_device_count())


def fp64_peak(device: int = 0, kind: int = 0, iters: int = 2000) -> float:
    """Measured FP64 TFLOP/s of the DFMA (kind 0) or DMMA m8n8k4 (kind 1) pipe."""
    v = float(load_library().fmpc_fp64_peak(device, kind, iters))
    if v < 0:
        raise FmpcError(-16, strerror(-16))
    return v
