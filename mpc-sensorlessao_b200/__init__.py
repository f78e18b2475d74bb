"""B200-native fastMPC hot path of jinsungkim96/MPC-SensorlessAO.

Host-side mirror (Python, ctypes) of the reference's MATLAB interface for this path:
  Fast_MPC2 / Fast_MPC2_VAR1  <->  Fast_MPC/VAR_2/Fast_MPC2.m, Fast_MPC/VAR_1/Fast_MPC2.m
  zernmodfit / ZernikeFitter  <->  zernmodfit.m (+ zernfun.m)
  Estimator                   <->  the lsqminnorm estimator step, README.md:478 (model_approx.mat)
  FastMPCBatch                 :   the batched C-ABI (include/fmpc.h) for many instances
  FastMPCMulti                 :   the same over every GPU of the box in one call (fmpc_multi_*)
All compute runs in the CUDA library `lib/libfmpc_b200.so`; there is no CPU fallback.
"""
from ._lib import (FmpcError, FmpcParams, build_library, device_count, fp64_peak, lib_path, load_library,
                   strerror)
from .fast_mpc2 import FastMPCBatch, FastMPCMulti, Fast_MPC2, Fast_MPC2_VAR1, deinterleave, interleave
from .zernike import SampleFitter, ZernikeFitter, zernmodfit
from .estimator import Estimator, identify_var

__all__ = ["FmpcError", "FmpcParams", "build_library", "device_count", "fp64_peak", "lib_path", "load_library",
           "strerror", "FastMPCBatch", "FastMPCMulti", "Fast_MPC2", "Fast_MPC2_VAR1", "deinterleave", "interleave",
           "ZernikeFitter", "SampleFitter", "zernmodfit", "Estimator", "identify_var"]
