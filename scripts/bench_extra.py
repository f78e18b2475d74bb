"""Side measurements of SURVEY.md 8d that are not bench.py's headline line (evidence under profiles/):
  zmf   : zernmodfit of 2000 and 32768 synthetic 128 x 128 frames (configs[2]) vs the HBM roofline
  c1    : configs[0], VAR(1) n = 27, m = 144, T = 10, ONE closed loop of K steps -- the reference's VAR_1 as written
          (ramp rows + literal C, general kernel) and the box-only fast kernel; literal dense numpy oracle beside it
  c5    : configs[4] shape, n = 66, m = 144, T = 30 (generic kernel)
      python scripts/bench_extra.py [zmf] [c1] [c5]
"""
import ctypes as C
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import mpc_sensorlessao_b200 as pk
from mpc_sensorlessao_b200 import synth
from mpc_sensorlessao_b200._lib import load_library

L = load_library()
dev = torch.device("cuda", 0)
peaks = {}
try:
    peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
except Exception:
    pass
HBM = float(peaks.get("hbm_gbs", 6550.4))
vp = lambda t: C.c_void_p(t.data_ptr())


def f_newton(n, m, T):
    return T * (n * n * m + (19.0 / 3.0) * n ** 3 + 8 * n * m + 26 * n * n)


def bench_zmf():
    nL = int(os.environ.get("ZMF_NL", "128"))
    for nf in (2000, 32768):
        zf = pk.ZernikeFitter(nL, 6, max_frames=8)
        nm = zf.nmodes
        frames = torch.randn((nf, nL * nL), dtype=torch.float64, device=dev)
        coef = torch.empty((nf, nm), dtype=torch.float64, device=dev)
        st = torch.cuda.Stream(dev)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        ts = []
        for it in range(8):
            flush.fill_(it)                               # 256 MB > L2: the frames of the 2000-frame case must come from HBM
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            rc = L.zmf_fit_d(zf._h, nf, vp(frames), vp(coef), C.c_void_p(st.cuda_stream))
            e1.record(st)
            torch.cuda.synchronize()
            assert rc == 0
            ts.append(e0.elapsed_time(e1))
        ms = float(np.median(ts[2:]))
        byts = nf * (8 * nL * nL + 8 * nm)
        print(json.dumps({"workload": f"zernmodfit N=6, {nf} frames {nL}x{nL}", "kernel_ms": ms, "frames_per_s": nf / ms * 1e3,
                          "roofline": {"bound": "hbm", "achieved": byts / ms / 1e6, "peak": HBM, "unit": "GB/s", "frac": byts / ms / 1e6 / HBM,
                                       "bytes_per_frame": byts // nf},
                          "fp64_tflops": 2 * nm * zf.npix_in * nf / ms / 1e9}), flush=True)
        # synthesis (README.md:592-598): frames = Z coef, HBM-write bound
        ts = []
        for it in range(8):
            flush.fill_(it)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            rc = L.zmf_synth_d(zf._h, nf, vp(coef), vp(frames), C.c_void_p(st.cuda_stream))
            e1.record(st)
            torch.cuda.synchronize()
            assert rc == 0
            ts.append(e0.elapsed_time(e1))
        ms = float(np.median(ts[2:]))
        print(json.dumps({"workload": f"Zernike synthesis N=6, {nf} frames {nL}x{nL}", "kernel_ms": ms, "frames_per_s": nf / ms * 1e3,
                          "roofline": {"bound": "hbm", "achieved": byts / ms / 1e6, "peak": HBM, "unit": "GB/s", "frac": byts / ms / 1e6 / HBM,
                                       "bytes_per_frame": byts // nf},
                          "fp64_tflops": 2 * nm * zf.npix_in * nf / ms / 1e9}), flush=True)
        zf.close()


def bench_c1(K=200):
    from oracle import fastmpc_dense as fd
    p = synth.make_problem(6, 10, var_order=1, drop_piston=True, u_bound=28.0)
    a = synth.aberrations(p, 1, K, seed=3, amp=0.3)
    nu0 = np.random.RandomState(5489).random_sample((K, 1, p.T * p.n))
    res = {}
    for name, kw in (("reference VAR_1 as written (ramp rows + literal C; general kernel)",
                      dict(du_min=p.du_min, du_max=p.du_max, ramp_rows=True, var1_literal_bug=True)),
                     ("box rows, corrected C (warp DMMA kernel)", dict())):
        hb = pk.FastMPCBatch(p.A1, None, p.B, p.Q, p.R, p.Qf, p.u_min, p.u_max, p.T, p.x_min, p.x_max, max_batch=1, **kw)
        hb.closed_loop(a[:, :5], nu0=nu0[:5], kappa=0.01, niters=5)
        t0 = time.perf_counter()
        out = hb.closed_loop(a, nu0=nu0, kappa=0.01, niters=5)
        wall = time.perf_counter() - t0
        res[name] = out
        print(json.dumps({"workload": f"C1 single closed loop, VAR(1) n={p.n} m={p.m} T={p.T}, {K} steps, niters=5: {name}",
                          "kernel_kind": hb.kernel_kind, "steps_per_s_device": K / out["telapsed"], "steps_per_s_wall": K / wall,
                          "ms_per_step": out["telapsed"] / K * 1e3, "newton_iters_per_step": float(out["iters"].mean()),
                          "rms_residual_first_last": [float(np.sqrt((out["X_acc"][0, 0] ** 2).mean())),
                                                      float(np.sqrt((out["X_acc"][0, -1] ** 2).mean()))]}), flush=True)
        hb.close()
    # literal dense oracle (what MATLAB executes) on the first steps of the same loop
    u_prev = np.zeros(p.m); z = None
    t0 = time.perf_counter(); nst = 3
    first = list(res.values())[0]
    for k in range(nst):
        x0 = a[0, k] + p.B @ u_prev
        if z is not None:
            Z = z.reshape(p.T, p.n + p.m); z = np.vstack([Z[1:], Z[-1:]]).reshape(-1)
        o = fd.Fast_MPC2_VAR1(p.Q, p.R, None, p.Qf, None, None, None, p.x_min, p.x_max, p.u_min, p.u_max, p.du_min, p.du_max,
                              p.T, x0, u_prev, p.A1, p.B, np.zeros(p.T * p.n), None, z)
        z = o.mpc_fixed_log_newton(5, 0.01, nu0=nu0[k, 0])
        u_prev = z[:p.m].copy()
        err = np.abs(first["U_acc"][0, k] - u_prev).max() / np.abs(u_prev).max()
        print(f"  step {k}: dense-oracle iters {o.last_stats['iters']} relerr(U applied) {err:.2e}", flush=True)
    dt = (time.perf_counter() - t0) / nst
    print(json.dumps({"cpu_baseline": {"kind": "port", "what": "literal dense numpy restatement of Fast_MPC/VAR_1 (what MATLAB executes)",
                                       "steps_per_s": 1 / dt, "cores": os.cpu_count(), "sample": f"{nst} closed-loop steps"}}), flush=True)
    # structured C port (oracle/fmpc_ref_general.c) on the same loop, one instance = one host thread
    from oracle import fmpc_ref
    u_prev = np.zeros(p.m); z = np.tile(np.concatenate([(p.u_min + p.u_max) / 2, (p.x_min + p.x_max) / 2]), p.T)
    t0 = time.perf_counter(); nst2 = 10
    for k in range(nst2):
        x0 = a[0, k] + p.B @ u_prev
        if k > 0:
            Z = z.reshape(p.T, p.n + p.m); z = np.vstack([Z[1:], Z[-1:]]).reshape(-1)
        out = fmpc_ref.solve_batch_general(p.A1, None, p.B, p.Q, p.R, p.Qf, p.u_min, p.u_max, 0.01, 5, x0[:, None], None, None,
                                           z[:, None], nu0[k, 0][:, None], u_prev=u_prev[:, None], du_min=p.du_min, du_max=p.du_max,
                                           ramp_rows=True, literal_bug=True)
        z = out["z"][:, 0].copy(); u_prev = z[:p.m].copy()
        if k < 3:
            print(f"  step {k}: C port relerr(U applied) {np.abs(first['U_acc'][0, k] - u_prev).max() / np.abs(u_prev).max():.2e}", flush=True)
    dt2 = (time.perf_counter() - t0) / nst2
    print(json.dumps({"cpu_baseline": {"kind": "port", "what": "structured C restatement of Fast_MPC/VAR_1 as written (oracle/fmpc_ref_general.c)",
                                       "steps_per_s": 1 / dt2, "cores": 1, "sample": f"{nst2} closed-loop steps, one instance"}}), flush=True)


def bench_c5(nb=int(os.environ.get("C5_NB", "2048"))):
    p = synth.make_problem(10, 30)
    wi = synth.warm_inputs(p, nb, seed=4)
    hb = pk.FastMPCBatch(p.A1, p.A2, p.B, p.Q, p.R, p.Qf, p.u_min, p.u_max, p.T, p.x_min, p.x_max, max_batch=nb)
    hb.step(wi["x0"][:64], wi["x0_pre"][:64], None, None, wi["X0"][:64], wi["U0"][:64], wi["nu0"][:64], kappa=0.01, niters=5)
    out = hb.step(wi["x0"], wi["x0_pre"], None, None, wi["X0"], wi["U0"], wi["nu0"], kappa=0.01, niters=5)
    its = float(out["iters"].sum())
    F = f_newton(p.n, p.m, p.T)
    print(json.dumps({"workload": f"C5 shape VAR(2) n={p.n} m={p.m} T={p.T}, {nb} instances, niters=5", "kernel_kind": hb.kernel_kind,
                      "kernel_ms": out["telapsed"] * 1e3, "solves_per_s": nb / out["telapsed"], "newton_iters_per_solve": its / nb,
                      "tflops_model": its * F / out["telapsed"] / 1e12}), flush=True)
    hb.close()


if __name__ == "__main__":
    what = sys.argv[1:] or ["zmf", "c1", "c5"]
    if "zmf" in what:
        bench_zmf()
    if "c1" in what:
        bench_c1()
    if "c5" in what:
        bench_c5()
