"""GPU probe of the resident closed-loop step (fmpc_step_r) at the C2 shape: iterations per step, e2e rate from pinned and
pageable host buffers, the full-surface fmpc_step next to it, and the device MT19937 generator rate."""
import os, sys, time, json, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import mpc_sensorlessao_b200 as pk
from mpc_sensorlessao_b200 import synth

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
K = int(sys.argv[2]) if len(sys.argv) > 2 else 24
ub = float(sys.argv[3]) if len(sys.argv) > 3 else 28.0
p = synth.make_problem(6, 20, u_bound=ub)
n, m, T = p.n, p.m, p.T
a = synth.aberrations(p, nb, K, seed=3)
hb = pk.FastMPCBatch(p.A1, p.A2, p.B, p.Q, p.R, p.Qf, p.u_min, p.u_max, T, p.x_min, p.x_max, max_batch=nb)
L = hb._L
# pass 1: run the loop (host computes x0 = a + B u_prev like the README loop), record the x0 sequence
x0s = np.empty((K, nb, n))
u_prev = np.zeros((nb, m))
its = []
t0 = time.perf_counter()
for k in range(K):
    x0s[k] = a[:, k] + u_prev @ p.B.T
    out = hb.step_resident(x0s[k], reset=(k == 0), niters=5)
    u_prev = out["u0"]
    its.append(float(out["iters"].mean()))
print("iters/solve per step:", [round(v, 3) for v in its], "status hist last", np.bincount(out["status"], minlength=5).tolist(), flush=True)
print("loop incl. host x0 update: %.1f ms/step" % ((time.perf_counter() - t0) / K * 1e3), flush=True)

params = hb.params(0.01, 5, 0)
def replay(x0buf, u0buf, label):
    vp = lambda t: C.c_void_p(t.ctypes.data if isinstance(t, np.ndarray) else t.data_ptr())
    st = np.zeros(nb, np.int32); it = np.zeros(nb, np.int32)
    for rep in range(2):
        t0 = time.perf_counter(); dev = 0.0
        for k in range(K):
            tel = C.c_double(0)
            rc = L.fmpc_step_r(hb._h, C.byref(params), nb, 1 if k == 0 else 0, vp(x0buf[k]), None, None, None, None, None,
                               vp(u0buf), None, None, vp(st), vp(it), C.cast(C.byref(tel), C.c_void_p))
            assert rc == 0, rc
            dev += tel.value
        dt = time.perf_counter() - t0
    print(f"{label}: {nb * K / dt:.0f} solves/s e2e ({dt / K * 1e3:.3f} ms/step), solve kernels {dev / K * 1e3:.3f} ms/step -> {nb * K / dev:.0f} solves/s", flush=True)

xp = [torch.from_numpy(x0s[k].copy()).pin_memory() for k in range(K)]
up = torch.empty((nb, m), dtype=torch.float64).pin_memory()
replay(xp, up, "resident, pinned host buffers, nu0 = device MATLAB stream")
xg = [x0s[k].copy() for k in range(K)]
ug = np.empty((nb, m))
replay(xg, ug, "resident, pageable host buffers")

# full-surface step on the same loop state for comparison (explicit warm start in, full horizon out)
wi = synth.warm_inputs(p, nb)
for label, pin in (("pinned", True), ("pageable", False)):
    arrs = {k: (torch.from_numpy(np.ascontiguousarray(wi[k])).pin_memory().numpy() if pin else np.ascontiguousarray(wi[k])) for k in wi}
    Xo = torch.empty((nb, T, n), dtype=torch.float64); Uo = torch.empty((nb, T, m), dtype=torch.float64)
    if pin:
        Xo, Uo = Xo.pin_memory(), Uo.pin_memory()
    vp = lambda t: C.c_void_p(t.ctypes.data if isinstance(t, np.ndarray) else t.data_ptr())
    for rep in range(2):
        t0 = time.perf_counter()
        for k in range(8):
            rc = L.fmpc_step(hb._h, C.byref(params), nb, vp(arrs["x0"]), vp(arrs["x0_pre"]), None, None, None, vp(arrs["X0"]), vp(arrs["U0"]),
                             vp(arrs["nu0"]), vp(Xo), vp(Uo), None, None, None)
            assert rc == 0
        dt = time.perf_counter() - t0
    print(f"full-surface fmpc_step, {label} host buffers: {nb * 8 / dt:.0f} solves/s ({dt / 8 * 1e3:.3f} ms/step)", flush=True)

# closed loop on the device, explicit nu0 vs device stream
nu0 = np.random.RandomState(1).random_sample((K, nb, T * n))
for label, nu in (("explicit nu0", nu0), ("nu0 = NULL (device MT19937)", None)):
    for rep in range(2):
        t0 = time.perf_counter()
        out = hb.closed_loop(a, nu0=nu, kappa=0.01, niters=5)
        wall = time.perf_counter() - t0
    print(f"fmpc_closed_loop {label}: device {nb * K / out['telapsed']:.0f} solves/s, wall {nb * K / wall:.0f}, iters/solve {out['iters'].mean():.3f}", flush=True)
hb.close()
