"""Full-surface fmpc_step with pageable vs pinned host buffers (C2, 4096 instances) for a given FMPC_COPY_THREADS."""
import os, sys, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mpc_sensorlessao_b200 as pk
from mpc_sensorlessao_b200 import synth
nb = 4096
p = synth.make_problem(6, 20)
wi = synth.warm_inputs(p, nb)
hb = pk.FastMPCBatch(p.A1, p.A2, p.B, p.Q, p.R, p.Qf, p.u_min, p.u_max, p.T, p.x_min, p.x_max, max_batch=nb)
params = hb.params(0.01, 5, 0)
for pin in (True, False):
    arrs = {k: (torch.from_numpy(np.ascontiguousarray(v)).pin_memory().numpy() if pin else np.ascontiguousarray(v).copy()) for k, v in wi.items()}
    X = torch.empty((nb, p.T, p.n), dtype=torch.float64); U = torch.empty((nb, p.T, p.m), dtype=torch.float64)
    if pin: X, U = X.pin_memory(), U.pin_memory()
    vp = lambda t: C.c_void_p(t.ctypes.data if isinstance(t, np.ndarray) else t.data_ptr())
    for rep in range(3):
        t0 = time.perf_counter()
        for k in range(10):
            rc = hb._L.fmpc_step(hb._h, C.byref(params), nb, vp(arrs["x0"]), vp(arrs["x0_pre"]), None, None, None, vp(arrs["X0"]), vp(arrs["U0"]),
                                 vp(arrs["nu0"]), vp(X), vp(U), None, None, None)
            assert rc == 0
        dt = (time.perf_counter() - t0) / 10
    print(f"threads {os.environ.get('FMPC_COPY_THREADS', 'default')} {'pinned' if pin else 'pageable'}: {dt*1e3:.3f} ms/step, {nb/dt:.0f} solves/s, {246e6/dt/1e9:.1f} GB/s host traffic", flush=True)
hb.close()
