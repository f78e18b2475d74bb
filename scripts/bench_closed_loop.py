"""fmpc_closed_loop at the C2 shape: 4096 instances x K steps entirely on the device (north-star item d)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mpc_sensorlessao_b200 as pk
from mpc_sensorlessao_b200 import synth
nb, K = 4096, 20
p = synth.make_problem(6, 20)
a = synth.aberrations(p, nb, K, seed=3)
nu0 = np.random.RandomState(1).random_sample((K, nb, p.T * p.n))
hb = pk.FastMPCBatch(p.A1, p.A2, p.B, p.Q, p.R, p.Qf, p.u_min, p.u_max, p.T, p.x_min, p.x_max, max_batch=nb)
hb.closed_loop(a[:, :3], nu0=nu0[:3], kappa=0.01, niters=5)
for label, nu in (("explicit nu0", nu0), ("MATLAB stream (nu0 = NULL, generated on the host)", None)):
    t0 = time.perf_counter()
    out = hb.closed_loop(a, nu0=nu, kappa=0.01, niters=5)
    wall = time.perf_counter() - t0
    print(json.dumps({"workload": f"closed loop C2 shape, {nb} instances x {K} steps, {label}", "device_s": out["telapsed"], "wall_s": wall,
                      "solves_per_s_device": nb * K / out["telapsed"], "solves_per_s_wall": nb * K / wall,
                      "newton_iters_per_solve": float(out["iters"].mean()),
                      "rms_residual_first_last": [float(np.sqrt((out["X_acc"][:, 0] ** 2).mean())), float(np.sqrt((out["X_acc"][:, -1] ** 2).mean()))]}), flush=True)
