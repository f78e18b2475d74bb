"""Device MT19937 generator: rate, and the stream against numpy's MT19937(5489) through the public API."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, ctypes as C, torch
import mpc_sensorlessao_b200 as pk
from mpc_sensorlessao_b200 import synth
p = synth.make_problem(6, 20)
nb = 4096
hb = pk.FastMPCBatch(p.A1, p.A2, p.B, p.Q, p.R, p.Qf, p.u_min, p.u_max, p.T, p.x_min, p.x_max, max_batch=nb)
x0 = np.zeros((nb, p.n)); x0[:, 0] = 0.1
for k in range(6):
    t0 = time.perf_counter()
    out = hb.step_resident(x0, reset=(k == 0), niters=5)
    dt = time.perf_counter() - t0
    print(f"step {k}: wall {dt*1e3:.3f} ms, solve kernel {out['telapsed']*1e3:.3f} ms", flush=True)
hb.close()
