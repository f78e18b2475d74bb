"""One C5-shape launch (n = 66, m = 144, T = 30) of the CTA DMMA kernel for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mpc_sensorlessao_b200 as pk
from mpc_sensorlessao_b200 import synth
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 296
p = synth.make_problem(10, 30)
wi = synth.warm_inputs(p, nb, seed=4)
hb = pk.FastMPCBatch(p.A1, p.A2, p.B, p.Q, p.R, p.Qf, p.u_min, p.u_max, p.T, p.x_min, p.x_max, max_batch=nb)
for _ in range(2):
    out = hb.step(wi["x0"], wi["x0_pre"], None, None, wi["X0"], wi["U0"], wi["nu0"], kappa=0.01, niters=5)
print("kernel ms", out["telapsed"] * 1e3, "iters", out["iters"].mean())
