"""Short single-kernel target for ncu: a few C2 steps (n=28, m=144, T=20) with nb instances."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mpc_sensorlessao_b200 as pk
from mpc_sensorlessao_b200 import synth
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 888
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
p = synth.make_problem(6, 20)
wi = synth.warm_inputs(p, nb)
hb = pk.FastMPCBatch(p.A1, p.A2, p.B, p.Q, p.R, p.Qf, p.u_min, p.u_max, 20, p.x_min, p.x_max, max_batch=nb)
for _ in range(reps):
    out = hb.step(wi['x0'], wi['x0_pre'], None, None, wi['X0'], wi['U0'], wi['nu0'], kappa=0.01, niters=5)
print("kernel ms", out['telapsed'] * 1e3, "kind", hb.kernel_kind)
hb.close()
