"""A few C1 closed-loop steps (VAR_1 as written: ramp rows + literal C, n = 27, m = 144, T = 10) on the general kernel, for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mpc_sensorlessao_b200 as pk
from mpc_sensorlessao_b200 import synth
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 1
p = synth.make_problem(6, 10, var_order=1, drop_piston=True, u_bound=28.0)
K = 4
a = synth.aberrations(p, nb, K, seed=3, amp=0.3)
nu0 = np.random.RandomState(5489).random_sample((K, nb, p.T * p.n))
hb = pk.FastMPCBatch(p.A1, None, p.B, p.Q, p.R, p.Qf, p.u_min, p.u_max, p.T, p.x_min, p.x_max, du_min=p.du_min, du_max=p.du_max,
                     ramp_rows=True, var1_literal_bug=True, max_batch=nb)
out = hb.closed_loop(a, nu0=nu0, kappa=0.01, niters=5)
print("ms per step", out["telapsed"] / K * 1e3, "iters", out["iters"].mean())
