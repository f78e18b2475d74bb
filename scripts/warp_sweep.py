"""Throughput of the C2 workload vs warps per CTA (FMPC_WARPS_PER_CTA), to separate latency-bound from resource-bound behaviour."""
import os, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, numpy as np
sys.path.insert(0, %r)
import mpc_sensorlessao_b200 as pk
from mpc_sensorlessao_b200 import synth
p = synth.make_problem(6, 20); nb = 4096
wi = synth.warm_inputs(p, nb)
hb = pk.FastMPCBatch(p.A1, p.A2, p.B, p.Q, p.R, p.Qf, p.u_min, p.u_max, 20, p.x_min, p.x_max, max_batch=nb)
best = 1e9
for rep in range(4):
    out = hb.step(wi['x0'], wi['x0_pre'], None, None, wi['X0'], wi['U0'], wi['nu0'], kappa=0.01, niters=5)
    best = min(best, out['telapsed'])
print("%%.3f ms  %%.0f solves/s" %% (best*1e3, nb/best))
''' % root
for w in (sys.argv[1:] or ["2", "4", "6", "7", "8"]):
    env = dict(os.environ, FMPC_WARPS_PER_CTA=w)
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    print(f"warps/CTA {w}: {out.stdout.strip()} {out.stderr.strip()[-200:]}", flush=True)
