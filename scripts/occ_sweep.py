"""Latency vs occupancy experiment: per-instance solve latency with 1..5 CTAs per SM."""
import os, sys, subprocess, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    import numpy as np
    import mpc_sensorlessao_b200 as pk
    from mpc_sensorlessao_b200 import synth
    c = int(sys.argv[1])
    p = synth.make_problem(6, 20)
    nb = 148 * c * 4
    wi = synth.warm_inputs(p, nb)
    hb = pk.FastMPCBatch(p.A1, p.A2, p.B, p.Q, p.R, p.Qf, p.u_min, p.u_max, 20, p.x_min, p.x_max, max_batch=nb)
    best = 1e9
    for _ in range(4):
        out = hb.step(wi['x0'], wi['x0_pre'], None, None, wi['X0'], wi['U0'], wi['nu0'], kappa=0.01, niters=5)
        best = min(best, out['telapsed'])
    print(f"ctas/SM {c}: batch {nb}, kernel {best*1e3:.3f} ms, per-instance latency {best/4*1.965e9/1e3:.0f} kcycles, {nb/best:.0f} solves/s", flush=True)
else:
    for c in (1, 2, 3, 4, 5):
        env = dict(os.environ, FMPC_CTAS_PER_SM=str(c))
        subprocess.run([sys.executable, __file__, str(c)], env=env)
