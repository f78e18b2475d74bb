"""Device MT19937 generator: duration of the fill kernel for one C2 step's worth of doubles (4096 x 560), per variant."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mpc_sensorlessao_b200 as pk
from mpc_sensorlessao_b200 import synth
p = synth.make_problem(6, 20)
nb = 4096
hb = pk.FastMPCBatch(p.A1, p.A2, p.B, p.Q, p.R, p.Qf, p.u_min, p.u_max, p.T, p.x_min, p.x_max, max_batch=nb)
x0 = torch.zeros((nb, p.n), dtype=torch.float64).pin_memory(); x0[:, 0] = 0.1
for k in range(8):
    t0 = time.perf_counter()
    out = hb.step_resident(x0.numpy(), reset=(k == 0), niters=5)
    dt = time.perf_counter() - t0
    if k >= 5:
        print(f"variant {os.environ.get('FMPC_MT_VARIANT', '0')} step {k}: wall {dt*1e3:.3f} ms, solve kernel {out['telapsed']*1e3:.3f} ms", flush=True)
# the generator alone: seed + one cold call that must generate before it can solve (nothing prefetched)
import ctypes as C
for rep in range(3):
    hb._L.fmpc_seed_stream(hb._h, 5489)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = hb.step_resident(x0.numpy(), reset=True, niters=5)
    dt = time.perf_counter() - t0
    print(f"  cold call (generate, then solve): wall {dt*1e3:.3f} ms, solve {out['telapsed']*1e3:.3f} ms -> generator ~ {(dt - out['telapsed'])*1e3 - 0.25:.2f} ms", flush=True)
hb.close()
