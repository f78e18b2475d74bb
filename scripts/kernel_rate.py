"""Device-side rate of the solve kernel alone (fmpc_step_d, inputs in HBM, CUDA events): C2 shape, `nb` instances.
usage: kernel_rate.py [nb] [reps] [u_bound]"""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mpc_sensorlessao_b200 as pk
from mpc_sensorlessao_b200 import synth
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
ub = float(sys.argv[3]) if len(sys.argv) > 3 else 28.0
N = int(os.environ.get("KR_N", "6")); T = int(os.environ.get("KR_T", "20"))
p = synth.make_problem(N, T, u_bound=ub)
wi = synth.warm_inputs(p, nb)
dev = torch.device("cuda", 0)
hb = pk.FastMPCBatch(p.A1, p.A2, p.B, p.Q, p.R, p.Qf, p.u_min, p.u_max, p.T, p.x_min, p.x_max, max_batch=nb)
d = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in wi.items()}
X = torch.empty((nb, p.T, p.n), dtype=torch.float64, device=dev); U = torch.empty((nb, p.T, p.m), dtype=torch.float64, device=dev)
st = torch.empty(nb, dtype=torch.int32, device=dev); it = torch.empty(nb, dtype=torch.int32, device=dev)
params = hb.params(0.01, 5, 0)
vp = lambda t: C.c_void_p(t.data_ptr())
s = torch.cuda.Stream(dev)
def step():
    rc = hb._L.fmpc_step_d(hb._h, C.byref(params), nb, vp(d["x0"]), vp(d["x0_pre"]), None, None, None, vp(d["X0"]), vp(d["U0"]), vp(d["nu0"]),
                           vp(X), vp(U), vp(st), vp(it), C.c_void_p(s.cuda_stream))
    assert rc == 0, rc
for _ in range(3): step()
torch.cuda.synchronize()
e = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
e[0].record(s)
for i in range(reps):
    step(); e[i + 1].record(s)
torch.cuda.synchronize()
ms = np.array([e[i].elapsed_time(e[i + 1]) for i in range(reps)])
its = int(it.sum().item())
F = p.T * (p.n ** 2 * p.m + 19 / 3 * p.n ** 3 + 8 * p.n * p.m + 26 * p.n ** 2)
print(f"{os.environ.get('KR_TAG', '')} nb={nb} kind={hb.kernel_kind}: {np.median(ms):.3f} ms (min {ms.min():.3f}), {nb / np.median(ms) * 1e3:.0f} solves/s, "
      f"iters/solve {its / nb:.2f}, {its * F / np.median(ms) / 1e9:.2f} TFLOP/s", flush=True)
hb.close()
