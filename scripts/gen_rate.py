"""Device-side rate of the general-structure kernel at the reference's own VAR_1 problem (BASELINE.json configs[0] shape:
n = 27, m = 144, T = 10, ramp rows + literal C), `nb` instances in one batched solve (fmpc_step_d, inputs in HBM, CUDA events).
With a -DFMPC_PROF build (FMPC_B200_LIB=.../libfmpc_b200_prof.so) also prints the phase shares of the last launch.
usage: gen_rate.py [nb] [reps]"""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mpc_sensorlessao_b200 as pk
from mpc_sensorlessao_b200 import synth
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
p = synth.make_problem(6, 10, var_order=1, drop_piston=True, u_bound=28.0)
wi = synth.warm_inputs(p, nb, seed=7)
u_prev = wi["U0"][:, 0] + 0.01 * np.random.RandomState(8).randn(nb, p.m)
dev = torch.device("cuda", 0)
hb = pk.FastMPCBatch(p.A1, None, p.B, p.Q, p.R, p.Qf, p.u_min, p.u_max, p.T, p.x_min, p.x_max, du_min=p.du_min, du_max=p.du_max,
                     ramp_rows=True, var1_literal_bug=True, max_batch=nb)
d = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in dict(x0=wi["x0"], X0=wi["X0"], U0=wi["U0"], nu0=wi["nu0"], u_prev=u_prev).items()}
X = torch.empty((nb, p.T, p.n), dtype=torch.float64, device=dev); U = torch.empty((nb, p.T, p.m), dtype=torch.float64, device=dev)
st = torch.empty(nb, dtype=torch.int32, device=dev); it = torch.empty(nb, dtype=torch.int32, device=dev)
params = hb.params(0.01, 5, 0)
vp = lambda t: C.c_void_p(t.data_ptr())
s = torch.cuda.Stream(dev)
def step():
    rc = hb._L.fmpc_step_d(hb._h, C.byref(params), nb, vp(d["x0"]), None, vp(d["u_prev"]), None, None, vp(d["X0"]), vp(d["U0"]), vp(d["nu0"]),
                           vp(X), vp(U), vp(st), vp(it), C.c_void_p(s.cuda_stream))
    assert rc == 0, rc
step(); torch.cuda.synchronize()
e = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
e[0].record(s)
for i in range(reps):
    step(); e[i + 1].record(s)
torch.cuda.synchronize()
ms = np.array([e[i].elapsed_time(e[i + 1]) for i in range(reps)])
its = int(it.sum().item())
print(f"{os.environ.get('KR_TAG', '')} C1 nb={nb} kind={hb.kernel_kind}: {np.median(ms):.3f} ms (min {ms.min():.3f}), {nb / np.median(ms) * 1e3:.0f} solves/s, "
      f"iters/solve {its / nb:.2f}, {its / np.median(ms) * 1e3:.0f} Newton iterations/s, status {np.bincount(st.cpu().numpy(), minlength=5).tolist()}, "
      f"checksum {float(U.double().sum().item()):.12e}", flush=True)
prof = np.array(hb.last_profile(), dtype=np.float64)
if prof.sum() > 0:
    names = ["init", "barrier+resid", "inv(Phi_uu)", "rhs", "Schur assembly", "potrf+fwd", "panel", "trailing", "backward", "dz", "line search", "copy-out"]
    print("  phase shares: " + ", ".join(f"{n} {100 * v / prof.sum():.1f}%" for n, v in zip(names, prof)) + f"  (cycles/iteration/CTA {prof.sum() / max(its, 1):.0f})", flush=True)
hb.close()
