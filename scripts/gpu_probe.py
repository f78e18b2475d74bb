"""Ad-hoc GPU probe: parity of the CUDA solve against the structured C oracle + FP64 peaks."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mpc_sensorlessao_b200 as pk
from mpc_sensorlessao_b200 import synth
from oracle import fmpc_ref as fr

def relerr(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))

def run_case(name, A1, A2, B, Q, R, Qf, umin, umax, T, x0, x0p, w, X0, U0, nu0, xf=None, niters=5, kappa=0.01):
    nb, n = x0.shape; m = B.shape[1]
    hb = pk.FastMPCBatch(A1, A2, B, Q, R, Qf, umin, umax, T, -100*np.ones(n), 100*np.ones(n), max_batch=nb)
    t0 = time.time()
    out = hb.step(x0, x0p, w, xf, X0, U0, nu0, kappa=kappa, niters=niters)
    t1 = time.time()
    if X0 is None:
        z0 = np.tile(np.concatenate([(umin+umax)/2, np.zeros(n)]), T)[None].repeat(nb, 0)
    else:
        z0 = np.concatenate([U0, X0], axis=2).reshape(nb, -1)
    ref = fr.solve_batch(A1, A2, B, Q, R, Qf, umin, umax, kappa, niters, x0.T, None if x0p is None else x0p.T,
                         None if w is None else w.T, z0.T, nu0.T, xf=None if xf is None else xf.T)
    Z = ref['z'].T.reshape(nb, T, n + m)
    Uref, Xref = Z[:, :, :m], Z[:, :, m:]
    eu = max(relerr(out['U'][b], Uref[b]) for b in range(nb)); ex = max(relerr(out['X'][b], Xref[b]) for b in range(nb))
    print(f"{name}: errU {eu:.2e} errX {ex:.2e} iters gpu {out['iters'][:6]} ref {ref['iters'][:6]} status {out['status'][:6]} ref {ref['status'][:6]} "
          f"halv {ref['halvings'][:6]} kernel {out['telapsed']*1e3:.2f} ms wall {1e3*(t1-t0):.1f} ms", flush=True)
    hb.close()
    return eu, ex

def small(seed, n, m, T, nb, umax, a2=True, xf=False, warm=False, niters=5):
    rs = np.random.RandomState(seed)
    A1 = 0.5*np.eye(n)+0.1*rs.randn(n,n); A2 = (0.2*np.eye(n)+0.05*rs.randn(n,n)) if a2 else None; B = rs.randn(n,m)
    Q = np.diag(1+rs.rand(n))*3; R = np.diag(1+rs.rand(m)); Qf = Q*2
    x0 = rs.randn(nb,n); x0p = rs.randn(nb,n) if a2 else None; w = 0.1*rs.randn(nb,T*n)
    xfv = 0.1*rs.randn(nb,n) if xf else None
    nu0 = rs.rand(nb,(T+(1 if xf else 0))*n)
    um = umax*np.ones(m)
    X0=U0=None
    if warm:
        X0 = 0.5*rs.randn(nb,T,n); U0 = np.clip(0.5*rs.randn(nb,T,m), -0.9*umax, 0.9*umax)
    return run_case(f"small n{n} m{m} T{T} nb{nb} umax{umax} a2={a2} xf={xf} warm={warm}", A1,A2,B,Q,R,Qf,-um,um,T,x0,x0p,w,X0,U0,nu0,xfv,niters)

print("devices", pk.device_count(), flush=True)
for kind, nm in ((0, "DFMA"), (1, "DMMA m8n8k4")):
    print(nm, "TFLOP/s:", [round(pk.fp64_peak(0, kind, 4000), 2) for _ in range(2)], flush=True)

small(1, 6, 4, 5, 3, 2.0)
small(2, 6, 4, 5, 3, 0.3, xf=True)
small(3, 6, 4, 5, 3, 0.3, a2=False)
small(4, 8, 5, 10, 5, 0.2, xf=True, warm=True)
small(5, 8, 5, 10, 5, 0.1, warm=True, niters=8)
small(6, 5, 7, 1, 2, 0.5)
small(7, 5, 7, 2, 2, 0.5, xf=True)
small(8, 33, 20, 6, 4, 0.5, warm=True)

for N, T, nb in ((6, 20, 64), (6, 10, 16)):
    p = synth.make_problem(N, T)
    wi = synth.warm_inputs(p, nb)
    run_case(f"README-size n{p.n} m{p.m} T{T} nb{nb}", p.A1, p.A2, p.B, p.Q, p.R, p.Qf, p.u_min, p.u_max, T,
             wi['x0'], wi['x0_pre'], None, wi['X0'], wi['U0'], wi['nu0'])
# tight bounds so the barrier matters
p = synth.make_problem(6, 20, u_bound=1.0)
wi = synth.warm_inputs(p, 32)
run_case("tight n28 T20 nb32", p.A1, p.A2, p.B, p.Q, p.R, p.Qf, p.u_min, p.u_max, 20, wi['x0'], wi['x0_pre'], None, wi['X0'], wi['U0'], wi['nu0'], niters=10)

# throughput at C2
p = synth.make_problem(6, 20)
nb = 4096
wi = synth.warm_inputs(p, nb)
hb = pk.FastMPCBatch(p.A1, p.A2, p.B, p.Q, p.R, p.Qf, p.u_min, p.u_max, 20, p.x_min, p.x_max, max_batch=nb)
for rep in range(3):
    out = hb.step(wi['x0'], wi['x0_pre'], None, None, wi['X0'], wi['U0'], wi['nu0'], kappa=0.01, niters=5)
    its = int(out['iters'].sum())
    F = 20*(28*28*144 + 19/3*28**3 + 8*28*144 + 26*28*28)
    print(f"C2 nb={nb}: kernel {out['telapsed']*1e3:.2f} ms, {nb/out['telapsed']:.0f} solves/s, iters total {its} (mean {its/nb:.2f}), "
          f"{its*F/out['telapsed']/1e12:.3f} TFLOP/s (model)", np.bincount(out['status']), flush=True)
hb.close()
# n = 66
p = synth.make_problem(10, 30)
wi = synth.warm_inputs(p, 8)
run_case("C5-size n66 m144 T30 nb8", p.A1, p.A2, p.B, p.Q, p.R, p.Qf, p.u_min, p.u_max, 30, wi['x0'], wi['x0_pre'], None, wi['X0'], wi['U0'], wi['nu0'])

# zernike
from oracle import zernike_ref as zr
zf = pk.ZernikeFitter(128, 6, max_frames=256)
rs = np.random.RandomState(0)
r, th, is_in = zr.pupil_grid(128)
n_, m_ = zr.mode_indices(6)
Z = zr.zernfun(n_, m_, r, th)
print("basis err", np.abs(zf.basis() - Z).max(), "mask equal", bool((zf.mask() == is_in).all()))
c = rs.randn(8, 28)
frames = np.full((8, 128*128), np.nan)
frames[:, is_in.T.reshape(-1)] = c @ Z.T + 0.01*rs.randn(8, Z.shape[0])
frames = frames.reshape(8, 128, 128).transpose(0, 2, 1)
coef, tel = zf.fit(frames)
ref = zr.fit_frames_literal(frames, 6)
print("zernmodfit err", relerr(coef, ref), "kernel ms", tel*1e3)
