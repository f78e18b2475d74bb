"""GPU probe for the warp-per-instance kernel: parity vs the structured C oracle on many shapes, then
C2 throughput; with FMPC_B200_LIB=<prof build> also the phase cycle breakdown."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import mpc_sensorlessao_b200 as pk
from mpc_sensorlessao_b200 import synth
from oracle import fmpc_ref as fr
from cases import small_problem, ref_solve, relerr

PH = ["init", "newton_pass", "fwd_sweep", "bwd_sweep", "Ct_pass", "linesearch", "accept", "copyout", "fwdA", "fwdB_potrf", "fwdC", "x"]
mode = sys.argv[1] if len(sys.argv) > 1 else "all"


def run_small(kw, niters=6, kappa=0.01):
    c = small_problem(**kw)
    hb = pk.FastMPCBatch(c["A1"], c["A2"], c["B"], c["Q"], c["R"], c["Qf"], c["u_min"], c["u_max"], c["T"], c["x_min"], c["x_max"],
                         max_batch=c["nb"])
    kind = hb.kernel_kind
    out = hb.step(c["x0"], c["x0_pre"], c["w"], c["xf"], c["X0"], c["U0"], c["nu0"], kappa=kappa, niters=niters)
    hb.close()
    ref = ref_solve(fr, c, niters, kappa)
    eu = max(relerr(out["U"][b], ref["U"][b]) for b in range(c["nb"]))
    ex = max(relerr(out["X"][b], ref["X"][b]) for b in range(c["nb"]))
    ok = eu < 1e-9 and ex < 1e-9 and np.array_equal(out["iters"], ref["iters"]) and np.array_equal(out["status"], ref["status"])
    print(f"{'OK ' if ok else 'BAD'} kind{kind} {kw}: errU {eu:.2e} errX {ex:.2e} iters {out['iters'].tolist()} ref {ref['iters'].tolist()} "
          f"status {out['status'].tolist()} ref {ref['status'].tolist()} halv {ref['halvings'].tolist()}", flush=True)
    return ok


if mode in ("all", "parity"):
    print("devices", pk.device_count(), flush=True)
    cases = [
        dict(seed=1, n=6, m=4, T=5, nb=3, umax=2.0),
        dict(seed=2, n=6, m=4, T=5, nb=3, umax=0.3, xf=True),
        dict(seed=3, n=6, m=4, T=5, nb=3, umax=0.3, a2=False),
        dict(seed=4, n=8, m=5, T=10, nb=5, umax=0.2, xf=True, warm=True),
        dict(seed=5, n=8, m=5, T=10, nb=5, umax=0.1, warm=True),
        dict(seed=6, n=5, m=7, T=1, nb=2, umax=0.5),
        dict(seed=7, n=5, m=7, T=2, nb=2, umax=0.5, xf=True),
        dict(seed=9, n=1, m=1, T=3, nb=2, umax=1.0),
        dict(seed=13, n=12, m=9, T=7, nb=3, umax=0.4, warm=True),
        dict(seed=14, n=16, m=11, T=9, nb=3, umax=0.4, warm=True, xf=True),
        dict(seed=15, n=20, m=17, T=26, nb=3, umax=0.4, warm=True),
        dict(seed=16, n=24, m=8, T=4, nb=2, umax=0.4, warm=True),
        dict(seed=17, n=30, m=33, T=5, nb=2, umax=0.4, warm=True),
        dict(seed=18, n=32, m=16, T=5, nb=2, umax=0.4, warm=True, xf=True),
        dict(seed=10, n=27, m=144, T=10, nb=4, umax=3.0, a2=False, warm=True),
        dict(seed=11, n=28, m=144, T=20, nb=6, umax=0.5, warm=True),
        dict(seed=19, n=28, m=144, T=20, nb=40, umax=0.5, warm=True, xf=True),
    ]
    nbad = sum(0 if run_small(kw) else 1 for kw in cases)
    # line-search regimes
    for seed in (59, 61, 63, 67, 40, 41):
        c = small_problem(seed, 6, 5, 6, 1, 0.05, warm=True)
        c["U0"] = np.clip(c["U0"] * 10, -0.0499, 0.0499)
        hb = pk.FastMPCBatch(c["A1"], c["A2"], c["B"], c["Q"], c["R"], c["Qf"], c["u_min"], c["u_max"], c["T"], c["x_min"], c["x_max"], max_batch=1)
        out = hb.step(c["x0"], c["x0_pre"], c["w"], c["xf"], c["X0"], c["U0"], c["nu0"], kappa=0.01, niters=4)
        hb.close()
        ref = ref_solve(fr, c, 4, 0.01)
        eu, ex = relerr(out["U"][0], ref["U"][0]), relerr(out["X"][0], ref["X"][0])
        ok = eu < 1e-9 and ex < 1e-9
        nbad += 0 if ok else 1
        print(f"{'OK ' if ok else 'BAD'} linesearch seed {seed}: errU {eu:.2e} errX {ex:.2e} halv {ref['halvings'].tolist()} iters {out['iters'].tolist()} ref {ref['iters'].tolist()}", flush=True)
    print("PARITY FAILURES:", nbad, flush=True)

if mode in ("all", "perf"):
    p = synth.make_problem(6, 20)
    for nb in (888, 4096):
        wi = synth.warm_inputs(p, nb)
        hb = pk.FastMPCBatch(p.A1, p.A2, p.B, p.Q, p.R, p.Qf, p.u_min, p.u_max, 20, p.x_min, p.x_max, max_batch=nb)
        F = 20 * (28 * 28 * 144 + 19 / 3 * 28 ** 3 + 8 * 28 * 144 + 26 * 28 * 28)
        for rep in range(3):
            out = hb.step(wi['x0'], wi['x0_pre'], None, None, wi['X0'], wi['U0'], wi['nu0'], kappa=0.01, niters=5)
            its = int(out['iters'].sum())
            print(f"C2 kind{hb.kernel_kind} nb={nb}: kernel {out['telapsed']*1e3:.3f} ms, {nb/out['telapsed']:.0f} solves/s, iters/solve {its/nb:.2f}, "
                  f"{its*F/out['telapsed']/1e12:.3f} TFLOP/s (model)", np.bincount(out['status']).tolist(), flush=True)
        prof = hb.last_profile()
        if prof.sum() > 0:
            tot = prof[:8].sum()
            print("phase cycles per solve (warp-local):", {PH[i]: int(prof[i] / nb) for i in range(11)}, "total", int(tot / nb), flush=True)
        if nb == 4096:
            ref = fr.solve_batch(p.A1, p.A2, p.B, p.Q, p.R, p.Qf, p.u_min, p.u_max, 0.01, 5, wi['x0'][:64].T, wi['x0_pre'][:64].T, None,
                                 np.concatenate([wi['U0'][:64], wi['X0'][:64]], axis=2).reshape(64, -1).T, wi['nu0'][:64].T)
            Z = ref['z'].T.reshape(64, 20, 28 + 144)
            eu = max(relerr(out['U'][b], Z[b, :, :144]) for b in range(64)); ex = max(relerr(out['X'][b], Z[b, :, 144:]) for b in range(64))
            print(f"C2 parity (64 of {nb}): errU {eu:.2e} errX {ex:.2e} iters equal {np.array_equal(out['iters'][:64], ref['iters'])}", flush=True)
        hb.close()
    # tight bounds: several Newton iterations per solve
    p = synth.make_problem(6, 20, u_bound=1.0)
    nb = 2048
    wi = synth.warm_inputs(p, nb)
    hb = pk.FastMPCBatch(p.A1, p.A2, p.B, p.Q, p.R, p.Qf, p.u_min, p.u_max, 20, p.x_min, p.x_max, max_batch=nb)
    for rep in range(2):
        out = hb.step(wi['x0'], wi['x0_pre'], None, None, wi['X0'], wi['U0'], wi['nu0'], kappa=0.01, niters=10)
        its = int(out['iters'].sum())
        print(f"tight kind{hb.kernel_kind} nb={nb}: kernel {out['telapsed']*1e3:.3f} ms, {nb/out['telapsed']:.0f} solves/s, iters/solve {its/nb:.2f}, "
              f"{its*F/out['telapsed']/1e12:.3f} TFLOP/s (model)", np.bincount(out['status']).tolist(), flush=True)
    hb.close()
