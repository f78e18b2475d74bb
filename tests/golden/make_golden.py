"""Generates the committed golden fixtures from the LITERAL DENSE oracle (oracle/fastmpc_dense.py,
oracle/zernike_ref.py).  The reference itself holds no golden vectors and cannot run here
(MATLAB), so these pin the oracle's output at commit time: any later change to
the oracle, the structured C port or the CUDA path that moves a result shows up against them.

    python tests/golden/make_golden.py        # rewrites tests/golden/*.npz  (~1-2 min)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from cases import dense_solve, small_problem  # noqa: E402
from oracle import fastmpc_dense as fd  # noqa: E402
from oracle import zernike_ref as zr  # noqa: E402
import mpc_sensorlessao_b200  # noqa: E402,F401
from mpc_sensorlessao_b200 import synth  # noqa: E402


def pack_case(c, niters, kappa):
    Us, Xs, its, ee = [], [], [], []
    for b in range(c["nb"]):
        z, st = dense_solve(fd, c, b, niters, kappa)
        U, X = fd.deinterleave(z, c["n"], c["m"], c["T"])
        Us.append(U.T), Xs.append(X.T), its.append(st["iters"]), ee.append(st["early_exit"])
    out = {k: v for k, v in c.items() if isinstance(v, np.ndarray)}
    out.update(n=c["n"], m=c["m"], T=c["T"], nb=c["nb"], niters=niters, kappa=kappa, U=np.array(Us), X=np.array(Xs),
               iters=np.array(its), early_exit=np.array(ee), has_a2=c["A2"] is not None, has_xf=c["xf"] is not None,
               warm=c["X0"] is not None)
    return out


def synth_case(N, T, nb, u_bound, var_order=2, drop_piston=False, seed=2):
    p = synth.make_problem(N, T, var_order=var_order, drop_piston=drop_piston, u_bound=u_bound)
    wi = synth.warm_inputs(p, nb, seed=seed)
    return dict(n=p.n, m=p.m, T=T, nb=nb, A1=p.A1, A2=p.A2, B=p.B, Q=p.Q, R=p.R, Qf=p.Qf, u_min=p.u_min, u_max=p.u_max,
                x_min=p.x_min, x_max=p.x_max, x0=wi["x0"], x0_pre=wi["x0_pre"] if var_order == 2 else None,
                w=np.zeros((nb, T * p.n)), xf=None, nu0=wi["nu0"], X0=wi["X0"], U0=wi["U0"])


def main():
    cases = {
        "small_var2_cold": (small_problem(101, 6, 4, 5, 3, 2.0), 5, 0.01),
        "small_var2_xf_warm": (small_problem(102, 8, 5, 10, 3, 0.2, xf=True, warm=True), 5, 0.01),
        "small_var1_warm": (small_problem(103, 7, 9, 6, 3, 0.3, a2=False, warm=True), 6, 0.01),
        "small_tight_ls": (small_problem(61, 6, 5, 6, 1, 0.05, warm=True), 4, 0.01),
        "readme_c2_n28_T20": (synth_case(6, 20, 2, 28.0), 5, 0.01),
        "tight_c2_n28_T20": (synth_case(6, 20, 2, 1.0), 4, 0.01),
        "readme_c1_var1_n27_T10": (synth_case(6, 10, 2, 28.0, var_order=1, drop_piston=True), 5, 0.01),
    }
    c = cases["small_tight_ls"][0]
    c["U0"] = np.clip(c["U0"] * 10, -0.0499, 0.0499)
    for name, (c, niters, kappa) in cases.items():
        out = pack_case(c, niters, kappa)
        out = {k: v for k, v in out.items() if v is not None}
        np.savez_compressed(os.path.join(HERE, f"fmpc_{name}.npz"), **out)
        print(name, "iters", out["iters"], "early", out["early_exit"], flush=True)
    # zernmodfit: 6 frames 128 x 128, N = 6 and 3 frames N = 10; frames stored as float32-exact
    # values (so the fixture compresses) with NaN outside the pupil
    for N, nf in ((6, 6), (10, 3)):
        rs = np.random.RandomState(N)
        r, th, is_in = zr.pupil_grid(128)
        n_, m_ = zr.mode_indices(N)
        Z = zr.zernfun(n_, m_, r, th)
        frames = np.full((nf, 128 * 128), np.nan)
        frames[:, is_in.T.reshape(-1)] = (rs.randn(nf, n_.shape[0]) @ Z.T + 0.05 * rs.randn(nf, Z.shape[0])).astype(np.float32)
        frames = frames.reshape(nf, 128, 128).transpose(0, 2, 1)
        coef = zr.fit_frames_literal(frames, N)
        np.savez_compressed(os.path.join(HERE, f"zernmodfit_N{N}.npz"), frames=frames.astype(np.float32), coef=coef, N=N)
        print("zernmodfit", N, coef.shape)


if __name__ == "__main__":
    main()
