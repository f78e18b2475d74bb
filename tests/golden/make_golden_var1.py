"""Golden fixtures of the reference's VAR_1 solver exactly as written -- ramp-rate rows
(VAR_1/fast_mpc_ineq_const.m:58-79) and the literal column placement of the second block row of C
(VAR_1/fast_mpc_eq_const.m:34-37) -- from the literal dense oracle (oracle/fastmpc_dense.py).

    python tests/golden/make_golden_var1.py      # rewrites tests/golden/var1lit_*.npz (~1 min)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from cases import var1_literal_case, var1_literal_dense  # noqa: E402
from oracle import fastmpc_dense as fd  # noqa: E402
import mpc_sensorlessao_b200  # noqa: E402,F401
from mpc_sensorlessao_b200 import synth  # noqa: E402


def readme_c1(nb, seed=7):
    """BASELINE.json configs[0] shape: VAR(1), n = 27 (piston dropped), m = 144, T = 10, README weights and bounds
    (|u| <= 28, du = +-0.2121, README.md:352-356), warm start, u_prev = first warm input."""
    p = synth.make_problem(6, 10, var_order=1, drop_piston=True, u_bound=28.0)
    wi = synth.warm_inputs(p, nb, seed=seed)
    rs = np.random.RandomState(seed)
    U0 = wi["U0"].copy()
    # a warm start that respects the ramp rows: a slow random walk from u_prev
    u_prev = U0[:, 0, :].copy()
    steps = 0.05 * rs.randn(nb, p.T, p.m)
    U0 = u_prev[:, None, :] + np.cumsum(steps, axis=1)
    return dict(n=p.n, m=p.m, T=p.T, nb=nb, A1=p.A1, A2=None, B=p.B, Q=p.Q, R=p.R, Qf=p.Qf, u_min=p.u_min, u_max=p.u_max,
                x_min=p.x_min, x_max=p.x_max, du_min=-0.2121 * np.ones(p.m), du_max=0.2121 * np.ones(p.m),
                x0=wi["x0"], x0_pre=None, u_prev=u_prev, w=np.zeros((nb, p.T * p.n)), xf=None, nu0=wi["nu0"],
                X0=wi["X0"], U0=U0)


def feasible_case():
    """Small states: the iterates stay inside the box and the ramp bounds (the regime an MPC is meant to run in)."""
    c = var1_literal_case(203, 6, 9, 6, 3, 0.6, 0.15)
    sc = 0.2
    c["x0"] *= sc; c["w"] *= sc; c["X0"] *= sc
    c["U0"] = 0.2 * c["U0"]; c["u_prev"] = 0.2 * c["u_prev"]
    return c


def main():
    cases = {
        "small_ramp_feasible": (feasible_case(), 8, 0.01, True, True),
        "small_ramp_bug": (var1_literal_case(201, 6, 9, 6, 3, 0.6, 0.15), 5, 0.01, True, True),
        "small_ramp_only_xf": (var1_literal_case(202, 7, 5, 8, 3, 0.5, 0.1, xf=True), 5, 0.01, True, False),
        "readme_c1_n27_T10": (readme_c1(2), 3, 0.01, True, True),
    }
    for name, (c, niters, kappa, ramp, bug) in cases.items():
        Us, Xs, its, ee = [], [], [], []
        for b in range(c["nb"]):
            z, st = var1_literal_dense(fd, c, b, niters, kappa, ramp=ramp, bug=bug)
            U, X = fd.deinterleave(z, c["n"], c["m"], c["T"])
            Us.append(U.T), Xs.append(X.T), its.append(st["iters"]), ee.append(st["early_exit"])
        out = {k: v for k, v in c.items() if isinstance(v, np.ndarray)}
        out.update(n=c["n"], m=c["m"], T=c["T"], nb=c["nb"], niters=niters, kappa=kappa, ramp=ramp, bug=bug,
                   U=np.array(Us), X=np.array(Xs), iters=np.array(its), early_exit=np.array(ee))
        np.savez_compressed(os.path.join(HERE, f"var1lit_{name}.npz"), **out)
        print(name, "iters", out["iters"], "early", out["early_exit"], flush=True)


if __name__ == "__main__":
    main()
