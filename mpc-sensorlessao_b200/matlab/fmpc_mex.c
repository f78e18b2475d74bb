/* fmpc_mex.c -- thin MEX gateway onto the C-ABI of include/fmpc.h (pure marshalling, no arithmetic).
 *
 * The reference has no MEX layer: its fastMPC path is the MATLAB class Fast_MPC2
 * (Fast_MPC/VAR_2/Fast_MPC2.m).  Fast_MPC2_b200.m keeps that class surface and calls this gateway, so
 * README.md:548-570 runs unmodified apart from the class name.
 *
 *   h = fmpc_mex('create', sys, max_batch, device)      sys: struct with the fields of fmpc_sys
 *   [z, status, iters, telapsed] = fmpc_mex('step', h, params, x0, x0_pre, u_prev, w, xf, z0, nu0)
 *   [z, status, iters, telapsed] = fmpc_mex('frontend', h, mode, params, kmin, kmax, x0, x0_pre, u_prev, w, xf, z0, nu0)
 *   [u0, status, iters, telapsed] = fmpc_mex('step_r', h, params, reset, x0, x0_pre, u_prev, w, xf, nu0)     resident closed-loop step
 *   x_next = fmpc_mex('state_update', h, x, x_pre, u, w)
 *   fmpc_mex('destroy', h)
 *   hz = fmpc_mex('zmf_create', nL, N, max_frames, device);  c = fmpc_mex('zmf_fit', hz, frames);  fmpc_mex('zmf_destroy', hz)
 *   frames = fmpc_mex('zmf_synth', hz, coef)             coef: nmodes x nf  ->  nL x nL x nf  (README.md:592-598)
 *   he = fmpc_mex('est_create', A_s, b_s, max_batch, device);  x_hat = fmpc_mex('est_apply', he, Y);  fmpc_mex('est_destroy', he)
 *
 * Build (on a machine with MATLAB):  mex -R2018a fmpc_mex.c -I../../include -L../lib -lfmpc_b200
 * Here it is only compile-checked against tests/stubs/mex.h (there is no MATLAB in the image).
 */
#include <string.h>
#include <stdint.h>
#include "mex.h"
#include "fmpc.h"

static const double *opt(const mxArray *a) { return (a && !mxIsEmpty(a)) ? mxGetPr(a) : NULL; }   /* [] -> NULL */

static const mxArray *field(const mxArray *s, const char *name)
{
    const mxArray *f = mxGetField(s, 0, name);
    return (f && !mxIsEmpty(f)) ? f : NULL;
}
static const double *fieldp(const mxArray *s, const char *name) { const mxArray *f = field(s, name); return f ? mxGetPr(f) : NULL; }
static int fieldi(const mxArray *s, const char *name, int dflt) { const mxArray *f = field(s, name); return f ? (int)mxGetScalar(f) : dflt; }
static double fieldd(const mxArray *s, const char *name, double dflt) { const mxArray *f = field(s, name); return f ? mxGetScalar(f) : dflt; }

static void *get_handle(const mxArray *a)
{
    if (!mxIsUint64(a) || mxGetNumberOfElements(a) != 1) mexErrMsgIdAndTxt("fmpc:handle", "invalid handle");
    return (void *)(uintptr_t)(*(const uint64_t *)mxGetData(a));
}
static mxArray *put_handle(void *h)
{
    mxArray *a = mxCreateNumericMatrix(1, 1, mxUINT64_CLASS, mxREAL);
    *(uint64_t *)mxGetData(a) = (uint64_t)(uintptr_t)h;
    return a;
}
static void check(int rc) { if (rc) mexErrMsgIdAndTxt("fmpc:error", "%s", fmpc_strerror(rc)); }   /* reference error() strings */

static void read_params(const mxArray *s, fmpc_params *p)
{
    fmpc_default_params(p);
    if (!s || mxIsEmpty(s)) return;
    p->kappa = fieldd(s, "kappa", p->kappa);   p->niters = fieldi(s, "niters", p->niters);
    p->ls_max = fieldi(s, "ls_max", p->ls_max); p->alpha = fieldd(s, "alpha", p->alpha);
    p->beta = fieldd(s, "beta", p->beta);       p->tol_r = fieldd(s, "tol_r", p->tol_r);
    p->tol_p = fieldd(s, "tol_p", p->tol_p);
}

/* shared by 'step' and 'frontend': a[] = x0, x0_pre, u_prev, w, xf, z0, nu0 */
static void run_solve(int nlhs, mxArray *plhs[], fmpc_handle *h, int mode, const fmpc_params *p, double kmin, double kmax,
                      const mxArray *const *a, int n, int m, int T)
{
    const int nb = (int)mxGetN(a[0]);
    const size_t N = (size_t)T * (n + m);
    double te = 0.0;
    (void)nlhs;
    plhs[0] = mxCreateDoubleMatrix(N, nb, mxREAL);
    plhs[1] = mxCreateNumericMatrix(nb, 1, mxINT32_CLASS, mxREAL);
    plhs[2] = mxCreateNumericMatrix(nb, 1, mxINT32_CLASS, mxREAL);
    if (a[5] && !mxIsEmpty(a[5]) && mxGetM(a[5]) != N) check(FMPC_ERR_INIT_SIZE);          /* fast_mpc_init.m:13-14 */
    if (mode == 0) {
        check(fmpc_step_z(h, p, nb, opt(a[0]), opt(a[1]), opt(a[2]), opt(a[3]), opt(a[4]), opt(a[5]), opt(a[6]),
                          mxGetPr(plhs[0]), (int *)mxGetData(plhs[1]), (int *)mxGetData(plhs[2]), &te));
    } else {
        /* front-ends work on (X, U); de-interleave in MATLAB order (README.md:558-570) */
        mxArray *X = mxCreateDoubleMatrix((size_t)n * T, nb, mxREAL), *U = mxCreateDoubleMatrix((size_t)m * T, nb, mxREAL);
        double *z = mxGetPr(plhs[0]);
        const double *z0 = opt(a[5]);
        int b, t;
        if (z0)
            for (b = 0; b < nb; ++b)
                for (t = 0; t < T; ++t) {
                    memcpy(mxGetPr(U) + ((size_t)b * T + t) * m, z0 + b * N + (size_t)t * (n + m), sizeof(double) * m);
                    memcpy(mxGetPr(X) + ((size_t)b * T + t) * n, z0 + b * N + (size_t)t * (n + m) + m, sizeof(double) * n);
                }
        check(fmpc_frontend(h, mode, p, kmin, kmax, nb, opt(a[0]), opt(a[1]), opt(a[2]), opt(a[3]), opt(a[4]),
                            z0 ? mxGetPr(X) : NULL, z0 ? mxGetPr(U) : NULL, opt(a[6]), mxGetPr(X), mxGetPr(U),
                            (int *)mxGetData(plhs[1]), (int *)mxGetData(plhs[2]), &te));
        for (b = 0; b < nb; ++b)
            for (t = 0; t < T; ++t) {
                memcpy(z + b * N + (size_t)t * (n + m), mxGetPr(U) + ((size_t)b * T + t) * m, sizeof(double) * m);
                memcpy(z + b * N + (size_t)t * (n + m) + m, mxGetPr(X) + ((size_t)b * T + t) * n, sizeof(double) * n);
            }
        mxDestroyArray(X); mxDestroyArray(U);
    }
    plhs[3] = mxCreateDoubleScalar(te);
}

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[])
{
    char cmd[32];
    if (nrhs < 1 || mxGetString(prhs[0], cmd, sizeof cmd)) mexErrMsgIdAndTxt("fmpc:usage", "first argument must be a command string");

    if (!strcmp(cmd, "create")) {
        const mxArray *s = prhs[1];
        fmpc_sys sys;
        fmpc_handle *h = NULL;
        memset(&sys, 0, sizeof sys);
        sys.n = fieldi(s, "n", 0); sys.m = fieldi(s, "m", 0); sys.T = fieldi(s, "T", 0);
        sys.var_order = fieldi(s, "var_order", 2); sys.ramp_rows = fieldi(s, "ramp_rows", 0);
        sys.var1_literal_bug = fieldi(s, "var1_literal_bug", 0);
        sys.A1 = fieldp(s, "A1"); sys.A2 = fieldp(s, "A2"); sys.B = fieldp(s, "B");
        sys.Q = fieldp(s, "Q"); sys.R = fieldp(s, "R"); sys.Qf = fieldp(s, "Qf");
        sys.q = fieldp(s, "q"); sys.r = fieldp(s, "r"); sys.qf = fieldp(s, "qf");
        sys.x_min = fieldp(s, "x_min"); sys.x_max = fieldp(s, "x_max"); sys.u_min = fieldp(s, "u_min"); sys.u_max = fieldp(s, "u_max");
        sys.du_min = fieldp(s, "du_min"); sys.du_max = fieldp(s, "du_max");
        check(fmpc_create(&h, &sys, nrhs > 2 ? (int)mxGetScalar(prhs[2]) : 1, nrhs > 3 ? (int)mxGetScalar(prhs[3]) : 0));
        plhs[0] = put_handle(h);
    } else if (!strcmp(cmd, "step") || !strcmp(cmd, "frontend")) {
        /* step:     h, params, x0, x0_pre, u_prev, w, xf, z0, nu0
         * frontend: h, mode, params, kmin, kmax, x0, ...                                                  */
        const int fe = !strcmp(cmd, "frontend");
        const int base = fe ? 6 : 3;
        const mxArray *ps = prhs[fe ? 3 : 2];
        fmpc_params p;
        const mxArray *a[7];
        int i;
        if (nrhs < base + 7) mexErrMsgIdAndTxt("fmpc:usage", "not enough arguments");
        read_params(ps, &p);
        for (i = 0; i < 7; ++i) a[i] = prhs[base + i];
        {
            fmpc_handle *h = (fmpc_handle *)get_handle(prhs[1]);
            int n = 0, m = 0, T = 0;
            check(fmpc_get_dims(h, &n, &m, &T));
            run_solve(nlhs, plhs, h, fe ? (int)mxGetScalar(prhs[2]) : 0, &p,
                      fe ? mxGetScalar(prhs[4]) : 0.0, fe ? mxGetScalar(prhs[5]) : 0.0, a, n, m, T);
        }
    } else if (!strcmp(cmd, "step_r")) {
        /* h, params, reset, x0, x0_pre, u_prev, w, xf, nu0  ->  U(:,0) per instance (README.md:589), status, iters, telapsed */
        fmpc_handle *h = (fmpc_handle *)get_handle(prhs[1]);
        fmpc_params p;
        int n = 0, m = 0, T = 0, nb;
        double te = 0.0;
        if (nrhs < 10) mexErrMsgIdAndTxt("fmpc:usage", "not enough arguments");
        read_params(prhs[2], &p);
        check(fmpc_get_dims(h, &n, &m, &T));
        nb = (int)mxGetN(prhs[4]);
        plhs[0] = mxCreateDoubleMatrix((mwSize)m, (mwSize)nb, mxREAL);
        plhs[1] = mxCreateNumericMatrix(nb, 1, mxINT32_CLASS, mxREAL);
        plhs[2] = mxCreateNumericMatrix(nb, 1, mxINT32_CLASS, mxREAL);
        check(fmpc_step_r(h, &p, nb, mxGetScalar(prhs[3]) != 0.0 ? FMPC_R_RESET : 0, opt(prhs[4]), opt(prhs[5]), opt(prhs[6]),
                          opt(prhs[7]), opt(prhs[8]), opt(prhs[9]), mxGetPr(plhs[0]), NULL, NULL, (int *)mxGetData(plhs[1]),
                          (int *)mxGetData(plhs[2]), &te));
        plhs[3] = mxCreateDoubleScalar(te);
    } else if (!strcmp(cmd, "state_update")) {
        const int nb = (int)mxGetN(prhs[2]);
        plhs[0] = mxCreateDoubleMatrix(mxGetM(prhs[2]), nb, mxREAL);
        check(fmpc_state_update((fmpc_handle *)get_handle(prhs[1]), nb, opt(prhs[2]), opt(prhs[3]), opt(prhs[4]),
                                nrhs > 5 ? opt(prhs[5]) : NULL, mxGetPr(plhs[0])));
    } else if (!strcmp(cmd, "destroy")) {
        fmpc_destroy((fmpc_handle *)get_handle(prhs[1]));
    } else if (!strcmp(cmd, "zmf_create")) {
        zmf_handle *h = NULL;
        check(zmf_create(&h, (int)mxGetScalar(prhs[1]), (int)mxGetScalar(prhs[2]), nrhs > 3 ? (int)mxGetScalar(prhs[3]) : 2048,
                         nrhs > 4 ? (int)mxGetScalar(prhs[4]) : 0));
        plhs[0] = put_handle(h);
    } else if (!strcmp(cmd, "zmf_fit")) {
        zmf_handle *h = (zmf_handle *)get_handle(prhs[1]);
        const mwSize *d = mxGetDimensions(prhs[2]);
        const int nf = mxGetNumberOfDimensions(prhs[2]) > 2 ? (int)d[2] : 1;
        plhs[0] = mxCreateDoubleMatrix(zmf_nmodes(h), nf, mxREAL);
        check(zmf_fit(h, nf, mxGetPr(prhs[2]), mxGetPr(plhs[0]), NULL));
    } else if (!strcmp(cmd, "zmf_synth")) {
        zmf_handle *h = (zmf_handle *)get_handle(prhs[1]);
        const int nf = (int)mxGetN(prhs[2]);
        const int nL = (int)mxGetScalar(prhs[3]);            /* frame size the handle was created with */
        mwSize dims[3];
        dims[0] = (mwSize)nL; dims[1] = (mwSize)nL; dims[2] = (mwSize)nf;
        plhs[0] = mxCreateNumericArray(3, dims, mxDOUBLE_CLASS, mxREAL);
        check(zmf_synth(h, nf, mxGetPr(prhs[2]), mxGetPr(plhs[0]), NULL));
    } else if (!strcmp(cmd, "zmf_destroy")) {
        zmf_destroy((zmf_handle *)get_handle(prhs[1]));
    } else if (!strcmp(cmd, "est_create")) {
        est_handle *h = NULL;
        check(est_create(&h, (int)mxGetM(prhs[1]), (int)mxGetN(prhs[1]), mxGetPr(prhs[1]), opt(prhs[2]),
                         nrhs > 3 ? (int)mxGetScalar(prhs[3]) : 1, nrhs > 4 ? (int)mxGetScalar(prhs[4]) : 0));
        plhs[0] = put_handle(h);
    } else if (!strcmp(cmd, "est_apply")) {
        est_handle *h = (est_handle *)get_handle(prhs[1]);
        const int nb = (int)mxGetN(prhs[2]);
        plhs[0] = mxCreateDoubleMatrix((mwSize)zmf_nmodes(h), (mwSize)nb, mxREAL);
        check(est_apply(h, nb, mxGetPr(prhs[2]), mxGetPr(plhs[0]), NULL));
    } else if (!strcmp(cmd, "est_destroy")) {
        est_destroy((est_handle *)get_handle(prhs[1]));
    } else {
        mexErrMsgIdAndTxt("fmpc:usage", "unknown command '%s'", cmd);
    }
}
