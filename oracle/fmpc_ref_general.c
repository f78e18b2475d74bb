/* ORACLE (test infrastructure, NOT product code).
 *
 * Second, independent CPU restatement of the reference's fastMPC Newton solver for the inputs the block-banded
 * restatement (fmpc_ref.c) does not cover -- VAR_1's ramp-rate rows (Fast_MPC/VAR_1/fast_mpc_ineq_const.m:58-79), the
 * literal column placement of the second block row of C (Fast_MPC/VAR_1/fast_mpc_eq_const.m:34-37), dense
 * Q / Qf -- so that the literal dense oracle (oracle/fastmpc_dense.py) is not the only thing that pins the CUDA
 * general-structure kernel.  Same algorithm as inf_newton_solver.m:1-43 / backtracking_inf_newton.m:2-11, different
 * route from both the dense oracle and the CUDA kernel:
 *   - C is built literally (dense, column-major), products with C use each row's nonzero column range;
 *   - Phi = 2H + k P'DP is never formed: its x blocks are 2Q / 2Qf (Cholesky solves), its u part is, for a diagonal R,
 *     one T x T tridiagonal system per actuator, solved by the Thomas algorithm for every right-hand side
 *     (no explicit inverse);
 *   - Y = C inv(Phi) C' dense, unblocked Cholesky;
 *   - the trial residual of the line search is evaluated literally, C'(nu + t dnu) included.
 * PARITY UNPINNED by the reference (MATLAB, no golden vectors): pinned against oracle/fastmpc_dense.py in
 * tests/test_oracle_fmpc_general.py.  Column-major doubles throughout.  Build: oracle/Makefile.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    int n, m, T, var_order, ramp_rows, literal_bug;
    const double *A1, *A2, *B, *Q, *R, *Qf, *q, *r, *qf, *u_min, *u_max, *du_min, *du_max;
} frefg_sys;

#define AT(M, ld, i, j) ((M)[(size_t)(j) * (ld) + (i)])
enum { ST_OK = 0, ST_EARLY_EXIT = 1, ST_NOT_PD = 2, ST_LS_MAX = 3, ST_NONFINITE = 4 };

static int chol(int n, double *A)
{
    for (int j = 0; j < n; ++j) {
        double d = AT(A, n, j, j);
        for (int k = 0; k < j; ++k) d -= AT(A, n, j, k) * AT(A, n, j, k);
        if (!(d > 0.0)) return j + 1;
        d = sqrt(d);
        AT(A, n, j, j) = d;
        for (int i = j + 1; i < n; ++i) {
            double s = AT(A, n, i, j);
            for (int k = 0; k < j; ++k) s -= AT(A, n, i, k) * AT(A, n, j, k);
            AT(A, n, i, j) = s / d;
        }
    }
    return 0;
}
static void chol_solve(int n, const double *L, double *x)
{
    for (int j = 0; j < n; ++j) { x[j] /= AT(L, n, j, j); for (int i = j + 1; i < n; ++i) x[i] -= AT(L, n, i, j) * x[j]; }
    for (int j = n - 1; j >= 0; --j) { double s = x[j]; for (int i = j + 1; i < n; ++i) s -= AT(L, n, i, j) * x[i]; x[j] = s / AT(L, n, j, j); }
}

typedef struct {
    int n, m, T, NB, N, NE, ramp;
    double *C;              /* NE x N */
    int *lo, *hi;           /* per row: nonzero column range [lo, hi) */
    double *LQ, *LQf;       /* Cholesky factors of 2Q, 2Qf */
    double *Q2, *Q2f;       /* 2Q, 2Qf (symmetrised) */
} pre_t;

static int pre_build(const frefg_sys *S, int has_xf, pre_t *P)
{
    const int n = S->n, m = S->m, T = S->T, st = n + m;
    P->n = n; P->m = m; P->T = T; P->NB = T + (has_xf ? 1 : 0); P->N = T * st; P->NE = P->NB * n; P->ramp = S->ramp_rows;
    const int N = P->N, NE = P->NE;
    P->C = calloc((size_t)NE * N, 8); P->lo = malloc(4 * NE); P->hi = malloc(4 * NE);
    P->LQ = malloc(8 * (size_t)n * n); P->LQf = malloc(8 * (size_t)n * n); P->Q2 = malloc(8 * (size_t)n * n); P->Q2f = malloc(8 * (size_t)n * n);
    double *C = P->C;
#define PUT(row0, col0, M, w, sgn) for (int r_ = 0; r_ < n; ++r_) for (int c_ = 0; c_ < (w); ++c_) AT(C, NE, (row0) + r_, (col0) + c_) = (sgn) * AT(M, n, r_, c_)
#define PUTI(row0, col0) for (int r_ = 0; r_ < n; ++r_) for (int c_ = 0; c_ < n; ++c_) AT(C, NE, (row0) + r_, (col0) + c_) = (r_ == c_) ? 1.0 : 0.0
    PUT(0, 0, S->B, m, -1.0); PUTI(0, m);                                   /* C(1:n,1:m+n) = [-B I] */
    for (int i = 1; i < T; ++i) {
        if (S->var_order == 2) {                                            /* VAR_2/fast_mpc_eq_const.m:41-49 */
            if (i == 1) { PUT(n, m, S->A1, n, -1.0); PUT(n, m + n, S->B, m, -1.0); PUTI(n, m + n + m); }
            else { int c0 = m + st * (i - 2); PUT(n * i, c0, S->A2, n, -1.0); PUT(n * i, c0 + n + m, S->A1, n, -1.0); PUT(n * i, c0 + 2 * n + m, S->B, m, -1.0); PUTI(n * i, c0 + 2 * n + 2 * m); }
        } else {                                                            /* VAR_1/fast_mpc_eq_const.m:34-43 */
            int c0 = (i - 1) * st + m;
            if (i == 1 && S->literal_bug) { c0 = n - 1; if (c0 + 2 * n + m > N || T < 3) return -3; }
            PUT(n * i, c0, S->A1, n, -1.0); PUT(n * i, c0 + n, S->B, m, -1.0); PUTI(n * i, c0 + n + m);
        }
    }
    if (has_xf) for (int k = 0; k < n; ++k) AT(C, NE, T * n + k, N - n + k) = 1.0;
    for (int r = 0; r < NE; ++r) {
        int lo = N, hi = 0;
        for (int c = 0; c < N; ++c) if (AT(C, NE, r, c) != 0.0) { if (c < lo) lo = c; hi = c + 1; }
        if (hi <= lo) { lo = 0; hi = 0; }
        P->lo[r] = lo; P->hi[r] = hi;
    }
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) {
        AT(P->Q2, n, i, j) = AT(S->Q, n, i, j) + AT(S->Q, n, j, i);
        AT(P->Q2f, n, i, j) = AT(S->Qf, n, i, j) + AT(S->Qf, n, j, i);
    }
    memcpy(P->LQ, P->Q2, 8 * (size_t)n * n); memcpy(P->LQf, P->Q2f, 8 * (size_t)n * n);
    if (chol(n, P->LQ) || chol(n, P->LQf)) return -1;
    return 0;
}
static void pre_free(pre_t *P) { free(P->C); free(P->lo); free(P->hi); free(P->LQ); free(P->LQf); free(P->Q2); free(P->Q2f); }

/* out = C v */
static void mulC(const pre_t *P, const double *v, double *out)
{
    for (int r = 0; r < P->NE; ++r) { double s = 0.0; for (int c = P->lo[r]; c < P->hi[r]; ++c) s += AT(P->C, P->NE, r, c) * v[c]; out[r] = s; }
}
/* out = C' v */
static void mulCt(const pre_t *P, const double *v, double *out)
{
    memset(out, 0, 8 * (size_t)P->N);
    for (int r = 0; r < P->NE; ++r) for (int c = P->lo[r]; c < P->hi[r]; ++c) out[c] += AT(P->C, P->NE, r, c) * v[r];
}
/* v <- inv(Phi) v : x blocks by Cholesky of 2Q / 2Qf, u part by one Thomas solve per actuator */
static void solve_phi(const pre_t *P, const double *tdiag, const double *toff, double *v, double *work /* 2 T */)
{
    const int n = P->n, m = P->m, T = P->T, st = n + m;
    for (int t = 0; t < T; ++t) chol_solve(n, t == T - 1 ? P->LQf : P->LQ, v + (size_t)t * st + m);
    double *cp = work, *dp = work + T;
    for (int j = 0; j < m; ++j) {
        if (!P->ramp) { for (int t = 0; t < T; ++t) v[(size_t)t * st + j] /= tdiag[(size_t)t * m + j]; continue; }
        /* tridiagonal: a_t x_t + o_{t-1} x_{t-1} + o_t x_{t+1} = v_t,  o_t = toff[t] couples t and t + 1 */
        double a0 = tdiag[j];
        cp[0] = (T > 1 ? toff[j] : 0.0) / a0; dp[0] = v[j] / a0;
        for (int t = 1; t < T; ++t) {
            const double o = toff[(size_t)(t - 1) * m + j];
            const double den = tdiag[(size_t)t * m + j] - o * cp[t - 1];
            cp[t] = (t + 1 < T ? toff[(size_t)t * m + j] : 0.0) / den;
            dp[t] = (v[(size_t)t * st + j] - o * dp[t - 1]) / den;
        }
        v[(size_t)(T - 1) * st + j] = dp[T - 1];
        for (int t = T - 2; t >= 0; --t) { dp[t] -= cp[t] * dp[t + 1]; v[(size_t)t * st + j] = dp[t]; }
    }
}
/* out = 2 H z + g  (fast_mpc_objective.m:50-65: H = blkdiag(R, [Q,R]..., Qf), cost z'Hz + g'z) */
static void grad_obj(const frefg_sys *S, const pre_t *P, const double *z, double *out)
{
    const int n = P->n, m = P->m, T = P->T, st = n + m;
    for (int t = 0; t < T; ++t) {
        for (int j = 0; j < m; ++j) out[(size_t)t * st + j] = 2.0 * AT(S->R, m, j, j) * z[(size_t)t * st + j] + (S->r ? S->r[j] : 0.0);
        const double *Q2 = (t == T - 1) ? P->Q2f : P->Q2, *lin = (t == T - 1) ? S->qf : S->q;
        const double *x = z + (size_t)t * st + m;
        for (int k = 0; k < n; ++k) { double s = 0.0; for (int kk = 0; kk < n; ++kk) s += AT(Q2, n, k, kk) * x[kk]; out[(size_t)t * st + m + k] = s + (lin ? lin[k] : 0.0); }
    }
}
static double nrm2sq(int k, const double *a) { double s = 0.0; for (int i = 0; i < k; ++i) s += a[i] * a[i]; return s; }

static int solve_one(const frefg_sys *S, const pre_t *P, double kappa, int niters, int ls_max, double alpha, double bt, double tol_r,
                     double tol_p, const double *x0, const double *x0_pre, const double *u_prev, const double *w, const double *xf,
                     const double *z0, const double *nu0, double *z, int *iters_out, int *halv_out)
{
    const int n = P->n, m = P->m, T = P->T, st = n + m, N = P->N, NE = P->NE;
    double *nu = malloc(8 * NE), *b = calloc(NE, 8), *pd = calloc(N, 8), *rd = malloc(8 * N), *rp = malloc(8 * NE), *tmpN = malloc(8 * N);
    double *tdiag = malloc(8 * (size_t)T * m), *toff = calloc((size_t)T * m, 8), *W = malloc(8 * (size_t)N * NE), *Y = malloc(8 * (size_t)NE * NE);
    double *rhs = malloc(8 * NE), *dnu = malloc(8 * NE), *dz = malloc(8 * N), *zt = malloc(8 * N), *nut = malloc(8 * NE), *work = malloc(8 * 2 * (size_t)T);
    double *rdt = malloc(8 * N), *rpt = malloc(8 * NE);
    int status = ST_OK, iters = 0, halv = 0;
    memcpy(z, z0, 8 * N); memcpy(nu, nu0, 8 * NE);
    if (w) memcpy(b, w, 8 * (size_t)T * n);
    for (int k = 0; k < n; ++k) { double s = 0.0; for (int kk = 0; kk < n; ++kk) s += AT(S->A1, n, k, kk) * x0[kk]; b[k] += s; }
    if (S->var_order == 2) {
        for (int k = 0; k < n; ++k) { double s = 0.0; for (int kk = 0; kk < n; ++kk) s += AT(S->A2, n, k, kk) * x0_pre[kk]; b[k] += s; }
        if (T > 1) for (int k = 0; k < n; ++k) { double s = 0.0; for (int kk = 0; kk < n; ++kk) s += AT(S->A2, n, k, kk) * x0[kk]; b[n + k] += s; }
    }
    if (P->NB > T) memcpy(b + (size_t)T * n, xf, 8 * n);
    for (int it = 0; it < niters; ++it) {
        /* s = h - P z, d = 1./s, Phi_uu = 2R + k P'diag(d.^2)P  (inf_newton_KKT_H.m:3-13; rows: fast_mpc_ineq_const.m:42-79) */
        for (int t = 0; t < T; ++t)
            for (int j = 0; j < m; ++j) {
                const double u = z[(size_t)t * st + j];
                const double dp = 1.0 / (S->u_max[j] - u), dm = 1.0 / (-S->u_min[j] + u);
                double g = dp - dm, dd = dp * dp + dm * dm, off = 0.0;
                if (P->ramp) {
                    double su, sl;
                    if (t == 0) { su = (u_prev[j] + S->du_max[j]) - u; sl = (-u_prev[j] - S->du_min[j]) + u; }
                    else { const double pz = u - z[(size_t)(t - 1) * st + j]; su = S->du_max[j] - pz; sl = -S->du_min[j] + pz; }
                    const double ru = 1.0 / su, rl = 1.0 / sl;
                    g += ru - rl; dd += ru * ru + rl * rl;
                    if (t + 1 < T) {
                        const double pzn = z[(size_t)(t + 1) * st + j] - u;
                        const double run = 1.0 / (S->du_max[j] - pzn), rln = 1.0 / (-S->du_min[j] + pzn);
                        g -= run - rln; dd += run * run + rln * rln; off = -kappa * (run * run + rln * rln);
                    }
                }
                pd[(size_t)t * st + j] = kappa * g;
                tdiag[(size_t)t * m + j] = 2.0 * AT(S->R, m, j, j) + kappa * dd;
                toff[(size_t)t * m + j] = off;
            }
        /* residuals (inf_newton_solver.m:12-22) */
        grad_obj(S, P, z, rd); mulCt(P, nu, tmpN);
        for (int i = 0; i < N; ++i) rd[i] += pd[i] + tmpN[i];
        mulC(P, z, rp); for (int i = 0; i < NE; ++i) rp[i] -= b[i];
        const double ssp = nrm2sq(NE, rp), nr0 = sqrt(nrm2sq(N, rd) + ssp);
        if (!isfinite(nr0)) { status = ST_NONFINITE; break; }
        if (nr0 <= tol_r && sqrt(ssp) <= tol_p) { status = ST_EARLY_EXIT; break; }
        /* W = inv(Phi) C' column by column ; Y = C W  (:24-27) */
        int bad = 0;
        for (int j = 0; j < T * m; ++j) if (!(tdiag[j] > 0.0)) bad = 1;
        if (bad) { status = ST_NOT_PD; break; }
        for (int r = 0; r < NE; ++r) {
            double *col = W + (size_t)r * N;
            memset(col, 0, 8 * N);
            for (int c = P->lo[r]; c < P->hi[r]; ++c) col[c] = AT(P->C, NE, r, c);
            solve_phi(P, tdiag, toff, col, work);
        }
        for (int r2 = 0; r2 < NE; ++r2) { mulC(P, W + (size_t)r2 * N, tmpN); for (int r = 0; r < NE; ++r) AT(Y, NE, r, r2) = tmpN[r]; }
        /* beta = -r_p + C inv(Phi) r_d ; dnu = -inv(Y) beta ; dz = inv(Phi)(-r_d - C'dnu)  (:28-35) */
        memcpy(tmpN, rd, 8 * N); solve_phi(P, tdiag, toff, tmpN, work); mulC(P, tmpN, rhs);
        for (int i = 0; i < NE; ++i) rhs[i] = rp[i] - rhs[i];
        if (chol(NE, Y)) { status = ST_NOT_PD; break; }
        memcpy(dnu, rhs, 8 * NE); chol_solve(NE, Y, dnu);
        mulCt(P, dnu, dz);
        for (int i = 0; i < N; ++i) dz[i] = -(rd[i] + dz[i]);
        solve_phi(P, tdiag, toff, dz, work);
        /* backtracking on ||[r_p; r_d]||, d frozen (backtracking_inf_newton.m:2-11) */
        double t = 1.0; int nh = 0;
        for (;;) {
            for (int i = 0; i < N; ++i) zt[i] = z[i] + t * dz[i];
            for (int i = 0; i < NE; ++i) nut[i] = nu[i] + t * dnu[i];
            grad_obj(S, P, zt, rdt); mulCt(P, nut, tmpN);
            for (int i = 0; i < N; ++i) rdt[i] += pd[i] + tmpN[i];
            mulC(P, zt, rpt); for (int i = 0; i < NE; ++i) rpt[i] -= b[i];
            const double nrt = sqrt(nrm2sq(N, rdt) + nrm2sq(NE, rpt));
            if (!(nrt > (1.0 - alpha * t) * nr0)) break;
            if (t == 0.0) break;
            if (ls_max > 0 && nh >= ls_max) { status = ST_LS_MAX; break; }
            t *= bt; ++nh;
        }
        halv += nh;
        memcpy(z, zt, 8 * N); memcpy(nu, nut, 8 * NE);
        ++iters;
    }
    *iters_out = iters; *halv_out = halv;
    free(nu); free(b); free(pd); free(rd); free(rp); free(tmpN); free(tdiag); free(toff); free(W); free(Y); free(rhs); free(dnu); free(dz);
    free(zt); free(nut); free(work); free(rdt); free(rpt);
    return status;
}

/* Batched solve, one column per instance: x0, x0_pre (n), u_prev (m), w (T n), xf (n), z0 (N), nu0 (NE); z_out (N).
 * Returns 0, -1 (Q / Qf not PD), -2 (bad arguments), -3 (a literal VAR_1 C that MATLAB itself would reject). */
int frefg_solve_batch(const frefg_sys *S, double kappa, int niters, int ls_max, double alpha, double beta, double tol_r, double tol_p,
                      int nbatch, const double *x0, const double *x0_pre, const double *u_prev, const double *w, const double *xf,
                      const double *z0, const double *nu0, double *z_out, int *status, int *iters, int *halvings, int nthreads)
{
    if (!S || !S->A1 || !S->B || !S->Q || !S->R || !S->Qf || !S->u_min || !S->u_max || !x0 || !z0 || !nu0 || !z_out) return -2;
    if (S->var_order == 2 && (!S->A2 || !x0_pre)) return -2;
    if (S->ramp_rows && (!S->du_min || !S->du_max || !u_prev)) return -2;
    pre_t P;
    const int rc = pre_build(S, xf != NULL, &P);
    if (rc) { pre_free(&P); return rc; }
    const int n = P.n, m = P.m, T = P.T, N = P.N, NE = P.NE;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
    for (int b = 0; b < nbatch; ++b) {
        int it = 0, hv = 0;
        const int st = solve_one(S, &P, kappa, niters, ls_max, alpha, beta, tol_r, tol_p, x0 + (size_t)b * n,
                                 x0_pre ? x0_pre + (size_t)b * n : NULL, u_prev ? u_prev + (size_t)b * m : NULL,
                                 w ? w + (size_t)b * T * n : NULL, xf ? xf + (size_t)b * n : NULL, z0 + (size_t)b * N,
                                 nu0 + (size_t)b * NE, z_out + (size_t)b * N, &it, &hv);
        if (status) status[b] = st;
        if (iters) iters[b] = it;
        if (halvings) halvings[b] = hv;
    }
    pre_free(&P);
    return 0;
}
