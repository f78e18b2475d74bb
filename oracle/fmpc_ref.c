/* ORACLE (test infrastructure, NOT product code).
 *
 * Structured (block-banded) CPU restatement of the reference's fastMPC Newton solver,
 * plain C + OpenMP over instances.  It computes what
 *   Fast_MPC/VAR_2/Fast_MPC2.m:124-130  (mpc_fixed_log_newton)
 *     -> fast_mpc_init.m:12-26, fast_mpc_objective.m:50-65, fast_mpc_ineq_const.m:42-56,
 *        fast_mpc_eq_const.m:38-71, inf_newton_KKT_H.m:3-13, inf_newton_solver.m:1-43,
 *        backtracking_inf_newton.m:2-11
 * computes, but never forms the dense H, P, C: Phi is block diagonal (box rows on u only),
 * the Schur complement C*inv(Phi)*C' is block penta-diagonal in n x n blocks
 * and is factored by a band-2 block Cholesky.  VAR(1) = the same code with A2 == NULL
 * (the CORRECTED structure of VAR_1/fast_mpc_eq_const.m, i.e. without the column bug F9,
 * and without VAR_1's ramp rows).
 *
 * PARITY UNPINNED: the reference is MATLAB with no golden vectors; this
 * file is pinned against oracle/fastmpc_dense.py (the literal dense restatement) in
 * tests/test_oracle_*.py.  It is also bench.py's CPU baseline (cpu_baseline.kind="port").
 *
 * All matrices column-major double, exactly as MATLAB hands them over.
 * Build: see oracle/Makefile  (gcc -O3 -fopenmp -shared -fPIC).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    int n, m, T;
    const double *A1, *A2;          /* n x n ; A2 may be NULL (VAR(1)) */
    const double *B;                /* n x m */
    const double *Q, *R, *Qf;       /* n x n, m x m, n x n (dense, symmetric PD) */
    const double *q, *r, *qf;       /* optional (NULL = zeros) */
    const double *u_min, *u_max;    /* m */
} fref_sys;

/* ------------------------------------------------------------------ small dense helpers */
#define AT(M, ld, i, j) ((M)[(size_t)(j) * (ld) + (i)])

static int potrf_lower(int n, double *A)
{   /* in place, lower triangle; upper triangle left untouched */
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
/* X := X * inv(L)'   (X is r x n, L lower n x n) */
static void trsm_right_lt(int r, int n, const double *L, double *X)
{
    for (int j = 0; j < n; ++j) {
        for (int k = 0; k < j; ++k) {
            double l = AT(L, n, j, k);
            for (int i = 0; i < r; ++i) AT(X, r, i, j) -= AT(X, r, i, k) * l;
        }
        double d = 1.0 / AT(L, n, j, j);
        for (int i = 0; i < r; ++i) AT(X, r, i, j) *= d;
    }
}
static void trsv_lower(int n, const double *L, double *x)
{
    for (int j = 0; j < n; ++j) {
        x[j] /= AT(L, n, j, j);
        double xj = x[j];
        for (int i = j + 1; i < n; ++i) x[i] -= AT(L, n, i, j) * xj;
    }
}
static void trsv_lower_t(int n, const double *L, double *x)
{
    for (int j = n - 1; j >= 0; --j) {
        double s = x[j];
        for (int i = j + 1; i < n; ++i) s -= AT(L, n, i, j) * x[i];
        x[j] = s / AT(L, n, j, j);
    }
}
/* C(r x c) += alpha * A(r x k) * B(c x k)' */
static void gemm_nt(int r, int c, int k, double alpha, const double *A, const double *Bm, double *C)
{
    for (int j = 0; j < c; ++j)
        for (int l = 0; l < k; ++l) {
            double b = alpha * AT(Bm, c, j, l);
            if (b == 0.0) continue;
            for (int i = 0; i < r; ++i) AT(C, r, i, j) += AT(A, r, i, l) * b;
        }
}
/* C(r x c) += alpha * A(r x k) * B(k x c) */
static void gemm_nn(int r, int c, int k, double alpha, const double *A, const double *Bm, double *C)
{
    for (int j = 0; j < c; ++j)
        for (int l = 0; l < k; ++l) {
            double b = alpha * AT(Bm, k, l, j);
            if (b == 0.0) continue;
            for (int i = 0; i < r; ++i) AT(C, r, i, j) += AT(A, r, i, l) * b;
        }
}
/* y += alpha * A(r x c) * x */
static void gemv_n(int r, int c, double alpha, const double *A, const double *x, double *y)
{
    for (int j = 0; j < c; ++j) {
        double xj = alpha * x[j];
        for (int i = 0; i < r; ++i) y[i] += AT(A, r, i, j) * xj;
    }
}
/* y += alpha * A(r x c)' * x   (y has c entries) */
static void gemv_t(int r, int c, double alpha, const double *A, const double *x, double *y)
{
    for (int j = 0; j < c; ++j) {
        double s = 0.0;
        for (int i = 0; i < r; ++i) s += AT(A, r, i, j) * x[i];
        y[j] += alpha * s;
    }
}
static int spd_inverse(int n, const double *A, double scale, double *Ainv, double *work)
{   /* Ainv = inv(scale*A) via Cholesky; work n*n */
    for (int i = 0; i < n * n; ++i) work[i] = scale * A[i];
    int info = potrf_lower(n, work);
    if (info) return info;
    memset(Ainv, 0, sizeof(double) * n * n);
    for (int j = 0; j < n; ++j) {
        double *col = Ainv + (size_t)j * n;
        col[j] = 1.0;
        trsv_lower(n, work, col);
        trsv_lower_t(n, work, col);
    }
    return 0;
}
static int is_diagonal(int n, const double *A)
{
    for (int j = 0; j < n; ++j)
        for (int i = 0; i < n; ++i)
            if (i != j && AT(A, n, i, j) != 0.0) return 0;
    return 1;
}

/* ------------------------------------------------------------------ shared precompute */
typedef struct {
    int n, m, T, NB, has_xf, r_diag;
    double *Qi, *Qfi;       /* inv(2Q), inv(2Qf)  n x n */
    double *Yd;             /* NB blocks: constant part of Y[i,i] */
    double *Y1;             /* NB blocks: Y[i+1,i] */
    double *Y2;             /* NB blocks: Y[i+2,i] */
} fref_pre;

static const double *Qi_of(const fref_pre *P, int j) { return j == P->T ? P->Qfi : P->Qi; }

static int precompute(const fref_sys *S, int has_xf, fref_pre *P)
{
    int n = S->n, T = S->T, nn = n * n;
    P->n = n; P->m = S->m; P->T = T; P->has_xf = has_xf; P->NB = T + (has_xf ? 1 : 0);
    P->r_diag = is_diagonal(S->m, S->R);
    P->Qi = calloc(nn, 8); P->Qfi = calloc(nn, 8);
    P->Yd = calloc((size_t)P->NB * nn, 8); P->Y1 = calloc((size_t)P->NB * nn, 8); P->Y2 = calloc((size_t)P->NB * nn, 8);
    double *w1 = calloc(nn, 8), *w2 = calloc(nn, 8);
    int info = spd_inverse(n, S->Q, 2.0, P->Qi, w1);
    if (!info) info = spd_inverse(n, S->Qf, 2.0, P->Qfi, w1);
    if (info) { free(w1); free(w2); return -1; }
    for (int i = 0; i < P->NB; ++i) {
        double *Yd = P->Yd + (size_t)i * nn, *Y1 = P->Y1 + (size_t)i * nn, *Y2 = P->Y2 + (size_t)i * nn;
        if (i == T) { memcpy(Yd, Qi_of(P, T), 8 * nn); continue; }           /* terminal row: [0 .. I] z = xf */
        memcpy(Yd, Qi_of(P, i + 1), 8 * nn);                                 /* I * Qi_{i+1} * I' */
        if (i >= 1) { memset(w1, 0, 8 * nn); gemm_nn(n, n, n, 1.0, S->A1, Qi_of(P, i), w1); gemm_nt(n, n, n, 1.0, w1, S->A1, Yd); }
        if (i >= 2 && S->A2) { memset(w1, 0, 8 * nn); gemm_nn(n, n, n, 1.0, S->A2, Qi_of(P, i - 1), w1); gemm_nt(n, n, n, 1.0, w1, S->A2, Yd); }
        if (i + 1 < T) {          /* Y[i+1,i] = -A1*Qi_{i+1} + A2*Qi_i*A1' */
            gemm_nn(n, n, n, -1.0, S->A1, Qi_of(P, i + 1), Y1);
            if (i >= 1 && S->A2) { memset(w1, 0, 8 * nn); gemm_nn(n, n, n, 1.0, S->A2, Qi_of(P, i), w1); gemm_nt(n, n, n, 1.0, w1, S->A1, Y1); }
        } else if (i + 1 == T && has_xf) memcpy(Y1, Qi_of(P, T), 8 * nn);    /* Y[T,T-1] = Qi_T */
        if (i + 2 < T && S->A2) gemm_nn(n, n, n, -1.0, S->A2, Qi_of(P, i + 1), Y2);
    }
    free(w1); free(w2);
    return 0;
}
static void pre_free(fref_pre *P) { free(P->Qi); free(P->Qfi); free(P->Yd); free(P->Y1); free(P->Y2); }

/* ------------------------------------------------------------------ per-instance workspace */
typedef struct {
    double *z, *nu, *dz, *dnu, *zt, *nut;
    double *dbar, *phi;              /* T*m: kappa*P'd on u, kappa*P'DP diagonal on u */
    double *rd, *rp, *rdt, *rpt;     /* N, NB*n */
    double *p;                       /* N: inv(Phi)*r_d */
    double *beta, *y;                /* NB*n */
    double *Lf, *L1, *L2;            /* NB blocks each */
    double *Rt;                      /* T blocks m x m (dense-R path: chol of Rtilde_t) */
    double *W;                       /* m x n scratch */
    double *b;                       /* NB*n */
} fref_ws;

static void ws_alloc(fref_ws *W, const fref_pre *P, int dense_r)
{
    int n = P->n, m = P->m, T = P->T, NB = P->NB;
    size_t N = (size_t)T * (n + m), nb = (size_t)NB * n, nn = (size_t)n * n;
    W->z = calloc(N, 8); W->dz = calloc(N, 8); W->zt = calloc(N, 8);
    W->nu = calloc(nb, 8); W->dnu = calloc(nb, 8); W->nut = calloc(nb, 8);
    W->dbar = calloc((size_t)T * m, 8); W->phi = calloc((size_t)T * m, 8);
    W->rd = calloc(N, 8); W->rdt = calloc(N, 8); W->p = calloc(N, 8);
    W->rp = calloc(nb, 8); W->rpt = calloc(nb, 8); W->beta = calloc(nb, 8); W->y = calloc(nb, 8); W->b = calloc(nb, 8);
    W->Lf = calloc(NB * nn, 8); W->L1 = calloc(NB * nn, 8); W->L2 = calloc(NB * nn, 8);
    W->Rt = dense_r ? calloc((size_t)T * m * m, 8) : NULL;
    W->W = calloc((size_t)m * n, 8);
}
static void ws_free(fref_ws *W)
{
    free(W->z); free(W->dz); free(W->zt); free(W->nu); free(W->dnu); free(W->nut); free(W->dbar); free(W->phi);
    free(W->rd); free(W->rdt); free(W->p); free(W->rp); free(W->rpt); free(W->beta); free(W->y); free(W->b);
    free(W->Lf); free(W->L1); free(W->L2); free(W->Rt); free(W->W);
}

/* z layout (fast_mpc_init.m:22-25): stage t holds [u_t (m); x_{t+1} (n)] */
#define UOF(v, t) ((v) + (size_t)(t) * (n + m))
#define XOF(v, j) ((v) + (size_t)((j) - 1) * (n + m) + m)      /* x_j, j = 1..T */

/* out = C z   (without the -b), NB*n.  Rows follow VAR_2/fast_mpc_eq_const.m:38-49,67-71 */
static void apply_C(const fref_sys *S, const fref_pre *P, const double *z, double *out)
{
    int n = P->n, m = P->m, T = P->T;
    for (int i = 0; i < T; ++i) {
        double *o = out + (size_t)i * n;
        memcpy(o, XOF(z, i + 1), 8 * n);
        gemv_n(n, m, -1.0, S->B, UOF(z, i), o);
        if (i >= 1) gemv_n(n, n, -1.0, S->A1, XOF(z, i), o);
        if (i >= 2 && S->A2) gemv_n(n, n, -1.0, S->A2, XOF(z, i - 1), o);
    }
    if (P->has_xf) memcpy(out + (size_t)T * n, XOF(z, T), 8 * n);
}
/* out = C' v, N */
static void apply_Ct(const fref_sys *S, const fref_pre *P, const double *v, double *out)
{
    int n = P->n, m = P->m, T = P->T;
    for (int t = 0; t < T; ++t) {
        double *ou = UOF(out, t);
        memset(ou, 0, 8 * m);
        gemv_t(n, m, -1.0, S->B, v + (size_t)t * n, ou);
        int j = t + 1;                                   /* x_j */
        double *ox = XOF(out, j);
        memcpy(ox, v + (size_t)(j - 1) * n, 8 * n);
        if (j <= T - 1) gemv_t(n, n, -1.0, S->A1, v + (size_t)j * n, ox);
        if (j <= T - 2 && S->A2) gemv_t(n, n, -1.0, S->A2, v + (size_t)(j + 1) * n, ox);
        if (j == T && P->has_xf) for (int k = 0; k < n; ++k) ox[k] += v[(size_t)T * n + k];
    }
}
/* out = 2Hz + g  (fast_mpc_objective.m:50-65, factor 2 from inf_newton_solver.m:12) */
static void apply_2H_g(const fref_sys *S, const fref_pre *P, const double *z, double *out)
{
    int n = P->n, m = P->m, T = P->T;
    for (int t = 0; t < T; ++t) {
        double *ou = UOF(out, t);
        for (int k = 0; k < m; ++k) ou[k] = S->r ? S->r[k] : 0.0;
        gemv_n(m, m, 2.0, S->R, UOF(z, t), ou);
        int j = t + 1;
        double *ox = XOF(out, j);
        const double *ql = (j == T) ? S->qf : S->q;
        for (int k = 0; k < n; ++k) ox[k] = ql ? ql[k] : 0.0;
        gemv_n(n, n, 2.0, (j == T) ? S->Qf : S->Q, XOF(z, j), ox);
    }
}
static double norm2sq(size_t k, const double *a) { double s = 0; for (size_t i = 0; i < k; ++i) s += a[i] * a[i]; return s; }

/* status codes (shared with include/fmpc.h) */
enum { ST_OK = 0, ST_EARLY_EXIT = 1, ST_NOT_PD = 2, ST_LS_MAX = 3, ST_NONFINITE = 4 };

static int solve_one(const fref_sys *S, const fref_pre *P, fref_ws *W, double kappa, int niters, int ls_max,
                     double alpha, double bt, double tol_r, double tol_p,
                     const double *x0, const double *x0_pre, const double *w, const double *xf,
                     const double *z0, const double *nu0, double *z_out, double *nu_out, int *iters_out, int *halv_out)
{
    int n = P->n, m = P->m, T = P->T, NB = P->NB;
    size_t N = (size_t)T * (n + m), nb = (size_t)NB * n, nn = (size_t)n * n;
    int status = ST_OK, iters = 0, halv = 0;
    memcpy(W->z, z0, 8 * N);
    memcpy(W->nu, nu0, 8 * nb);
    /* b : VAR_2/fast_mpc_eq_const.m:39,44,47,68 */
    memset(W->b, 0, 8 * nb);
    if (w) memcpy(W->b, w, 8 * (size_t)T * n);
    gemv_n(n, n, 1.0, S->A1, x0, W->b);
    if (S->A2) { gemv_n(n, n, 1.0, S->A2, x0_pre, W->b); if (T > 1) gemv_n(n, n, 1.0, S->A2, x0, W->b + n); }
    if (P->has_xf) memcpy(W->b + (size_t)T * n, xf, 8 * n);

    for (int it = 0; it < niters; ++it) {
        /* barrier terms, inf_newton_KKT_H.m:3-13 (rows [I;-I] on u_t, h=[u_max;-u_min]) */
        for (int t = 0; t < T; ++t)
            for (int k = 0; k < m; ++k) {
                double u = UOF(W->z, t)[k];
                double sp = S->u_max[k] - u, sm = -S->u_min[k] + u;
                double dp = 1.0 / sp, dm = 1.0 / sm;
                W->dbar[(size_t)t * m + k] = kappa * (dp - dm);
                W->phi[(size_t)t * m + k] = kappa * (dp * dp + dm * dm);
            }
        /* r_d = 2Hz + g + k P'd + C'nu ; r_p = Cz - b   (inf_newton_solver.m:12-13) */
        apply_2H_g(S, P, W->z, W->rd);
        apply_Ct(S, P, W->nu, W->p);
        for (size_t i = 0; i < N; ++i) W->rd[i] += W->p[i];
        for (int t = 0; t < T; ++t) for (int k = 0; k < m; ++k) UOF(W->rd, t)[k] += W->dbar[(size_t)t * m + k];
        apply_C(S, P, W->z, W->rp);
        for (size_t i = 0; i < nb; ++i) W->rp[i] -= W->b[i];
        double nrp2 = norm2sq(nb, W->rp), nr0 = sqrt(norm2sq(N, W->rd) + nrp2);
        if (!isfinite(nr0)) { status = ST_NONFINITE; break; }
        if (nr0 <= tol_r && sqrt(nrp2) <= tol_p) { status = ST_EARLY_EXIT; break; }      /* :19-22 */

        /* p = inv(Phi) r_d ;  D_t = B inv(Rtilde_t) B' */
        for (int t = 0; t < T; ++t) {
            double *Lf = W->Lf + (size_t)t * nn;
            memcpy(Lf, P->Yd + (size_t)t * nn, 8 * nn);
            if (P->r_diag) {
                for (int k = 0; k < m; ++k) {
                    double ph = 1.0 / (2.0 * AT(S->R, m, k, k) + W->phi[(size_t)t * m + k]);
                    UOF(W->p, t)[k] = UOF(W->rd, t)[k] * ph;
                    for (int c = 0; c < n; ++c) {
                        double bc = AT(S->B, n, c, k) * ph;
                        for (int r2 = 0; r2 < n; ++r2) AT(Lf, n, r2, c) += AT(S->B, n, r2, k) * bc;
                    }
                }
            } else {
                double *Rt = W->Rt + (size_t)t * m * m;
                for (size_t i = 0; i < (size_t)m * m; ++i) Rt[i] = 2.0 * S->R[i];
                for (int k = 0; k < m; ++k) AT(Rt, m, k, k) += W->phi[(size_t)t * m + k];
                if (potrf_lower(m, Rt)) { status = ST_NOT_PD; goto done; }
                memcpy(UOF(W->p, t), UOF(W->rd, t), 8 * m);
                trsv_lower(m, Rt, UOF(W->p, t)); trsv_lower_t(m, Rt, UOF(W->p, t));
                for (int c = 0; c < n; ++c) {               /* W = inv(Lr) B'  (m x n) */
                    double *col = W->W + (size_t)c * m;
                    for (int k = 0; k < m; ++k) col[k] = AT(S->B, n, c, k);
                    trsv_lower(m, Rt, col);
                }
                for (int c = 0; c < n; ++c) for (int r2 = 0; r2 < n; ++r2) {
                    double s = 0; for (int k = 0; k < m; ++k) s += W->W[(size_t)r2 * m + k] * W->W[(size_t)c * m + k];
                    AT(Lf, n, r2, c) += s;
                }
            }
            int j = t + 1;
            memset(XOF(W->p, j), 0, 8 * n);
            gemv_n(n, n, 1.0, Qi_of(P, j), XOF(W->rd, j), XOF(W->p, j));
        }
        if (P->has_xf) memcpy(W->Lf + (size_t)T * nn, P->Yd + (size_t)T * nn, 8 * nn);
        /* beta = -r_p + C inv(Phi) r_d  (:28-29) */
        apply_C(S, P, W->p, W->beta);
        for (size_t i = 0; i < nb; ++i) W->beta[i] -= W->rp[i];

        /* band-2 block Cholesky of Y (:30) fused with the forward solve of  Y dnu = -beta (:31) */
        for (int i = 0; i < NB; ++i) {
            double *Lf = W->Lf + (size_t)i * nn, *L1 = W->L1 + (size_t)i * nn, *L2 = W->L2 + (size_t)i * nn;
            if (i >= 1) gemm_nt(n, n, n, -1.0, W->L1 + (size_t)(i - 1) * nn, W->L1 + (size_t)(i - 1) * nn, Lf);
            if (i >= 2) gemm_nt(n, n, n, -1.0, W->L2 + (size_t)(i - 2) * nn, W->L2 + (size_t)(i - 2) * nn, Lf);
            if (potrf_lower(n, Lf)) { status = ST_NOT_PD; goto done; }
            if (i + 1 < NB) {
                memcpy(L1, P->Y1 + (size_t)i * nn, 8 * nn);
                if (i >= 1) gemm_nt(n, n, n, -1.0, W->L2 + (size_t)(i - 1) * nn, W->L1 + (size_t)(i - 1) * nn, L1);
                trsm_right_lt(n, n, Lf, L1);
            }
            if (i + 2 < NB) { memcpy(L2, P->Y2 + (size_t)i * nn, 8 * nn); trsm_right_lt(n, n, Lf, L2); }
            double *y = W->y + (size_t)i * n;
            for (int k = 0; k < n; ++k) y[k] = -W->beta[(size_t)i * n + k];
            if (i >= 1) gemv_n(n, n, -1.0, W->L1 + (size_t)(i - 1) * nn, W->y + (size_t)(i - 1) * n, y);
            if (i >= 2) gemv_n(n, n, -1.0, W->L2 + (size_t)(i - 2) * nn, W->y + (size_t)(i - 2) * n, y);
            trsv_lower(n, Lf, y);
        }
        for (int i = NB - 1; i >= 0; --i) {                    /* (:32) */
            double *dn = W->dnu + (size_t)i * n;
            memcpy(dn, W->y + (size_t)i * n, 8 * n);
            if (i + 1 < NB) gemv_t(n, n, -1.0, W->L1 + (size_t)i * nn, W->dnu + (size_t)(i + 1) * n, dn);
            if (i + 2 < NB) gemv_t(n, n, -1.0, W->L2 + (size_t)i * nn, W->dnu + (size_t)(i + 2) * n, dn);
            trsv_lower_t(n, W->Lf + (size_t)i * nn, dn);
        }
        /* dz = inv(Phi)(-r_d - C' dnu)  (:34-35) */
        apply_Ct(S, P, W->dnu, W->dz);
        for (size_t i = 0; i < N; ++i) W->dz[i] = -W->rd[i] - W->dz[i];
        for (int t = 0; t < T; ++t) {
            if (P->r_diag) {
                for (int k = 0; k < m; ++k) UOF(W->dz, t)[k] /= (2.0 * AT(S->R, m, k, k) + W->phi[(size_t)t * m + k]);
            } else {
                const double *Rt = W->Rt + (size_t)t * m * m;
                trsv_lower(m, Rt, UOF(W->dz, t)); trsv_lower_t(m, Rt, UOF(W->dz, t));
            }
            int j = t + 1;
            double tmp[n];
            memset(tmp, 0, sizeof tmp);
            gemv_n(n, n, 1.0, Qi_of(P, j), XOF(W->dz, j), tmp);
            memcpy(XOF(W->dz, j), tmp, 8 * n);
        }
        /* backtracking on ||[r_p; r_d]|| with d frozen (backtracking_inf_newton.m:2-11) */
        double t_ls = 1.0;
        int nh = 0;
        for (;;) {
            for (size_t i = 0; i < N; ++i) W->zt[i] = W->z[i] + t_ls * W->dz[i];
            for (size_t i = 0; i < nb; ++i) W->nut[i] = W->nu[i] + t_ls * W->dnu[i];
            apply_2H_g(S, P, W->zt, W->rdt);
            apply_Ct(S, P, W->nut, W->p);
            for (size_t i = 0; i < N; ++i) W->rdt[i] += W->p[i];
            for (int t = 0; t < T; ++t) for (int k = 0; k < m; ++k) UOF(W->rdt, t)[k] += W->dbar[(size_t)t * m + k];
            apply_C(S, P, W->zt, W->rpt);
            for (size_t i = 0; i < nb; ++i) W->rpt[i] -= W->b[i];
            double nrt = sqrt(norm2sq(N, W->rdt) + norm2sq(nb, W->rpt));
            if (!(nrt > (1.0 - alpha * t_ls) * nr0)) break;
            if (ls_max > 0 && nh >= ls_max) { status = ST_LS_MAX; break; }
            t_ls *= bt; ++nh;
        }
        halv += nh;
        memcpy(W->z, W->zt, 8 * N);
        memcpy(W->nu, W->nut, 8 * nb);
        ++iters;
    }
done:
    memcpy(z_out, W->z, 8 * N);
    if (nu_out) memcpy(nu_out, W->nu, 8 * nb);
    if (iters_out) *iters_out = iters;
    if (halv_out) *halv_out = halv;
    return status;
}

/* ------------------------------------------------------------------ exported entry points */
/* Batched solve: instance b uses column b of each per-instance array.
 * x0, x0_pre: n x nb ; w: (T n) x nb or NULL ; xf: n x nb or NULL ; z0: N x nb ;
 * nu0: (NB n) x nb ; z_out: N x nb ; nu_out: (NB n) x nb or NULL ; status/iters/halvings: nb.
 * Returns 0, or -1 if Q/Qf are not PD, -2 on bad arguments. */
int fref_solve_batch(const fref_sys *S, double kappa, int niters, int ls_max, double alpha, double beta,
                     double tol_r, double tol_p, int nbatch,
                     const double *x0, const double *x0_pre, const double *w, const double *xf,
                     const double *z0, const double *nu0,
                     double *z_out, double *nu_out, int *status, int *iters, int *halvings, int nthreads)
{
    if (!S || !S->A1 || !S->B || !S->Q || !S->R || !S->Qf || !S->u_min || !S->u_max || !x0 || !z0 || !nu0 || !z_out)
        return -2;
    if (S->A2 && !x0_pre) return -2;
    fref_pre P;
    if (precompute(S, xf != NULL, &P)) { pre_free(&P); return -1; }
    int n = P.n, m = P.m, T = P.T;
    size_t N = (size_t)T * (n + m), nb = (size_t)P.NB * n;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
#pragma omp parallel num_threads(nthreads)
    {
        fref_ws W;
        ws_alloc(&W, &P, !P.r_diag);
#pragma omp for schedule(dynamic, 1)
        for (int b = 0; b < nbatch; ++b) {
            int it = 0, hv = 0;
            int st = solve_one(S, &P, &W, kappa, niters, ls_max, alpha, beta, tol_r, tol_p,
                               x0 + (size_t)b * n, x0_pre ? x0_pre + (size_t)b * n : NULL,
                               w ? w + (size_t)b * T * n : NULL, xf ? xf + (size_t)b * n : NULL,
                               z0 + (size_t)b * N, nu0 + (size_t)b * nb,
                               z_out + (size_t)b * N, nu_out ? nu_out + (size_t)b * nb : NULL, &it, &hv);
            if (status) status[b] = st;
            if (iters) iters[b] = it;
            if (halvings) halvings[b] = hv;
        }
        ws_free(&W);
    }
    pre_free(&P);
    return 0;
}

int fref_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
