// fastMPC batched Newton solve -- sm_100a CUDA kernels (fp64).
//
// One persistent CTA per MPC instance (dynamic instance counter).  What the reference does with
// dense N x N matrices (inf_newton_solver.m:10-41) is done here on the block structure:
//   Phi is block diagonal (box rows on u only, fast_mpc_ineq_const.m:42-56), the Schur complement
//   Y = C inv(Phi) C' is block penta-diagonal in n x n blocks (two-lag C, fast_mpc_eq_const.m:38-49),
//   factored by a band-2 block Cholesky along the horizon with the stage blocks in shared memory.
// The iterate (U, X) lives in the output arrays; per-CTA vectors and the band factor live in an
// L2-resident scratch area (grid-sized, not batch-sized).
#include <cuda_runtime.h>
#include <math.h>
#include <stdlib.h>
#include "fmpc_internal.h"
#include "fmpc_device.cuh"

using namespace fmpc_dev;


// =============================================================================================
// The solve kernel
// =============================================================================================
__global__ void fmpc_solve_kernel_v1(const DevSys S, const StepArgs A)
{
    extern __shared__ double smem[];
    const int n = S.n, m = S.m, T = S.T;
    const int NB = T + (A.has_xf ? 1 : 0);
    const int ld = n | 1;                         // odd leading dimension: conflict-free column walks
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5;
    const Ctx c{S, n, m, T, NB, tid, nt};
    const WsLayout L = WsLayout::make(n, m, T);

    // shared memory carve-up: 6 stage blocks + pinv_i (m) + yprev (2n) + reduce scratch
    double *blk[6];
    for (int i = 0; i < 6; ++i) blk[i] = smem + (size_t)i * n * ld;
    double *sm_pinv = smem + (size_t)6 * n * ld;
    double *sm_vec = sm_pinv + m;                 // 3n : rhs / y_{i-1} / y_{i-2}
    double *red = sm_vec + 3 * n;                 // 34
    __shared__ int s_inst, s_flag;

    double *ws = A.ws + (size_t)blockIdx.x * A.ws_stride;
    double *nu = ws + L.nu, *dnu = ws + L.dnu, *yv = ws + L.yv, *rp = ws + L.rp, *rpt = ws + L.rpt, *bv = ws + L.bv;
    double *hx = ws + L.hx, *hdx = ws + L.hdx, *dx = ws + L.dx, *xt = ws + L.xt, *rdx = ws + L.rdx;
    double *hu = ws + L.hu, *hdu = ws + L.hdu, *du = ws + L.du, *ut = ws + L.ut, *dbar = ws + L.dbar;
    double *pinv = ws + L.pinv, *rdu = ws + L.rdu;
    double *gLf = ws + L.Lf, *gL1 = ws + L.L1, *gL2 = ws + L.L2;
    const size_t nn = (size_t)n * n;

    for (;;) {
        __syncthreads();
        if (tid == 0) s_inst = (int)atomicAdd(A.counter, 1u);
        __syncthreads();
        const int b = s_inst;
        if (b >= A.nbatch) break;

        double *u = A.U + (size_t)b * m * T;
        double *x = A.X + (size_t)b * n * T;
        const double *x0 = A.x0 + (size_t)b * n;
        const double *x0p = A.x0_pre ? A.x0_pre + (size_t)b * n : nullptr;

        // ---- initial iterate: warm start or midpoint cold start (fast_mpc_init.m:12-26) ----
        if (A.cold) {
            for (int e = tid; e < T * m; e += nt) { const int j = e % m; u[e] = (S.umin[j] + S.umax[j]) / 2; }
            for (int e = tid; e < T * n; e += nt) { const int k = e % n; x[e] = (S.xmin[k] + S.xmax[k]) / 2; }
        } else if (A.U0 != A.U || A.X0 != A.X) {
            const double *u0 = A.U0 + (size_t)b * m * T, *xx0 = A.X0 + (size_t)b * n * T;
            for (int e = tid; e < T * m; e += nt) u[e] = u0[e];
            for (int e = tid; e < T * n; e += nt) x[e] = xx0[e];
        }
        for (int e = tid; e < NB * n; e += nt) nu[e] = A.nu0[(size_t)b * NB * n + e];
        // ---- b (fast_mpc_eq_const.m:39,44,47,68) ----
        for (int e = tid; e < NB * n; e += nt) {
            const int i = e / n, k = e - i * n;
            double v;
            if (i < T) {
                v = A.w ? A.w[(size_t)b * T * n + e] : 0.0;
                if (i == 0) {
                    double s = 0.0;
                    for (int kk = 0; kk < n; ++kk) s = fma(S.A1[k + n * kk], x0[kk], s);
                    if (S.has_a2) for (int kk = 0; kk < n; ++kk) s = fma(S.A2[k + n * kk], x0p[kk], s);
                    v += s;
                } else if (i == 1 && S.has_a2) {
                    double s = 0.0;
                    for (int kk = 0; kk < n; ++kk) s = fma(S.A2[k + n * kk], x0[kk], s);
                    v += s;
                }
            } else {
                v = A.xf[(size_t)b * n + k];
            }
            bv[e] = v;
        }
        __syncthreads();
        apply_Ct(c, nu, hu, hx);
        apply_C_minus_b(c, u, x, bv, rp);
        __syncthreads();

        int status = ST_OK, iters = 0;
        for (int it = 0; it < A.niters; ++it) {
            // ---- barrier terms (inf_newton_KKT_H.m:3-13) ----
            for (int e = tid; e < T * m; e += nt) {
                const int j = e % m;
                const double uu = u[e];
                const double sp = S.umax[j] - uu, sm = -S.umin[j] + uu;
                const double dp = 1.0 / sp, dm = 1.0 / sm;
                dbar[e] = A.kappa * (dp - dm);
                pinv[e] = 1.0 / (S.r2[j] + A.kappa * (dp * dp + dm * dm));
            }
            __syncthreads();
            // ---- residuals + early exit (inf_newton_solver.m:12-22) ----
            double ssp;
            const double ss0 = resid_sumsq(c, u, x, hu, nullptr, hx, nullptr, 0.0, dbar, rp, red, rdu, rdx, &ssp);
            const double nr0 = sqrt(ss0);
            if (!isfinite(nr0)) { status = ST_NONFINITE; break; }
            if (nr0 <= A.tol_r && sqrt(ssp) <= A.tol_p) { status = ST_EARLY_EXIT; break; }
            __syncthreads();
            // ---- beta = -r_p + C inv(Phi) r_d  (:28-29); p reuses du/dx as scratch ----
            for (int e = tid; e < T * m; e += nt) du[e] = rdu[e] * pinv[e];
            for (int e = tid; e < T * n; e += nt) {
                const int jm1 = e / n, k = e - jm1 * n;
                dx[e] = rdx[e] * ((jm1 == T - 1) ? S.qif[k] : S.qi[k]);
            }
            __syncthreads();
            for (int e = tid; e < NB * n; e += nt) {
                const int i = e / n, k = e - i * n;
                double acc;
                if (i < T) {
                    acc = dx[i * n + k];
                    const double *pu = du + (size_t)i * m;
                    double s = 0.0;
                    for (int j = 0; j < m; ++j) s = fma(__ldg(S.B + k + n * j), pu[j], s);
                    acc -= s;
                    if (i >= 1) {
                        const double *px = dx + (size_t)(i - 1) * n;
                        s = 0.0;
                        for (int kk = 0; kk < n; ++kk) s = fma(__ldg(S.A1 + k + n * kk), px[kk], s);
                        acc -= s;
                    }
                    if (i >= 2 && S.has_a2) {
                        const double *px = dx + (size_t)(i - 2) * n;
                        s = 0.0;
                        for (int kk = 0; kk < n; ++kk) s = fma(__ldg(S.A2 + k + n * kk), px[kk], s);
                        acc -= s;
                    }
                } else {
                    acc = dx[(T - 1) * n + k];
                }
                yv[e] = -(acc - rp[e]);           // rhs of  Y dnu = -beta
            }
            __syncthreads();

            // ---- band-2 block Cholesky of Y fused with the forward solve (:30-31) ----
            // roles of the six shared blocks rotate every stage
            double *bS = blk[0], *bM1 = blk[1], *bM2 = blk[2], *bL1p = blk[3], *bL2p = blk[4], *bL2pp = blk[5];
            bool fail = false;
            for (int i = 0; i < NB; ++i) {
                const bool has1 = (i + 1 < NB), has2 = (i + 2 < NB) && S.has_a2;
                if (i < T) for (int j = tid; j < m; j += nt) sm_pinv[j] = pinv[(size_t)i * m + j];
                __syncthreads();
                // 1. assemble S (lower triangle), M1, M2, rhs
                const double *Yd = S.ypool + (size_t)S.ydi[i] * nn;
                for (int e = tid; e < n * n; e += nt) {
                    const int r = e / n, cc = e - r * n;
                    if (cc > r) continue;
                    double acc = Yd[e];
                    if (i < T) {
                        double s = 0.0;
                        for (int j = 0; j < m; ++j)
                            s = fma(__ldg(S.B + r + n * j) * sm_pinv[j], __ldg(S.B + cc + n * j), s);
                        acc += s;
                    }
                    if (i >= 1) {
                        double s = 0.0;
                        for (int k = 0; k < n; ++k) s = fma(bL1p[r * ld + k], bL1p[cc * ld + k], s);
                        acc -= s;
                    }
                    if (i >= 2 && S.has_a2) {
                        double s = 0.0;
                        for (int k = 0; k < n; ++k) s = fma(bL2pp[r * ld + k], bL2pp[cc * ld + k], s);
                        acc -= s;
                    }
                    bS[r * ld + cc] = acc;
                }
                if (has1) {
                    const int yi = S.y1i[i];
                    for (int e = tid; e < n * n; e += nt) {
                        const int r = e / n, cc = e - r * n;
                        double acc = (yi >= 0) ? S.ypool[(size_t)yi * nn + e] : 0.0;
                        if (i >= 1 && S.has_a2) {
                            double s = 0.0;
                            for (int k = 0; k < n; ++k) s = fma(bL2p[r * ld + k], bL1p[cc * ld + k], s);
                            acc -= s;
                        }
                        bM1[r * ld + cc] = acc;
                    }
                }
                if (has2) {
                    const int yi = S.y2i[i];
                    for (int e = tid; e < n * n; e += nt) {
                        const int r = e / n, cc = e - r * n;
                        bM2[r * ld + cc] = (yi >= 0) ? S.ypool[(size_t)yi * nn + e] : 0.0;
                    }
                }
                for (int k = tid; k < n; k += nt) {     // rhs_i = yv_i - L1p y_{i-1} - L2pp y_{i-2}
                    double acc = yv[i * n + k];
                    if (i >= 1) { double s = 0.0; for (int kk = 0; kk < n; ++kk) s = fma(bL1p[k * ld + kk], sm_vec[n + kk], s); acc -= s; }
                    if (i >= 2 && S.has_a2) { double s = 0.0; for (int kk = 0; kk < n; ++kk) s = fma(bL2pp[k * ld + kk], sm_vec[2 * n + kk], s); acc -= s; }
                    sm_vec[k] = acc;
                }
                __syncthreads();
                // 2. potrf (warp 0)
                if (wid == 0) {
                    const int info = warp_potrf(bS, n, ld, lane);
                    if (lane == 0) s_flag = info;
                }
                __syncthreads();
                if (s_flag) { fail = true; break; }
                // 3. triangular solves: rows of M1, M2 (X L' = M), and L y = rhs
                {
                    const int nrows = (has1 ? n : 0) + (has2 ? n : 0);
                    for (int rr = tid; rr < nrows; rr += nt) {
                        double *row = (rr < n && has1) ? (bM1 + rr * ld) : (bM2 + (rr - (has1 ? n : 0)) * ld);
                        for (int j = 0; j < n; ++j) {
                            double s = row[j];
                            for (int k = 0; k < j; ++k) s = fma(-row[k], bS[j * ld + k], s);
                            row[j] = s / bS[j * ld + j];
                        }
                    }
                    // forward substitution by the last warp (keeps it off the threads doing rows when nt > nrows)
                    if (wid == (nt >> 5) - 1) {
                        for (int j = 0; j < n; ++j) {
                            __syncwarp();
                            const double yj = sm_vec[j] / bS[j * ld + j];
                            __syncwarp();
                            if (lane == 0) sm_vec[j] = yj;
                            for (int k = j + 1 + lane; k < n; k += 32) sm_vec[k] = fma(-bS[k * ld + j], yj, sm_vec[k]);
                        }
                    }
                }
                __syncthreads();
                // 4. spill factor blocks for the backward pass, publish y_i, rotate
                for (int e = tid; e < n * n; e += nt) {
                    const int r = e / n, cc = e - r * n;
                    gLf[(size_t)i * nn + e] = (cc <= r) ? bS[r * ld + cc] : 0.0;
                    if (has1) gL1[(size_t)i * nn + e] = bM1[r * ld + cc];
                    if (has2) gL2[(size_t)i * nn + e] = bM2[r * ld + cc];
                }
                for (int k = tid; k < n; k += nt) {
                    const double yk = sm_vec[k];
                    yv[i * n + k] = yk;
                    sm_vec[2 * n + k] = sm_vec[n + k];
                    sm_vec[n + k] = yk;
                }
                {   // rotate: L2pp <- L2p, L2p <- M2, L1p <- M1 ; freed buffers become S, M1, M2
                    double *oS = bS, *oL1p = bL1p, *oL2pp = bL2pp;
                    bL2pp = bL2p; bL2p = bM2; bL1p = bM1;
                    bS = oS; bM1 = oL1p; bM2 = oL2pp;
                }
                __syncthreads();
            }
            if (fail) { status = ST_NOT_PD; break; }

            // ---- backward solve  L' dnu = y  (:32) ----
            for (int i = NB - 1; i >= 0; --i) {
                const bool has1 = (i + 1 < NB), has2 = (i + 2 < NB) && S.has_a2;
                for (int k = tid; k < n; k += nt) {
                    double acc = yv[i * n + k];
                    if (has1) { double s = 0.0; for (int r = 0; r < n; ++r) s = fma(gL1[(size_t)i * nn + r * n + k], dnu[(i + 1) * n + r], s); acc -= s; }
                    if (has2) { double s = 0.0; for (int r = 0; r < n; ++r) s = fma(gL2[(size_t)i * nn + r * n + k], dnu[(i + 2) * n + r], s); acc -= s; }
                    sm_vec[k] = acc;
                }
                __syncthreads();
                if (wid == 0) {
                    const double *Lf = gLf + (size_t)i * nn;
                    for (int j = n - 1; j >= 0; --j) {
                        __syncwarp();
                        const double xj = sm_vec[j] / Lf[j * n + j];
                        __syncwarp();
                        if (lane == 0) sm_vec[j] = xj;
                        for (int k = lane; k < j; k += 32) sm_vec[k] = fma(-Lf[j * n + k], xj, sm_vec[k]);
                    }
                }
                __syncthreads();
                for (int k = tid; k < n; k += nt) dnu[i * n + k] = sm_vec[k];
                __syncthreads();
            }

            // ---- dz = inv(Phi)(-r_d - C' dnu)  (:34-35) ----
            apply_Ct(c, dnu, hdu, hdx);
            __syncthreads();
            for (int e = tid; e < T * m; e += nt) du[e] = -(rdu[e] - hdu[e]) * pinv[e];
            for (int e = tid; e < T * n; e += nt) {
                const int jm1 = e / n, k = e - jm1 * n;
                dx[e] = -(rdx[e] + hdx[e]) * ((jm1 == T - 1) ? S.qif[k] : S.qi[k]);
            }
            __syncthreads();

            // ---- backtracking on ||[r_p; r_d]||, d frozen (backtracking_inf_newton.m:2-11) ----
            double t = 1.0;
            int nh = 0;
            for (;;) {
                for (int e = tid; e < T * m; e += nt) ut[e] = __fma_rn(t, du[e], u[e]);
                for (int e = tid; e < T * n; e += nt) xt[e] = __fma_rn(t, dx[e], x[e]);
                __syncthreads();
                apply_C_minus_b(c, ut, xt, bv, rpt);
                __syncthreads();
                const double sst = resid_sumsq(c, ut, xt, hu, hdu, hx, hdx, t, dbar, rpt, red, nullptr, nullptr, nullptr);
                const double nrt = sqrt(sst);
                if (!(nrt > (1.0 - A.alpha * t) * nr0)) break;
                if (t == 0.0) break;
                if (A.ls_max > 0 && nh >= A.ls_max) { status = ST_LS_MAX; break; }
                t *= A.beta;
                ++nh;
                __syncthreads();
            }
            __syncthreads();
            // accept
            for (int e = tid; e < T * m; e += nt) { u[e] = ut[e]; hu[e] = __fma_rn(t, hdu[e], hu[e]); }
            for (int e = tid; e < T * n; e += nt) { x[e] = xt[e]; hx[e] = __fma_rn(t, hdx[e], hx[e]); }
            for (int e = tid; e < NB * n; e += nt) { nu[e] = __fma_rn(t, dnu[e], nu[e]); rp[e] = rpt[e]; }
            ++iters;
            __syncthreads();
        }
        if (tid == 0) {
            if (A.status) A.status[b] = status;
            if (A.iters) A.iters[b] = iters;
            atomicAdd(A.iters_total, (unsigned long long)iters);
        }
    }
}

// =============================================================================================
// state update  x+ = A1 x + A2 x- + B u (+ w)      (VAR_2/fast_mpc_eq_const.m:39-47 as a recurrence)
// =============================================================================================
__global__ void fmpc_state_update_kernel(const DevSys S, int nbatch, const double *__restrict__ x,
                                         const double *__restrict__ xpre, const double *__restrict__ u,
                                         const double *__restrict__ w, double *__restrict__ xnext)
{
    const int n = S.n, m = S.m;
    extern __shared__ double sm[];          // x (n), xpre (n), u (m) of this instance
    for (int b = blockIdx.x; b < nbatch; b += gridDim.x) {
        __syncthreads();
        for (int k = threadIdx.x; k < n; k += blockDim.x) {
            sm[k] = x[(size_t)b * n + k];
            sm[n + k] = (S.has_a2 && xpre) ? xpre[(size_t)b * n + k] : 0.0;
        }
        for (int j = threadIdx.x; j < m; j += blockDim.x) sm[2 * n + j] = u[(size_t)b * m + j];
        __syncthreads();
        for (int k = threadIdx.x; k < n; k += blockDim.x) {
            double s = w ? w[(size_t)b * n + k] : 0.0;
            for (int kk = 0; kk < n; ++kk) s = fma(__ldg(S.A1 + k + n * kk), sm[kk], s);
            if (S.has_a2) for (int kk = 0; kk < n; ++kk) s = fma(__ldg(S.A2 + k + n * kk), sm[n + kk], s);
            for (int j = 0; j < m; ++j) s = fma(__ldg(S.B + k + n * j), sm[2 * n + j], s);
            xnext[(size_t)b * n + k] = s;
        }
    }
}

// closed-loop glue: x0 = a_k + B u_prev ; x0_pre <- old x0 ; shift the warm start one stage
__global__ void fmpc_shift_warm_kernel(const DevSys S, int nbatch, const double *__restrict__ a_k, int a_stride,
                                       double *__restrict__ X, double *__restrict__ U, double *__restrict__ x0,
                                       double *__restrict__ x0_pre, double *__restrict__ u_prev, int first)
{
    const int n = S.n, m = S.m, T = S.T;
    extern __shared__ double sm[];          // u_prev (m)
    for (int b = blockIdx.x; b < nbatch; b += gridDim.x) {
        __syncthreads();
        double *Ub = U + (size_t)b * m * T, *Xb = X + (size_t)b * n * T;
        for (int j = threadIdx.x; j < m; j += blockDim.x) {
            const double up = first ? 0.0 : Ub[j];      // U(:,0) of the previous solve = applied input
            sm[j] = up;
            u_prev[(size_t)b * m + j] = up;
        }
        __syncthreads();
        for (int k = threadIdx.x; k < n; k += blockDim.x) {
            double s = a_k[(size_t)b * a_stride + k];
            for (int j = 0; j < m; ++j) s = fma(__ldg(S.B + k + n * j), sm[j], s);
            x0_pre[(size_t)b * n + k] = first ? 0.0 : x0[(size_t)b * n + k];
            x0[(size_t)b * n + k] = s;
        }
        __syncthreads();
        if (!first) {
            // shift: stage t <- stage t+1, last stage repeated (per-thread strided copy, ascending t)
            for (int j = threadIdx.x; j < m; j += blockDim.x)
                for (int t = 0; t + 1 < T; ++t) Ub[(size_t)t * m + j] = Ub[(size_t)(t + 1) * m + j];
            for (int k = threadIdx.x; k < n; k += blockDim.x)
                for (int t = 0; t + 1 < T; ++t) Xb[(size_t)t * n + k] = Xb[(size_t)(t + 1) * n + k];
        }
    }
}

// resident closed loop, kernels without the fused variants: warm start <- previous solution shifted one stage, in place
__global__ void fmpc_shift_inplace_kernel(int n, int m, int T, int nbatch, double *__restrict__ X, double *__restrict__ U)
{
    for (int b = blockIdx.x; b < nbatch; b += gridDim.x) {
        double *Ub = U + (size_t)b * m * T, *Xb = X + (size_t)b * n * T;
        for (int j = threadIdx.x; j < m; j += blockDim.x)
            for (int t = 0; t + 1 < T; ++t) Ub[(size_t)t * m + j] = Ub[(size_t)(t + 1) * m + j];
        for (int k = threadIdx.x; k < n; k += blockDim.x)
            for (int t = 0; t + 1 < T; ++t) Xb[(size_t)t * n + k] = Xb[(size_t)(t + 1) * n + k];
    }
}
// u_first[:,b] = U(:,0,b): the input the loop applies (README.md:589)
__global__ void fmpc_extract_first_kernel(int m, int T, int nbatch, const double *__restrict__ U, double *__restrict__ u_first)
{
    const size_t tot = (size_t)nbatch * m;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += (size_t)gridDim.x * blockDim.x) {
        const size_t b = e / m;
        u_first[e] = U[b * (size_t)m * T + (e - b * m)];
    }
}

// closed-loop logs of step k: U_acc[:,k,b] = U(:,0,b) (the applied input), X_acc[:,k,b] = x0[:,b], iters_acc[k,b] = iters[b]
__global__ void fmpc_log_step_kernel(int n, int m, int T, int nbatch, int K, int k, const double *__restrict__ U,
                                     const double *__restrict__ x0, const int *__restrict__ iters, double *__restrict__ Uacc,
                                     double *__restrict__ Xacc, int *__restrict__ itacc)
{
    const size_t tot = (size_t)nbatch * (m + n);
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += (size_t)gridDim.x * blockDim.x) {
        const size_t b = e / (m + n);
        const int j = (int)(e - b * (m + n));
        if (j < m) Uacc[(b * K + k) * m + j] = U[b * (size_t)m * T + j];
        else Xacc[(b * K + k) * n + (j - m)] = x0[b * n + (j - m)];
        if (j == 0 && itacc) itacc[b * K + k] = iters[b];
    }
}

// =============================================================================================
// MATLAB's default global stream on the device: `nu = rand(length(b),1)` (inf_newton_solver.m:2) is MT19937 seeded with
// 5489, doubles by genrand_res53.  The stream is ONE dependency chain,
//     x[k+624] = x[k+397] ^ f(x[k], x[k+1]),
// which reaches back only 227 words (three barrier-separated phases per 624-word block); substituting the recurrence into itself
// expresses EVERY word of the next block by words of the current one,
//     n[e] = o[e+397] ^ f(o[e],o[e+1])                                                        e <  227
//     n[e] = o[e+170] ^ f(o[e-227],o[e-226]) ^ f(o[e],o[e+1])                           227 <= e <  454
//     n[e] = o[e- 57] ^ f(o[e-454],o[e-453]) ^ f(o[e-227],o[e-226]) ^ f(o[e],o[e+1])    454 <= e <  623
//     n[623] = n[396] ^ f(o[623], n[0])      (both expanded the same way)
// so a block costs ONE barrier.  Two kernels:
//   fmpc_mt_twist_kernel   : ONE CTA, one word per thread (operand indices are loop invariants, no divergence except word 623),
//                            walks the chain and writes the raw blocks to global memory: ~500 cycles per block, the floor of one
//                            LDS -> ALU -> STS -> barrier round trip (scripts/ubench/mt.cu, profiles/r02_mt_generator.log);
//   fmpc_mt_convert_kernel : tempering + genrand_res53 of the raw words, embarrassingly parallel.
// state[0..623] = the last generated block; the index of the next unread word lives on the host (fmpc_api.cu).
// =============================================================================================
__device__ __forceinline__ unsigned mt_f(unsigned a, unsigned b)
{
    const unsigned y = (a & 0x80000000u) | (b & 0x7fffffffu);
    return (y >> 1) ^ ((0u - (y & 1u)) & 0x9908b0dfu);
}
__device__ __forceinline__ unsigned mt_temper(unsigned y)
{
    y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
    return y;
}
// raw[0..624) <- the stored block, raw[624 (b + 1) ..) <- the b-th new block, b < nblocks; state <- the last block
__global__ void __launch_bounds__(640, 1) fmpc_mt_twist_kernel(unsigned *__restrict__ state, unsigned *__restrict__ raw, int nblocks)
{
    __shared__ __align__(16) unsigned mt[2][624];
    const int e = threadIdx.x;
    const bool act = e < 624, last = (e == 623);
    if (act) { const unsigned v = state[e]; mt[0][e] = v; raw[e] = v; }
    // loop-invariant operand indices of word e: x ^ f(a0,a1) ^ (f(b0,b1) & mb) ^ (f(c0,c1) & mc)
    int ix = 0, ia0 = 0, ia1 = 0, ib0 = 0, ib1 = 0, ic0 = 0, ic1 = 0;
    unsigned mb = 0u, mc = 0u;
    if (e < 227) { ix = e + 397; ia0 = e; ia1 = e + 1; }
    else if (e < 454) { ix = e + 170; ia0 = e; ia1 = e + 1; ib0 = e - 227; ib1 = e - 226; mb = ~0u; }
    else if (e < 623) { ix = e - 57; ia0 = e; ia1 = e + 1; ib0 = e - 227; ib1 = e - 226; mb = ~0u; ic0 = e - 454; ic1 = e - 453; mc = ~0u; }
    __syncthreads();
    int cur = 0;
    for (int b = 0; b < nblocks; ++b) {
        const unsigned *o = mt[cur];
        unsigned v;
        if (!last) {
            v = o[ix] ^ mt_f(o[ia0], o[ia1]) ^ (mt_f(o[ib0], o[ib1]) & mb) ^ (mt_f(o[ic0], o[ic1]) & mc);
        } else {
            const unsigned n0 = o[397] ^ mt_f(o[0], o[1]);
            const unsigned n396 = o[566] ^ mt_f(o[169], o[170]) ^ mt_f(o[396], o[397]);
            v = n396 ^ mt_f(o[623], n0);
        }
        if (act) { mt[cur ^ 1][e] = v; raw[(size_t)(b + 1) * 624 + e] = v; }
        __syncthreads();
        cur ^= 1;
    }
    if (act) state[e] = mt[cur][e];
}
// out[i] = genrand_res53(raw[2 i], raw[2 i + 1])
__global__ void fmpc_mt_convert_kernel(const unsigned *__restrict__ raw, double *__restrict__ out, unsigned long long count)
{
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned a = mt_temper(raw[2 * i]) >> 5, b = mt_temper(raw[2 * i + 1]) >> 6;
        // (a 2^26 + b) 2^-53 without integer -> double conversions: both pieces are exact in the mantissa of 2^52 + v
        const double da = __longlong_as_double(0x4330000000000000ll | (long long)a) - 4503599627370496.0;
        const double db = __longlong_as_double(0x4330000000000000ll | (long long)b) - 4503599627370496.0;
        out[i] = (da * 67108864.0 + db) * (1.0 / 9007199254740992.0);
    }
}

// =============================================================================================
// host-side launch helpers
// =============================================================================================
static size_t solve_smem_bytes(int n, int m)
{
    const int ld = n | 1;
    return ((size_t)6 * n * ld + m + 3 * n + 34) * sizeof(double);
}

int fmpc_solve_config(const DevSys &S, int device, SolveLaunchCfg *cfg)
{
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return -1;
    const size_t smem = solve_smem_bytes(S.n, S.m);
    if (smem > (size_t)prop.sharedMemPerBlockOptin) return -2;
    const int block = (S.n <= 32) ? 128 : 256;
    if (cudaFuncSetAttribute(fmpc_solve_kernel_v1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return -3;
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fmpc_solve_kernel_v1, block, smem) != cudaSuccess || per_sm < 1)
        return -4;
    cfg->grid = prop.multiProcessorCount * per_sm;
    cfg->block = block;
    cfg->smem = smem;
    cfg->use_mma = 0;
    return 0;
}

void fmpc_launch_solve(const DevSys &S, const StepArgs &A, const SolveLaunchCfg &cfg, void *stream)
{
    int grid = cfg.grid < A.nbatch ? cfg.grid : A.nbatch;
    if (grid < 1) grid = 1;
    fmpc_solve_kernel_v1<<<grid, cfg.block, cfg.smem, (cudaStream_t)stream>>>(S, A);
}

void fmpc_launch_state_update(const DevSys &S, int nbatch, const double *x, const double *xpre, const double *u,
                              const double *w, double *xnext, void *stream)
{
    int grid = nbatch < 148 * 8 ? nbatch : 148 * 8;
    if (grid < 1) grid = 1;
    const size_t smem = (size_t)(2 * S.n + S.m) * sizeof(double);
    fmpc_state_update_kernel<<<grid, 64, smem, (cudaStream_t)stream>>>(S, nbatch, x, xpre, u, w, xnext);
}

void fmpc_launch_shift_warm(const DevSys &S, int nbatch, const double *a_k, int a_stride, double *X, double *U,
                            double *x0, double *x0_pre, double *u_prev, int first, void *stream)
{
    int grid = nbatch < 148 * 8 ? nbatch : 148 * 8;
    if (grid < 1) grid = 1;
    fmpc_shift_warm_kernel<<<grid, 64, (size_t)S.m * sizeof(double), (cudaStream_t)stream>>>(S, nbatch, a_k, a_stride, X, U,
                                                                                            x0, x0_pre, u_prev, first);
}

void fmpc_launch_log_step(int n, int m, int T, int nbatch, int K, int k, const double *U, const double *x0, const int *iters,
                          double *Uacc, double *Xacc, int *itacc, void *stream)
{
    const size_t tot = (size_t)nbatch * (m + n);
    int grid = (int)((tot + 255) / 256);
    if (grid > 148 * 16) grid = 148 * 16;
    if (grid < 1) grid = 1;
    fmpc_log_step_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(n, m, T, nbatch, K, k, U, x0, iters, Uacc, Xacc, itacc);
}

// Appends `count` doubles of the stream at `out`.  *idx = next unread word of the stored block (624 = none), updated.
// `raw` holds at least 2 count + 1248 words.
void fmpc_launch_mt_fill(unsigned *state, unsigned *raw, int *idx, double *out, unsigned long long count, void *stream)
{
    if (count == 0) return;
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned long long need = 2 * count, avail = (unsigned long long)(624 - *idx);
    const int nblk = (need > avail) ? (int)((need - avail + 623) / 624) : 0;
    fmpc_mt_twist_kernel<<<1, 640, 0, st>>>(state, raw, nblk);
    unsigned long long cb = (count + 255) / 256;
    if (cb > 148 * 4) cb = 148 * 4;                       // a small grid: it shares the GPU with a running solve
    fmpc_mt_convert_kernel<<<(int)cb, 256, 0, st>>>(raw + *idx, out, count);
    *idx = (int)((unsigned long long)*idx + need - 624ull * (unsigned long long)nblk);
}

void fmpc_launch_shift_inplace(int n, int m, int T, int nbatch, double *X, double *U, void *stream)
{
    int grid = nbatch < 148 * 8 ? nbatch : 148 * 8;
    if (grid < 1) grid = 1;
    fmpc_shift_inplace_kernel<<<grid, 64, 0, (cudaStream_t)stream>>>(n, m, T, nbatch, X, U);
}

void fmpc_launch_extract_first(int m, int T, int nbatch, const double *U, double *u_first, void *stream)
{
    const size_t tot = (size_t)nbatch * m;
    int grid = (int)((tot + 255) / 256);
    if (grid > 148 * 8) grid = 148 * 8;
    if (grid < 1) grid = 1;
    fmpc_extract_first_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(m, T, nbatch, U, u_first);
}
