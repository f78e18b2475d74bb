// Device helpers shared by the fastMPC solve kernels (v1 generic path and the DMMA path).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include "fmpc_internal.h"

namespace fmpc_dev {


static __device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// CTA-wide sum; every thread gets the result.  `red` holds >= 33 doubles of shared memory.
static __device__ double block_sum(double v, double *red)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();                    // protect `red` from the previous use
    if (lane == 0) red[wid] = v;
    __syncthreads();
    if (wid == 0) {
        double s = (lane < nw) ? red[lane] : 0.0;
        s = warp_sum(s);
        if (lane == 0) red[32] = s;
    }
    __syncthreads();
    return red[32];
}

// The two residual expressions are kept in ONE place each so that the line search compares norms
// computed by bit-identical arithmetic at t -> 0 (backtracking_inf_newton.m:4 terminates that way).
static __device__ __forceinline__ double rdu_expr(double r2, double rl, double u, double hu, double dbar)
{
    return __dadd_rn(__dsub_rn(__fma_rn(r2, u, rl), hu), dbar);      // 2R u + r - B'nu + k P'd
}
static __device__ __forceinline__ double rdx_expr(double q2, double ql, double x, double hx)
{
    return __dadd_rn(__fma_rn(q2, x, ql), hx);                       // 2Q x + q + (C'nu)_x
}

struct Ctx {
    const DevSys &S;
    int n, m, T, NB, tid, nt;
};

// out = C z - bv   (NB*n), rows follow VAR_2/fast_mpc_eq_const.m:38-49,67-71
// NOTE: u/x/bv/out are written inside this kernel -> plain pointers (no __restrict__/ld.global.nc).
static __device__ void apply_C_minus_b(const Ctx &c, const double *u, const double *x, const double *bv, double *out)
{
    const int n = c.n, m = c.m, T = c.T;
    const DevSys &S = c.S;
    for (int e = c.tid; e < c.NB * n; e += c.nt) {
        const int i = e / n, k = e - i * n;
        double acc;
        if (i < T) {
            acc = x[i * n + k];
            const double *ui = u + (size_t)i * m;
            double s = 0.0;
            for (int j = 0; j < m; ++j) s = fma(__ldg(S.B + k + n * j), ui[j], s);
            acc -= s;
            if (i >= 1) {
                const double *xi = x + (size_t)(i - 1) * n;
                s = 0.0;
                for (int kk = 0; kk < n; ++kk) s = fma(__ldg(S.A1 + k + n * kk), xi[kk], s);
                acc -= s;
            }
            if (i >= 2 && S.has_a2) {
                const double *xi = x + (size_t)(i - 2) * n;
                s = 0.0;
                for (int kk = 0; kk < n; ++kk) s = fma(__ldg(S.A2 + k + n * kk), xi[kk], s);
                acc -= s;
            }
        } else {
            acc = x[(T - 1) * n + k];
        }
        out[e] = acc - bv[e];
    }
}

// hu = B' v_t (T*m)  [so (C'v)_u = -hu],  hx = (C'v)_x (T*n)
static __device__ void apply_Ct(const Ctx &c, const double *v, double *hu, double *hx)
{
    const int n = c.n, m = c.m, T = c.T;
    const DevSys &S = c.S;
    for (int e = c.tid; e < T * m; e += c.nt) {
        const int t = e / m, j = e - t * m;
        const double *vt = v + (size_t)t * n;
        double s = 0.0;
        for (int k = 0; k < n; ++k) s = fma(__ldg(S.Bt + j + m * k), vt[k], s);
        hu[e] = s;
    }
    for (int e = c.tid; e < T * n; e += c.nt) {
        const int jm1 = e / n, k = e - jm1 * n, j = jm1 + 1;      // x_j
        double acc = v[jm1 * n + k];
        if (j <= T - 1) {
            const double *vj = v + (size_t)j * n;
            double s = 0.0;
            for (int kk = 0; kk < n; ++kk) s = fma(__ldg(S.A1t + k + n * kk), vj[kk], s);
            acc -= s;
        }
        if (j <= T - 2 && S.has_a2) {
            const double *vj = v + (size_t)(j + 1) * n;
            double s = 0.0;
            for (int kk = 0; kk < n; ++kk) s = fma(__ldg(S.A2t + k + n * kk), vj[kk], s);
            acc -= s;
        }
        if (j == T && c.NB > T) acc += v[T * n + k];
        hx[e] = acc;
    }
}

// sum of squares of the full residual [r_d; r_p] at (u, x) with dual images hu + t*hdu, hx + t*hdx
static __device__ double resid_sumsq(const Ctx &c, const double *u, const double *x, const double *hu, const double *hdu,
                              const double *hx, const double *hdx, double t, const double *dbar, const double *rp,
                              double *red, double *rdu_out, double *rdx_out, double *sumsq_p)
{
    const int n = c.n, m = c.m, T = c.T;
    const DevSys &S = c.S;
    double s = 0.0;
    for (int e = c.tid; e < T * m; e += c.nt) {
        const int j = e % m;
        const double h = hdu ? __fma_rn(t, hdu[e], hu[e]) : hu[e];
        const double r = rdu_expr(__ldg(S.r2 + j), __ldg(S.rl + j), u[e], h, dbar[e]);
        if (rdu_out) rdu_out[e] = r;
        s = fma(r, r, s);
    }
    for (int e = c.tid; e < T * n; e += c.nt) {
        const int jm1 = e / n, k = e - jm1 * n;
        const bool last = (jm1 == T - 1);
        const double h = hdx ? __fma_rn(t, hdx[e], hx[e]) : hx[e];
        const double r = rdx_expr(__ldg((last ? S.q2f : S.q2) + k), __ldg((last ? S.qfl : S.ql) + k), x[e], h);
        if (rdx_out) rdx_out[e] = r;
        s = fma(r, r, s);
    }
    double sp = 0.0;
    for (int e = c.tid; e < c.NB * n; e += c.nt) sp = fma(rp[e], rp[e], sp);
    const double tot_p = block_sum(sp, red);
    const double tot_d = block_sum(s, red);
    if (sumsq_p) *sumsq_p = tot_p;
    return tot_d + tot_p;
}

// ---- single-warp dense helpers on shared-memory blocks (row-major, leading dimension ld) ----
// in-place lower Cholesky; returns 0 or (failing column + 1)
static __device__ int warp_potrf(double *Sm, int n, int ld, int lane)
{
    int info = 0;
    for (int k = 0; k < n; ++k) {
        const double d = Sm[k * ld + k];
        if (!(d > 0.0)) { info = k + 1; break; }           // uniform across the warp
        const double dk = sqrt(d);
        __syncwarp();
        for (int r = k + lane; r < n; r += 32) Sm[r * ld + k] = (r == k) ? dk : Sm[r * ld + k] / dk;
        __syncwarp();
        for (int r = k + 1 + lane; r < n; r += 32) {
            const double l = Sm[r * ld + k];
            for (int cc = k + 1; cc <= r; ++cc) Sm[r * ld + cc] = fma(-l, Sm[cc * ld + k], Sm[r * ld + cc]);
        }
        __syncwarp();
    }
    return info;
}


} // namespace fmpc_dev
