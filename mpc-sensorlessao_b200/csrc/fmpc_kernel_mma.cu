// fastMPC batched Newton solve -- DMMA path for n <= 32 (sm_100a, fp64).
//
// Same algorithm and the same per-instance control flow as the generic kernel in fmpc_kernels.cu
// (one persistent CTA per MPC instance; inf_newton_solver.m:10-41 on the block structure), but every
// dense contraction runs on the FP64 tensor pipe (mma.sync.aligned.m8n8k4.f64 = DMMA, measured
// 37.2 TFLOP/s on B200 = the FP64 roofline):
//   * B diag(w_t) B' for ALL stages at once as ONE GEMM  G (npairs x m) * W (m x T), G[p][j] = B(r,j) B(c,j)
//     precomputed on the host: no 28 -> 32 padding waste, symmetric half only;
//   * band-2 block Cholesky of the Schur complement: syrk / gemm updates as 8 x 8 x 4 DMMA tiles on
//     shared-memory blocks (row-major, leading dimension = 4 mod 8 doubles: conflict-free fragment loads);
//   * the diagonal block is factored by ONE warp with rows in registers (right-looking, column broadcast
//     through a 32-double shared vector), then inverted explicitly (column per lane), so that both
//     triangular solves with n right-hand sides become DMMA products with inv(L)' and the forward /
//     backward substitutions become GEMVs.
#include <cuda_runtime.h>
#include <math.h>
#include "fmpc_internal.h"
#include "fmpc_device.cuh"

using namespace fmpc_dev;

namespace {

constexpr int NTHREADS = 128;
constexpr int NWARPS = NTHREADS / 32;
constexpr int TCHUNK = 24;          // stages per pass of the G * W GEMM (3 DMMA column tiles)

__device__ __forceinline__ void dmma(double &c0, double &c1, const double a, const double b)
{
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

struct Geom {
    int n, KP, ld, RP, nt, ks, lds, mp, ldp;
    __host__ __device__ static Geom make(int n, int m)
    {
        Geom g;
        g.n = n;
        g.KP = (n + 3) & ~3;
        g.ld = (g.KP % 8 == 4) ? g.KP : g.KP + 4;
        g.RP = (n + 7) & ~7;
        g.nt = g.RP / 8;
        g.ks = g.KP / 4;
        g.lds = g.RP + 1;
        g.mp = (m + 3) & ~3;
        g.ldp = (g.mp % 8 == 4) ? g.mp : g.mp + 4;
        return g;
    }
    __host__ __device__ size_t blk() const { return (size_t)RP * ld; }
    // doubles of dynamic shared memory
    __host__ __device__ size_t smem_doubles() const
    {
        size_t ops = 6 * blk();
        const size_t pch = (size_t)TCHUNK * ldp;
        if (pch > ops) ops = pch;
        return ops + (size_t)RP * lds + 32 * 3 + 64 + 32 + 34;
    }
};

// C(8x8 tile) += A(rows ra.., k) * B(rows rb.., k)'   over ksteps k-steps of 4, operands in shared memory
__device__ __forceinline__ void tile_nt(double &c0, double &c1, const double *A, const double *Bm, int ksteps)
{
#pragma unroll 4
    for (int k = 0; k < ksteps; ++k) dmma(c0, c1, A[4 * k], Bm[4 * k]);
}

// ---------------------------------------------------------------------------------------------
// One warp: in-register Cholesky of the n x n block in bS (lower triangle, leading dimension lds),
// then explicit inverse of the factor.  Outputs: bLinv (RP x ld, zero padded, operand layout) and
// gLinv (n x n row-major, global scratch for the backward pass).  Returns 0 or failing column + 1.
// ---------------------------------------------------------------------------------------------
__device__ __noinline__ int warp_potrf_inverse(double *bS, int lds, double *bLinv, int ld, int KP, double *gLinv, int n,
                                               double *colbuf, double *rsv, int lane)
{
    double a[32];
    const int r = lane;
#pragma unroll
    for (int c = 0; c < 32; ++c) a[c] = (r < n && c <= r) ? bS[r * lds + c] : ((c == r) ? 1.0 : 0.0);
    int info = 0;
#pragma unroll
    for (int k = 0; k < 32; ++k) {
        if (k < n) {
            double *cb = colbuf + (k & 1) * 32;
            cb[lane] = a[k];                               // unscaled column k (valid for lanes >= k)
            __syncwarp();
            const double d = cb[k];
            if (!(d > 0.0) || !isfinite(d)) { if (!info) info = k + 1; }
            const double dinv = 1.0 / d;
            const double rs = rsqrt(d);
            const double t = a[k] * dinv;
#pragma unroll
            for (int c = k + 1; c < 32; ++c)
                if (c < n) a[c] = fma(-t, cb[c], a[c]);   // rank-1 update; entries above the diagonal are never read
            a[k] *= rs;                                    // L(r,k)
            if (lane == 0) rsv[k] = rs;                    // 1 / L(k,k)
        }
    }
    if (info) return info;
    // L rows -> bS (row r by lane r; lds odd: conflict-free), then inverse column j by lane j
#pragma unroll
    for (int c = 0; c < 32; ++c)
        if (c < n && r < n && c <= r) bS[r * lds + c] = a[c];
    __syncwarp();
    double x[32];
    const int j = lane;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        x[i] = 0.0;
        if (i < n) {
            double s0 = 0.0, s1 = 0.0;
#pragma unroll
            for (int k = 0; k + 1 < i; k += 2) {
                s0 = fma(bS[i * lds + k], x[k], s0);
                s1 = fma(bS[i * lds + k + 1], x[k + 1], s1);
            }
            if (i & 1) s0 = fma(bS[i * lds + i - 1], x[i - 1], s0);
            const double rs = rsv[i];
            x[i] = (i < j) ? 0.0 : ((i == j) ? rs : -(s0 + s1) * rs);
        }
    }
    const bool live = (j < n);
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        if (i < n) {
            const double v = live ? x[i] : 0.0;
            if (j < KP) bLinv[i * ld + j] = v;
            if (live) gLinv[i * n + j] = v;
        }
    }
    return 0;
}

} // namespace

// =============================================================================================
__global__ void __launch_bounds__(NTHREADS, 4) fmpc_solve_kernel_mma(const DevSys S, const StepArgs A)
{
    extern __shared__ double smem[];
    const int n = S.n, m = S.m, T = S.T;
    const int NB = T + (A.has_xf ? 1 : 0);
    const Geom G = Geom::make(n, m);
    const int ld = G.ld, KP = G.KP, RP = G.RP, nt = G.nt, ks = G.ks, lds = G.lds;
    const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, wid = tid >> 5;
    const int gq = lane >> 2, q = lane & 3;                 // DMMA fragment coordinates
    const Ctx c{S, n, m, T, NB, tid, nthr};
    const WsLayout L = WsLayout::make(n, m, T);

    double *ops = smem;                                     // 6 operand blocks | W chunk of the G*W GEMM
    size_t opsz = 6 * G.blk();
    if ((size_t)TCHUNK * G.ldp > opsz) opsz = (size_t)TCHUNK * G.ldp;
    double *bS = smem + opsz;                               // RP x lds
    double *sm_rhs = bS + (size_t)RP * lds;                 // 32
    double *sm_y1 = sm_rhs + 32, *sm_y2 = sm_y1 + 32;       // y_{i-1}, y_{i-2}
    double *colbuf = sm_y2 + 32;                            // 64
    double *rsv = colbuf + 64;                              // 32
    double *red = rsv + 32;                                 // 34
    __shared__ int s_inst, s_flag;

    double *ws = A.ws + (size_t)blockIdx.x * A.ws_stride;
    double *nu = ws + L.nu, *dnu = ws + L.dnu, *yv = ws + L.yv, *rp = ws + L.rp, *rpt = ws + L.rpt, *bv = ws + L.bv;
    double *hx = ws + L.hx, *hdx = ws + L.hdx, *dx = ws + L.dx, *xt = ws + L.xt, *rdx = ws + L.rdx;
    double *hu = ws + L.hu, *hdu = ws + L.hdu, *du = ws + L.du, *ut = ws + L.ut, *dbar = ws + L.dbar;
    double *pinv = ws + L.pinv, *rdu = ws + L.rdu;
    double *gLi = ws + L.Lf, *gL1 = ws + L.L1, *gL2 = ws + L.L2, *Dsc = ws + L.Dsc;
    const size_t nn = (size_t)n * n;
    const int Mp = S.Mp, mp = S.mp, ldp = G.ldp;

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

        // ---- initial iterate (fast_mpc_init.m:12-26), nu, b (fast_mpc_eq_const.m:39,44,47,68) ----
        if (A.cold) {
            for (int e = tid; e < T * m; e += nthr) { const int j = e % m; u[e] = (S.umin[j] + S.umax[j]) / 2; }
            for (int e = tid; e < T * n; e += nthr) { const int k = e % n; x[e] = (S.xmin[k] + S.xmax[k]) / 2; }
        } else if (A.U0 != A.U || A.X0 != A.X) {
            const double *u0 = A.U0 + (size_t)b * m * T, *xx0 = A.X0 + (size_t)b * n * T;
            for (int e = tid; e < T * m; e += nthr) u[e] = u0[e];
            for (int e = tid; e < T * n; e += nthr) x[e] = xx0[e];
        }
        for (int e = tid; e < NB * n; e += nthr) nu[e] = A.nu0[(size_t)b * NB * n + e];
        for (int e = tid; e < NB * n; e += nthr) {
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
            for (int e = tid; e < T * m; e += nthr) {
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
            // ---- beta = -r_p + C inv(Phi) r_d  (:28-29) ----
            for (int e = tid; e < T * m; e += nthr) du[e] = rdu[e] * pinv[e];
            for (int e = tid; e < T * n; e += nthr) {
                const int jm1 = e / n, k = e - jm1 * n;
                dx[e] = rdx[e] * ((jm1 == T - 1) ? S.qif[k] : S.qi[k]);
            }
            __syncthreads();
            apply_C_minus_b(c, du, dx, rp, yv);             // yv = C p - r_p = beta
            __syncthreads();
            for (int e = tid; e < NB * n; e += nthr) yv[e] = -yv[e];   // rhs of  Y dnu = -beta

            // ---- D_t = B diag(w_t) B' for all stages: Dsc(t, pair) = G * W, DMMA ----
            for (int t0 = 0; t0 < T; t0 += TCHUNK) {
                __syncthreads();
                for (int e = tid; e < TCHUNK * mp; e += nthr) {
                    const int tl = e / mp, j = e - tl * mp, t = t0 + tl;
                    ops[tl * ldp + j] = (t < T && j < m) ? pinv[(size_t)t * m + j] : 0.0;
                }
                __syncthreads();
                const int ntt = min(TCHUNK / 8, (T - t0 + 7) / 8);
                const int kst = mp / 4;
                for (int g = wid; g < Mp / 8; g += NWARPS) {
                    double acc[TCHUNK / 8][2];
#pragma unroll
                    for (int tt = 0; tt < TCHUNK / 8; ++tt) acc[tt][0] = acc[tt][1] = 0.0;
                    const double *Grow = S.G + (size_t)(8 * g + gq) * mp + q;
                    const double *Wrow = ops + gq * ldp + q;
                    for (int k0 = 0; k0 < kst; k0 += 12) {
                        double af[12];
#pragma unroll
                        for (int kk = 0; kk < 12; ++kk) af[kk] = (k0 + kk < kst) ? __ldg(Grow + 4 * (k0 + kk)) : 0.0;
#pragma unroll
                        for (int kk = 0; kk < 12; ++kk) {
                            if (k0 + kk < kst) {
#pragma unroll
                                for (int tt = 0; tt < TCHUNK / 8; ++tt)
                                    if (tt < ntt) dmma(acc[tt][0], acc[tt][1], af[kk], Wrow[(size_t)(8 * tt) * ldp + 4 * (k0 + kk)]);
                            }
                        }
                    }
                    const int pair = 8 * g + gq;
#pragma unroll
                    for (int tt = 0; tt < TCHUNK / 8; ++tt) {
                        const int t = t0 + 8 * tt + 2 * q;
                        if (tt < ntt) {
                            if (t < T) Dsc[(size_t)t * Mp + pair] = acc[tt][0];
                            if (t + 1 < T) Dsc[(size_t)(t + 1) * Mp + pair] = acc[tt][1];
                        }
                    }
                }
            }
            __syncthreads();
            // zero the operand blocks (padding rows / columns must stay exactly zero)
            for (int e = tid; e < (int)(6 * G.blk()); e += nthr) ops[e] = 0.0;
            __syncthreads();

            // ---- band-2 block Cholesky of Y fused with the forward solve (:30-31) ----
            double *bLinv = ops, *bM1 = ops + G.blk(), *bL1p = ops + 2 * G.blk();
            double *bM2 = ops + 3 * G.blk(), *bL2p = ops + 4 * G.blk(), *bL2pp = ops + 5 * G.blk();
            bool fail = false;
            const int nS = nt * (nt + 1) / 2;
            for (int i = 0; i < NB; ++i) {
                const bool has1 = (i + 1 < NB), has2 = (i + 2 < NB) && S.has_a2;
                const bool up1 = (i >= 1), up2 = (i >= 2) && S.has_a2;
                // -- phase 1: S (lower tiles) and M1 tiles --
                const double *Yd = S.ypool + (size_t)S.ydi[i] * nn;
                const int y1 = S.y1i[i];
                const int ntask = nS + (has1 ? nt * nt : 0);
                for (int task = wid; task < ntask; task += NWARPS) {
                    if (task < nS) {
                        int rt = 0;
                        while (task >= (rt + 1) * (rt + 2) / 2) ++rt;
                        const int ct = task - rt * (rt + 1) / 2;
                        const int r = 8 * rt + gq, cc = 8 * ct + 2 * q;
                        double i0 = 0.0, i1 = 0.0;
                        if (r < n) {
                            if (cc <= r) { i0 = Yd[r * n + cc]; if (i < T) i0 += Dsc[(size_t)i * Mp + r * (r + 1) / 2 + cc]; }
                            if (cc + 1 <= r) { i1 = Yd[r * n + cc + 1]; if (i < T) i1 += Dsc[(size_t)i * Mp + r * (r + 1) / 2 + cc + 1]; }
                        }
                        double p0 = 0.0, p1 = 0.0;
                        if (up1) tile_nt(p0, p1, bL1p + (8 * rt + gq) * ld + q, bL1p + (8 * ct + gq) * ld + q, ks);
                        if (up2) tile_nt(p0, p1, bL2pp + (8 * rt + gq) * ld + q, bL2pp + (8 * ct + gq) * ld + q, ks);
                        if (r < n) {
                            if (cc <= r) bS[r * lds + cc] = i0 - p0;
                            if (cc + 1 <= r) bS[r * lds + cc + 1] = i1 - p1;
                        }
                    } else {
                        const int tk = task - nS, rt = tk / nt, ct = tk - rt * nt;
                        const int r = 8 * rt + gq, cc = 8 * ct + 2 * q;
                        double i0 = 0.0, i1 = 0.0;
                        if (r < n && y1 >= 0) {
                            if (cc < n) i0 = S.ypool[(size_t)y1 * nn + r * n + cc];
                            if (cc + 1 < n) i1 = S.ypool[(size_t)y1 * nn + r * n + cc + 1];
                        }
                        double p0 = 0.0, p1 = 0.0;
                        if (up1 && S.has_a2) tile_nt(p0, p1, bL2p + (8 * rt + gq) * ld + q, bL1p + (8 * ct + gq) * ld + q, ks);
                        if (cc < KP) bM1[r * ld + cc] = i0 - p0;
                        if (cc + 1 < KP) bM1[r * ld + cc + 1] = i1 - p1;
                    }
                }
                // rhs_i = yv_i - L1p y_{i-1} - L2pp y_{i-2}   (4 lanes per row)
                {
                    const int r = tid >> 2;
                    if (r < RP) {
                        double s = 0.0;
                        if (up1) for (int k = q; k < KP; k += 4) s = fma(bL1p[r * ld + k], sm_y1[k], s);
                        if (up2) for (int k = q; k < KP; k += 4) s = fma(bL2pp[r * ld + k], sm_y2[k], s);
                        s += __shfl_xor_sync(0xffffffffu, s, 1);
                        s += __shfl_xor_sync(0xffffffffu, s, 2);
                        if (q == 0) sm_rhs[r] = (r < n) ? yv[i * n + r] - s : 0.0;
                    }
                }
                __syncthreads();
                // -- phase 2: factor + invert the diagonal block (warp 0) --
                if (wid == 0) {
                    const int info = warp_potrf_inverse(bS, lds, bLinv, ld, KP, gLi + (size_t)i * nn, n, colbuf, rsv, lane);
                    if (lane == 0) s_flag = info;
                }
                __syncthreads();
                if (s_flag) { fail = true; break; }
                // -- phase 3: L1_i = M1 inv(L)' (in place), L2_i = Y2 inv(L)', y_i = inv(L) rhs --
                for (int rt = wid; rt < nt; rt += NWARPS) {
                    const int r = 8 * rt + gq;
                    if (has1) {
                        double af[8];
#pragma unroll
                        for (int k = 0; k < 8; ++k) af[k] = (k < ks) ? bM1[r * ld + 4 * k + q] : 0.0;
                        __syncwarp();
                        for (int ct = 0; ct < nt; ++ct) {
                            double c0 = 0.0, c1 = 0.0;
                            const int kmax = min(ks, 2 * (ct + 1));          // inv(L) is lower triangular
                            const double *Bp = bLinv + (8 * ct + gq) * ld + q;
#pragma unroll
                            for (int k = 0; k < 8; ++k) if (k < kmax) dmma(c0, c1, af[k], Bp[4 * k]);
                            const int cc = 8 * ct + 2 * q;
                            if (cc < KP) bM1[r * ld + cc] = c0;
                            if (cc + 1 < KP) bM1[r * ld + cc + 1] = c1;
                            if (r < n) {
                                if (cc < n) gL1[(size_t)i * nn + r * n + cc] = c0;
                                if (cc + 1 < n) gL1[(size_t)i * nn + r * n + cc + 1] = c1;
                            }
                        }
                    }
                    if (has2) {
                        const int y2 = S.y2i[i];
                        double af[8];
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const int kk = 4 * k + q;
                            af[k] = (y2 >= 0 && r < n && kk < n) ? __ldg(S.ypool + (size_t)y2 * nn + r * n + kk) : 0.0;
                        }
                        for (int ct = 0; ct < nt; ++ct) {
                            double c0 = 0.0, c1 = 0.0;
                            const int kmax = min(ks, 2 * (ct + 1));
                            const double *Bp = bLinv + (8 * ct + gq) * ld + q;
#pragma unroll
                            for (int k = 0; k < 8; ++k) if (k < kmax) dmma(c0, c1, af[k], Bp[4 * k]);
                            const int cc = 8 * ct + 2 * q;
                            if (cc < KP) bM2[r * ld + cc] = c0;
                            if (cc + 1 < KP) bM2[r * ld + cc + 1] = c1;
                            if (r < n) {
                                if (cc < n) gL2[(size_t)i * nn + r * n + cc] = c0;
                                if (cc + 1 < n) gL2[(size_t)i * nn + r * n + cc + 1] = c1;
                            }
                        }
                    }
                }
                {
                    const int r = tid >> 2;
                    double s = 0.0;
                    if (r < RP) for (int k = q; k < KP; k += 4) s = fma(bLinv[r * ld + k], sm_rhs[k], s);
                    s += __shfl_xor_sync(0xffffffffu, s, 1);
                    s += __shfl_xor_sync(0xffffffffu, s, 2);
                    __syncthreads();                               // everyone is done with sm_y1 / sm_y2 / bL*p of this stage
                    if (r < RP && q == 0) {
                        if (r < n) yv[i * n + r] = s;
                        sm_y2[r] = sm_y1[r];
                        sm_y1[r] = s;
                    }
                }
                {   // rotate: L2pp <- L2p, L2p <- M2, L1p <- M1 ; freed buffers become M1, M2
                    double *oL1p = bL1p, *oL2pp = bL2pp;
                    bL2pp = bL2p; bL2p = bM2; bL1p = bM1;
                    bM1 = oL1p; bM2 = oL2pp;
                }
                __syncthreads();
            }
            if (fail) { status = ST_NOT_PD; break; }

            // ---- backward solve  dnu_i = inv(L_i)' (y_i - L1_i' dnu_{i+1} - L2_i' dnu_{i+2})  (:32) ----
            for (int i = NB - 1; i >= 0; --i) {
                const bool has1 = (i + 1 < NB), has2 = (i + 2 < NB) && S.has_a2;
                {   // v[k] : 4 threads per column k, rows split by q
                    const int k = tid >> 2;
                    double s = 0.0;
                    if (k < n) {
                        if (has1) for (int r = q; r < n; r += 4) s = fma(gL1[(size_t)i * nn + r * n + k], dnu[(i + 1) * n + r], s);
                        if (has2) for (int r = q; r < n; r += 4) s = fma(gL2[(size_t)i * nn + r * n + k], dnu[(i + 2) * n + r], s);
                    }
                    s += __shfl_xor_sync(0xffffffffu, s, 1);
                    s += __shfl_xor_sync(0xffffffffu, s, 2);
                    if (k < n && q == 0) sm_rhs[k] = yv[i * n + k] - s;
                }
                __syncthreads();
                {
                    const int k = tid >> 2;
                    double s = 0.0;
                    if (k < n) for (int r = k + q; r < n; r += 4) s = fma(gLi[(size_t)i * nn + r * n + k], sm_rhs[r], s);
                    s += __shfl_xor_sync(0xffffffffu, s, 1);
                    s += __shfl_xor_sync(0xffffffffu, s, 2);
                    if (k < n && q == 0) dnu[i * n + k] = s;
                }
                __syncthreads();
            }

            // ---- dz = inv(Phi)(-r_d - C' dnu)  (:34-35) ----
            apply_Ct(c, dnu, hdu, hdx);
            __syncthreads();
            for (int e = tid; e < T * m; e += nthr) du[e] = -(rdu[e] - hdu[e]) * pinv[e];
            for (int e = tid; e < T * n; e += nthr) {
                const int jm1 = e / n, k = e - jm1 * n;
                dx[e] = -(rdx[e] + hdx[e]) * ((jm1 == T - 1) ? S.qif[k] : S.qi[k]);
            }
            __syncthreads();

            // ---- backtracking on ||[r_p; r_d]||, d frozen (backtracking_inf_newton.m:2-11) ----
            double t = 1.0;
            int nh = 0;
            for (;;) {
                for (int e = tid; e < T * m; e += nthr) ut[e] = __fma_rn(t, du[e], u[e]);
                for (int e = tid; e < T * n; e += nthr) xt[e] = __fma_rn(t, dx[e], x[e]);
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
            for (int e = tid; e < T * m; e += nthr) { u[e] = ut[e]; hu[e] = __fma_rn(t, hdu[e], hu[e]); }
            for (int e = tid; e < T * n; e += nthr) { x[e] = xt[e]; hx[e] = __fma_rn(t, hdx[e], hx[e]); }
            for (int e = tid; e < NB * n; e += nthr) { nu[e] = __fma_rn(t, dnu[e], nu[e]); rp[e] = rpt[e]; }
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
int fmpc_mma_config(const DevSys &S, int device, SolveLaunchCfg *cfg)
{
    if (S.n > 32) return -1;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return -2;
    const Geom G = Geom::make(S.n, S.m);
    const size_t smem = G.smem_doubles() * sizeof(double);
    if (smem > (size_t)prop.sharedMemPerBlockOptin) return -3;
    if (cudaFuncSetAttribute(fmpc_solve_kernel_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -4;
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fmpc_solve_kernel_mma, NTHREADS, smem) != cudaSuccess || per_sm < 1)
        return -5;
    cfg->grid = prop.multiProcessorCount * per_sm;
    cfg->block = NTHREADS;
    cfg->smem = smem;
    cfg->use_mma = 1;
    return 0;
}

void fmpc_launch_solve_mma(const DevSys &S, const StepArgs &A, const SolveLaunchCfg &cfg, void *stream)
{
    int grid = cfg.grid < A.nbatch ? cfg.grid : A.nbatch;
    if (grid < 1) grid = 1;
    fmpc_solve_kernel_mma<<<grid, cfg.block, cfg.smem, (cudaStream_t)stream>>>(S, A);
}
