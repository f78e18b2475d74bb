// fastMPC batched Newton solve -- general-structure kernel (sm_100a, fp64).
//
// Covers the reference inputs the block-banded fast kernels do not:
//   * VAR_1 ramp-rate rows (VAR_1/fast_mpc_ineq_const.m:58-79): Phi couples u_{t-1} and u_t, so with a diagonal R
//     the u part of Phi is, per actuator j, a T x T scalar tridiagonal matrix; its inverse is dense across stages
//     and the Schur complement Y = C inv(Phi) C' loses its band (dense (T n) x (T n));
//   * the literal VAR_1 column placement of the second block row of C (VAR_1/fast_mpc_eq_const.m:34-37):
//     C is kept as the reference builds it, one dense row window per block row, and every C / C' application and
//     the Schur assembly work on those windows (nothing assumes [-A -B I] sits on stage boundaries);
//   * dense (non-diagonal) SPD Q / Qf: inv(2Q) is a dense constant, the x part of Y is precomputed on the host.
// One persistent CTA per MPC instance (dynamic instance counter), same Newton / early-exit / back-tracking control
// flow as the other kernels (inf_newton_solver.m:10-41, backtracking_inf_newton.m:2-11).  Per iteration:
//   barrier terms -> residuals -> per-actuator tridiagonal LDL' + explicit inverse M_j (T x T) ->
//   Y = Yx + sum Cu_a diag(M[t_a,t_b,:]) Cu_b'  -> dense blocked Cholesky (n x n diagonal blocks in shared memory,
//   panel in shared memory, 4 x 4 register tiles for the trailing update) fused with the forward substitution ->
//   backward substitution -> dz = inv(Phi)(-r_d - C' dnu) -> residual-norm back-tracking.
#include <cuda_runtime.h>
#include <math.h>
#include <cstring>
#include <vector>
#include "../../include/fmpc.h"
#include "fmpc_internal.h"
#include "fmpc_device.cuh"

using namespace fmpc_dev;

namespace {

constexpr int GEN_THREADS = 256;

__device__ __forceinline__ void dmma_gen(double &c0, double &c1, const double a, const double b)
{
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

struct GenWs {      // per-CTA scratch layout (doubles)
    size_t z, zt, dz, rd, h, hd, pd, nu, dnu, rp, rpt, bv, yv, tdiag, toff, minv, E, Y, total;
    __host__ __device__ static GenWs make(int n, int m, int T, int ramp, int dense_r = 0)
    {
        GenWs L;
        const size_t N = (size_t)T * (n + m), NE = (size_t)(T + 1) * n, tm = (size_t)T * m;
        size_t o = 0;
        L.z = o; o += N; L.zt = o; o += N; L.dz = o; o += N; L.rd = o; o += N; L.h = o; o += N; L.hd = o; o += N; L.pd = o; o += N;
        L.nu = o; o += NE; L.dnu = o; o += NE; L.rp = o; o += NE; L.rpt = o; o += NE; L.bv = o; o += NE; L.yv = o; o += NE;
        L.tdiag = o; o += tm; L.toff = o; o += tm;
        L.minv = o; o += dense_r ? tm * m : (ramp ? tm * T : tm);       // dense R: inv(Phi_uu) of every stage, m x m each
        L.E = o; o += dense_r ? (size_t)(T + 1) * 4 * n * m : 0;          // dense R: C_u inv(Phi_uu) per (block row, u block)
        o = (o + 15) & ~(size_t)15;
        L.Y = o; o += NE * NE;
        L.total = (o + 15) & ~(size_t)15;
        return L;
    }
};

// out = C v (- bv)   : one warp per scalar row, lanes stride over the row window (coalesced)
__device__ void gen_apply_C(const GenSys &G, int NB, const double *v, const double *bv, double *out)
{
    const int n = G.n, lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int row = wid; row < NB * n; row += nw) {
        const int i = row / n, k = row - i * n, len = G.cw_len[i];
        const double *c = G.cw + G.cw_ptr[i] + (size_t)k * len;
        const double *vv = v + G.cw_off[i];
        double s = 0.0;
        for (int j = lane; j < len; j += 32) s = fma(__ldg(c + j), vv[j], s);
        s = warp_sum(s);
        if (lane == 0) out[row] = bv ? s - bv[row] : s;
    }
}

// out = C' v   : one thread per column of C
__device__ void gen_apply_Ct(const GenSys &G, int NB, const double *v, double *out)
{
    const int n = G.n, N = G.N;
    for (int c = threadIdx.x; c < N; c += blockDim.x) {
        double s = 0.0;
        for (int i = 0; i < NB; ++i) {
            const int off = G.cw_off[i], len = G.cw_len[i];
            if (c >= off && c < off + len) {
                const double *cc = G.cw + G.cw_ptr[i] + (c - off);
                const double *vi = v + (size_t)i * n;
                for (int k = 0; k < n; ++k) s = fma(__ldg(cc + (size_t)k * len), vi[k], s);
            }
        }
        out[c] = s;
    }
}

// One element of r_d = 2 H z + g + k P'd + C'nu at column c of z (the single place this expression lives: the
// line search compares norms computed by identical arithmetic)
__device__ __forceinline__ double gen_rd_elem(const DevSys &S, const GenSys &G, int c, const double *z, double hv, const double *pd)
{
    const int n = G.n, m = G.m, st = n + m;
    const int t = c / st, j = c - t * st;
    if (j < m) {
        if (G.dense_r) {                                       // 2 R u + r : dense row of R + R'
            const double *R2 = G.R2 + (size_t)j * m, *u = z + (size_t)t * st;
            double s = 0.0;
            for (int jj = 0; jj < m; ++jj) s = fma(__ldg(R2 + jj), u[jj], s);
            return __dadd_rn(__dadd_rn(__dadd_rn(s, S.rl[j]), hv), pd[c]);
        }
        return __dadd_rn(__dadd_rn(__fma_rn(S.r2[j], z[c], S.rl[j]), hv), pd[c]);
    }
    const int k = j - m;
    const bool last = (t == G.T - 1);
    const double *Q2 = (last ? G.Q2f : G.Q2) + (size_t)k * n;
    const double *x = z + (size_t)t * st + m;
    double s = 0.0;
    for (int kk = 0; kk < n; ++kk) s = fma(__ldg(Q2 + kk), x[kk], s);
    return __dadd_rn(__dadd_rn(s, (last ? S.qfl : S.ql)[k]), hv);
}

// p = inv(Phi) v : u part through the explicit per-actuator inverse (ramp) or the diagonal, x part through inv(2Q)
__device__ void gen_apply_phi_inv(const DevSys &S, const GenSys &G, const double *minv, const double *v, double *p, double sign)
{
    const int n = G.n, m = G.m, T = G.T, st = n + m;
    for (int e = threadIdx.x; e < T * m; e += blockDim.x) {
        const int t = e / m, j = e - t * m;
        double s;
        if (G.dense_r) {
            const double *M = minv + ((size_t)t * m + j) * m, *vu = v + (size_t)t * st;
            s = 0.0;
            for (int jj = 0; jj < m; ++jj) s = fma(M[jj], vu[jj], s);
        } else if (G.ramp) {
            s = 0.0;
            for (int tp = 0; tp < T; ++tp) s = fma(minv[((size_t)t * T + tp) * m + j], v[(size_t)tp * st + j], s);
        } else {
            s = minv[e] * v[(size_t)t * st + j];
        }
        p[(size_t)t * st + j] = sign * s;
    }
    for (int e = threadIdx.x; e < T * n; e += blockDim.x) {
        const int t = e / n, k = e - t * n;
        const double *Qi = ((t == T - 1) ? G.Qif : G.Qi) + (size_t)k * n;
        const double *vx = v + (size_t)t * st + m;
        double s = 0.0;
        for (int kk = 0; kk < n; ++kk) s = fma(__ldg(Qi + kk), vx[kk], s);
        p[(size_t)t * st + m + k] = sign * s;
    }
}

__global__ void __launch_bounds__(GEN_THREADS) fmpc_solve_kernel_gen(const DevSys S, const GenSys G, const StepArgs A)
{
    extern __shared__ double smem[];
    const int n = G.n, m = G.m, T = G.T, N = G.N, st = n + m;
    const int NB = T + (A.has_xf ? 1 : 0), NE = NB * n;
    const int ld = n | 1;
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5;
    const GenWs L = GenWs::make(n, m, T, G.ramp, G.dense_r);

    double *bS = smem;                                  // n x ld : diagonal block
    double *sm_y = bS + (size_t)n * ld;                 // n
    double *sm_vec = sm_y + n;                          // n
    double *sm_part = sm_vec + n;                       // nt
    double *red = sm_part + nt;                         // 34
    double *panel = red + 34;                           // panel_rows x ld
    __shared__ int s_inst, s_flag;

    double *ws = A.ws + (size_t)blockIdx.x * A.ws_stride;
    double *z = ws + L.z, *zt = ws + L.zt, *dz = ws + L.dz, *rd = ws + L.rd, *h = ws + L.h, *hd = ws + L.hd, *pd = ws + L.pd;
    double *nu = ws + L.nu, *dnu = ws + L.dnu, *rp = ws + L.rp, *rpt = ws + L.rpt, *bv = ws + L.bv, *yv = ws + L.yv;
    double *tdiag = ws + L.tdiag, *toff = ws + L.toff, *minv = ws + L.minv, *Y = ws + L.Y, *Eb = ws + L.E;
    const size_t ldy = (size_t)NE;

    for (;;) {
        __syncthreads();
        if (tid == 0) s_inst = (int)atomicAdd(A.counter, 1u);
        __syncthreads();
        const int b = s_inst;
        if (b >= A.nbatch) break;

        const double *x0 = A.x0 + (size_t)b * n;
        const double *x0p = A.x0_pre ? A.x0_pre + (size_t)b * n : nullptr;
        const double *uprev = A.u_prev ? A.u_prev + (size_t)b * m : nullptr;

        // ---- initial iterate (fast_mpc_init.m:12-26), dual start, b ----
        for (int c = tid; c < N; c += nt) {
            const int t = c / st, j = c - t * st;
            double v;
            if (A.cold) v = (j < m) ? (S.umin[j] + S.umax[j]) / 2 : (S.xmin[j - m] + S.xmax[j - m]) / 2;
            else v = (j < m) ? A.U0[(size_t)b * m * T + (size_t)t * m + j] : A.X0[(size_t)b * n * T + (size_t)t * n + (j - m)];
            z[c] = v;
        }
        for (int e = tid; e < NE; e += nt) nu[e] = A.nu0[(size_t)b * NE + e];
        for (int e = tid; e < NE; e += nt) {
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
        gen_apply_Ct(G, NB, nu, h);
        gen_apply_C(G, NB, z, bv, rp);
        __syncthreads();

        int status = ST_OK, iters = 0;
        for (int it = 0; it < A.niters; ++it) {
            // ---- barrier terms: s = h - P z, d = 1./s, Phi_uu = 2R + k P'diag(d.^2)P  (inf_newton_KKT_H.m:3-13) ----
            for (int e = tid; e < T * m; e += nt) {
                const int t = e / m, j = e - t * m;
                const size_t c = (size_t)t * st + j;
                const double u = z[c];
                const double dp = 1.0 / (S.umax[j] - u), dm = 1.0 / (-S.umin[j] + u);
                double g = dp - dm, dd = dp * dp + dm * dm, off = 0.0;
                if (G.ramp) {
                    // ramp rows of stage t: [I;-I] u_0 <= [u_prev+du_max; -u_prev-du_min] (t = 0),
                    // [-I 0 I; I 0 -I][u_{t-1}; x_t; u_t] <= [du_max; -du_min] (t >= 1)   (VAR_1/fast_mpc_ineq_const.m:62-76)
                    double su, sl;
                    if (t == 0) { su = (uprev[j] + G.dumax[j]) - u; sl = (-uprev[j] - G.dumin[j]) + u; }
                    else { const double pz = u - z[c - st]; su = G.dumax[j] - pz; sl = -G.dumin[j] + pz; }
                    const double ru = 1.0 / su, rl = 1.0 / sl;
                    g += ru - rl;
                    dd += ru * ru + rl * rl;
                    if (t + 1 < T) {
                        const double pzn = z[c + st] - u;
                        const double run = 1.0 / (G.dumax[j] - pzn), rln = 1.0 / (-G.dumin[j] + pzn);
                        g -= run - rln;
                        const double wn = run * run + rln * rln;
                        dd += wn;
                        off = -A.kappa * wn;
                    }
                }
                pd[c] = A.kappa * g;
                tdiag[e] = (G.dense_r ? 0.0 : S.r2[j]) + A.kappa * dd;      // dense R: the barrier part only, R + R' is added below
                toff[e] = off;
            }
            for (int e = tid; e < T * n; e += nt) { const int t = e / n; pd[(size_t)t * st + m + (e - t * n)] = 0.0; }
            __syncthreads();
            // ---- residuals + early exit (inf_newton_solver.m:12-22) ----
            double sd = 0.0, sp = 0.0;
            for (int c = tid; c < N; c += nt) { const double r = gen_rd_elem(S, G, c, z, h[c], pd); rd[c] = r; sd = fma(r, r, sd); }
            for (int e = tid; e < NE; e += nt) sp = fma(rp[e], rp[e], sp);
            const double ssp = block_sum(sp, red);
            const double ss0 = block_sum(sd, red) + ssp;
            const double nr0 = sqrt(ss0);
            if (!isfinite(nr0)) { status = ST_NONFINITE; break; }
            if (nr0 <= A.tol_r && sqrt(ssp) <= A.tol_p) { status = ST_EARLY_EXIT; break; }

            // ---- inv(Phi_uu): per-actuator tridiagonal LDL' (in place: tdiag <- d, toff <- l) and explicit inverse ----
            if (tid == 0) s_flag = 0;
            __syncthreads();
            if (G.dense_r) {
                // dense R (fast_mpc_objective.m:20-21 takes any square R): Phi_uu of stage t = (R + R') + k diag(d.^2) is a dense
                // SPD m x m matrix.  Packed lower Cholesky in shared memory (the panel area is free here), explicit inverse of
                // the factor column by column, inv(Phi_uu) = inv(L)' inv(L) to the scratch; then E = C_u inv(Phi_uu) per
                // (block row, u block) for the Schur assembly.
                double *Lp = panel, *Xp = panel + (size_t)m * (m + 1) / 2;
                auto ix = [](int r, int c) { return (size_t)r * (r + 1) / 2 + c; };
                for (int t = 0; t < T; ++t) {
                    __syncthreads();
                    for (int e = tid; e < m * m; e += nt) {
                        const int r = e / m, c = e - r * m;
                        if (c <= r) Lp[ix(r, c)] = __ldg(G.R2 + e) + (r == c ? tdiag[(size_t)t * m + r] : 0.0);
                    }
                    __syncthreads();
                    for (int k = 0; k < m; ++k) {
                        const double akk = Lp[ix(k, k)];
                        __syncthreads();
                        if (!(akk > 0.0)) { if (tid == 0) s_flag = 1; break; }
                        const double dk = sqrt(akk);
                        if (tid == 0) Lp[ix(k, k)] = dk;
                        for (int r = k + 1 + tid; r < m; r += nt) Lp[ix(r, k)] /= dk;
                        __syncthreads();
                        const int rem = m - 1 - k;                      // trailing lower triangle of size rem
                        for (int e = tid; e < rem * (rem + 1) / 2; e += nt) {
                            int rr = (int)((sqrt(8.0 * (double)e + 1.0) - 1.0) * 0.5);
                            while ((rr + 1) * (rr + 2) / 2 <= e) ++rr;
                            while (rr * (rr + 1) / 2 > e) --rr;
                            const int cc = e - rr * (rr + 1) / 2;
                            const int r = k + 1 + rr, c = k + 1 + cc;
                            Lp[ix(r, c)] = fma(-Lp[ix(r, k)], Lp[ix(c, k)], Lp[ix(r, c)]);
                        }
                        __syncthreads();
                    }
                    __syncthreads();
                    if (s_flag) break;
                    for (int c = tid; c < m; c += nt) {                 // column c of inv(L)
                        Xp[ix(c, c)] = 1.0 / Lp[ix(c, c)];
                        for (int r = c + 1; r < m; ++r) {
                            double sacc = 0.0;
                            for (int k = c; k < r; ++k) sacc = fma(Lp[ix(r, k)], Xp[ix(k, c)], sacc);
                            Xp[ix(r, c)] = -sacc / Lp[ix(r, r)];
                        }
                    }
                    __syncthreads();
                    double *Mt = minv + (size_t)t * m * m;
                    for (int e = tid; e < m * m; e += nt) {
                        const int r = e / m, c = e - r * m;
                        if (c > r) continue;
                        double sacc = 0.0;
                        for (int k = r; k < m; ++k) sacc = fma(Xp[ix(k, r)], Xp[ix(k, c)], sacc);
                        Mt[(size_t)r * m + c] = sacc;
                        Mt[(size_t)c * m + r] = sacc;
                    }
                }
                __syncthreads();
                if (!s_flag) {
                    for (int i = 0; i < NB; ++i)
                        for (int a = 0; a < G.ue_cnt[i]; ++a) {
                            const double *Ca = G.cu + G.ue_ptr[4 * i + a];
                            const double *Mt = minv + (size_t)G.ue_t[4 * i + a] * m * m;
                            double *Ea = Eb + ((size_t)i * 4 + a) * n * m;
                            for (int e = tid; e < n * m; e += nt) {
                                const int r = e / m, j = e - r * m;
                                double sacc = 0.0;
                                for (int jj = 0; jj < m; ++jj) sacc = fma(__ldg(Ca + (size_t)r * m + jj), Mt[(size_t)jj * m + j], sacc);
                                Ea[e] = sacc;
                            }
                        }
                }
            } else if (G.ramp) {
                for (int j = tid; j < m; j += nt) {
                    double d = tdiag[j];
                    bool bad = !(d > 0.0);
                    for (int t = 1; t < T; ++t) {
                        const double o = toff[(size_t)(t - 1) * m + j];
                        const double l = o / d;
                        toff[(size_t)(t - 1) * m + j] = l;
                        d = tdiag[(size_t)t * m + j] - l * o;
                        tdiag[(size_t)t * m + j] = d;
                        if (!(d > 0.0)) bad = true;
                    }
                    if (bad) s_flag = 1;
                    for (int tc = 0; tc < T; ++tc) {          // column tc of the inverse: L D L' x = e_tc
                        double y = 1.0;
                        for (int t = 0; t < tc; ++t) minv[((size_t)t * T + tc) * m + j] = 0.0;
                        minv[((size_t)tc * T + tc) * m + j] = 1.0;
                        for (int t = tc + 1; t < T; ++t) { y = -toff[(size_t)(t - 1) * m + j] * y; minv[((size_t)t * T + tc) * m + j] = y; }
                        double x = minv[((size_t)(T - 1) * T + tc) * m + j] / tdiag[(size_t)(T - 1) * m + j];
                        minv[((size_t)(T - 1) * T + tc) * m + j] = x;
                        for (int t = T - 2; t >= 0; --t) {
                            x = minv[((size_t)t * T + tc) * m + j] / tdiag[(size_t)t * m + j] - toff[(size_t)t * m + j] * x;
                            minv[((size_t)t * T + tc) * m + j] = x;
                        }
                    }
                }
            } else {
                for (int e = tid; e < T * m; e += nt) { const double d = tdiag[e]; if (!(d > 0.0)) s_flag = 1; minv[e] = 1.0 / d; }
            }
            __syncthreads();
            if (s_flag) { status = ST_NOT_PD; break; }

            // ---- rhs of  Y dnu = -beta,  beta = -r_p + C inv(Phi) r_d  (:28-29) ----
            gen_apply_phi_inv(S, G, minv, rd, dz, 1.0);
            __syncthreads();
            gen_apply_C(G, NB, dz, nullptr, yv);
            __syncthreads();
            for (int e = tid; e < NE; e += nt) yv[e] = rp[e] - yv[e];

            // ---- Y = Yx + C_u inv(Phi_uu) C_u'  (lower block triangle, full diagonal blocks) ----
            // One warp task per (block pair, 8 x 8 tile): the tile of  sum_ab C_a diag(cv_ab) C_b'  is a chain of FP64 tensor-pipe
            // products over the m actuators (A = C_a[r][j] cv[j] scaled on the fly, B = C_b[c][j], both L1-resident).
            {
                const int lane = tid & 31, wid = tid >> 5, nw = nt >> 5, gq = lane >> 2, q = lane & 3;
                const int ntl = (n + 7) / 8, npair = NB * (NB + 1) / 2, ntask = npair * ntl * ntl;
                for (int task = wid; task < ntask; task += nw) {
                    const int pr = task / (ntl * ntl), tl = task - pr * ntl * ntl, rt = tl / ntl, ct = tl - rt * ntl;
                    int i = 0;
                    while ((i + 1) * (i + 2) / 2 <= pr) ++i;
                    const int k = pr - i * (i + 1) / 2;
                    const int ra = 8 * rt + gq, cb = 8 * ct + gq;             // A-operand row / B-operand row of this lane
                    double c0 = 0.0, c1 = 0.0;
                    for (int a = 0; a < G.ue_cnt[i]; ++a) {
                        const int ta = G.ue_t[4 * i + a];
                        const double *Ca = G.dense_r ? Eb + ((size_t)i * 4 + a) * n * m + (size_t)min(ra, n - 1) * m
                                                     : G.cu + G.ue_ptr[4 * i + a] + (size_t)min(ra, n - 1) * m;
                        for (int bb = 0; bb < G.ue_cnt[k]; ++bb) {
                            const int tb = G.ue_t[4 * k + bb];
                            if (!G.ramp && ta != tb) continue;
                            // dense R: the A operand is a row of E = C_a inv(Phi_uu) already; `ones` keeps one code path
                            const double *cv = G.dense_r ? nullptr : (G.ramp ? minv + ((size_t)ta * T + tb) * m : minv + (size_t)ta * m);
                            const double *Cb = G.cu + G.ue_ptr[4 * k + bb] + (size_t)min(cb, n - 1) * m;
                            double e0 = 0.0, e1 = 0.0;
                            int jb = 0;                                       // warp-uniform loop bounds (mma.sync needs all lanes)
                            for (; jb + 8 <= m; jb += 8) {                    // two accumulation chains
                                const int j = jb + q;
                                dmma_gen(c0, c1, cv ? __ldg(Ca + j) * cv[j] : Ca[j], __ldg(Cb + j));
                                dmma_gen(e0, e1, cv ? __ldg(Ca + j + 4) * cv[j + 4] : Ca[j + 4], __ldg(Cb + j + 4));
                            }
                            for (; jb < m; jb += 4) {                         // remaining k-steps, columns >= m contribute zeros
                                const int j = jb + q;
                                const bool ok = j < m;
                                dmma_gen(c0, c1, ok ? (cv ? __ldg(Ca + j) * cv[j] : Ca[j]) : 0.0, ok ? __ldg(Cb + j) : 0.0);
                            }
                            c0 += e0; c1 += e1;
                        }
                    }
                    const int r = 8 * rt + gq, c = 8 * ct + 2 * q;
                    if (r < n) {
                        const double *yx = G.Yx + ((size_t)i * n + r) * G.ldyx + (size_t)k * n;
                        double *yo = Y + ((size_t)i * n + r) * ldy + (size_t)k * n;
                        if (c < n) yo[c] = __ldg(yx + c) + c0;
                        if (c + 1 < n) yo[c + 1] = __ldg(yx + c + 1) + c1;
                    }
                }
            }
            __syncthreads();

            // ---- dense blocked Cholesky of Y fused with the forward substitution (:30-31) ----
            bool fail = false;
            for (int K = 0; K < NB; ++K) {
                const int r0 = (K + 1) * n, R = NE - r0;          // trailing rows
                double *Ykk = Y + ((size_t)K * n) * ldy + (size_t)K * n;
                for (int e = tid; e < n * n; e += nt) { const int r = e / n, c = e - r * n; bS[r * ld + c] = Ykk[(size_t)r * ldy + c]; }
                for (int k = tid; k < n; k += nt) sm_y[k] = yv[K * n + k];
                __syncthreads();
                if (wid == 0) {
                    const int info = warp_potrf(bS, n, ld, lane);
                    if (lane == 0) s_flag = info;
                }
                __syncthreads();
                if (s_flag) { fail = true; break; }
                // y_K = inv(L_KK) rhs_K by the last warp, while the others write the factor back
                if (wid == (nt >> 5) - 1) {
                    for (int j = 0; j < n; ++j) {
                        __syncwarp();
                        const double yj = sm_y[j] / bS[j * ld + j];
                        __syncwarp();
                        if (lane == 0) sm_y[j] = yj;
                        for (int k = j + 1 + lane; k < n; k += 32) sm_y[k] = fma(-bS[k * ld + j], yj, sm_y[k]);
                    }
                } else {
                    for (int e = tid; e < n * n; e += nt - 32) { const int r = e / n, c = e - r * n; Ykk[(size_t)r * ldy + c] = (c <= r) ? bS[r * ld + c] : 0.0; }
                }
                __syncthreads();
                for (int k = tid; k < n; k += nt) yv[K * n + k] = sm_y[k];
                // panel: rows r0.. of block column K  <-  row * inv(L_KK)'  ; chunks of panel_rows rows through shared memory
                const bool whole = (R <= G.panel_rows);
                for (int c0 = 0; c0 < R; c0 += G.panel_rows) {
                    const int rc = min(G.panel_rows, R - c0);
                    double *Yp = Y + ((size_t)(r0 + c0)) * ldy + (size_t)K * n;
                    __syncthreads();
                    for (int e = tid; e < rc * n; e += nt) { const int r = e / n, c = e - r * n; panel[r * ld + c] = Yp[(size_t)r * ldy + c]; }
                    __syncthreads();
                    for (int r = tid; r < rc; r += nt) {
                        double *row = panel + r * ld;
                        double dot = 0.0;
                        for (int j = 0; j < n; ++j) {
                            double s = row[j];
                            for (int k = 0; k < j; ++k) s = fma(-row[k], bS[j * ld + k], s);
                            s /= bS[j * ld + j];
                            row[j] = s;
                            dot = fma(s, sm_y[j], dot);
                        }
                        yv[r0 + c0 + r] -= dot;                  // forward substitution rides along
                    }
                    __syncthreads();
                    for (int e = tid; e < rc * n; e += nt) { const int r = e / n, c = e - r * n; Yp[(size_t)r * ldy + c] = panel[r * ld + c]; }
                }
                __syncthreads();
                // trailing update  Y[r,c] -= sum_j P[r,j] P[c,j]   (r >= c), 4 x 4 register tiles
                if (R > 0) {
                    const double *P = whole ? panel : (Y + (size_t)r0 * ldy + (size_t)K * n);
                    const size_t ldp = whole ? (size_t)ld : ldy;
                    const int NT4 = (R + 3) / 4;
                    const long long ntiles = (long long)NT4 * (NT4 + 1) / 2;
                    for (long long e = tid; e < ntiles; e += nt) {
                        int ti = (int)((sqrt(8.0 * (double)e + 1.0) - 1.0) * 0.5);
                        while ((long long)(ti + 1) * (ti + 2) / 2 <= e) ++ti;
                        while ((long long)ti * (ti + 1) / 2 > e) --ti;
                        const int tj = (int)(e - (long long)ti * (ti + 1) / 2);
                        const double *pa[4], *pb[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            pa[u] = P + (size_t)min(4 * ti + u, R - 1) * ldp;
                            pb[u] = P + (size_t)min(4 * tj + u, R - 1) * ldp;
                        }
                        double acc[4][4];
#pragma unroll
                        for (int u = 0; u < 4; ++u)
#pragma unroll
                            for (int v = 0; v < 4; ++v) acc[u][v] = 0.0;
                        for (int j = 0; j < n; ++j) {
                            double av[4], bw[4];
#pragma unroll
                            for (int u = 0; u < 4; ++u) { av[u] = pa[u][j]; bw[u] = pb[u][j]; }
#pragma unroll
                            for (int u = 0; u < 4; ++u)
#pragma unroll
                                for (int v = 0; v < 4; ++v) acc[u][v] = fma(av[u], bw[v], acc[u][v]);
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u)
#pragma unroll
                            for (int v = 0; v < 4; ++v) {
                                const int r = 4 * ti + u, c = 4 * tj + v;
                                if (r < R && c <= r) Y[((size_t)(r0 + r)) * ldy + (size_t)(r0 + c)] -= acc[u][v];
                            }
                    }
                }
                __syncthreads();
            }
            if (fail) { status = ST_NOT_PD; break; }

            // ---- backward substitution  L' dnu = y  (:32) ----
            for (int K = NB - 1; K >= 0; --K) {
                const int r0 = (K + 1) * n;
                {   // sm_vec[c] = y_K[c] - sum_{r >= r0} L[r, K n + c] dnu[r] : column dots split over nt / n row groups
                    const int groups = max(1, nt / n);
                    const int c = tid % n, g = tid / n;
                    double s = 0.0;
                    if (g < groups)
                        for (int r = r0 + g; r < NE; r += groups) s = fma(Y[(size_t)r * ldy + (size_t)K * n + c], dnu[r], s);
                    sm_part[tid] = (g < groups) ? s : 0.0;
                    __syncthreads();
                    for (int k = tid; k < n; k += nt) {
                        double acc = yv[K * n + k];
                        for (int gg = 0; gg < groups; ++gg) acc -= sm_part[gg * n + k];
                        sm_vec[k] = acc;
                    }
                }
                __syncthreads();
                if (wid == 0) {
                    const double *Lf = Y + ((size_t)K * n) * ldy + (size_t)K * n;
                    for (int j = n - 1; j >= 0; --j) {
                        __syncwarp();
                        const double xj = sm_vec[j] / Lf[(size_t)j * ldy + j];
                        __syncwarp();
                        if (lane == 0) sm_vec[j] = xj;
                        for (int k = lane; k < j; k += 32) sm_vec[k] = fma(-Lf[(size_t)j * ldy + k], xj, sm_vec[k]);
                    }
                }
                __syncthreads();
                for (int k = tid; k < n; k += nt) dnu[K * n + k] = sm_vec[k];
                __syncthreads();
            }

            // ---- dz = inv(Phi)(-r_d - C' dnu)  (:34-35) ----
            gen_apply_Ct(G, NB, dnu, hd);
            __syncthreads();
            for (int c = tid; c < N; c += nt) zt[c] = rd[c] + hd[c];
            __syncthreads();
            gen_apply_phi_inv(S, G, minv, zt, dz, -1.0);
            __syncthreads();

            // ---- backtracking on ||[r_p; r_d]||, d frozen (backtracking_inf_newton.m:2-11) ----
            double t = 1.0;
            int nh = 0;
            for (;;) {
                for (int c = tid; c < N; c += nt) zt[c] = __fma_rn(t, dz[c], z[c]);
                __syncthreads();
                gen_apply_C(G, NB, zt, bv, rpt);
                double sdt = 0.0;
                for (int c = tid; c < N; c += nt) { const double r = gen_rd_elem(S, G, c, zt, __fma_rn(t, hd[c], h[c]), pd); sdt = fma(r, r, sdt); }
                __syncthreads();
                double spt = 0.0;
                for (int e = tid; e < NE; e += nt) spt = fma(rpt[e], rpt[e], spt);
                const double sspt = block_sum(spt, red);
                const double nrt = sqrt(block_sum(sdt, red) + sspt);
                if (!(nrt > (1.0 - A.alpha * t) * nr0)) break;
                if (t == 0.0) break;
                if (A.ls_max > 0 && nh >= A.ls_max) { status = ST_LS_MAX; break; }
                t *= A.beta;
                ++nh;
                __syncthreads();
            }
            __syncthreads();
            for (int c = tid; c < N; c += nt) { z[c] = zt[c]; h[c] = __fma_rn(t, hd[c], h[c]); }
            for (int e = tid; e < NE; e += nt) { nu[e] = __fma_rn(t, dnu[e], nu[e]); rp[e] = rpt[e]; }
            ++iters;
            __syncthreads();
        }
        __syncthreads();
        // ---- de-interleave (README.md:558-570) ----
        for (int c = tid; c < N; c += nt) {
            const int t = c / st, j = c - t * st;
            if (j < m) A.U[(size_t)b * m * T + (size_t)t * m + j] = z[c];
            else A.X[(size_t)b * n * T + (size_t)t * n + (j - m)] = z[c];
        }
        if (tid == 0) {
            if (A.status) A.status[b] = status;
            if (A.iters) A.iters[b] = iters;
            atomicAdd(A.iters_total, (unsigned long long)iters);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// host side: literal C (as the reference builds it), row windows, u-column blocks, x part of Y
// ---------------------------------------------------------------------------------------------
// dense SPD inverse (row-major n x n) by Cholesky; returns false if not positive definite
bool spd_inverse(const std::vector<double> &A, int n, std::vector<double> &Ainv)
{
    std::vector<double> Lm(A);
    for (int k = 0; k < n; ++k) {
        double d = Lm[(size_t)k * n + k];
        for (int j = 0; j < k; ++j) d -= Lm[(size_t)k * n + j] * Lm[(size_t)k * n + j];
        if (!(d > 0.0)) return false;
        d = std::sqrt(d);
        Lm[(size_t)k * n + k] = d;
        for (int r = k + 1; r < n; ++r) {
            double s = Lm[(size_t)r * n + k];
            for (int j = 0; j < k; ++j) s -= Lm[(size_t)r * n + j] * Lm[(size_t)k * n + j];
            Lm[(size_t)r * n + k] = s / d;
        }
    }
    // inv(L) column by column, then inv(A) = inv(L)' inv(L)
    std::vector<double> Li((size_t)n * n, 0.0);
    for (int c = 0; c < n; ++c) {
        Li[(size_t)c * n + c] = 1.0 / Lm[(size_t)c * n + c];
        for (int r = c + 1; r < n; ++r) {
            double s = 0.0;
            for (int j = c; j < r; ++j) s -= Lm[(size_t)r * n + j] * Li[(size_t)j * n + c];
            Li[(size_t)r * n + c] = s / Lm[(size_t)r * n + r];
        }
    }
    Ainv.assign((size_t)n * n, 0.0);
    for (int r = 0; r < n; ++r)
        for (int c = 0; c <= r; ++c) {
            double s = 0.0;
            for (int k = r; k < n; ++k) s += Li[(size_t)k * n + r] * Li[(size_t)k * n + c];
            Ainv[(size_t)r * n + c] = Ainv[(size_t)c * n + r] = s;
        }
    return true;
}

template <class T> T *gen_upload(std::vector<void *> &allocs, const std::vector<T> &v)
{
    void *p = nullptr;
    if (cudaMalloc(&p, v.size() * sizeof(T) + 16) != cudaSuccess) return nullptr;
    if (!v.empty() && cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice) != cudaSuccess) { cudaFree(p); return nullptr; }
    allocs.push_back(p);
    return (T *)p;
}

size_t gen_fixed_smem_doubles(int n) { return (size_t)n * (n | 1) + 2 * (size_t)n + GEN_THREADS + 34; }

} // namespace

// Builds the device tables of the general kernel.  Returns FMPC_OK, FMPC_ERR_NOT_PD, FMPC_ERR_UNSUPPORTED or FMPC_ERR_CUDA.
int fmpc_gen_create(const fmpc_sys *s, int device, GenSys *out, std::vector<void *> &allocs, SolveLaunchCfg *cfg)
{
    const int n = s->n, m = s->m, T = s->T, st = n + m, N = T * st, NBm = T + 1, NE = NBm * n;
    const bool a2 = (s->var_order == 2);
    const bool bug = (s->var_order == 1) && s->var1_literal_bug;
    auto Bm = [&](int r, int c) { return s->B[(size_t)c * n + r]; };
    auto A1m = [&](int r, int c) { return s->A1[(size_t)c * n + r]; };
    auto A2m = [&](int r, int c) { return s->A2[(size_t)c * n + r]; };

    // ---- C exactly as the reference writes it, with the terminal block row appended (row-major NE x N) ----
    std::vector<double> Cd((size_t)NE * N, 0.0);
    auto put = [&](int row0, int col0, int kind /*0:-A2 1:-A1 2:-B 3:I*/) {
        const int w = (kind == 2) ? m : n;
        for (int r = 0; r < n; ++r)
            for (int c = 0; c < w; ++c) {
                double v;
                if (kind == 0) v = -A2m(r, c); else if (kind == 1) v = -A1m(r, c); else if (kind == 2) v = -Bm(r, c); else v = (r == c) ? 1.0 : 0.0;
                Cd[(size_t)(row0 + r) * N + col0 + c] = v;
            }
    };
    put(0, 0, 2); put(0, m, 3);                                       // C(1:n,1:m+n) = [-B I]
    for (int i = 1; i < T; ++i) {
        if (a2) {                                                     // VAR_2/fast_mpc_eq_const.m:41-49
            if (i == 1) { put(n, m, 1); put(n, m + n, 2); put(n, m + n + m, 3); }
            else { const int c0 = m + st * (i - 2); put(n * i, c0, 0); put(n * i, c0 + n + m, 1); put(n * i, c0 + 2 * n + m, 2); put(n * i, c0 + 2 * n + 2 * m, 3); }
        } else {                                                      // VAR_1/fast_mpc_eq_const.m:34-43
            int c0 = (i - 1) * st + m;                                // :40 (0-based)
            if (i == 1 && bug) {
                c0 = n - 1;                                           // :36  columns n : 3n+m-1 (1-based)
                if (c0 + 2 * n + m > N) return FMPC_ERR_UNSUPPORTED;  // MATLAB would grow C and then fail on C*z
            }
            put(n * i, c0, 1); put(n * i, c0 + n, 2); put(n * i, c0 + n + m, 3);
        }
    }
    for (int k = 0; k < n; ++k) Cd[(size_t)(T * n + k) * N + (N - n + k)] = 1.0;     // terminal row x_T = xf (:67-71)

    // ---- row windows (de-duplicated by content) ----
    std::vector<double> cw;
    std::vector<int> cw_ptr(NBm), cw_off(NBm), cw_len(NBm);
    for (int i = 0; i < NBm; ++i) {
        int lo = N, hi = 0;
        for (int r = 0; r < n; ++r)
            for (int c = 0; c < N; ++c)
                if (Cd[(size_t)(i * n + r) * N + c] != 0.0) { if (c < lo) lo = c; if (c + 1 > hi) hi = c + 1; }
        if (hi <= lo) { lo = 0; hi = 1; }
        const int len = hi - lo;
        std::vector<double> wmat((size_t)n * len);
        for (int r = 0; r < n; ++r) for (int c = 0; c < len; ++c) wmat[(size_t)r * len + c] = Cd[(size_t)(i * n + r) * N + lo + c];
        int found = -1;
        for (int p = 0; p < i && found < 0; ++p)
            if (cw_len[p] == len && std::memcmp(&cw[cw_ptr[p]], wmat.data(), wmat.size() * 8) == 0) found = cw_ptr[p];
        if (found < 0) { found = (int)cw.size(); cw.insert(cw.end(), wmat.begin(), wmat.end()); }
        cw_ptr[i] = found; cw_off[i] = lo; cw_len[i] = len;
    }
    // ---- u-column blocks per block row: stages whose u columns this row touches ----
    std::vector<double> cu, cut;
    std::vector<int> ue_cnt(NBm, 0), ue_t(4 * NBm, 0), ue_ptr(4 * NBm, 0);
    for (int i = 0; i < NBm; ++i)
        for (int t = 0; t < T; ++t) {
            std::vector<double> blk((size_t)n * m);
            bool nz = false;
            for (int r = 0; r < n; ++r)
                for (int j = 0; j < m; ++j) { const double v = Cd[(size_t)(i * n + r) * N + (size_t)t * st + j]; blk[(size_t)r * m + j] = v; if (v != 0.0) nz = true; }
            if (!nz) continue;
            if (ue_cnt[i] >= 4) return FMPC_ERR_UNSUPPORTED;
            int found = -1;
            for (size_t p = 0; p + blk.size() <= cu.size() && found < 0; p += blk.size())
                if (std::memcmp(&cu[p], blk.data(), blk.size() * 8) == 0) found = (int)p;
            if (found < 0) {
                found = (int)cu.size();
                cu.insert(cu.end(), blk.begin(), blk.end());
                cut.resize(cu.size());
                for (int r = 0; r < n; ++r) for (int j = 0; j < m; ++j) cut[(size_t)found + (size_t)j * n + r] = blk[(size_t)r * m + j];
            }
            ue_t[4 * i + ue_cnt[i]] = t; ue_ptr[4 * i + ue_cnt[i]] = found; ++ue_cnt[i];
        }
    // ---- dense 2Q, 2Qf and their inverses (row-major) ----
    std::vector<double> Q2((size_t)n * n), Q2f((size_t)n * n), Qi, Qif;
    bool qdiag = true;
    for (int r = 0; r < n; ++r)
        for (int c = 0; c < n; ++c) {
            // H is used as z'Hz: only the symmetric part matters for the gradient 2Hz when Q is symmetric (the reference
            // assumes it: chol(Phi) reads one triangle).  Symmetrise so that both agree for any input.
            Q2[(size_t)r * n + c] = s->Q[(size_t)c * n + r] + s->Q[(size_t)r * n + c];
            Q2f[(size_t)r * n + c] = s->Qf[(size_t)c * n + r] + s->Qf[(size_t)r * n + c];
            if (r != c && (Q2[(size_t)r * n + c] != 0.0 || Q2f[(size_t)r * n + c] != 0.0)) qdiag = false;
        }
    if (qdiag) {
        Qi.assign((size_t)n * n, 0.0); Qif.assign((size_t)n * n, 0.0);
        for (int k = 0; k < n; ++k) {
            if (!(Q2[(size_t)k * n + k] > 0.0) || !(Q2f[(size_t)k * n + k] > 0.0)) return FMPC_ERR_NOT_PD;
            Qi[(size_t)k * n + k] = 1.0 / Q2[(size_t)k * n + k]; Qif[(size_t)k * n + k] = 1.0 / Q2f[(size_t)k * n + k];
        }
    } else if (!spd_inverse(Q2, n, Qi) || !spd_inverse(Q2f, n, Qif)) return FMPC_ERR_NOT_PD;

    // ---- Yx = C_x inv(Phi_xx) C_x'  (row-major NE x NE) ----
    std::vector<double> Yx((size_t)NE * NE, 0.0), Wk((size_t)n * n);
    for (int t = 0; t < T; ++t) {
        const int xc = t * st + m;                                   // columns of x_{t+1}
        const std::vector<double> &Qinv = (t == T - 1) ? Qif : Qi;
        std::vector<int> rowsb;
        for (int i = 0; i < NBm; ++i) if (cw_off[i] < xc + n && cw_off[i] + cw_len[i] > xc) rowsb.push_back(i);
        for (int i : rowsb) {
            // Wk = C[i, x_t] inv(2Q)   (n x n)
            for (int r = 0; r < n; ++r)
                for (int c = 0; c < n; ++c) {
                    double sacc = 0.0;
                    for (int k = 0; k < n; ++k) sacc += Cd[(size_t)(i * n + r) * N + xc + k] * Qinv[(size_t)k * n + c];
                    Wk[(size_t)r * n + c] = sacc;
                }
            for (int k2 : rowsb)
                for (int r = 0; r < n; ++r)
                    for (int c = 0; c < n; ++c) {
                        double sacc = 0.0;
                        for (int k = 0; k < n; ++k) sacc += Wk[(size_t)r * n + k] * Cd[(size_t)(k2 * n + c) * N + xc + k];
                        Yx[(size_t)(i * n + r) * NE + (size_t)k2 * n + c] += sacc;
                    }
        }
    }

    // ---- dense R (box rows only): R + R' row-major; SPD checked once here (the per-stage factorisation re-checks) ----
    bool rdiag = true;
    std::vector<double> R2((size_t)m * m);
    for (int r = 0; r < m; ++r)
        for (int c = 0; c < m; ++c) {
            R2[(size_t)r * m + c] = s->R[(size_t)c * m + r] + s->R[(size_t)r * m + c];
            if (r != c && R2[(size_t)r * m + c] != 0.0) rdiag = false;
        }
    if (!rdiag) {
        if (s->ramp_rows) return FMPC_ERR_UNSUPPORTED;           // dense R + ramp rows: Phi_uu is block tridiagonal in m x m blocks
        std::vector<double> Rinv;
        if (!spd_inverse(R2, m, Rinv)) return FMPC_ERR_NOT_PD;
    }

    GenSys G{};
    G.n = n; G.m = m; G.T = T; G.N = N; G.ramp = s->ramp_rows ? 1 : 0; G.ldyx = NE;
    G.dense_r = rdiag ? 0 : 1;
    bool ok = true;
#define GUP(field, vec) do { G.field = gen_upload(allocs, vec); if (!G.field) ok = false; } while (0)
    GUP(cw, cw); GUP(cw_ptr, cw_ptr); GUP(cw_off, cw_off); GUP(cw_len, cw_len);
    GUP(cu, cu); GUP(cut, cut); GUP(ue_cnt, ue_cnt); GUP(ue_t, ue_t); GUP(ue_ptr, ue_ptr);
    GUP(Yx, Yx); GUP(Q2, Q2); GUP(Q2f, Q2f); GUP(Qi, Qi); GUP(Qif, Qif);
    std::vector<double> dumin(m, 0.0), dumax(m, 0.0);
    if (s->ramp_rows) { dumin.assign(s->du_min, s->du_min + m); dumax.assign(s->du_max, s->du_max + m); }
    GUP(dumin, dumin); GUP(dumax, dumax);
    if (G.dense_r) GUP(R2, R2);
#undef GUP
    if (!ok) return FMPC_ERR_CUDA;

    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return FMPC_ERR_CUDA;
    const int ld = n | 1;
    const size_t fixed = gen_fixed_smem_doubles(n) * 8;
    const size_t budget = (size_t)prop.sharedMemPerBlockOptin - 1024;
    if (fixed + 8 * (size_t)ld * 8 > budget) return FMPC_ERR_UNSUPPORTED;
    int prow = (int)((budget - fixed) / ((size_t)ld * 8));
    if (prow > NE - n) prow = NE - n;
    if (prow < 8) prow = 8;
    G.panel_rows = prow;
    size_t smem = fixed + (size_t)prow * ld * 8;
    if (G.dense_r) {        // the panel area doubles as the packed Cholesky factor of Phi_uu and its inverse (2 x m(m+1)/2 doubles)
        const size_t need = fixed + (size_t)m * (m + 1) * 8;
        if (need > budget) return FMPC_ERR_UNSUPPORTED;
        if (need > smem) smem = need;
    }
    if (cudaFuncSetAttribute(fmpc_solve_kernel_gen, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return FMPC_ERR_CUDA;
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fmpc_solve_kernel_gen, GEN_THREADS, smem) != cudaSuccess || per_sm < 1)
        return FMPC_ERR_CUDA;
    cfg->grid = prop.multiProcessorCount * per_sm;
    cfg->block = GEN_THREADS;
    cfg->smem = smem;
    cfg->use_mma = 3;
    cfg->slots = cfg->grid;
    cfg->ws_stride = GenWs::make(n, m, T, G.ramp, G.dense_r).total;
    *out = G;
    return FMPC_OK;
}

void fmpc_launch_solve_gen(const DevSys &S, const GenSys &G, const StepArgs &A, const SolveLaunchCfg &cfg, void *stream)
{
    int grid = cfg.grid < A.nbatch ? cfg.grid : A.nbatch;
    if (grid < 1) grid = 1;
    fmpc_solve_kernel_gen<<<grid, cfg.block, cfg.smem, (cudaStream_t)stream>>>(S, G, A);
}
