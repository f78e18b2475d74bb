// fastMPC batched Newton solve -- warp-per-instance DMMA kernel for n <= 32 (sm_100a, fp64).
//
// One WARP owns one MPC instance from the first residual to the last line search
// (inf_newton_solver.m:10-41 on the block structure of DESIGN.md section 2).  Nothing in the solve needs
// a CTA barrier: the warps of a CTA only share the problem constants (B, A1, A2, bounds, weights) that
// are staged once in shared memory.  Every dense contraction is an FP64 tensor-pipe instruction
// (mma.sync.aligned.m8n8k4.f64 = DMMA):
//   * C z, C inv(Phi) r_d and the trial residuals are horizon GEMMs  [T x (m+2n)] * [(m+2n) x n]  with
//     the stage index as the M dimension; the barrier terms / trial point are computed by the thread
//     that owns the A-fragment element, in registers, straight from global memory (no staging);
//   * C' v is the transposed pair  [T x n] * [n x m],  [T x n] * [n x n];
//   * band-2 block Cholesky of Y = C inv(Phi) C' per stage i:
//       S_i  = Yd_i + B diag(w_i) B' - L1_{i-1} L1_{i-1}' - L2_{i-2} L2_{i-2}'     (lower tiles, registers)
//       L_i  = chol(S_i), inv(L_i)                                                  (one warp, registers)
//       L1_i = (Y1_i - L2_{i-1} L1_{i-1}') inv(L_i)' ,  L2_i = Y2 inv(L_i)'
//     The accumulator tiles of (Y1 - L2 L1') are fed back as the A operand of the next product without
//     leaving registers (the k index of an m8n8k4 contraction may be permuted freely, and the
//     accumulator layout {row gq, cols 2q, 2q+1} is a valid A layout for k = {2q+e}); inv(L_i) is stored
//     in that fragment order.  The forward substitution costs nothing: row n of every block (a padding
//     row of the last 8-row tile) carries y_{i-1}', so rhs_i and y_i = inv(L_i) rhs_i fall out of the
//     same tiles.
//   * the factor (inv(L_i), L1_i, L2_i) streams to a per-warp global scratch and comes back through a
//     4-slot cp.async ring for the backward substitution.
// Shared memory per warp: 3 history blocks + 1 work block + the staged w_i row; per CTA: the constants.
#include <cuda_runtime.h>
#include <math.h>
#include <stdlib.h>
#include "fmpc_internal.h"

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int TTMAX = 3;                 // 8-stage row tiles accumulated together in the horizon GEMMs

#ifdef FMPC_PROF
#define PROF_DECL long long p_acc[12]; long long p_last = clock64(); for (int i_ = 0; i_ < 12; ++i_) p_acc[i_] = 0;
#define PROF_T(idx) do { const long long now_ = clock64(); p_acc[idx] += now_ - p_last; p_last = now_; } while (0)
#else
#define PROF_DECL
#define PROF_T(idx) do { } while (0)
#endif

__device__ __forceinline__ void dmma(double (&c)[2], const double a, const double b)
{
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
__device__ __forceinline__ double dneg(const double x)
{
    return __hiloint2double(__double2hiint(x) ^ (int)0x80000000, __double2loint(x));
}
// 1/x : MUFU.RCP64H seed + two Newton steps (<= 1 ulp for normal x of either sign).
__device__ __forceinline__ double rcp_nr(const double x)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    return fma(y, e, y);
}
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
__device__ __forceinline__ double ldp(const double *p, const bool ok) { return ok ? *p : 0.0; }

__device__ __forceinline__ void cp_async16(double *smem_dst, const double *gsrc)
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// The two residual expressions live in ONE place each: the line search compares norms computed by
// bit-identical arithmetic as t -> 0 (backtracking_inf_newton.m:4 terminates that way).
__device__ __forceinline__ double rdu_expr(double r2, double rl, double u, double hu, double dbar)
{
    return __dadd_rn(__dsub_rn(__fma_rn(r2, u, rl), hu), dbar);      // 2R u + r - B'nu + k P'd
}
__device__ __forceinline__ double rdx_expr(double q2, double ql, double x, double hx)
{
    return __dadd_rn(__fma_rn(q2, x, ql), hx);                       // 2Q x + q + (C'nu)_x
}

// ---------------------------------------------------------------------------------------------
// Geometry (host + device).  NPOT = compile-time size of the diagonal-block factorization,
// CT = 8-column tiles, RT = 8-row tiles including the row that carries the forward substitution.
// ---------------------------------------------------------------------------------------------
__host__ __device__ inline int npot_of(int n) { return n <= 8 ? 8 : (n <= 16 ? 16 : (n <= 24 ? 24 : (n <= 28 ? 28 : 32))); }
__host__ __device__ inline int ld_of(int npot) { return (npot % 8 == 4) ? npot : npot + 4; }

struct WGeom {
    int n, m, T, NPOT, CT, RT, KS, LD, BLK, WSZ, mpad, MK, MT8, npad, LDB, NTT, NN;
    size_t const_doubles, warp_doubles;
    __host__ __device__ static WGeom make(int n, int m, int T)
    {
        WGeom g;
        g.n = n; g.m = m; g.T = T;
        g.NPOT = npot_of(n);
        g.CT = (g.NPOT + 7) / 8;
        g.RT = n / 8 + 1;
        g.KS = g.NPOT / 4;
        g.LD = ld_of(g.NPOT);
        g.BLK = ((g.NPOT + 1) * g.LD + 1) & ~1;
        g.WSZ = g.NPOT * (g.NPOT + 2) + 32;
        if (g.WSZ < 64) g.WSZ = 64;
        g.MT8 = (m + 7) / 8;
        g.mpad = 8 * g.MT8;
        g.MK = g.mpad / 4;
        g.npad = 8 * g.CT;
        g.LDB = (g.mpad % 8 == 4) ? g.mpad : g.mpad + 4;
        g.NTT = (T + 7) / 8;
        g.NN = (n * n + 1) & ~1;
        //          B            A1, A2 (+ overrun pad)   umax umin r2 rl     q2 q2f ql qfl qi qif
        g.const_doubles = (size_t)n * g.LDB + 2 * ((size_t)n * g.LD + 8) + 4 * (size_t)g.mpad + 6 * (size_t)g.npad;
        g.const_doubles = (g.const_doubles + 1) & ~(size_t)1;
        size_t wb = 2 * (size_t)g.mpad;
        if (wb < 128) wb = 128;                        // also holds the 4 x 32 vectors of the backward sweep
        g.warp_doubles = 3 * (size_t)g.BLK + (size_t)g.WSZ + wb;
        return g;
    }
};

// per-warp global scratch (doubles)
struct WsW {
    size_t UC, UT, HU, HDU, DU, WV, DB, RDU;       // T x mpad
    size_t XC, XT, HX, HDX, DX, RDX;               // T x npad
    size_t RP, RPT, YV, DNU, BV;                   // (T+1) x npad
    size_t Li, L1, L2;                             // (T+1) x NN
    size_t total;
    __host__ __device__ static WsW make(const WGeom &g)
    {
        WsW L;
        const size_t tu = (size_t)g.T * g.mpad, tx = (size_t)g.T * g.npad, tb = (size_t)(g.T + 1) * g.npad, bl = (size_t)(g.T + 1) * g.NN;
        size_t o = 0;
        L.UC = o; o += tu; L.UT = o; o += tu; L.HU = o; o += tu; L.HDU = o; o += tu; L.DU = o; o += tu; L.WV = o; o += tu;
        L.DB = o; o += tu; L.RDU = o; o += tu;
        L.XC = o; o += tx; L.XT = o; o += tx; L.HX = o; o += tx; L.HDX = o; o += tx; L.DX = o; o += tx; L.RDX = o; o += tx;
        L.RP = o; o += tb; L.RPT = o; o += tb; L.YV = o; o += tb; L.DNU = o; o += tb; L.BV = o; o += tb;
        L.Li = o; o += bl; L.L1 = o; o += bl; L.L2 = o; o += bl;
        L.total = (o + 31) & ~(size_t)31;
        return L;
    }
};

struct WCtx {
    int n, m, T, NB, a2, has_xf;
    int mpad, MK, MT8, npad, LDB, NTT, NN;
    int lane, gq, q;
    double kappa;
    // shared constants
    const double *sB, *sA1, *sA2, *sUmax, *sUmin, *sR2, *sRl, *sQ2, *sQl, *sQi;   // sQ*: [2][npad] (0: stages < T, 1: stage T)
    // per-warp shared
    double *blk0, *blk1, *blk2, *bW, *wbuf;
    // per-warp global scratch
    double *UC, *UT, *HU, *HDU, *DU, *WV, *DB, *RDU, *XC, *XT, *HX, *HDX, *DX, *RDX, *RP, *RPT, *YV, *DNU, *BV, *gLi, *gL1, *gL2;
    // pool of iterate-independent Schur blocks
    const double *ypool;
    const int *ydi, *y1i, *y2i;
};

enum { K_RP = 0, K_NEWTON = 1, K_TRIAL = 2 };

// ---------------------------------------------------------------------------------------------
// Horizon GEMM  acc[t][k] = (B v_u,t + A1 v_x,t-1 + A2 v_x,t-2)[k]  and its three uses
//   K_RP     : v = z              ; RP  = x_t+1 - acc - b                     (r_p = C z - b, inf_newton_solver.m:14)
//   K_NEWTON : v = inv(Phi) r_d   ; YV  = r_p - C v ; barrier terms, r_d, norms (inf_newton_KKT_H.m:3-13, :12, :28-29)
//   K_TRIAL  : v = z + ts dz      ; RPT = C v - b ; r_d(ts) with d frozen, norms (backtracking_inf_newton.m:4)
// ss_d / ss_p return this lane's partial sums of squares (same accumulation order in NEWTON and TRIAL).
// ---------------------------------------------------------------------------------------------
template <int NPOT, int KIND>
__device__ __forceinline__ void pass_Cv(const WCtx &c, const double ts, double &ss_d, double &ss_p)
{
    constexpr int CT = (NPOT + 7) / 8, KS = NPOT / 4, LD = (NPOT % 8 == 4) ? NPOT : NPOT + 4;
    const int n = c.n, T = c.T, mpad = c.mpad, npad = c.npad, gq = c.gq, q = c.q, lane = c.lane;
    ss_d = 0.0; ss_p = 0.0;
    // ---- x-space elementwise (linear ownership), then visible to the whole warp ----
    if (KIND == K_NEWTON) {
        for (int e = lane; e < T * npad; e += 32) {
            const int t = e / npad, k = e - t * npad, st = (t == T - 1) ? npad : 0;
            const double r = rdx_expr(c.sQ2[st + k], c.sQl[st + k], c.XC[e], c.HX[e]);
            c.RDX[e] = r;
            ss_d = fma(r, r, ss_d);
            c.DX[e] = r * c.sQi[st + k];                                   // p_x = inv(2Q) r_dx
        }
    } else if (KIND == K_TRIAL) {
        for (int e = lane; e < T * npad; e += 32) {
            const int t = e / npad, k = e - t * npad, st = (t == T - 1) ? npad : 0;
            const double xv = __fma_rn(ts, c.DX[e], c.XC[e]);
            c.XT[e] = xv;
            const double r = rdx_expr(c.sQ2[st + k], c.sQl[st + k], xv, __fma_rn(ts, c.HDX[e], c.HX[e]));
            ss_d = fma(r, r, ss_d);
        }
    }
    __syncwarp();
    const double *xsrc = (KIND == K_RP) ? c.XC : (KIND == K_NEWTON ? c.DX : c.XT);

    for (int tt0 = 0; tt0 < c.NTT; tt0 += TTMAX) {
        const int TT = min(TTMAX, c.NTT - tt0);
        double acc[TTMAX][CT][2];
#pragma unroll
        for (int tt = 0; tt < TTMAX; ++tt)
#pragma unroll
            for (int nt = 0; nt < CT; ++nt) acc[tt][nt][0] = acc[tt][nt][1] = 0.0;
        // ---- u part: A fragment element (t = 8 tt + gq, j = 4 kk + q) is produced by its owner ----
#pragma unroll 2
        for (int kk = 0; kk < c.MK; ++kk) {
            const int j = 4 * kk + q;
            double a[TTMAX];
#pragma unroll
            for (int tt = 0; tt < TTMAX; ++tt) {
                const int t = 8 * (tt0 + tt) + gq;
                a[tt] = 0.0;
                if (t < T) {
                    const size_t idx = (size_t)t * mpad + j;
                    if (KIND == K_RP) {
                        a[tt] = c.UC[idx];
                    } else if (KIND == K_NEWTON) {
                        const double uu = c.UC[idx], h = c.HU[idx];
                        const double sp = c.sUmax[j] - uu, sm = uu - c.sUmin[j];
                        const double dp = rcp_nr(sp), dm = rcp_nr(sm);
                        const double db = c.kappa * (dp - dm);
                        const double r2 = c.sR2[j];
                        const double w = rcp_nr(fma(c.kappa, fma(dp, dp, dm * dm), r2));
                        const double r = rdu_expr(r2, c.sRl[j], uu, h, db);
                        c.DB[idx] = db; c.WV[idx] = w; c.RDU[idx] = r;
                        ss_d = fma(r, r, ss_d);
                        a[tt] = r * w;                                      // p_u = inv(Phi_u) r_du
                    } else {
                        const double uv = __fma_rn(ts, c.DU[idx], c.UC[idx]);
                        c.UT[idx] = uv;
                        const double r = rdu_expr(c.sR2[j], c.sRl[j], uv, __fma_rn(ts, c.HDU[idx], c.HU[idx]), c.DB[idx]);
                        ss_d = fma(r, r, ss_d);
                        a[tt] = uv;
                    }
                }
            }
            double bf[CT];
#pragma unroll
            for (int nt = 0; nt < CT; ++nt) bf[nt] = ldp(c.sB + (size_t)(8 * nt + gq) * c.LDB + j, 8 * nt + gq < n);
#pragma unroll
            for (int tt = 0; tt < TTMAX; ++tt)
                if (tt < TT) {
#pragma unroll
                    for (int nt = 0; nt < CT; ++nt) dmma(acc[tt][nt], a[tt], bf[nt]);
                }
        }
        // ---- x part: A1 v_{t-1} + A2 v_{t-2} (shifted rows read from the scratch arrays) ----
#pragma unroll
        for (int kk = 0; kk < KS; ++kk) {
            const int kc = 4 * kk + q;
            double a1[TTMAX], a2[TTMAX];
#pragma unroll
            for (int tt = 0; tt < TTMAX; ++tt) {
                const int t = 8 * (tt0 + tt) + gq;
                a1[tt] = (t < T && t >= 1 && kc < n) ? xsrc[(size_t)(t - 1) * npad + kc] : 0.0;
                a2[tt] = (c.a2 && t < T && t >= 2 && kc < n) ? xsrc[(size_t)(t - 2) * npad + kc] : 0.0;
            }
            double b1[CT], b2[CT];
#pragma unroll
            for (int nt = 0; nt < CT; ++nt) {
                const bool ok = (8 * nt + gq < n) && (kc < n);
                b1[nt] = ldp(c.sA1 + (size_t)(8 * nt + gq) * LD + kc, ok);
                b2[nt] = ldp(c.sA2 + (size_t)(8 * nt + gq) * LD + kc, ok && c.a2);
            }
#pragma unroll
            for (int tt = 0; tt < TTMAX; ++tt)
                if (tt < TT) {
#pragma unroll
                    for (int nt = 0; nt < CT; ++nt) {
                        dmma(acc[tt][nt], a1[tt], b1[nt]);
                        if (c.a2) dmma(acc[tt][nt], a2[tt], b2[nt]);
                    }
                }
        }
        // ---- epilogue in the accumulator layout: (t = 8 tt + gq, k = 8 nt + 2 q + e) ----
#pragma unroll
        for (int tt = 0; tt < TTMAX; ++tt) {
            const int t = 8 * (tt0 + tt) + gq;
            if (tt < TT && t < T) {
#pragma unroll
                for (int nt = 0; nt < CT; ++nt) {
                    const size_t idx = (size_t)t * npad + 8 * nt + 2 * q;
                    if (KIND == K_RP || KIND == K_TRIAL) {
                        const double2 xv = *reinterpret_cast<const double2 *>(xsrc + idx);
                        const double2 bb = *reinterpret_cast<const double2 *>(c.BV + idx);
                        double2 r;
                        r.x = xv.x - acc[tt][nt][0] - bb.x;
                        r.y = xv.y - acc[tt][nt][1] - bb.y;
                        *reinterpret_cast<double2 *>((KIND == K_RP ? c.RP : c.RPT) + idx) = r;
                        ss_p = fma(r.x, r.x, ss_p);
                        ss_p = fma(r.y, r.y, ss_p);
                    } else {
                        const double2 rp = *reinterpret_cast<const double2 *>(c.RP + idx);
                        const double2 px = *reinterpret_cast<const double2 *>(c.DX + idx);
                        ss_p = fma(rp.x, rp.x, ss_p);
                        ss_p = fma(rp.y, rp.y, ss_p);
                        double2 y;
                        y.x = rp.x - px.x + acc[tt][nt][0];                  // -beta = r_p - C p
                        y.y = rp.y - px.y + acc[tt][nt][1];
                        *reinterpret_cast<double2 *>(c.YV + idx) = y;
                    }
                }
            }
        }
    }
    // ---- terminal row x_T = xf (fast_mpc_eq_const.m:67-71) ----
    if (c.has_xf && lane < n) {
        const size_t iT = (size_t)T * npad + lane, iL = (size_t)(T - 1) * npad + lane;
        if (KIND == K_RP || KIND == K_TRIAL) {
            const double r = xsrc[iL] - c.BV[iT];
            (KIND == K_RP ? c.RP : c.RPT)[iT] = r;
            ss_p = fma(r, r, ss_p);
        } else {
            const double rp = c.RP[iT];
            ss_p = fma(rp, rp, ss_p);
            c.YV[iT] = rp - c.DX[iL];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// C' v for v = V (rows t < NB, leading dimension ldv):  hu_t = B' v_t ,  hx_t = v_t - A1' v_{t+1} - A2' v_{t+2} (+ v_T)
//   MODE 0 : HU, HX stored (images of the dual start)
//   MODE 1 : HDU, HDX stored and  du = -(r_du - hdu) w ,  dx = -(r_dx + hdx) inv(2Q)      (inf_newton_solver.m:34-35)
// ---------------------------------------------------------------------------------------------
template <int NPOT, int MODE>
__device__ __forceinline__ void pass_Ct(const WCtx &c, const double *V, const int ldv)
{
    constexpr int CT = (NPOT + 7) / 8, KS = NPOT / 4, LD = (NPOT % 8 == 4) ? NPOT : NPOT + 4;
    const int n = c.n, T = c.T, mpad = c.mpad, npad = c.npad, gq = c.gq, q = c.q;
    for (int tt0 = 0; tt0 < c.NTT; tt0 += TTMAX) {
        const int TT = min(TTMAX, c.NTT - tt0);
        {   // ---- u part ----
            double av[TTMAX][KS];
#pragma unroll
            for (int tt = 0; tt < TTMAX; ++tt)
#pragma unroll
                for (int kk = 0; kk < KS; ++kk) {
                    const int t = 8 * (tt0 + tt) + gq, kc = 4 * kk + q;
                    av[tt][kk] = (t < T && kc < n) ? V[(size_t)t * ldv + kc] : 0.0;
                }
            for (int jt = 0; jt < c.MT8; ++jt) {
                double acc[TTMAX][2];
#pragma unroll
                for (int tt = 0; tt < TTMAX; ++tt) acc[tt][0] = acc[tt][1] = 0.0;
#pragma unroll
                for (int kk = 0; kk < KS; ++kk) {
                    const double bf = ldp(c.sB + (size_t)(4 * kk + q) * c.LDB + 8 * jt + gq, 4 * kk + q < n);
#pragma unroll
                    for (int tt = 0; tt < TTMAX; ++tt)
                        if (tt < TT) dmma(acc[tt], av[tt][kk], bf);
                }
#pragma unroll
                for (int tt = 0; tt < TTMAX; ++tt) {
                    const int t = 8 * (tt0 + tt) + gq;
                    if (tt < TT && t < T) {
                        const size_t idx = (size_t)t * mpad + 8 * jt + 2 * q;
                        const double2 h = make_double2(acc[tt][0], acc[tt][1]);
                        if (MODE == 0) {
                            *reinterpret_cast<double2 *>(c.HU + idx) = h;
                        } else {
                            *reinterpret_cast<double2 *>(c.HDU + idx) = h;
                            const double2 r = *reinterpret_cast<const double2 *>(c.RDU + idx);
                            const double2 w = *reinterpret_cast<const double2 *>(c.WV + idx);
                            double2 d;
                            d.x = -(r.x - h.x) * w.x;
                            d.y = -(r.y - h.y) * w.y;
                            *reinterpret_cast<double2 *>(c.DU + idx) = d;
                        }
                    }
                }
            }
        }
        {   // ---- x part ----
            double acc[TTMAX][CT][2];
#pragma unroll
            for (int tt = 0; tt < TTMAX; ++tt)
#pragma unroll
                for (int nt = 0; nt < CT; ++nt) acc[tt][nt][0] = acc[tt][nt][1] = 0.0;
#pragma unroll
            for (int kk = 0; kk < KS; ++kk) {
                const int kc = 4 * kk + q;
                double a1[TTMAX], a2[TTMAX];
#pragma unroll
                for (int tt = 0; tt < TTMAX; ++tt) {
                    const int t = 8 * (tt0 + tt) + gq;
                    a1[tt] = (t + 1 < T && kc < n) ? V[(size_t)(t + 1) * ldv + kc] : 0.0;
                    a2[tt] = (c.a2 && t + 2 < T && kc < n) ? V[(size_t)(t + 2) * ldv + kc] : 0.0;
                }
                double b1[CT], b2[CT];
#pragma unroll
                for (int nt = 0; nt < CT; ++nt) {
                    const bool ok = (kc < n) && (8 * nt + gq < n);
                    b1[nt] = ldp(c.sA1 + (size_t)kc * LD + 8 * nt + gq, ok);
                    b2[nt] = ldp(c.sA2 + (size_t)kc * LD + 8 * nt + gq, ok && c.a2);
                }
#pragma unroll
                for (int tt = 0; tt < TTMAX; ++tt)
                    if (tt < TT) {
#pragma unroll
                        for (int nt = 0; nt < CT; ++nt) {
                            dmma(acc[tt][nt], a1[tt], b1[nt]);
                            if (c.a2) dmma(acc[tt][nt], a2[tt], b2[nt]);
                        }
                    }
            }
#pragma unroll
            for (int tt = 0; tt < TTMAX; ++tt) {
                const int t = 8 * (tt0 + tt) + gq;
                if (tt < TT && t < T) {
                    const int st = (t == T - 1) ? npad : 0;
#pragma unroll
                    for (int nt = 0; nt < CT; ++nt) {
                        const int k0 = 8 * nt + 2 * q;
                        double h[2];
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int k = k0 + e;
                            double v = (k < n) ? V[(size_t)t * ldv + k] : 0.0;
                            v -= acc[tt][nt][e];
                            if (c.has_xf && t == T - 1 && k < n) v += V[(size_t)T * ldv + k];
                            h[e] = v;
                        }
                        const size_t idx = (size_t)t * npad + k0;
                        if (MODE == 0) {
                            *reinterpret_cast<double2 *>(c.HX + idx) = make_double2(h[0], h[1]);
                        } else {
                            *reinterpret_cast<double2 *>(c.HDX + idx) = make_double2(h[0], h[1]);
                            const double2 r = *reinterpret_cast<const double2 *>(c.RDX + idx);
                            double2 d;
                            d.x = -(r.x + h[0]) * c.sQi[st + k0];
                            d.y = -(r.y + h[1]) * c.sQi[st + k0 + 1];
                            *reinterpret_cast<double2 *>(c.DX + idx) = d;
                        }
                    }
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// One warp: Cholesky of the n x n block held (lower triangle, leading dimension NP+1) in bW, and the
// explicit inverse of the factor.  S = U D U' right-looking with row r in the registers of lane r
// (pivot chain through shuffles), then V = inv(U) column j by lane j, inv(L) = diag(1/sqrt(d)) V.
// Outputs: inv(L) in DMMA B-fragment order in bW (tile pair (ct, jt <= ct): 64 doubles, slot
// 2 * (4 gq + q) + e  <->  inv(L)(8 ct + gq, 8 jt + 2 q + e)), and row-major n x n in gLinv.
// Returns 0 or failing column + 1 (uniform).
// ---------------------------------------------------------------------------------------------
template <int NP>
__device__ __forceinline__ int warp_potrf_inverse(double *bW, const int n, double *gLinv, const int lane)
{
    constexpr int LDS_ = NP + 1, LDU = NP + 2, CT = (NP + 7) / 8;
    const int r = lane;
    double a[NP];
#pragma unroll
    for (int cc = 0; cc < NP; ++cc) a[cc] = (r < n && cc < r) ? bW[r * LDS_ + cc] : 0.0;
    double diag = (r < n) ? bW[r * LDS_ + r] : 1.0;
    double dpiv = 1.0;
    int info = 0;
    __syncwarp();                                           // S is in registers: bW is free
    double *colbuf = bW;                                    // 2 x 32
#pragma unroll
    for (int k = 0; k < NP; ++k) {
        const double d = __shfl_sync(FULL, diag, k);        // pivot of column k
        if (!(d > 0.0) || !(d < 1.0e300)) { if (!info) info = k + 1; }
        if (lane == k) dpiv = d;
        const double at = a[k];
        if (k + 1 < NP) {
            double *cb = colbuf + (k & 1) * 32;
            cb[lane] = at;
            const double dinv = rcp_nr(d);
            const double t = at * dinv;                     // U(r,k)
            a[k] = t;
            diag = fma(-t, at, diag);
            __syncwarp();
            if ((k + 1) & 1) a[k + 1] = fma(-t, cb[k + 1], a[k + 1]);
#pragma unroll
            for (int c2 = (k + 2) & ~1; c2 + 1 < NP; c2 += 2) {
                const double2 p = *reinterpret_cast<const double2 *>(cb + c2);
                a[c2] = fma(-t, p.x, a[c2]);
                a[c2 + 1] = fma(-t, p.y, a[c2 + 1]);
            }
        }
    }
    if (info) return info;
    __syncwarp();                                           // column buffer dead
    double *bU = bW, *rsv = bW + NP * LDU;
    rsv[lane] = rsqrt(dpiv);
    if (r < NP) {
#pragma unroll
        for (int cc = 0; cc + 1 < NP; cc += 2) *reinterpret_cast<double2 *>(bU + r * LDU + cc) = make_double2(a[cc], a[cc + 1]);
    }
    __syncwarp();
    const int j = lane;
    double v[NP];
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int k = 0; k + 1 < i; k += 2) {
            const double2 p = *reinterpret_cast<const double2 *>(bU + i * LDU + k);
            s0 = fma(p.x, v[k], s0);
            s1 = fma(p.y, v[k + 1], s1);
        }
        if (i & 1) s0 = fma(bU[i * LDU + i - 1], v[i - 1], s0);
        v[i] = (i < j) ? 0.0 : ((i == j) ? 1.0 : -(s0 + s1));
    }
    double rs[NP];
#pragma unroll
    for (int i = 0; i < NP; ++i) rs[i] = rsv[i];
    __syncwarp();                                           // U and rsv consumed: the fragment area may be written
    const bool live = (j < n);
    const int jt = j >> 3, jslot = 2 * ((j & 7) >> 1) + (j & 1);           // 2 q' + e
#pragma unroll
    for (int i = 0; i < 8 * CT; ++i) {
        const int ct = i >> 3, gqp = i & 7;
        double val = 0.0;
        if (i < NP) val = (live && i < n) ? v[i < NP ? i : 0] * rs[i < NP ? i : 0] : 0.0;
        if (j < 8 * CT && jt <= ct) bW[(ct * (ct + 1) / 2 + jt) * 64 + 8 * gqp + jslot] = val;
        if (live && i < n) gLinv[i * n + j] = val;
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Band-2 block Cholesky of Y fused with the forward substitution (inf_newton_solver.m:27-31).
// On return YV holds y = inv(L) (-beta).  Returns 0 or the failing stage + 1.
// ---------------------------------------------------------------------------------------------
template <int NPOT, int RT>
__device__ __forceinline__ int forward_sweep(const WCtx &c)
{
    constexpr int CT = (NPOT + 7) / 8, KS = NPOT / 4, LD = (NPOT % 8 == 4) ? NPOT : NPOT + 4, LDS_ = NPOT + 1;
    const int n = c.n, T = c.T, NB = c.NB, mpad = c.mpad, npad = c.npad, gq = c.gq, q = c.q, lane = c.lane;
    const int yr_gq = n & 7;                                // row n lives in tile RT-1, fragment row yr_gq
    const bool yrow = (gq == yr_gq);
    const size_t nn = (size_t)n * n;
    double *bL1 = c.blk0, *bL2p = c.blk1, *bL2pp = c.blk2, *bW = c.bW;

    // stage the first w row
    for (int ch = lane; ch < mpad / 2; ch += 32) cp_async16(c.wbuf + 2 * ch, c.WV + 2 * ch);
    cp_async_commit();

    for (int i = 0; i < NB; ++i) {
        const bool has1 = (i + 1 < NB), has2 = (i + 2 < NB) && c.a2;
        const bool up1 = (i >= 1), up2 = (i >= 2) && c.a2;
        // ================= phase A: S_i (lower tiles) + rhs row =================
        double sacc[RT][CT][2];
#pragma unroll
        for (int rt = 0; rt < RT; ++rt)
#pragma unroll
            for (int ct = 0; ct < CT; ++ct) sacc[rt][ct][0] = sacc[rt][ct][1] = 0.0;
        if (i < T) {
            cp_async_wait<0>();
            __syncwarp();
            const double *wb = c.wbuf + (size_t)(i & 1) * mpad;
            if (i + 1 < T) {
                double *wn = c.wbuf + (size_t)((i + 1) & 1) * mpad;
                const double *src = c.WV + (size_t)(i + 1) * mpad;
                for (int ch = lane; ch < mpad / 2; ch += 32) cp_async16(wn + 2 * ch, src + 2 * ch);
            }
            cp_async_commit();
            // B diag(w_i) B'
#pragma unroll 2
            for (int kk = 0; kk < c.MK; ++kk) {
                const int j = 4 * kk + q;
                const double wv = wb[j];
                double fr[CT];
#pragma unroll
                for (int rt = 0; rt < CT; ++rt) fr[rt] = ldp(c.sB + (size_t)(8 * rt + gq) * c.LDB + j, 8 * rt + gq < n);
#pragma unroll
                for (int rt = 0; rt < CT; ++rt) {
                    const double a = fr[rt] * wv;
#pragma unroll
                    for (int ct = 0; ct <= rt; ++ct) dmma(sacc[rt][ct], a, fr[ct]);
                }
            }
        }
        {   // + iterate-independent part of Y[i,i]; rhs row <- -beta_i
            const int yd = c.ydi[i];
            if (yd >= 0) {
                const double *Yd = c.ypool + (size_t)yd * nn;
#pragma unroll
                for (int rt = 0; rt < CT; ++rt) {
                    const int r = 8 * rt + gq;
#pragma unroll
                    for (int ct = 0; ct <= rt; ++ct) {
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int cc = 8 * ct + 2 * q + e;
                            if (r < n && cc <= r) sacc[rt][ct][e] += __ldg(Yd + (size_t)r * n + cc);
                        }
                    }
                }
            }
            if (yrow) {
#pragma unroll
                for (int ct = 0; ct < CT; ++ct) {
                    const double2 y = *reinterpret_cast<const double2 *>(c.YV + (size_t)i * npad + 8 * ct + 2 * q);
                    sacc[RT - 1][ct][0] = y.x;
                    sacc[RT - 1][ct][1] = y.y;
                }
            }
        }
        if (up1) {
#pragma unroll
            for (int kk = 0; kk < KS; ++kk) {
                double fr[RT];
#pragma unroll
                for (int rt = 0; rt < RT; ++rt) fr[rt] = ldp(bL1 + (8 * rt + gq) * LD + 4 * kk + q, 8 * rt + gq <= n);
#pragma unroll
                for (int rt = 0; rt < RT; ++rt) {
                    const double a = dneg(fr[rt]);
#pragma unroll
                    for (int ct = 0; ct < CT; ++ct)
                        if (ct <= rt) dmma(sacc[rt][ct], a, fr[ct]);
                }
            }
        }
        if (up2) {
#pragma unroll
            for (int kk = 0; kk < KS; ++kk) {
                double fr[RT];
#pragma unroll
                for (int rt = 0; rt < RT; ++rt) fr[rt] = ldp(bL2pp + (8 * rt + gq) * LD + 4 * kk + q, 8 * rt + gq <= n);
#pragma unroll
                for (int rt = 0; rt < RT; ++rt) {
                    const double a = dneg(fr[rt]);
#pragma unroll
                    for (int ct = 0; ct < CT; ++ct)
                        if (ct <= rt) dmma(sacc[rt][ct], a, fr[ct]);
                }
            }
        }
        // S -> work block (row per lane layout for the factorization); rhs row stays in registers
        double rhsv[CT][2];
#pragma unroll
        for (int ct = 0; ct < CT; ++ct) { rhsv[ct][0] = sacc[RT - 1][ct][0]; rhsv[ct][1] = sacc[RT - 1][ct][1]; }
#pragma unroll
        for (int rt = 0; rt < CT; ++rt) {
            const int r = 8 * rt + gq;
#pragma unroll
            for (int ct = 0; ct <= rt; ++ct) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int cc = 8 * ct + 2 * q + e;
                    if (r < n && cc <= r) bW[r * LDS_ + cc] = sacc[rt][ct][e];
                }
            }
        }
        __syncwarp();
        // ================= phase B: L_i, inv(L_i) =================
        const int info = warp_potrf_inverse<NPOT>(bW, n, c.gLi + (size_t)i * c.NN, lane);
        if (info) return i + 1;
        __syncwarp();
        // ================= phase C: L1_i, y_i, L2_i =================
        double macc[RT][CT][2];
        {
            const int y1 = c.y1i[i];
            const double *Y1 = c.ypool + (size_t)(y1 >= 0 ? y1 : 0) * nn;
            const bool ld1 = has1 && (y1 >= 0);
#pragma unroll
            for (int rt = 0; rt < RT; ++rt) {
                const int r = 8 * rt + gq;
#pragma unroll
                for (int ct = 0; ct < CT; ++ct) {
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int cc = 8 * ct + 2 * q + e;
                        macc[rt][ct][e] = (ld1 && r < n && cc < n) ? __ldg(Y1 + (size_t)r * n + cc) : 0.0;
                    }
                }
            }
        }
        if (has1 && up1 && c.a2) {
#pragma unroll
            for (int kk = 0; kk < KS; ++kk) {
                double fa[RT], fb[CT];
#pragma unroll
                for (int rt = 0; rt < RT; ++rt) fa[rt] = dneg(ldp(bL2p + (8 * rt + gq) * LD + 4 * kk + q, 8 * rt + gq < n));
#pragma unroll
                for (int ct = 0; ct < CT; ++ct) fb[ct] = ldp(bL1 + (8 * ct + gq) * LD + 4 * kk + q, 8 * ct + gq < n);
#pragma unroll
                for (int rt = 0; rt < RT; ++rt)
#pragma unroll
                    for (int ct = 0; ct < CT; ++ct) dmma(macc[rt][ct], fa[rt], fb[ct]);
            }
        }
        if (yrow) {
#pragma unroll
            for (int ct = 0; ct < CT; ++ct) { macc[RT - 1][ct][0] = rhsv[ct][0]; macc[RT - 1][ct][1] = rhsv[ct][1]; }
        }
        __syncwarp();                                        // every read of L1_{i-1} is done: its block receives L1_i
        double yv[CT][2];
#pragma unroll
        for (int rt = 0; rt < RT; ++rt) {
            if (has1 || rt == RT - 1) {
                const int r = 8 * rt + gq;
#pragma unroll
                for (int ct = 0; ct < CT; ++ct) {
                    double o[2] = {0.0, 0.0};
#pragma unroll
                    for (int jt = 0; jt <= ct; ++jt) {
                        const double2 lf = *reinterpret_cast<const double2 *>(bW + (ct * (ct + 1) / 2 + jt) * 64 + 2 * lane);
                        dmma(o, macc[rt][jt][0], lf.x);
                        dmma(o, macc[rt][jt][1], lf.y);
                    }
                    if (rt == RT - 1) { yv[ct][0] = o[0]; yv[ct][1] = o[1]; }
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int cc = 8 * ct + 2 * q + e;
                        if (r <= n && cc < NPOT) bL1[r * LD + cc] = o[e];
                        if (has1 && r < n && cc < n) c.gL1[(size_t)i * c.NN + (size_t)r * n + cc] = o[e];
                        if (r == n && cc < n) c.YV[(size_t)i * npad + cc] = o[e];
                    }
                }
            }
        }
        if (has2) {
            const bool y2ok = (c.y2i[i] >= 0);
            double nqi[CT][2];                                // -inv(2Q)(k) for k = 8 jt + 2 q + e
#pragma unroll
            for (int jt = 0; jt < CT; ++jt)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int k = 8 * jt + 2 * q + e;
                    nqi[jt][e] = (k < n) ? -c.sQi[k] : 0.0;
                }
#pragma unroll
            for (int rt = 0; rt < CT; ++rt) {
                const int r = 8 * rt + gq;
                double af[CT][2];
#pragma unroll
                for (int jt = 0; jt < CT; ++jt)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int k = 8 * jt + 2 * q + e;
                        af[jt][e] = ldp(c.sA2 + (size_t)r * LD + k, y2ok && r < n && k < n) * nqi[jt][e];   // Y2 = -A2 inv(2Q)
                    }
#pragma unroll
                for (int ct = 0; ct < CT; ++ct) {
                    double o[2] = {0.0, 0.0};
#pragma unroll
                    for (int jt = 0; jt <= ct; ++jt) {
                        const double2 lf = *reinterpret_cast<const double2 *>(bW + (ct * (ct + 1) / 2 + jt) * 64 + 2 * lane);
                        dmma(o, af[jt][0], lf.x);
                        dmma(o, af[jt][1], lf.y);
                    }
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int cc = 8 * ct + 2 * q + e;
                        if (r < n && cc < NPOT) bL2pp[r * LD + cc] = o[e];
                        if (r < n && cc < n) c.gL2[(size_t)i * c.NN + (size_t)r * n + cc] = o[e];
                    }
                }
            }
            if (yrow) {                                       // row n of the L2 block carries y_i as well
#pragma unroll
                for (int ct = 0; ct < CT; ++ct)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int cc = 8 * ct + 2 * q + e;
                        if (cc < NPOT) bL2pp[n * LD + cc] = yv[ct][e];
                    }
            }
        }
        { double *t2 = bL2p; bL2p = bL2pp; bL2pp = t2; }
        __syncwarp();
    }
    cp_async_wait<0>();
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Backward substitution  dnu_i = inv(L_i)' (y_i - L1_i' dnu_{i+1} - L2_i' dnu_{i+2})  (inf_newton_solver.m:32).
// The factor comes back from the global scratch through a 4-slot cp.async ring (3 entries per stage).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void ring_issue(const WCtx &c, double *const (&slot)[4], const int e, const int nent)
{
    if (e < nent) {
        const int i = c.NB - 1 - e / 3, kind = e % 3;
        const bool on = (kind == 2) || (kind == 0 && i + 1 < c.NB) || (kind == 1 && i + 2 < c.NB && c.a2);
        if (on) {
            const double *src = (kind == 0 ? c.gL1 : (kind == 1 ? c.gL2 : c.gLi)) + (size_t)i * c.NN;
            double *dst = slot[e & 3];
            for (int ch = c.lane; ch < c.NN / 2; ch += 32) cp_async16(dst + 2 * ch, src + 2 * ch);
        }
    }
    cp_async_commit();
}

__device__ __forceinline__ void backward_sweep(const WCtx &c)
{
    const int n = c.n, NB = c.NB, npad = c.npad, lane = c.lane;
    double *const slot[4] = {c.blk0, c.blk1, c.blk2, c.bW};
    double *vec = c.wbuf;                                   // [3][32] dnu ring + [32] tmp
    const int nent = 3 * NB;
    __syncwarp();
    for (int e = 0; e < 4; ++e) ring_issue(c, slot, e, nent);
    double acc = 0.0;
    for (int e = 0; e < nent; ++e) {
        const int i = NB - 1 - e / 3, kind = e % 3;
        cp_async_wait<3>();
        __syncwarp();
        const double *M = slot[e & 3];
        if (kind == 0) {
            acc = 0.0;
            if (i + 1 < NB && lane < n) {
                const double *d1 = vec + ((i + 1) % 3) * 32;
                double s0 = 0.0, s1 = 0.0;
                int r = 0;
                for (; r + 1 < n; r += 2) { s0 = fma(M[r * n + lane], d1[r], s0); s1 = fma(M[(r + 1) * n + lane], d1[r + 1], s1); }
                if (r < n) s0 = fma(M[r * n + lane], d1[r], s0);
                acc = s0 + s1;
            }
        } else if (kind == 1) {
            if (i + 2 < NB && c.a2 && lane < n) {
                const double *d2 = vec + ((i + 2) % 3) * 32;
                double s0 = 0.0, s1 = 0.0;
                int r = 0;
                for (; r + 1 < n; r += 2) { s0 = fma(M[r * n + lane], d2[r], s0); s1 = fma(M[(r + 1) * n + lane], d2[r + 1], s1); }
                if (r < n) s0 = fma(M[r * n + lane], d2[r], s0);
                acc += s0 + s1;
            }
        } else {
            double *tmp = vec + 96;
            if (lane < n) tmp[lane] = c.YV[(size_t)i * npad + lane] - acc;
            __syncwarp();
            if (lane < n) {
                double s0 = 0.0, s1 = 0.0;
                int r = lane;
                for (; r + 1 < n; r += 2) { s0 = fma(M[r * n + lane], tmp[r], s0); s1 = fma(M[(r + 1) * n + lane], tmp[r + 1], s1); }
                if (r < n) s0 = fma(M[r * n + lane], tmp[r], s0);
                const double d = s0 + s1;
                vec[(i % 3) * 32 + lane] = d;
                c.DNU[(size_t)i * npad + lane] = d;
            }
        }
        __syncwarp();                                        // slot consumed by every lane
        ring_issue(c, slot, e + 4, nent);
    }
    cp_async_wait<0>();
    __syncwarp();
}

} // namespace

// =============================================================================================
template <int NPOT, int RT>
__global__ void __launch_bounds__(256, 1) fmpc_solve_kernel_warp(const DevSys S, const StepArgs A)
{
    extern __shared__ double smem[];
    const int n = S.n, m = S.m, T = S.T;
    const WGeom G = WGeom::make(n, m, T);
    const WsW L = WsW::make(G);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nwarps = blockDim.x >> 5;
    constexpr int LD = (NPOT % 8 == 4) ? NPOT : NPOT + 4;

    // ---- shared memory: zero everything (padding and overrun reads must see finite values), then the constants ----
    {
        const size_t tot = G.const_doubles + (size_t)nwarps * G.warp_doubles;
        for (size_t e = tid; e < tot; e += blockDim.x) smem[e] = 0.0;
    }
    __syncthreads();
    double *sB = smem;
    double *sA1 = sB + (size_t)n * G.LDB;
    double *sA2 = sA1 + (size_t)n * LD + 8;
    double *sUmax = sA2 + (size_t)n * LD + 8;
    double *sUmin = sUmax + G.mpad, *sR2 = sUmin + G.mpad, *sRl = sR2 + G.mpad;
    double *sQ2 = sRl + G.mpad, *sQl = sQ2 + 2 * G.npad, *sQi = sQl + 2 * G.npad;
    for (int e = tid; e < n * m; e += blockDim.x) { const int k = e % n, j = e / n; sB[(size_t)k * G.LDB + j] = S.B[e]; }      // B is column-major n x m
    for (int e = tid; e < n * n; e += blockDim.x) {
        const int k = e % n, kc = e / n;
        sA1[(size_t)k * LD + kc] = S.A1[e];
        sA2[(size_t)k * LD + kc] = S.has_a2 ? S.A2[e] : 0.0;
    }
    for (int j = tid; j < G.mpad; j += blockDim.x) {
        sUmax[j] = (j < m) ? S.umax[j] : 1.0;
        sUmin[j] = (j < m) ? S.umin[j] : -1.0;
        sR2[j] = (j < m) ? S.r2[j] : 1.0;
        sRl[j] = (j < m) ? S.rl[j] : 0.0;
    }
    for (int k = tid; k < G.npad; k += blockDim.x) {
        const bool ok = k < n;
        sQ2[k] = ok ? S.q2[k] : 0.0;  sQ2[G.npad + k] = ok ? S.q2f[k] : 0.0;
        sQl[k] = ok ? S.ql[k] : 0.0;  sQl[G.npad + k] = ok ? S.qfl[k] : 0.0;
        sQi[k] = ok ? S.qi[k] : 0.0;  sQi[G.npad + k] = ok ? S.qif[k] : 0.0;
    }
    __syncthreads();

    WCtx c;
    c.n = n; c.m = m; c.T = T; c.NB = T + (A.has_xf ? 1 : 0); c.a2 = S.has_a2; c.has_xf = A.has_xf;
    c.mpad = G.mpad; c.MK = G.MK; c.MT8 = G.MT8; c.npad = G.npad; c.LDB = G.LDB; c.NTT = G.NTT; c.NN = G.NN;
    c.lane = lane; c.gq = lane >> 2; c.q = lane & 3;
    c.kappa = A.kappa;
    c.sB = sB; c.sA1 = sA1; c.sA2 = sA2; c.sUmax = sUmax; c.sUmin = sUmin; c.sR2 = sR2; c.sRl = sRl; c.sQ2 = sQ2; c.sQl = sQl; c.sQi = sQi;
    double *wsm = smem + G.const_doubles + (size_t)wid * G.warp_doubles;
    c.blk0 = wsm; c.blk1 = wsm + G.BLK; c.blk2 = wsm + 2 * G.BLK; c.bW = wsm + 3 * G.BLK; c.wbuf = c.bW + G.WSZ;
    double *ws = A.ws + ((size_t)blockIdx.x * nwarps + wid) * A.ws_stride;
    c.HU = ws + L.HU; c.HDU = ws + L.HDU; c.DU = ws + L.DU; c.WV = ws + L.WV; c.DB = ws + L.DB; c.RDU = ws + L.RDU;
    c.HX = ws + L.HX; c.HDX = ws + L.HDX; c.DX = ws + L.DX; c.RDX = ws + L.RDX;
    c.YV = ws + L.YV; c.DNU = ws + L.DNU; c.BV = ws + L.BV;
    c.gLi = ws + L.Li; c.gL1 = ws + L.L1; c.gL2 = ws + L.L2;
    c.ypool = S.ypool; c.ydi = S.ydi; c.y1i = S.y1i; c.y2i = S.y2i;
    const int NB = c.NB, mpad = G.mpad, npad = G.npad;

    for (;;) {
        int b = 0;
        if (lane == 0) b = (int)atomicAdd(A.counter, 1u);
        b = __shfl_sync(FULL, b, 0);
        if (b >= A.nbatch) break;
        PROF_DECL
        c.UC = ws + L.UC; c.UT = ws + L.UT; c.XC = ws + L.XC; c.XT = ws + L.XT; c.RP = ws + L.RP; c.RPT = ws + L.RPT;
        const double *x0 = A.x0 + (size_t)b * n;
        const double *x0p = A.x0_pre ? A.x0_pre + (size_t)b * n : nullptr;

        // ---- b (fast_mpc_eq_const.m:39,44,47,68), initial iterate (fast_mpc_init.m:12-26) ----
        if (lane < n) {
            double s0 = 0.0, s1 = 0.0;
            for (int kc = 0; kc < n; ++kc) {
                const double xv = x0[kc];
                s0 = fma(sA1[(size_t)lane * LD + kc], xv, s0);
                if (S.has_a2) { s0 = fma(sA2[(size_t)lane * LD + kc], x0p[kc], s0); s1 = fma(sA2[(size_t)lane * LD + kc], xv, s1); }
            }
            for (int i = 0; i < NB; ++i) {
                double v;
                if (i < T) {
                    v = A.w ? A.w[(size_t)b * T * n + (size_t)i * n + lane] : 0.0;
                    if (i == 0) v += s0;
                    else if (i == 1) v += s1;
                } else {
                    v = A.xf[(size_t)b * n + lane];
                }
                c.BV[(size_t)i * npad + lane] = v;
            }
        }
        {
            const double *u0 = A.cold ? nullptr : A.U0 + (size_t)b * m * T;
            const double *xx0 = A.cold ? nullptr : A.X0 + (size_t)b * n * T;
            for (int e = lane; e < T * mpad; e += 32) {
                const int t = e / mpad, j = e - t * mpad;
                c.UC[e] = (j < m) ? (A.cold ? (S.umin[j] + S.umax[j]) / 2 : u0[(size_t)t * m + j]) : 0.0;
            }
            for (int e = lane; e < T * npad; e += 32) {
                const int t = e / npad, k = e - t * npad;
                c.XC[e] = (k < n) ? (A.cold ? (S.xmin[k] + S.xmax[k]) / 2 : xx0[(size_t)t * n + k]) : 0.0;
            }
        }
        __syncwarp();
        double ssd, ssp;
        pass_Cv<NPOT, K_RP>(c, 0.0, ssd, ssp);                         // r_p = C z - b
        __syncwarp();
        pass_Ct<NPOT, 0>(c, A.nu0 + (size_t)b * NB * n, n);            // images of the dual start nu
        __syncwarp();
        PROF_T(0);

        int status = ST_OK, iters = 0;
        for (int it = 0; it < A.niters; ++it) {
            pass_Cv<NPOT, K_NEWTON>(c, 0.0, ssd, ssp);
            const double tot_p = warp_sum(ssp), tot_d = warp_sum(ssd);
            const double nr0 = sqrt(tot_d + tot_p);
            PROF_T(1);
            // ---- early exit (inf_newton_solver.m:19-22) ----
            if (!isfinite(nr0)) { status = ST_NONFINITE; break; }
            if (nr0 <= A.tol_r && sqrt(tot_p) <= A.tol_p) { status = ST_EARLY_EXIT; break; }
            __syncwarp();
            const int fail = forward_sweep<NPOT, RT>(c);
            PROF_T(2);
            if (fail) { status = ST_NOT_PD; break; }
            backward_sweep(c);
            PROF_T(3);
            pass_Ct<NPOT, 1>(c, c.DNU, npad);                          // dz = inv(Phi)(-r_d - C' dnu)
            __syncwarp();
            PROF_T(4);
            // ---- backtracking on ||[r_p; r_d]||, d frozen (backtracking_inf_newton.m:2-11) ----
            double t = 1.0;
            int nh = 0;
            for (;;) {
                pass_Cv<NPOT, K_TRIAL>(c, t, ssd, ssp);
                const double tp = warp_sum(ssp), td = warp_sum(ssd);
                const double nrt = sqrt(td + tp);
                __syncwarp();
                if (!(nrt > (1.0 - A.alpha * t) * nr0)) break;
                if (t == 0.0) break;
                if (A.ls_max > 0 && nh >= A.ls_max) { status = ST_LS_MAX; break; }
                t *= A.beta;
                ++nh;
            }
            PROF_T(5);
            // accept: swap iterate / r_p buffers, advance the dual images
            { double *s1 = c.UC; c.UC = c.UT; c.UT = s1; double *s2 = c.XC; c.XC = c.XT; c.XT = s2; double *s3 = c.RP; c.RP = c.RPT; c.RPT = s3; }
            for (int e = lane; e < T * mpad; e += 32) c.HU[e] = __fma_rn(t, c.HDU[e], c.HU[e]);
            for (int e = lane; e < T * npad; e += 32) c.HX[e] = __fma_rn(t, c.HDX[e], c.HX[e]);
            ++iters;
            __syncwarp();
            PROF_T(6);
        }
        __syncwarp();
        {
            double *uo = A.U + (size_t)b * m * T, *xo = A.X + (size_t)b * n * T;
            for (int e = lane; e < T * m; e += 32) { const int t = e / m, j = e - t * m; uo[e] = c.UC[(size_t)t * mpad + j]; }
            for (int e = lane; e < T * n; e += 32) { const int t = e / n, k = e - t * n; xo[e] = c.XC[(size_t)t * npad + k]; }
        }
        if (lane == 0) {
            if (A.status) A.status[b] = status;
            if (A.iters) A.iters[b] = iters;
            atomicAdd(A.iters_total, (unsigned long long)iters);
        }
        __syncwarp();
        PROF_T(7);
#ifdef FMPC_PROF
        if (A.prof && lane == 0)
            for (int i_ = 0; i_ < 12; ++i_) atomicAdd((unsigned long long *)A.prof + i_, (unsigned long long)p_acc[i_]);
#endif
    }
}

// =============================================================================================
template <int NPOT, int RT>
static int config_warp(const WGeom &G, SolveLaunchCfg *cfg, const cudaDeviceProp &prop)
{
    const size_t avail = prop.sharedMemPerBlockOptin;
    if (G.const_doubles * 8 + G.warp_doubles * 8 > avail) return -3;
    int warps = (int)((avail - G.const_doubles * 8) / (G.warp_doubles * 8));
    if (warps > 8) warps = 8;
    if (const char *e = getenv("FMPC_WARPS_PER_CTA")) { const int v = atoi(e); if (v >= 1 && v < warps) warps = v; }
    const size_t smem = (G.const_doubles + (size_t)warps * G.warp_doubles) * 8;
    if (cudaFuncSetAttribute(fmpc_solve_kernel_warp<NPOT, RT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -4;
    cfg->grid = prop.multiProcessorCount;
    cfg->block = 32 * warps;
    cfg->smem = smem;
    cfg->use_mma = 2;
    cfg->slots = cfg->grid * warps;
    cfg->ws_stride = WsW::make(G).total;
    return 0;
}

#define WARP_DISPATCH(G, CALL)                                                                         \
    switch ((G).NPOT * 8 + (G).RT) {                                                                   \
    case 8 * 8 + 1: CALL(8, 1); break;                                                                 \
    case 8 * 8 + 2: CALL(8, 2); break;                                                                 \
    case 16 * 8 + 2: CALL(16, 2); break;                                                               \
    case 16 * 8 + 3: CALL(16, 3); break;                                                               \
    case 24 * 8 + 3: CALL(24, 3); break;                                                               \
    case 24 * 8 + 4: CALL(24, 4); break;                                                               \
    case 28 * 8 + 4: CALL(28, 4); break;                                                               \
    case 32 * 8 + 4: CALL(32, 4); break;                                                               \
    case 32 * 8 + 5: CALL(32, 5); break;                                                               \
    default: break;                                                                                    \
    }

int fmpc_warp_config(const DevSys &S, int device, SolveLaunchCfg *cfg)
{
    if (S.n > 32) return -1;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return -2;
    const WGeom G = WGeom::make(S.n, S.m, S.T);
    int rc = -6;
#define CALL_CFG(NP_, RT_) rc = config_warp<NP_, RT_>(G, cfg, prop)
    WARP_DISPATCH(G, CALL_CFG)
#undef CALL_CFG
    return rc;
}

void fmpc_launch_solve_warp(const DevSys &S, const StepArgs &A, const SolveLaunchCfg &cfg, void *stream)
{
    const WGeom G = WGeom::make(S.n, S.m, S.T);
    const int warps = cfg.block / 32;
    int grid = (A.nbatch + warps - 1) / warps;
    if (grid > cfg.grid) grid = cfg.grid;
    if (grid < 1) grid = 1;
    cudaStream_t st = (cudaStream_t)stream;
#define CALL_LAUNCH(NP_, RT_) fmpc_solve_kernel_warp<NP_, RT_><<<grid, cfg.block, cfg.smem, st>>>(S, A)
    WARP_DISPATCH(G, CALL_LAUNCH)
#undef CALL_LAUNCH
}
