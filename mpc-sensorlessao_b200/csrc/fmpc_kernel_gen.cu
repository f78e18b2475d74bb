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
//   barrier terms -> residuals -> per-actuator tridiagonal LDL' + explicit inverse M_j (T x T, by recurrences) ->
//   Y = Yx + sum Cu_a diag(M[t_a,t_b,:]) Cu_b'  (warp tasks of 2 x 4 DMMA tiles, -B staged in shared memory) ->
//   dense blocked Cholesky with the right-hand side as one more row: inv(L_KK) by one warp (registers, look-ahead: it
//   runs during the trailing update of the block column before), panel <- panel inv(L_KK)' and the trailing update as DMMA
//   tiles from shared memory -> backward substitution (products with inv(L_KK)') -> dz = inv(Phi)(-r_d - C' dnu) ->
//   residual-norm back-tracking (window of C + trial point in shared memory).
// Two block sizes: 256 threads (two CTAs per SM) for batches, 512 threads per instance for batches below half the resident
// CTAs (FMPC_GEN_NARROW=1 forces 256; FMPC_GEN_GRID=g caps the resident CTAs: experiments).  DESIGN.md 3a has the numbers.
#include <cuda_runtime.h>
#include <math.h>
#include <cstring>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <utility>
#include "../../include/fmpc.h"
#include "fmpc_internal.h"
#include "fmpc_device.cuh"

using namespace fmpc_dev;

namespace {

#ifndef GEN_THREADS_N
#define GEN_THREADS_N 256
#endif
constexpr int GEN_THREADS = GEN_THREADS_N;       // batches: GEN_MIN_CTAS CTAs per SM
constexpr int GEN_THREADS_WIDE = 512;  // fewer instances than SMs: one wide CTA per instance (every phase is spread over the warps at run time)
#ifndef GEN_MIN_CTAS
#define GEN_MIN_CTAS 2
#endif
#ifdef FMPC_PROF   /* phase cycles of thread 0 of every CTA (fmpc_last_profile): init, barrier + residuals, inv(Phi_uu), rhs, Schur assembly,
                      potrf + forward solve, panel, trailing update, backward substitution, dz, line search, accept + copy-out */
#define GPROF_DECL long long p_acc[12]; long long p_last = clock64(); for (int i_ = 0; i_ < 12; ++i_) p_acc[i_] = 0;
#define GPROF_T(idx) do { const long long now_ = clock64(); p_acc[idx] += now_ - p_last; p_last = now_; } while (0)
#else
#define GPROF_DECL
#define GPROF_T(idx) do { } while (0)
#endif

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
        L.Y = o; o += (NE + 1) * NE;                                       // + 1 row: the right-hand side of Y dnu = -beta
        L.total = (o + 15) & ~(size_t)15;
        return L;
    }
};

// leading dimension of a shared-memory operand of 8 x 4 DMMA fragments: a multiple of 4 with ld % 8 == 4, so that the 16 lanes of a
// half-warp (rows gq = 0..3 / 4..7, columns q = 0..3) hit 16 different 8-byte banks
__host__ __device__ __forceinline__ int gen_ld(int n) { int l = (n + 3) & ~3; if ((l & 7) != 4) l += 4; return l; }

// e / d for 0 <= e < 2^32 / d with the precomputed magic = ceil(2^32 / d) (the copy loops divide by a run-time leading dimension)
__device__ __forceinline__ unsigned gen_magic(int d) { return (unsigned)((0x100000000ull + (unsigned)d - 1) / (unsigned)d); }
__device__ __forceinline__ int gen_div(int e, unsigned magic) { return (int)__umulhi((unsigned)e, magic); }

// NR x NC accumulation chains: acc[r][u] += A_r(8 x K) B_u(8 x K)' over K columns, 4 per step -- NR + NC fragment loads feed NR * NC
// tensor-pipe products (2 x 4: 0.75 loads per product).  Ca[r] / Cb[u] point at this lane's fragment row; the main loop is
// straight-line (no guards) so that the loads of the next steps are issued ahead of the products; only a last partial step
// (K % 4) is guarded.  SCALE: the A fragments are scaled by cv (the diagonal of the middle factor).
template <int NR, int NC, bool SCALE>
__device__ __forceinline__ void gen_chains(double (&acc)[2][4][2], const double *(&Ca)[2], const double *cv, const double *(&Cb)[4], int K, int q, int k0 = 0)
{
    const int kf = K & ~3;                                  // columns [k0, K), k0 a multiple of 4
#pragma unroll 2
    for (int jb = k0; jb < kf; jb += 4) {
        const int j = jb + q;
        double av[NR];
#pragma unroll
        for (int r = 0; r < NR; ++r) av[r] = Ca[r][j];
        if (SCALE) {
            const double c = cv[j];
#pragma unroll
            for (int r = 0; r < NR; ++r) av[r] *= c;
        }
#pragma unroll
        for (int u = 0; u < NC; ++u) {
            const double bv = Cb[u][j];
#pragma unroll
            for (int r = 0; r < NR; ++r) dmma_gen(acc[r][u][0], acc[r][u][1], av[r], bv);
        }
    }
    if (kf < K) {
        const int j = kf + q;
        const bool ok = j < K;
        const double c = SCALE ? (ok ? cv[j] : 0.0) : 1.0;
        double av[NR];
#pragma unroll
        for (int r = 0; r < NR; ++r) av[r] = ok ? Ca[r][j] * c : 0.0;
#pragma unroll
        for (int u = 0; u < NC; ++u) {
            const double bv = ok ? Cb[u][j] : 0.0;
#pragma unroll
            for (int r = 0; r < NR; ++r) dmma_gen(acc[r][u][0], acc[r][u][1], av[r], bv);
        }
    }
}
template <int NR, bool SCALE>
__device__ __forceinline__ void gen_chains_n(int nc, double (&acc)[2][4][2], const double *(&Ca)[2], const double *cv, const double *(&Cb)[4], int K, int q, int k0 = 0)
{
    if (nc == 4) gen_chains<NR, 4, SCALE>(acc, Ca, cv, Cb, K, q, k0);
    else if (nc == 3) gen_chains<NR, 3, SCALE>(acc, Ca, cv, Cb, K, q, k0);
    else if (nc == 2) gen_chains<NR, 2, SCALE>(acc, Ca, cv, Cb, K, q, k0);
    else gen_chains<NR, 1, SCALE>(acc, Ca, cv, Cb, K, q, k0);
}

// The SCALE chains with the diagonal cv (K <= 256 doubles, K % 4 == 0) read once, coalesced, into registers (lane l holds cv[32 b + l])
// and handed to the k-steps by shuffles: the per-step cv load from the global scratch (L1 holds little next to the streaming Y)
// was the top stall of the Schur assembly.
template <int NR, int NC>
__device__ __forceinline__ void gen_chains_cv(double (&acc)[2][4][2], const double *(&Ca)[2], const double *cv, const double *(&Cb)[4], int K, int lane,
                                              int k0, int k1)
{
    // columns [k0, k1) only (multiples of 4): the operand blocks of the literal VAR_1 second row are zero outside a column range
    const int q = lane & 3;
    double cr[8];
#pragma unroll
    for (int b = 0; b < 8; ++b) cr[b] = (32 * b + lane < K) ? cv[32 * b + lane] : 0.0;
#pragma unroll
    for (int b = 0; b < 8; ++b) {
        if (32 * b >= k0 && 32 * b + 32 <= k1) {            // a whole block of 8 k-steps: straight-line, the loads run ahead of the products
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
                const int j = 32 * b + 4 * kk + q;
                const double c = __shfl_sync(0xffffffffu, cr[b], 4 * kk + q);
                double av[NR];
#pragma unroll
                for (int r = 0; r < NR; ++r) av[r] = Ca[r][j] * c;
#pragma unroll
                for (int u = 0; u < NC; ++u) {
                    const double bv = Cb[u][j];
#pragma unroll
                    for (int r = 0; r < NR; ++r) dmma_gen(acc[r][u][0], acc[r][u][1], av[r], bv);
                }
            }
        } else if (32 * b < k1 && 32 * b + 32 > k0) {
            const int kbeg = max(0, (k0 - 32 * b) >> 2), kend = min(8, (k1 - 32 * b) >> 2);
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
                const bool on = kk >= kbeg && kk < kend;    // warp-uniform; off-steps are skipped (addresses stay inside the row)
                const int j = 32 * b + (on ? 4 * kk : 4 * kbeg) + q;
                const double c = __shfl_sync(0xffffffffu, cr[b], 4 * kk + q);
                if (on) {
                    double av[NR];
#pragma unroll
                    for (int r = 0; r < NR; ++r) av[r] = Ca[r][j] * c;
#pragma unroll
                    for (int u = 0; u < NC; ++u) {
                        const double bv = Cb[u][j];
#pragma unroll
                        for (int r = 0; r < NR; ++r) dmma_gen(acc[r][u][0], acc[r][u][1], av[r], bv);
                    }
                }
            }
        }
    }
}
template <int NR>
__device__ __forceinline__ void gen_chains_cv_n(int nc, double (&acc)[2][4][2], const double *(&Ca)[2], const double *cv, const double *(&Cb)[4], int K, int lane,
                                                int k0, int k1)
{
    if (nc == 4) gen_chains_cv<NR, 4>(acc, Ca, cv, Cb, K, lane, k0, k1);
    else if (nc == 3) gen_chains_cv<NR, 3>(acc, Ca, cv, Cb, K, lane, k0, k1);
    else if (nc == 2) gen_chains_cv<NR, 2>(acc, Ca, cv, Cb, K, lane, k0, k1);
    else gen_chains_cv<NR, 1>(acc, Ca, cv, Cb, K, lane, k0, k1);
}

// In-place lower Cholesky of the n x n block Sm (leading dimension ld, lower triangle read), then in-place inverse of the factor:
// on return the lower triangle holds inv(L) and the strict upper triangle zeros.  One warp, a lane per row; left-looking so that
// column k costs one pass over the finished columns instead of a rank-1 update of the whole trailing block; the reciprocal
// square root replaces the square root + division pair.  rdiag (n) and tmp (n) are shared scratch.  Returns 0 or failing column + 1.
__device__ int gen_potrf_inv(double *Sm, int n, int ld, int lane, double *rdiag, double *tmp)
{
    for (int k = 0; k < n; ++k) {
        const double *rk = Sm + (size_t)k * ld;
        for (int r = k + lane; r < n; r += 32) {
            double *rr = Sm + (size_t)r * ld;
            double s0 = rr[k], s1 = 0.0;
            int j = 0;
            for (; j + 1 < k; j += 2) { s0 = fma(-rr[j], rk[j], s0); s1 = fma(-rr[j + 1], rk[j + 1], s1); }
            if (j < k) s0 = fma(-rr[j], rk[j], s0);
            rr[k] = s0 + s1;
        }
        __syncwarp();
        const double d = Sm[(size_t)k * ld + k];
        if (!(d > 0.0)) return k + 1;                           // uniform across the warp
        const double rs = rsqrt(d);
        __syncwarp();
        for (int r = k + lane; r < n; r += 32) Sm[(size_t)r * ld + k] *= rs;
        if (lane == 0) rdiag[k] = rs;
        __syncwarp();
    }
    // inv(L), last column first: X[j+1:, j] = -X[j+1:, j+1:] L[j+1:, j] / L[j][j]  (the trailing block is inverted already)
    for (int j = n - 1; j >= 0; --j) {
        const double xjj = rdiag[j];
        for (int r = j + 1 + lane; r < n; r += 32) {
            const double *rr = Sm + (size_t)r * ld;
            double s0 = 0.0, s1 = 0.0;
            int k = j + 1;
            for (; k + 1 <= r; k += 2) { s0 = fma(rr[k], Sm[(size_t)k * ld + j], s0); s1 = fma(rr[k + 1], Sm[(size_t)(k + 1) * ld + j], s1); }
            if (k <= r) s0 = fma(rr[k], Sm[(size_t)k * ld + j], s0);
            tmp[r] = -(s0 + s1) * xjj;
        }
        __syncwarp();
        for (int r = j + lane; r < n; r += 32) Sm[(size_t)r * ld + j] = (r == j) ? xjj : tmp[r];
        __syncwarp();
    }
    return 0;
}

// 1/x for a positive normal x: MUFU.RCP64H seed + two Newton steps, no slow path
__device__ __forceinline__ double gen_rcp_pos(double x)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    return fma(y, e, y);
}

// Same result as gen_potrf_inv for blocks whose padded size NP (= the leading dimension, <= 32) fits one lane per row: the row lives
// in registers, S = U D U' right-looking with the pivot chain through shuffles (per column: shfl -> rcp -> 2 FMAs; the rank-1
// update reads the column from a 32-double shared vector), then V = inv(U) column j by lane j and inv(L) = diag(1/sqrt(d)) V.
// The shared-memory version pays three shared-memory round trips and a reciprocal square root per column (24 k + 42 k cycles
// for n = 27, scripts/ubench/potrf.cu).  In place: U, then inv(L), overwrite Sm.  colbuf: 64 doubles, rsv: 32 doubles.
template <int NP>
__device__ __forceinline__ int gen_potrf_inv_reg(double *Sm, int n, int nrows, int lane, double *colbuf, double *rsv)
{
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int ld = NP;
    const int r = lane;
    double a[NP];
#pragma unroll
    for (int c = 0; c < NP; ++c) a[c] = (r < n && c < r) ? Sm[r * ld + c] : 0.0;
    double diag = (r < n) ? Sm[r * ld + r] : 1.0;
    double dpiv = 1.0;
    int info = 0;
#pragma unroll
    for (int k = 0; k < NP; ++k) {
        const double d = __shfl_sync(FULL, diag, k);        // pivot of column k (lane k's diagonal entry)
        if (!(d > 0.0) || !(d < 1.0e300)) { if (!info) info = k + 1; }
        if (lane == k) dpiv = d;
        const double at = a[k];                             // unscaled column entry
        if (k + 1 < NP) {
            double *cb = colbuf + (k & 1) * 32;
            cb[lane] = at;
            const double dinv = gen_rcp_pos(d);
            const double t = at * dinv;                     // U(r,k)
            a[k] = t;
            diag = fma(-t, at, diag);                       // own diagonal entry (only lanes > k use it)
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
    if (info) return info;                                  // uniform: d is the same in every lane
    rsv[lane] = rsqrt(dpiv);
    __syncwarp();
#pragma unroll
    for (int c = 0; c + 1 < NP; c += 2)
        if (r < NP) *reinterpret_cast<double2 *>(Sm + r * ld + c) = make_double2(a[c], a[c + 1]);
    __syncwarp();
    // V = inv(U): column j by lane j;  v[i] = -(sum_{k<i} U(i,k) v[k]) for i > j, v[j] = 1, v[i<j] = 0
    const int j = lane;
    double v[NP];
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int k = 0; k + 1 < i; k += 2) {
            const double2 p = *reinterpret_cast<const double2 *>(Sm + i * ld + k);
            s0 = fma(p.x, v[k], s0);
            s1 = fma(p.y, v[k + 1], s1);
        }
        if (i & 1) s0 = fma(Sm[i * ld + i - 1], v[i - 1], s0);
        v[i] = (i < j) ? 0.0 : ((i == j) ? 1.0 : -(s0 + s1));
    }
    __syncwarp();
    const bool live = (j < n);
#pragma unroll
    for (int i = 0; i < NP; ++i)
        if (j < NP && i < nrows) Sm[i * ld + j] = (live && i < n) ? v[i] * rsv[i] : 0.0;      // inv(L)(i,j) = v(i,j) / sqrt(d_i)
    return 0;
}

// one warp: factor + invert the diagonal block in bS (register version when the padded size fits a lane per row)
__device__ __noinline__ int gen_potrf_dispatch(double *bS, int n, int ld, int nrb, int lane, double *sm_part, double *sm_y, double *sm_vec)
{
    if (ld == 28) return gen_potrf_inv_reg<28>(bS, n, nrb, lane, sm_part, sm_part + 64);       // sm_part (>= 96 doubles): column buffer + 1/sqrt(d)
    if (ld == 20) return gen_potrf_inv_reg<20>(bS, n, nrb, lane, sm_part, sm_part + 64);
    if (ld == 12) return gen_potrf_inv_reg<12>(bS, n, nrb, lane, sm_part, sm_part + 64);
    if (ld == 4) return gen_potrf_inv_reg<4>(bS, n, nrb, lane, sm_part, sm_part + 64);
    return gen_potrf_inv(bS, n, ld, lane, sm_y, sm_vec);
}

// out = C v (- bv)   : one warp per group of 4 scalar rows of a block row (they share the window of v: 5 loads per 4 products and 4
// independent sums in flight), lanes stride over the row window (coalesced); per row the same summation order as a row per warp
// win_s: shared-memory copy of the window matrix at offset win_ptr of G.cw (or nullptr)
__device__ void gen_apply_C(const GenSys &G, int NB, const double *v, const double *bv, double *out, const double *win_s = nullptr, int win_ptr = -1)
{
    const int n = G.n, lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int ngrp = (n + 3) >> 2;
    for (int task = wid; task < NB * ngrp; task += nw) {
        const int i = task / ngrp, k0 = 4 * (task - i * ngrp), len = G.cw_len[i];
        const bool ws = win_s && G.cw_ptr[i] == win_ptr;
        const double *c = ws ? win_s : G.cw + G.cw_ptr[i];
        const double *vv = v + G.cw_off[i];
        const double *c0 = c + (size_t)k0 * len, *c1 = c + (size_t)min(k0 + 1, n - 1) * len;
        const double *c2 = c + (size_t)min(k0 + 2, n - 1) * len, *c3 = c + (size_t)min(k0 + 3, n - 1) * len;
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll 2
        for (int j = lane; j < len; j += 32) {
            const double x = vv[j];
            s0 = fma(c0[j], x, s0);
            s1 = fma(c1[j], x, s1);
            s2 = fma(c2[j], x, s2);
            s3 = fma(c3[j], x, s3);
        }
        s0 = warp_sum(s0); s1 = warp_sum(s1); s2 = warp_sum(s2); s3 = warp_sum(s3);
        if (lane < 4 && k0 + lane < n) {
            const int row = i * n + k0 + lane;
            const double sv = lane == 0 ? s0 : (lane == 1 ? s1 : (lane == 2 ? s2 : s3));
            out[row] = bv ? sv - bv[row] : sv;
        }
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
    if (G.qdiag) s = fma(__ldg(Q2 + k), x[k], s);               // the other products are exact zeros: same value as the full dot
    else for (int kk = 0; kk < n; ++kk) s = fma(__ldg(Q2 + kk), x[kk], s);
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
        if (G.qdiag) s = fma(__ldg(Qi + k), vx[k], s);
        else for (int kk = 0; kk < n; ++kk) s = fma(__ldg(Qi + kk), vx[kk], s);
        p[(size_t)t * st + m + k] = sign * s;
    }
}

template <int NT>
__global__ void __launch_bounds__(NT, NT == GEN_THREADS ? GEN_MIN_CTAS : 1) fmpc_solve_kernel_gen(const DevSys S, const GenSys G, const StepArgs A)
{
    extern __shared__ double smem[];
    const int n = G.n, m = G.m, T = G.T, N = G.N, st = n + m;
    const int NB = T + (A.has_xf ? 1 : 0), NE = NB * n;
    const int ld = gen_ld(n), n8 = (n + 7) & ~7, nrb = max(n8, ld);     // nrb: rows of the diagonal block buffer
    const unsigned mg_ld = gen_magic(ld), mg_n = gen_magic(n);
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5;
    const GenWs L = GenWs::make(n, m, T, G.ramp, G.dense_r);

    double *bS = smem;                                  // nrb x ld : diagonal block (rows / columns beyond n are zero)
    double *sm_y = bS + (size_t)nrb * ld;               // n
    double *sm_vec = sm_y + n;                          // n
    double *sm_part = sm_vec + n;                       // nt
    double *red = sm_part + GEN_THREADS_WIDE;           // 34  (one layout for both block sizes)
    double *panel = red + 34;                           // panel_rows x ld
    __shared__ int s_inst, s_flag, s_task;

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
        GPROF_DECL

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
        GPROF_T(0);
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

            GPROF_T(1);
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
                // per actuator: T = L D L' (unit lower bidiagonal L), then the inverse from  M[T-1][T-1] = 1/d_{T-1},
                // M[t][t] = 1/d_t + l_t^2 M[t+1][t+1]  (all terms positive),  M[a][b] = -l_a M[a+1][b] for a < b, mirrored.
                // A thread per actuator for the two T-step recurrences, then a thread per (column, actuator) for the products:
                // no thread reads back what it stored and nothing is divided in the second pass.
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
                    double mm = 1.0 / d;
                    minv[((size_t)(T - 1) * T + (T - 1)) * m + j] = mm;
                    for (int t = T - 2; t >= 0; --t) {
                        const double l = toff[(size_t)t * m + j];
                        mm = fma(l * l, mm, 1.0 / tdiag[(size_t)t * m + j]);
                        minv[((size_t)t * T + t) * m + j] = mm;
                    }
                }
                __syncthreads();
                for (int e = tid; e < T * m; e += nt) {
                    const int bcol = e / m, j = e - bcol * m;
                    double x = minv[((size_t)bcol * T + bcol) * m + j];
                    for (int a2 = bcol - 1; a2 >= 0; --a2) {
                        x *= -toff[(size_t)a2 * m + j];
                        minv[((size_t)a2 * T + bcol) * m + j] = x;
                        minv[((size_t)bcol * T + a2) * m + j] = x;
                    }
                }
            } else {
                for (int e = tid; e < T * m; e += nt) { const double d = tdiag[e]; if (!(d > 0.0)) s_flag = 1; minv[e] = 1.0 / d; }
            }
            __syncthreads();
            if (s_flag) { status = ST_NOT_PD; break; }
            GPROF_T(2);

            // ---- rhs of  Y dnu = -beta,  beta = -r_p + C inv(Phi) r_d  (:28-29) ----
            gen_apply_phi_inv(S, G, minv, rd, dz, 1.0);
            __syncthreads();
            gen_apply_C(G, NB, dz, nullptr, yv);
            __syncthreads();
            for (int e = tid; e < NE; e += nt) yv[e] = rp[e] - yv[e];
            GPROF_T(3);

            // ---- Y = Yx + C_u inv(Phi_uu) C_u'  (lower block triangle, full diagonal blocks) ----
            // One warp task per (block pair, tile row, group of 4 tile columns): the tiles of  sum_ab C_a diag(cv_ab) C_b'  are chains
            // of FP64 tensor-pipe products over the m actuators; the scaled A fragment is shared by the 4 tile columns (4 independent
            // accumulation chains).  The u block most block rows share (G.cu_main: -B) is staged in shared memory (the panel area is
            // free here) with a bank-conflict-free leading dimension -- from global memory every fragment load touches 8 cache lines
            // and the L1 tag stage, not the FP64 pipe, sets the pace.  When both operands are the same block the n x n block is
            // symmetric (diag(cv) is): only tiles on or below its diagonal are computed and mirrored.
            {
                const int gq = lane >> 2, q = lane & 3;
                const int ntl = n8 >> 3, ngrp = (ntl + 3) >> 2, nrp = (ntl + 1) >> 1;
                const int ldm = gen_ld(m);
                const bool staged = G.cu_main >= 0;                       // host side: fits the panel area, diagonal R
                if (staged) {
                    const double *src = G.cu + G.cu_main;
                    for (int e = tid; e < n8 * ldm; e += nt) {
                        const int r = e / ldm, j = e - r * ldm;
                        panel[e] = (r < n && j < m) ? __ldg(src + (size_t)r * m + j) : 0.0;
                    }
                }
                for (int e = tid; e < NE; e += nt) Y[(size_t)NE * ldy + e] = yv[e];       // the right-hand side: last row of the factorisation
                if (tid == 0) s_task = 0;
                __syncthreads();
                // tasks are handed out dynamically, most expensive first (host-side order): block rows with several u blocks (the
                // literal VAR_1 second row) cost up to 16 x a symmetric single-tile task
                for (;;) {
                    int slot = 0;
                    if (lane == 0) slot = atomicAdd(&s_task, 1);
                    slot = __shfl_sync(0xffffffffu, slot, 0);
                    if (slot >= G.nsch) break;
                    const int task = G.sch[slot];
                    const int pr = task / (nrp * ngrp), tl = task - pr * nrp * ngrp, rp2 = tl / ngrp, ct0 = 4 * (tl - rp2 * ngrp);
                    int i = 0;
                    while ((i + 1) * (i + 2) / 2 <= pr) ++i;
                    const int k = pr - i * (i + 1) / 2;
                    if (i >= NB) continue;                                    // terminal block row of a problem solved without xf
                    const int na = G.ue_cnt[i], nbk = G.ue_cnt[k];
                    const bool sym = na == 1 && nbk == 1 && G.ue_ptr[4 * i] == G.ue_ptr[4 * k] && (G.ramp || G.ue_t[4 * i] == G.ue_t[4 * k]);
                    const int rt0 = 2 * rp2, nr = min(2, ntl - rt0);          // two tile rows x up to four tile columns per task
                    int nc = min(4, ntl - ct0);
                    if (sym) nc = min(nc, rt0 + nr - 1 - ct0 + 1);
                    if (nc <= 0) continue;                                    // the whole group lies above the diagonal of a symmetric block
                    // the products accumulate on top of the constant x part of the tiles (read here, ahead of the chains)
                    double acc[2][4][2];
#pragma unroll
                    for (int r2 = 0; r2 < 2; ++r2) {
                        const int ra = 8 * (rt0 + r2) + gq;
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int c = 8 * (ct0 + u) + 2 * q;
#pragma unroll
                            for (int v = 0; v < 2; ++v)
                                acc[r2][u][v] = (r2 < nr && u < nc && ra < n && c + v < n) ? __ldg(G.Yx + ((size_t)i * n + ra) * G.ldyx + (size_t)k * n + c + v) : 0.0;
                        }
                    }
                    for (int a = 0; a < na; ++a) {
                        const int ta = G.ue_t[4 * i + a], pa = G.ue_ptr[4 * i + a];
                        const bool sa = staged && pa == G.cu_main;
                        const double *Ca[2];
#pragma unroll
                        for (int r2 = 0; r2 < 2; ++r2) {
                            const int ra = min(8 * (rt0 + min(r2, nr - 1)) + gq, n8 - 1);
                            Ca[r2] = G.dense_r ? Eb + ((size_t)i * 4 + a) * n * m + (size_t)min(ra, n - 1) * m
                                   : sa ? panel + (size_t)ra * ldm : G.cu + pa + (size_t)min(ra, n - 1) * m;
                        }
                        for (int bb = 0; bb < nbk; ++bb) {
                            const int tb = G.ue_t[4 * k + bb], pb = G.ue_ptr[4 * k + bb];
                            if (!G.ramp && ta != tb) continue;
                            // dense R: the A operand is a row of E = C_a inv(Phi_uu) already
                            const double *cv = G.dense_r ? nullptr : (G.ramp ? minv + ((size_t)ta * T + tb) * m : minv + (size_t)ta * m);
                            const bool sb = staged && pb == G.cu_main;
                            const double *Cb[4];
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                const int cb = min(8 * (ct0 + u) + gq, n8 - 1);
                                Cb[u] = sb ? panel + (size_t)cb * ldm : G.cu + pb + (size_t)min(cb, n - 1) * m;
                            }
                            // columns where both blocks have entries (the blocks of the literal second row are zero outside a range);
                            // dense R: the A operand is E = C_a inv(Phi_uu), dense in every column
                            const int k0 = G.dense_r ? G.ue_lo[4 * k + bb] : max(G.ue_lo[4 * i + a], G.ue_lo[4 * k + bb]);
                            const int k1 = G.dense_r ? G.ue_hi[4 * k + bb] : min(G.ue_hi[4 * i + a], G.ue_hi[4 * k + bb]);
                            if (k0 >= k1) continue;
                            if (cv && m <= 256 && (m & 3) == 0) gen_chains_cv_n<2>(nc, acc, Ca, cv, Cb, m, lane, k0, k1);
                            else if (cv) gen_chains_n<2, true>(nc, acc, Ca, cv, Cb, min(k1, m), q, k0);
                            else gen_chains_n<2, false>(nc, acc, Ca, cv, Cb, min(k1, m), q, k0);
                        }
                    }
#pragma unroll
                    for (int r2 = 0; r2 < 2; ++r2) {
                        if (r2 >= nr) continue;
                        const int rt = rt0 + r2, r = 8 * rt + gq;
                        if (r >= n) continue;
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            if (u >= nc) continue;
                            const int ct = ct0 + u, c = 8 * ct + 2 * q;
                            if (sym && ct > rt) continue;                     // upper tile of a symmetric block: its mirror image is written below
#pragma unroll
                            for (int v = 0; v < 2; ++v) {
                                if (c + v >= n) continue;
                                Y[((size_t)i * n + r) * ldy + (size_t)k * n + c + v] = acc[r2][u][v];
                                if (sym && ct < rt)                           // mirror inside the symmetric u part of the block
                                    Y[((size_t)i * n + c + v) * ldy + (size_t)k * n + r] =
                                        acc[r2][u][v] + __ldg(G.Yxd + ((size_t)i * n + r) * G.ldyx + (size_t)k * n + c + v);
                            }
                        }
                    }
                }
            }
            __syncthreads();
            GPROF_T(4);

            // ---- dense blocked Cholesky of Y with the right-hand side as one more row (:30-31) ----
            // Per block column K: the diagonal block goes to shared memory, warp 0 factors it and inverts the factor in place while the
            // other warps stage the panel (all rows below, right-hand-side row included); panel <- panel inv(L_KK)' and the trailing
            // update  Y[r,c] -= sum_j P[r,j] P[c,j]  run on the FP64 tensor pipe from shared memory (8 x 8 tiles, conflict-free
            // leading dimension).  inv(L_KK) replaces L_KK in Y: the backward substitution multiplies by it instead of solving with
            // L_KK.  After the last block the right-hand-side row holds y = inv(L) rhs.
            bool fail = false;
            for (int K = 0; K < NB; ++K) {
                const int gq = lane >> 2, q = lane & 3, nw = nt >> 5, ntl = n8 >> 3, nks = (n + 3) >> 2;
                const int r0 = (K + 1) * n, R = NE - r0, Rp = R + 1;      // trailing rows; + the right-hand-side row
                double *Ykk = Y + ((size_t)K * n) * ldy + (size_t)K * n;
                double *Yp0 = Y + (size_t)r0 * ldy + (size_t)K * n;
                const bool whole = (Rp <= G.panel_rows);
                // K > 0: the diagonal block was updated, factored and inverted by warp 0 during the trailing update of the step before
                // (look-ahead: the serial factorisation of the small block is off the critical path); all warps stage the panel
                if (K == 0) {
                    for (int e = tid; e < nrb * ld; e += nt) {
                        const int r = gen_div(e, mg_ld), c = e - r * ld;
                        const double v = Ykk[(size_t)min(r, n - 1) * ldy + min(c, n - 1)];
                        bS[e] = (r < n && c <= r) ? v : 0.0;
                    }
                    if (tid == 0) s_flag = 0;
                    __syncthreads();
                }
                if (K == 0 && wid == 0) {
                    const int info = gen_potrf_dispatch(bS, n, ld, nrb, lane, sm_part, sm_y, sm_vec);
                    if (lane == 0) s_flag = info;
                } else {
                    const int rc = min(G.panel_rows, Rp);
                    const int t0 = (K == 0) ? tid - 32 : tid, ts = (K == 0) ? nt - 32 : nt;
#pragma unroll 4
                    for (int e = t0; e < rc * ld; e += ts) {
                        const int r = gen_div(e, mg_ld), c = e - r * ld;
                        const double v = Yp0[(size_t)r * ldy + min(c, n - 1)];
                        panel[e] = (c < n) ? v : 0.0;
                    }
                }
                __syncthreads();
                if (s_flag) { fail = true; break; }
                GPROF_T(5);
                for (int e = tid; e < n * n; e += nt) { const int r = gen_div(e, mg_n), c = e - r * n; Ykk[(size_t)r * ldy + c] = bS[r * ld + c]; }
                for (int c0 = 0; c0 < Rp; c0 += G.panel_rows) {
                    const int rc = min(G.panel_rows, Rp - c0);
                    double *Yp = Yp0 + (size_t)c0 * ldy;
                    if (c0 > 0) {
                        __syncthreads();
#pragma unroll 4
                        for (int e = tid; e < rc * ld; e += nt) {
                            const int r = gen_div(e, mg_ld), c = e - r * ld;
                            const double v = Yp[(size_t)r * ldy + min(c, n - 1)];
                            panel[e] = (c < n) ? v : 0.0;
                        }
                        __syncthreads();
                    }
                    // in place, tile columns from the last to the first: tile column ct reads columns < 8 ct + 8 only
                    for (int t8 = wid; 8 * t8 < rc; t8 += nw) {
                        const int prow = 8 * t8 + gq;
                        double *pr = panel + (size_t)min(prow, rc - 1) * ld;
                        for (int ct = ntl - 1; ct >= 0; --ct) {
                            const double *bl = bS + (size_t)(8 * ct + gq) * ld;
                            const int ne = min(2 * (ct + 1), nks);
                            double c0a = 0.0, c1a = 0.0, c0b = 0.0, c1b = 0.0;
                            int ks = 0;
                            for (; ks + 1 < ne; ks += 2) {
                                dmma_gen(c0a, c1a, pr[4 * ks + q], bl[4 * ks + q]);
                                dmma_gen(c0b, c1b, pr[4 * ks + 4 + q], bl[4 * ks + 4 + q]);
                            }
                            if (ks < ne) dmma_gen(c0a, c1a, pr[4 * ks + q], bl[4 * ks + q]);
                            __syncwarp();
                            const int c = 8 * ct + 2 * q;
                            if (prow < rc) {
                                if (c < n) pr[c] = c0a + c0b;
                                if (c + 1 < n) pr[c + 1] = c1a + c1b;
                            }
                            __syncwarp();
                        }
                    }
                    __syncthreads();
                    for (int e = tid; e < rc * n; e += nt) { const int r = gen_div(e, mg_n), c = e - r * n; Yp[(size_t)r * ldy + c] = panel[r * ld + c]; }
                }
                __syncthreads();
                GPROF_T(6);
                // trailing update: a warp per tile row, 4 tile columns at a time (4 independent chains), Y tiles read ahead of the products
                auto trailing = [&](const double *P, const size_t ldp) {
                    const int NT8 = (Rp + 7) >> 3;
                    double *Yt = Y + (size_t)r0 * ldy + r0;
                    // a warp task: two tile rows x four tile columns (8 chains fed by 6 fragment loads per step); its Y tiles are read
                    // ahead of the 56 products.  Warp 0 takes the row pairs of the next diagonal block, then factors it (look-ahead).
                    const int NP2 = (NT8 + 1) >> 1, npl = (ntl + 1) >> 1;
                    const int pstep = (wid == 0) ? 1 : nw - 1, plim = (wid == 0) ? min(npl, NP2) : NP2;
                    for (int p2 = (wid == 0) ? 0 : npl + wid - 1; p2 < plim; p2 += pstep) {
                        const int ti0 = 2 * p2, nr = min(2, NT8 - ti0), til = ti0 + nr - 1;
                        const double *pa[2];
#pragma unroll
                        for (int r2 = 0; r2 < 2; ++r2) pa[r2] = P + (size_t)min(8 * (ti0 + min(r2, nr - 1)) + gq, Rp - 1) * ldp;
                        for (int tj0 = 0; tj0 <= til; tj0 += 4) {
                            const int nc = min(4, til - tj0 + 1);
                            const double *pb[4];
                            double acc[2][4][2], yo[2][4][2];
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                pb[u] = P + (size_t)min(8 * (tj0 + u) + gq, Rp - 1) * ldp;
                                const int c = 8 * (tj0 + u) + 2 * q;
#pragma unroll
                                for (int r2 = 0; r2 < 2; ++r2) {
                                    const int r = 8 * (ti0 + r2) + gq;
                                    acc[r2][u][0] = acc[r2][u][1] = 0.0;
#pragma unroll
                                    for (int v = 0; v < 2; ++v)
                                        yo[r2][u][v] = (r2 < nr && u < nc && r < Rp && c + v <= r && c + v < R) ? Yt[(size_t)r * ldy + c + v] : 0.0;
                                }
                            }
                            gen_chains_n<2, false>(nc, acc, pa, nullptr, pb, n, q);
#pragma unroll
                            for (int r2 = 0; r2 < 2; ++r2) {
                                const int r = 8 * (ti0 + r2) + gq;
#pragma unroll
                                for (int u = 0; u < 4; ++u) {
                                    const int c = 8 * (tj0 + u) + 2 * q;
#pragma unroll
                                    for (int v = 0; v < 2; ++v)
                                        if (r2 < nr && u < nc && r < Rp && c + v <= r && c + v < R) {
                                            const double val = yo[r2][u][v] - acc[r2][u][v];
                                            if (r < n) bS[r * ld + c + v] = val;          // the next diagonal block stays on chip (inv(L_KK) is dead)
                                            else Yt[(size_t)r * ldy + c + v] = val;
                                        }
                                }
                            }
                        }
                    }
                    if (wid == 0) {
                        __syncwarp();
                        const int info = gen_potrf_dispatch(bS, n, ld, nrb, lane, sm_part, sm_y, sm_vec);
                        if (lane == 0 && info) s_flag = info;
                    }
                };
                if (R > 0) {                                                  // two instances: the shared-memory one uses shared loads
#ifndef GEN_LDS
                    trailing(whole ? panel : Yp0, whole ? (size_t)ld : ldy);
#else
                    if (whole) trailing(panel, (size_t)ld);
                    else trailing(Yp0, ldy);
#endif
                }
                __syncthreads();
                GPROF_T(7);
            }
            if (fail) { status = ST_NOT_PD; break; }

            // ---- backward substitution  L' dnu = y  (:32): dnu_K = inv(L_KK)' (y_K - sum_{r below} L[r, K]' dnu[r]) ----
            const int parts = (8 * n <= nt) ? 8 : ((4 * n <= nt) ? 4 : ((2 * n <= nt) ? 2 : 1));     // lanes per entry of the inv(L_KK)' product
            for (int K = NB - 1; K >= 0; --K) {
                const int r0 = (K + 1) * n;
                const int groups = max(1, nt / n);
                const int c = tid % n, g = tid / n;
                double s = 0.0;
                if (g < groups) {
                    const double *yc = Y + (size_t)K * n + c;
#pragma unroll 8
                    for (int r = r0 + g; r < NE; r += groups) s = fma(yc[(size_t)r * ldy], dnu[r], s);
                }
                sm_part[tid] = (g < groups) ? s : 0.0;
                __syncthreads();
                for (int k = tid; k < n; k += nt) {
                    double acc = Y[(size_t)NE * ldy + (size_t)K * n + k];
                    for (int gg = 0; gg < groups; ++gg) acc -= sm_part[gg * n + k];
                    sm_vec[k] = acc;
                }
                __syncthreads();
                const double *Xd = Y + ((size_t)K * n) * ldy + (size_t)K * n;
                for (int e0 = 0; e0 < n * parts; e0 += nt) {
                    const int e = e0 + tid, k = e / parts, part = e - k * parts;
                    double acc = 0.0;
                    if (k < n)
#pragma unroll 4
                        for (int j = k + part; j < n; j += parts) acc = fma(Xd[(size_t)j * ldy + k], sm_vec[j], acc);
                    for (int o = parts >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                    if (k < n && part == 0) dnu[K * n + k] = acc;
                }
                __syncthreads();
            }

            GPROF_T(8);
            // ---- dz = inv(Phi)(-r_d - C' dnu)  (:34-35) ----
            gen_apply_Ct(G, NB, dnu, hd);
            __syncthreads();
            for (int c = tid; c < N; c += nt) zt[c] = rd[c] + hd[c];
            __syncthreads();
            gen_apply_phi_inv(S, G, minv, zt, dz, -1.0);
            __syncthreads();

            GPROF_T(9);
            // ---- backtracking on ||[r_p; r_d]||, d frozen (backtracking_inf_newton.m:2-11) ----
            // The trial point and the window of C most block rows share live in shared memory for the whole search (the panel
            // area is free): every trial re-reads both, and from the global scratch each pass waits on L2.
            double t = 1.0;
            int nh = 0;
            const double *win_s = nullptr;
            double *zts = zt;
            if (G.ls_stage) {
                int wl = 0;
                for (int i = 0; i < NB; ++i) if (G.cw_ptr[i] == G.cw_main) wl = G.cw_len[i];
                for (int e = tid; e < n * wl; e += nt) panel[e] = __ldg(G.cw + G.cw_main + e);
                win_s = panel;
                zts = panel + (size_t)n * wl;
            }
            for (;;) {
                for (int c = tid; c < N; c += nt) zts[c] = __fma_rn(t, dz[c], z[c]);
                __syncthreads();
                gen_apply_C(G, NB, zts, bv, rpt, win_s, G.cw_main);
                double sdt = 0.0;
                for (int c = tid; c < N; c += nt) { const double r = gen_rd_elem(S, G, c, zts, __fma_rn(t, hd[c], h[c]), pd); sdt = fma(r, r, sdt); }
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
            GPROF_T(10);
            for (int c = tid; c < N; c += nt) { z[c] = zts[c]; h[c] = __fma_rn(t, hd[c], h[c]); }
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
        GPROF_T(11);
#ifdef FMPC_PROF
        if (A.prof && tid == 0)
            for (int i_ = 0; i_ < 12; ++i_) atomicAdd((unsigned long long *)A.prof + i_, (unsigned long long)p_acc[i_]);
#endif
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

size_t gen_fixed_smem_doubles(int n) { const int n8 = (n + 7) & ~7, ld = gen_ld(n); return (size_t)(n8 > ld ? n8 : ld) * ld + 2 * (size_t)n + GEN_THREADS_WIDE + 34; }

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
    std::vector<int> ue_cnt(NBm, 0), ue_t(4 * NBm, 0), ue_ptr(4 * NBm, 0), ue_lo(4 * NBm, 0), ue_hi(4 * NBm, 0);
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
            int jlo = m, jhi = 0;                                 // non-zero column range, widened to multiples of 4
            for (int r = 0; r < n; ++r)
                for (int j = 0; j < m; ++j) if (blk[(size_t)r * m + j] != 0.0) { jlo = std::min(jlo, j); jhi = std::max(jhi, j + 1); }
            ue_lo[4 * i + ue_cnt[i]] = jlo & ~3; ue_hi[4 * i + ue_cnt[i]] = std::min((jhi + 3) & ~3, (m + 3) & ~3);
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
    G.qdiag = qdiag ? 1 : 0;
    bool ok = true;
#define GUP(field, vec) do { G.field = gen_upload(allocs, vec); if (!G.field) ok = false; } while (0)
    GUP(cw, cw); GUP(cw_ptr, cw_ptr); GUP(cw_off, cw_off); GUP(cw_len, cw_len);
    GUP(cu, cu); GUP(cut, cut); GUP(ue_cnt, ue_cnt); GUP(ue_t, ue_t); GUP(ue_ptr, ue_ptr); GUP(ue_lo, ue_lo); GUP(ue_hi, ue_hi);
    GUP(Yx, Yx); GUP(Q2, Q2); GUP(Q2f, Q2f); GUP(Qi, Qi); GUP(Qif, Qif);
    std::vector<double> dumin(m, 0.0), dumax(m, 0.0);
    if (s->ramp_rows) { dumin.assign(s->du_min, s->du_min + m); dumax.assign(s->du_max, s->du_max + m); }
    GUP(dumin, dumin); GUP(dumax, dumax);
    if (G.dense_r) GUP(R2, R2);
#undef GUP
    if (!ok) return FMPC_ERR_CUDA;

    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return FMPC_ERR_CUDA;
    const int ld = gen_ld(n);
    const size_t fixed = gen_fixed_smem_doubles(n) * 8;
    const size_t budget = (size_t)prop.sharedMemPerBlockOptin - 1024;
    if (fixed + 8 * (size_t)ld * 8 > budget) return FMPC_ERR_UNSUPPORTED;
    int prow = (int)((budget - fixed) / ((size_t)ld * 8));
    if (prow > NE - n + 1) prow = NE - n + 1;             // every row below the first block column + the right-hand-side row
    if (prow < 8) prow = 8;
    G.panel_rows = prow;
    size_t smem = fixed + (size_t)prow * ld * 8;
    if (G.dense_r) {        // the panel area doubles as the packed Cholesky factor of Phi_uu and its inverse (2 x m(m+1)/2 doubles)
        const size_t need = fixed + (size_t)m * (m + 1) * 8;
        if (need > budget) return FMPC_ERR_UNSUPPORTED;
        if (need > smem) smem = need;
    }
    // Schur assembly: the u block most block rows reference (-B) is staged in the panel area when it fits
    G.cu_main = -1;
    if (!G.dense_r) {
        int best = 0;
        for (int i = 0; i < NBm; ++i)
            for (int a = 0; a < ue_cnt[i]; ++a) {
                int cnt = 0;
                for (int k = 0; k < NBm; ++k) for (int b = 0; b < ue_cnt[k]; ++b) cnt += (ue_ptr[4 * k + b] == ue_ptr[4 * i + a]);
                if (cnt > best) { best = cnt; G.cu_main = ue_ptr[4 * i + a]; }
            }
        const size_t need = fixed + (size_t)((n + 7) & ~7) * gen_ld(m) * 8;
        if (need > budget) G.cu_main = -1;
        else if (need > smem) smem = need;
    }
    {   // line search: the window most block rows share + the trial point in the panel area
        int best = 0, wl = 0;
        G.cw_main = -1;
        for (int i = 0; i < NBm; ++i) {
            int cnt = 0;
            for (int k = 0; k < NBm; ++k) cnt += (cw_ptr[k] == cw_ptr[i]);
            if (cnt > best) { best = cnt; G.cw_main = cw_ptr[i]; wl = cw_len[i]; }
        }
        const size_t need = fixed + ((size_t)n * wl + (size_t)N) * 8;
        G.ls_stage = (G.cw_main >= 0 && need <= budget) ? 1 : 0;
        if (G.ls_stage && need > smem) {
            if (need <= smem + 16384) smem = need;               // a little more shared memory is fine, a lot would cost a resident CTA
            else G.ls_stage = 0;
        }
    }
    {   // Schur assembly tasks (pair, two tile rows, four tile columns), most expensive first
        const int ntl = ((n + 7) & ~7) >> 3, ngrp = (ntl + 3) >> 2, nrp = (ntl + 1) >> 1;
        std::vector<std::pair<int, int>> tk;                     // (-cost, task)
        for (int i = 0, pr = 0; i < NBm; ++i)
            for (int k = 0; k <= i; ++k, ++pr) {
                const bool sym = ue_cnt[i] == 1 && ue_cnt[k] == 1 && ue_ptr[4 * i] == ue_ptr[4 * k] && (G.ramp || ue_t[4 * i] == ue_t[4 * k]);
                int combos = 0;
                for (int a = 0; a < ue_cnt[i]; ++a) for (int b = 0; b < ue_cnt[k]; ++b) combos += (G.ramp || ue_t[4 * i + a] == ue_t[4 * k + b]);
                for (int rp = 0; rp < nrp; ++rp)
                    for (int cg = 0; cg < ngrp; ++cg) {
                        const int nr = std::min(2, ntl - 2 * rp);
                        int nc = std::min(4, ntl - 4 * cg);
                        if (sym) nc = std::min(nc, 2 * rp + nr - 4 * cg);
                        if (nc <= 0) continue;
                        const bool fast = (G.cu_main >= 0 && sym && ue_ptr[4 * i] == G.cu_main);       // both operands staged
                        tk.push_back({-(combos * nc * (fast ? 2 : 3) + 1), (pr * nrp + rp) * ngrp + cg});
                    }
            }
        std::sort(tk.begin(), tk.end());
        std::vector<int> sch(tk.size());
        for (size_t e = 0; e < tk.size(); ++e) sch[e] = tk[e].second;
        G.nsch = (int)sch.size();
        G.sch = gen_upload(allocs, sch);
        if (!G.sch) return FMPC_ERR_CUDA;
        std::vector<double> Yxd((size_t)NE * NE, 0.0);           // per block: transpose minus the block
        for (int i = 0; i < NBm; ++i)
            for (int k = 0; k < NBm; ++k)
                for (int r = 0; r < n; ++r)
                    for (int c = 0; c < n; ++c)
                        Yxd[(size_t)(i * n + r) * NE + (size_t)k * n + c] = Yx[(size_t)(i * n + c) * NE + (size_t)k * n + r] - Yx[(size_t)(i * n + r) * NE + (size_t)k * n + c];
        G.Yxd = gen_upload(allocs, Yxd);
        if (!G.Yxd) return FMPC_ERR_CUDA;
    }
    if (cudaFuncSetAttribute(fmpc_solve_kernel_gen<GEN_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
        cudaFuncSetAttribute(fmpc_solve_kernel_gen<GEN_THREADS_WIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return FMPC_ERR_CUDA;
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fmpc_solve_kernel_gen<GEN_THREADS>, GEN_THREADS, smem) != cudaSuccess || per_sm < 1)
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
    if (const char *e = getenv("FMPC_GEN_GRID")) { const int g = atoi(e); if (g > 0 && g < grid) grid = g; }     // experiments: fewer resident CTAs
    if (grid < 1) grid = 1;
    // fewer instances than half the resident CTAs: every instance gets a 512-thread CTA of its own SM (single closed loops)
    const bool wide = A.nbatch * 2 <= cfg.grid && !getenv("FMPC_GEN_NARROW");
    if (wide) fmpc_solve_kernel_gen<GEN_THREADS_WIDE><<<grid, GEN_THREADS_WIDE, cfg.smem, (cudaStream_t)stream>>>(S, G, A);
    else fmpc_solve_kernel_gen<GEN_THREADS><<<grid, cfg.block, cfg.smem, (cudaStream_t)stream>>>(S, G, A);
}
