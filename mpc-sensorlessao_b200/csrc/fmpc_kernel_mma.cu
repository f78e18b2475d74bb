// fastMPC batched Newton solve -- CTA-per-instance DMMA path (sm_100a, fp64): n <= 72.
//   n <= 32 : 128 threads, 5 CTAs per SM (kept for A/B experiments: the warp-per-instance kernel is the default there)
//   32 < n <= 72 : 256 threads, 1 CTA per SM, five 72 x 68 operand blocks in shared memory (196 KB) -- the kernel for the
//                  66-mode / horizon-30 configuration (BASELINE.json configs[4]); the diagonal block is factored and
//                  inverted by the whole CTA (cta_potrf_inverse) instead of one warp with a row per lane.
//
// Same algorithm and per-instance control flow as the generic kernel in fmpc_kernels.cu (one persistent
// CTA per MPC instance; inf_newton_solver.m:10-41 on the block structure), but every dense contraction
// runs on the FP64 tensor pipe (mma.sync.aligned.m8n8k4.f64 = DMMA, measured 37.2 TFLOP/s on B200 = the
// FP64 roofline) with operands staged in shared memory (row-major, leading dimension = 4 mod 8 doubles:
// conflict-free fragment loads):
//   * every application of C / C' (r_p = Cz - b, C inv(Phi) r_d, C' dnu, trial residuals) is a set of GEMMs
//     over the horizon:  B U, A1 X, A2 X (n x T) and B' N, A1' N, A2' N (m x T, n x T); the elementwise
//     work (barrier terms, inv(Phi) scaling, trial point) is fused into the staging / epilogues;
//   * B diag(w_t) B' for ALL stages as ONE GEMM  G (npairs x m) * W (m x T),  G[p][j] = B(r,j) B(c,j)
//     precomputed on the host: symmetric half only, no 28 -> 32 padding;
//   * band-2 block Cholesky of the Schur complement: syrk / gemm updates as 8 x 8 x 4 tiles on five
//     shared-memory blocks; the diagonal block is factored by ONE warp with rows in registers
//     (S = U D U', pivot chain through shuffles) and inverted explicitly, so both triangular solves with
//     n right-hand sides become DMMA products with inv(L)' and the substitutions become GEMVs.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "fmpc_internal.h"
#include "fmpc_device.cuh"

using namespace fmpc_dev;

namespace {

#ifdef FMPC_PROF_BIG     /* phase timing of the n > 32 kernel: thread 0 of CTA 0, first instance, printed once */
#define PB_DECL long long pb_acc[10] = {0,0,0,0,0,0,0,0,0,0}; long long pb_last = clock64();
#define PB_T(idx) do { if (NP > 32 && tid == 0) { const long long now_ = clock64(); pb_acc[idx] += now_ - pb_last; pb_last = now_; } } while (0)
#define PB_PRINT do { if (NP > 32 && tid == 0 && blockIdx.x == 0) printf("phase cycles: pre %lld gemmG %lld zero %lld ph1 %lld potrf %lld ph3 %lld back %lld dz %lld ls %lld rest %lld\n", pb_acc[0], pb_acc[1], pb_acc[2], pb_acc[3], pb_acc[4], pb_acc[5], pb_acc[6], pb_acc[7], pb_acc[8], pb_acc[9]); } while (0)
#else
#define PB_DECL
#define PB_T(idx) do { } while (0)
#define PB_PRINT do { } while (0)
#endif
#ifndef FMPC_BIG_THREADS
#define FMPC_BIG_THREADS 256    // threads per CTA of the n > 32 kernel
#endif
#ifndef FMPC_BW_SMEM
#define FMPC_BW_SMEM 0      // backward substitution of the n > 32 kernel with the factor blocks prefetched by cp.async into shared
                            // memory: parity-tested, but measured no faster on B200 at the C5 shape (51.5 k vs 49-53 k solves/s)
#endif
constexpr int MAXTT = 4;             // horizon (column) tiles accumulated together by one warp (T <= 32 in one pass)
template <int NP> struct KCfg {
    static constexpr int NTH = (NP > 32) ? FMPC_BIG_THREADS : 128;       // threads per CTA (9 warps for the 9 row tiles of a 72-row block measured no faster: registers)
    static constexpr int MINB = (NP > 32) ? 1 : 5;          // resident CTAs per SM the register budget is set for
    static constexpr int KSMAX = (NP + 3) / 4;              // k-steps of a block product
    static constexpr int RMAX = (NP > 32 && FMPC_BIG_THREADS < 288) ? 2 : 1;          // rows per 4-lane group in the GEMV phases
};

__device__ __forceinline__ void dmma(double &c0, double &c1, const double a, const double b)
{
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

struct Geom {
    int n, KP, ld, RP, nt, ks, lds, ldu, mp, ldp, TP, ntt;
    __host__ __device__ static int kp_of(int n) { return n <= 8 ? 8 : (n <= 16 ? 16 : (n <= 28 ? 28 : (n <= 32 ? 32 : ((n + 7) & ~7)))); }   // n > 32: whole 8-column tiles (cta_potrf_inverse)
    __host__ __device__ static Geom make(int n, int m, int T)
    {
        Geom g;
        g.n = n;
        g.KP = kp_of(n);                                  // compile-time sizes of the kernel instantiations
        g.ld = (g.KP % 8 == 4) ? g.KP : g.KP + 4;
        g.RP = (g.KP + 7) & ~7;
        g.nt = g.RP / 8;
        g.ks = g.KP / 4;
        g.lds = (n <= 32) ? g.KP + 1 : g.ld;              // S scratch (n <= 32: odd, row-per-lane loads are conflict-free)
        g.ldu = g.KP + 2;                                 // U scratch (16 B aligned rows, conflict-free 128-bit row stores)
        g.mp = (m + 3) & ~3;
        g.ldp = (g.mp % 8 == 4) ? g.mp : g.mp + 4;
        g.TP = (T + 7) & ~7;
        g.ntt = g.TP / 8;
        return g;
    }
    __host__ __device__ size_t blk() const { return (size_t)RP * ld; }
    __host__ __device__ size_t ops_doubles() const
    {
        size_t a = 5 * blk();
        const size_t b = (size_t)TP * ldp + (size_t)(TP + 2) * ld;   // staging: U-like operand + X-like operand
        if (b > a) a = b;
        return (a + 1) & ~(size_t)1;
    }
    __host__ __device__ int vlen() const { return RP < 32 ? 32 : RP; }
    __host__ __device__ size_t smem_doubles() const { return ops_doubles() + (size_t)vlen() * 4 + 64 + 36 + (n > 32 ? 128 + 64 * (FMPC_BIG_THREADS / 32) : 0); }
};

// C(8x8 tile) += A(rows.., k) * B(rows.., k)'   over ksteps k-steps of 4, operands in shared memory
__device__ __forceinline__ void tile_nt(double &c0, double &c1, const double *A, const double *Bm, int ksteps)
{
    // two accumulation chains (even / odd k-steps): a dependent DMMA costs 28 cycles, an independent one 16
    double e0 = 0.0, e1 = 0.0;
    int k = 0;
#pragma unroll 2
    for (; k + 1 < ksteps; k += 2) {
        dmma(c0, c1, A[4 * k], Bm[4 * k]);
        dmma(e0, e1, A[4 * k + 4], Bm[4 * k + 4]);
    }
    if (k < ksteps) dmma(c0, c1, A[4 * k], Bm[4 * k]);
    c0 += e0; c1 += e1;
}

// acc[tt] += A(row arow of a GLOBAL row-major matrix) * Bs(shared rows 8 tt + .., ldb)'
//   Ag : points at A(arow, q);  Bs : points at Bs(row of tile 0, q);  entries with column >= Kv read as 0
__device__ __forceinline__ void mma_gA_sB(double (&acc)[MAXTT][2], const double *Ag, bool arow_ok, int Kv, int ksteps,
                                          const double *Bs, int ldb, int ntt_live, int q)
{
    for (int k0 = 0; k0 < ksteps; k0 += 12) {
        double af[12];
#pragma unroll
        for (int kk = 0; kk < 12; ++kk) {
            const int kc = 4 * (k0 + kk) + q;
            af[kk] = (k0 + kk < ksteps && arow_ok && kc < Kv) ? __ldg(Ag + 4 * (k0 + kk)) : 0.0;
        }
#pragma unroll
        for (int kk = 0; kk < 12; ++kk) {
            if (k0 + kk < ksteps) {
#pragma unroll
                for (int tt = 0; tt < MAXTT; ++tt)
                    if (tt < ntt_live) dmma(acc[tt][0], acc[tt][1], af[kk], Bs[(size_t)(8 * tt) * ldb + 4 * (k0 + kk)]);
            }
        }
    }
}

// 1/x for a positive normal x: MUFU.RCP64H seed + two Newton steps (<= 1 ulp), no slow path.
__device__ __forceinline__ double rcp_pos(double x)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    return fma(y, e, y);
}

// ---------------------------------------------------------------------------------------------
// One warp: Cholesky of the n x n block in bS (lower triangle, leading dimension NP+1) and explicit
// inverse of the factor, NP = padded size (compile time; rows/cols n..NP-1 behave as identity).
//   S = U D U'  (U unit lower, D = diag(d_k)) right-looking with row r in the registers of lane r; the
//   pivot chain runs through shuffles and each lane updates its own diagonal entry, so the serial work
//   per column is shfl -> rcp -> 2 FMAs; the rank-1 update reads the column from a 32-double shared
//   vector (128-bit broadcast loads).  Then V = inv(U) column j by lane j and
//   inv(L) = diag(1/sqrt(d)) V  with ONE rsqrt per lane.
// bS is consumed before bLinv is written (they may alias); bU must not alias either.
// Outputs: bLinv (RP x ld operand block, zero padded) and gLinv (n x n row-major, global scratch).
// Returns 0 or failing column + 1.
// ---------------------------------------------------------------------------------------------
template <int NP>
__device__ __forceinline__ int warp_potrf_inverse(const double *bS, double *bU, double *bLinv, int ld, double *gLinv, int n,
                                                  double *colbuf, double *rsv, int lane)
{
    constexpr int LDS_ = NP + 1, LDU = NP + 2, RP_ = (NP + 7) & ~7;
    constexpr unsigned FULL = 0xffffffffu;
    const int r = lane;
    double a[NP];
#pragma unroll
    for (int c = 0; c < NP; ++c) a[c] = (r < n && c < r) ? bS[r * LDS_ + c] : 0.0;
    double diag = (r < n) ? bS[r * LDS_ + r] : 1.0;
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
            const double dinv = rcp_pos(d);
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
#pragma unroll
    for (int c = 0; c + 1 < NP; c += 2)
        if (r < NP) *reinterpret_cast<double2 *>(bU + r * LDU + c) = make_double2(a[c], a[c + 1]);
    __syncwarp();
    // V = inv(U): column j by lane j;  v[i] = -(sum_{k<i} U(i,k) v[k]) for i > j, v[j] = 1, v[i<j] = 0
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
    const bool live = (j < n);
#pragma unroll
    for (int i = 0; i < RP_; ++i) {
        double val = 0.0;
        if (i < NP) val = (live && i < n) ? v[i < NP ? i : 0] * rsv[i] : 0.0;    // inv(L)(i,j) = v(i,j) / sqrt(d_i)
        if (j < NP) bLinv[i * ld + j] = val;
        if (live && i < n) gLinv[i * n + j] = val;
    }
    return 0;
}

// 1/sqrt(x), x > 0 normal : MUFU.RSQ64H seed and one cubic (Halley) step: the seed's relative error 2^-20 becomes ~2^-58, below the
// rounding of the result (the quadratic polish step that followed cost four more FP64 operations on the serial pivot chain of every
// column of every diagonal tile for nothing measurable; the warp kernel dropped it earlier in the round)
__device__ __forceinline__ double rsqrt_fast(const double x)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-x * y, y, 1.0);
    return fma(y * e, fma(0.375, e, 0.5), y);
}

// One warp: Cholesky of the 8 x 8 diagonal tile kb of bS and its inverse, in registers (lane r & 7 owns row r, columns
// travel by shuffles).  Rows / columns >= n behave as identity.  Writes inv(L_kk) to tLi (8 x 8 row-major) and to the
// diagonal tile of bT.  Returns 0 or failing column + 1 (uniform over the warp).
__device__ __forceinline__ int diag_tile_potrf_inverse(const double *bS, double *bT, int ld, int n, int kb, double *tLi, int lane)
{
    constexpr unsigned FULLM = 0xffffffffu;
    const double *D = bS + (size_t)(8 * kb) * ld + 8 * kb;
    const int r = lane & 7;
    const bool rpad = (8 * kb + r >= n);
    double a[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) a[c] = rpad ? ((c == r) ? 1.0 : 0.0) : ((c <= r) ? D[r * ld + c] : 0.0);
    int info = 0;
    double dinv = 1.0;                                          // 1 / L(r,r) = rsqrt of the r-th pivot
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const double d = __shfl_sync(FULLM, a[k], k);
        if ((!(d > 0.0) || !(d < 1.0e300)) && !info) info = 8 * kb + k + 1;       // uniform: same d in every lane
        const double rs = rsqrt_fast(d);
        if (r == k) dinv = rs;
        const double l = a[k] * rs;                             // L(r,k) for r >= k (lane k: sqrt(d))
        a[k] = l;
#pragma unroll
        for (int c = k + 1; c < 8; ++c) {
            const double lc = __shfl_sync(FULLM, l, c);
            if (r >= c) a[c] = fma(-l, lc, a[c]);
        }
    }
    if (info) return info;
    // inverse of the 8 x 8 lower factor: lane j (< 8) builds column j, rows of L arrive by shuffles
    const int j = r;
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        double sacc = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (k < i) sacc = fma(__shfl_sync(FULLM, a[k], i), x[k], sacc);
        const double di = __shfl_sync(FULLM, dinv, i);
        x[i] = (i < j) ? 0.0 : ((i == j) ? di : -sacc * di);
    }
    if (lane < 8) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { tLi[i * 8 + j] = x[i]; bT[(size_t)(8 * kb + i) * ld + 8 * kb + j] = x[i]; }
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Whole CTA (n > 32): Cholesky of the n x n block in bS (lower triangle, leading dimension ld) and explicit inverse of
// the factor, blocked by the 8 x 8 DMMA tile, right-looking with look-ahead:
//   step kb:  warp 0 updates the next diagonal tile (kb+1, kb+1) and immediately factors / inverts it (registers,
//             shuffles) while the other warps do the rest of the trailing update of step kb -- the serial diagonal
//             tiles, not the tile products, are the critical path;
//   a tile task (I, J), kb < J <= I, forms its two panel tiles P_I = S[I,kb] inv(L_kk)', P_J in accumulator layout
//             (4 DMMAs) and multiplies them straight from registers: an accumulator tile {row gq, cols 2q, 2q+1} is a
//             valid A operand and a valid B operand of X Y' with the contraction index taken as k = 2q + e.  Column kb
//             of S is only read; the diagonal task (I, I) files the panel tile L[I,kb] at the mirrored (otherwise
//             unused) upper position (kb, I);
//   inv(L) by block columns: X[i,k] = -inv(L_ii) sum_{j=k}^{i-1} L[i,j] X[j,k], one warp per column, no CTA barrier.
// Rows / columns >= n of the last tile behave as identity.  bT is a scratch block (receives X); on exit
// bS <- inv(L) zero padded to RP x ld and gLinv <- n x n row-major.  tmp: 128 + 64 doubles per warp.
// All threads must call it; returns 0 or failing column + 1 (the same value in every thread).
// ---------------------------------------------------------------------------------------------
__device__ __noinline__ int cta_potrf_inverse(double *bS, double *bT, int ld, int RP, double *gLinv, int n, int tid, int nth,
                                              double *tmp, int *s_info)
{
    const int lane = tid & 31, wid = tid >> 5, nw = nth >> 5, gq = lane >> 2, q = lane & 3;
    const int nt = RP / 8;
    double *wtmp = tmp + 128 + 64 * wid;                // per-warp 8 x 8 scratch; tmp[0..127]: inv(L_kk), double buffered
    // s_info[2]: the failure flag of the diagonal tile kb + 1 goes to slot (kb + 1) & 1, so that the write of one step cannot race with
    // the threads still reading the flag of the step before (racecheck: WAR on a single flag; an early return of a few threads
    // would have shifted the barriers of the failure path)
    if (tid == 0) { s_info[0] = 0; s_info[1] = 0; }
    __syncthreads();
    if (wid == 0) {
        const int info = diag_tile_potrf_inverse(bS, bT, ld, n, 0, tmp, lane);
        if (info && lane == 0) s_info[0] = info;
    }
    __syncthreads();
    if (s_info[0]) return s_info[0];
    for (int kb = 0; kb + 1 < nt; ++kb) {
        const double *tLi = tmp + 64 * (kb & 1);
        const int R = nt - kb - 1, ntask = R * (R + 1) / 2;
        auto tile_task = [&](int task) {
            int it = 0;
            while (task >= (it + 1) * (it + 2) / 2) ++it;
            const int jt = task - it * (it + 1) / 2;
            const int I = kb + 1 + it, J = kb + 1 + jt;
            const double *Bl = tLi + gq * 8 + q;                 // B[k][n] = inv(L_kk)[n][k]
            const double *Ai = bS + (size_t)(8 * I + gq) * ld + 8 * kb + q;
            double pi0 = 0.0, pi1 = 0.0, pj0 = 0.0, pj1 = 0.0;
            dmma(pi0, pi1, Ai[0], Bl[0]);
            if (J != I) {
                const double *Aj = bS + (size_t)(8 * J + gq) * ld + 8 * kb + q;
                dmma(pj0, pj1, Aj[0], Bl[0]);
                dmma(pi0, pi1, Ai[4], Bl[4]);
                dmma(pj0, pj1, Aj[4], Bl[4]);
            } else {
                dmma(pi0, pi1, Ai[4], Bl[4]);
                pj0 = pi0; pj1 = pi1;
            }
            double t0 = 0.0, t1 = 0.0;
            dmma(t0, t1, pi0, pj0);
            dmma(t0, t1, pi1, pj1);
            double *o = bS + (size_t)(8 * I + gq) * ld + 8 * J + 2 * q;
            o[0] -= t0; o[1] -= t1;
            if (J == I) {
                double *mo = bS + (size_t)(8 * kb + gq) * ld + 8 * I + 2 * q;
                mo[0] = pi0; mo[1] = pi1;
            }
        };
        if (wid == 0) {
            tile_task(0);                                        // (kb+1, kb+1): the next diagonal tile ...
            __syncwarp();
            const int info = diag_tile_potrf_inverse(bS, bT, ld, n, kb + 1, tmp + 64 * ((kb + 1) & 1), lane);   // ... factored at once
            if (info && lane == 0) s_info[(kb + 1) & 1] = info;
        } else {
            for (int task = wid; task < ntask; task += nw - 1) tile_task(task);
        }
        __syncthreads();
        if (s_info[(kb + 1) & 1]) return s_info[(kb + 1) & 1];
    }
    // ---- inv(L) by block columns: column k is an independent forward substitution over its block rows, one warp per
    //      column and no CTA barrier inside (L[i,j] is read from the mirrored tile (j, i)) ----
    for (int k = wid; k + 1 < nt; k += nw) {
        for (int i = k + 1; i < nt; ++i) {
            double c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0;
            for (int j = k; j < i; ++j) {
                const double *Ap = bS + (size_t)(8 * j + gq) * ld + 8 * i + q;
                const double *Bp = bT + (size_t)(8 * j + q) * ld + 8 * k + gq;
                dmma(c0, c1, Ap[0], Bp[0]);
                dmma(e0, e1, Ap[4], Bp[(size_t)4 * ld]);
            }
            __syncwarp();
            wtmp[gq * 8 + 2 * q] = c0 + e0;
            wtmp[gq * 8 + 2 * q + 1] = c1 + e1;
            __syncwarp();
            double x0 = 0.0, x1 = 0.0;
            const double *Lp = bT + (size_t)(8 * i + gq) * ld + 8 * i + q;
            dmma(x0, x1, Lp[0], wtmp[q * 8 + gq]);
            dmma(x0, x1, Lp[4], wtmp[(4 + q) * 8 + gq]);
            double *o = bT + (size_t)(8 * i + gq) * ld + 8 * k + 2 * q;
            o[0] = -x0; o[1] = -x1;
            __syncwarp();
        }
    }
    __syncthreads();
    for (int e = tid; e < RP * ld; e += nth) {
        const int i = e / ld, j = e - i * ld;
        const double v = (i < n && j <= i) ? bT[e] : 0.0;
        bS[e] = v;
        if (i < n && j < n) gLinv[(size_t)i * n + j] = v;
    }
    return 0;
}

// Everything a phase needs; lives in registers / constant bank.
struct KC {
    const DevSys &S;
    int n, m, T, NB, has_xf;
    int tid, lane, wid, gq, q;
    int ld, nt, ks, mp, ldp, TP, ntt;
    double *Us, *Xs, *red;
    int nth, nw;                        // threads / warps per CTA
};

// ---- out = sgn * (C v - sub)  for v staged as  Us [TP][ldp] (u-like rows t)  and  Xs [(TP+2)][ld]
//      (row rr <-> x-like source row rr - 2; source row j-1 holds x_j).  Rows follow
//      VAR_2/fast_mpc_eq_const.m:38-49,67-71.  out / sub : NB*n global vectors.
__device__ __noinline__ void mma_apply_C(const KC &k, const double *sub, double *out, bool negate)
{
    const DevSys &S = k.S;
    const int n = k.n, m = k.m, T = k.T;
    for (int mt = k.wid; mt < k.nt; mt += k.nw) {
        const int row = 8 * mt + k.gq;
        const bool rok = row < n;
        for (int tt0 = 0; tt0 < k.ntt; tt0 += MAXTT) {
            const int live = min(MAXTT, k.ntt - tt0);
            double acc[MAXTT][2];
#pragma unroll
            for (int tt = 0; tt < MAXTT; ++tt) acc[tt][0] = acc[tt][1] = 0.0;
            // B u_i
            mma_gA_sB(acc, S.Bt + (size_t)row * m + k.q, rok, m, k.mp / 4, k.Us + (size_t)(8 * tt0 + k.gq) * k.ldp + k.q, k.ldp, live, k.q);
            // A1 x_i  (x_i = source row i-1 = staged row i+1)
            mma_gA_sB(acc, S.A1t + (size_t)row * n + k.q, rok, n, k.ks, k.Xs + (size_t)(8 * tt0 + k.gq + 1) * k.ld + k.q, k.ld, live, k.q);
            // A2 x_{i-1}  (source row i-2 = staged row i)
            if (S.has_a2)
                mma_gA_sB(acc, S.A2t + (size_t)row * n + k.q, rok, n, k.ks, k.Xs + (size_t)(8 * tt0 + k.gq) * k.ld + k.q, k.ld, live, k.q);
#pragma unroll
            for (int tt = 0; tt < MAXTT; ++tt) {
                if (tt < live) {
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int i = 8 * (tt0 + tt) + 2 * k.q + e;
                        if (rok && i < T) {
                            const double v = k.Xs[(size_t)(i + 2) * k.ld + row] - acc[tt][e] - sub[(size_t)i * n + row];
                            out[(size_t)i * n + row] = negate ? -v : v;
                        }
                    }
                }
            }
        }
    }
    if (k.has_xf)
        for (int r = k.tid; r < n; r += k.nth) {
            const double v = k.Xs[(size_t)(T - 1 + 2) * k.ld + r] - sub[(size_t)T * n + r];
            out[(size_t)T * n + r] = negate ? -v : v;
        }
}

// ---- C' v for v staged as Ns = Xs region [(TP+2)][ld], row rr <-> v_rr for rr < T, rows >= T zero.
//      MODE 0 (init):  hu = B' nu_t ,  hx = (C' nu)_x                                  (stored)
//      MODE 1 (step):  hdu, hdx stored and  du = -(rdu - hdu) w ,  dx = -(rdx + hdx) qi     (inf_newton_solver.m:34-35)
template <int MODE>
__device__ __noinline__ void mma_apply_Ct(const KC &k, const double *vglob, double *hu, double *hx, const double *rdu,
                                          const double *rdx, const double *w, double *du, double *dx)
{
    const DevSys &S = k.S;
    const int n = k.n, m = k.m, T = k.T;
    const int mtu = (m + 7) / 8;
    for (int task = k.wid; task < mtu + k.nt; task += k.nw) {
        const bool upart = task < mtu;
        const int mt = upart ? task : task - mtu;
        const int row = 8 * mt + k.gq;
        const bool rok = row < (upart ? m : n);
        for (int tt0 = 0; tt0 < k.ntt; tt0 += MAXTT) {
            const int live = min(MAXTT, k.ntt - tt0);
            double acc[MAXTT][2];
#pragma unroll
            for (int tt = 0; tt < MAXTT; ++tt) acc[tt][0] = acc[tt][1] = 0.0;
            if (upart) {
                // (B' v_t)(j) : A = B' as row-major [j][k] = column-major B
                mma_gA_sB(acc, S.B + (size_t)row * n + k.q, rok, n, k.ks, k.Xs + (size_t)(8 * tt0 + k.gq) * k.ld + k.q, k.ld, live, k.q);
            } else {
                // A1' v_{t+1} + A2' v_{t+2}
                mma_gA_sB(acc, S.A1 + (size_t)row * n + k.q, rok, n, k.ks, k.Xs + (size_t)(8 * tt0 + k.gq + 1) * k.ld + k.q, k.ld, live, k.q);
                if (S.has_a2)
                    mma_gA_sB(acc, S.A2 + (size_t)row * n + k.q, rok, n, k.ks, k.Xs + (size_t)(8 * tt0 + k.gq + 2) * k.ld + k.q, k.ld, live, k.q);
            }
#pragma unroll
            for (int tt = 0; tt < MAXTT; ++tt) {
                if (tt < live) {
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int t = 8 * (tt0 + tt) + 2 * k.q + e;
                        if (rok && t < T) {
                            if (upart) {
                                const size_t idx = (size_t)t * m + row;
                                const double h = acc[tt][e];
                                hu[idx] = h;
                                if (MODE == 1) du[idx] = -(rdu[idx] - h) * w[idx];
                            } else {
                                const size_t idx = (size_t)t * n + row;
                                double h = k.Xs[(size_t)t * k.ld + row] - acc[tt][e];      // v_t - A1' v_{t+1} - A2' v_{t+2}
                                if (k.has_xf && t == T - 1) h += vglob[(size_t)T * n + row];
                                hx[idx] = h;
                                if (MODE == 1) dx[idx] = -(rdx[idx] + h) * __ldg((t == T - 1 ? S.qif : S.qi) + row);
                            }
                        }
                    }
                }
            }
        }
    }
}

// Stage an x-like global array (rows of n) into Xs with a row offset: Xs row rr <- src row rr - shift.
__device__ __forceinline__ void stage_rows(const KC &k, const double *src, int nrows, int shift)
{
    const int tot = (k.TP + 2) * k.ld;
    for (int e = k.tid; e < tot; e += k.nth) {
        const int rr = e / k.ld, c = e - rr * k.ld, sr = rr - shift;
        k.Xs[e] = (sr >= 0 && sr < nrows && c < k.n) ? src[(size_t)sr * k.n + c] : 0.0;
    }
}

} // namespace

// =============================================================================================
template <int NP>
__global__ void __launch_bounds__(KCfg<NP>::NTH, KCfg<NP>::MINB) fmpc_solve_kernel_mma(const DevSys S, const StepArgs A)
{
    constexpr int NTHREADS = KCfg<NP>::NTH, NWARPS = NTHREADS / 32, KSMAX = KCfg<NP>::KSMAX, RMAX = KCfg<NP>::RMAX;
    extern __shared__ double smem[];
    const int n = S.n, m = S.m, T = S.T;
    const int NB = T + (A.has_xf ? 1 : 0);
    const Geom G = Geom::make(n, m, T);
    const int ld = G.ld, KP = G.KP, RP = G.RP, nt = G.nt, ks = G.ks, lds = G.lds;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int gq = lane >> 2, q = lane & 3;                 // DMMA fragment coordinates
    const WsLayout L = WsLayout::make(n, m, T);
    const int Mp = S.Mp, mp = S.mp, ldp = G.ldp, TP = G.TP;

    double *ops = smem;                                     // 5 operand blocks | staging (Us + Xs)
    double *Us = ops, *Xs = ops + (size_t)TP * ldp;
    const int VL = G.vlen();
    double *sm_rhs = smem + G.ops_doubles();                // VL
    double *sm_y1 = sm_rhs + VL, *sm_y2 = sm_y1 + VL;       // y_{i-1}, y_{i-2}
    double *colbuf = sm_y2 + VL;                            // 64
    double *rsv = colbuf + 64;                              // VL
    double *red = rsv + VL;                                 // 36
    __shared__ int s_inst, s_flag, s_info2[2];
    double *ptmp = red + 36;                                // potrf scratch (n > 32): 64 + 64 per warp
    const KC kc{S, n, m, T, NB, A.has_xf, tid, lane, wid, gq, q, ld, nt, ks, mp, ldp, TP, G.ntt, Us, Xs, red, NTHREADS, NWARPS};

    double *ws = A.ws + (size_t)blockIdx.x * A.ws_stride;
    double *nu = ws + L.nu, *dnu = ws + L.dnu, *yv = ws + L.yv, *bv = ws + L.bv;
    double *rp = ws + L.rp, *rpt = ws + L.rpt;
    double *hx = ws + L.hx, *hdx = ws + L.hdx, *dx = ws + L.dx, *xt = ws + L.xt, *rdx = ws + L.rdx;
    double *hu = ws + L.hu, *hdu = ws + L.hdu, *du = ws + L.du, *ut = ws + L.ut, *dbar = ws + L.dbar;
    double *pinv = ws + L.pinv, *rdu = ws + L.rdu;
    double *gLi = ws + L.Lf, *gL1 = ws + L.L1, *gL2 = ws + L.L2, *Dsc = ws + L.Dsc;
    const size_t nn = (size_t)n * n;
    const int usp = TP * ldp;                               // padded u-space
    const int xsp = (TP + 2) * ld;                          // padded x-space (staged rows)

    for (;;) {
        __syncthreads();
        if (tid == 0) s_inst = (int)atomicAdd(A.counter, 1u);
        __syncthreads();
        const int b = s_inst;
        if (b >= A.nbatch) break;

        double *uo = A.U + (size_t)b * m * T, *xo = A.X + (size_t)b * n * T;      // output arrays
        double *u = uo, *x = xo;                                                    // current iterate (ping-pongs with ut / xt)
        double *un = ut, *xn = xt;
        rp = ws + L.rp; rpt = ws + L.rpt;
        const double *x0 = A.x0 + (size_t)b * n;
        const double *x0p = A.x0_pre ? A.x0_pre + (size_t)b * n : nullptr;

        // ---- initial iterate (fast_mpc_init.m:12-26) staged straight into Us / Xs; nu; b (eq_const.m:39,44,47,68) ----
        {
            const double *u0 = A.cold ? nullptr : A.U0 + (size_t)b * m * T;
            const double *xx0 = A.cold ? nullptr : A.X0 + (size_t)b * n * T;
            for (int e = tid; e < usp; e += NTHREADS) {
                const int t = e / ldp, j = e - t * ldp;
                double v = 0.0;
                if (t < T && j < m) {
                    v = A.cold ? (S.umin[j] + S.umax[j]) / 2 : u0[(size_t)t * m + j];
                    u[(size_t)t * m + j] = v;
                }
                Us[e] = v;
            }
            for (int e = tid; e < xsp; e += NTHREADS) {
                const int rr = e / ld, k = e - rr * ld, sr = rr - 2;
                double v = 0.0;
                if (sr >= 0 && sr < T && k < n) {
                    v = A.cold ? (S.xmin[k] + S.xmax[k]) / 2 : xx0[(size_t)sr * n + k];
                    x[(size_t)sr * n + k] = v;
                }
                Xs[e] = v;
            }
        }
        for (int e = tid; e < NB * n; e += NTHREADS) nu[e] = A.nu0[(size_t)b * NB * n + e];
        for (int e = tid; e < NB * n; e += NTHREADS) {
            const int i = e / n, k = e - i * n;
            double v;
            if (i < T) {
                v = A.w ? A.w[(size_t)b * T * n + e] : 0.0;
                if (i == 0) {
                    double s = 0.0;
                    for (int kk = 0; kk < n; ++kk) s = fma(__ldg(S.A1 + k + n * kk), x0[kk], s);
                    if (S.has_a2) for (int kk = 0; kk < n; ++kk) s = fma(__ldg(S.A2 + k + n * kk), x0p[kk], s);
                    v += s;
                } else if (i == 1 && S.has_a2) {
                    double s = 0.0;
                    for (int kk = 0; kk < n; ++kk) s = fma(__ldg(S.A2 + k + n * kk), x0[kk], s);
                    v += s;
                }
            } else {
                v = A.xf[(size_t)b * n + k];
            }
            bv[e] = v;
        }
        __syncthreads();
        mma_apply_C(kc, bv, rp, false);                     // r_p = C z - b
        __syncthreads();
        stage_rows(kc, A.nu0 + (size_t)b * NB * n, T, 0);
        __syncthreads();
        mma_apply_Ct<0>(kc, nu, hu, hx, nullptr, nullptr, nullptr, nullptr, nullptr);
        __syncthreads();

        int status = ST_OK, iters = 0;
        PB_DECL
        for (int it = 0; it < A.niters; ++it) {
            PB_T(9);
            // ---- barrier terms (inf_newton_KKT_H.m:3-13), r_d (:12), p = inv(Phi) r_d staged for C p ----
            double ssd = 0.0;
            for (int e = tid; e < usp; e += NTHREADS) {
                const int t = e / ldp, j = e - t * ldp;
                double pv = 0.0;
                if (t < T && j < m) {
                    const size_t idx = (size_t)t * m + j;
                    const double uu = u[idx];
                    const double sp = __ldg(S.umax + j) - uu, sm = -__ldg(S.umin + j) + uu;
                    const double dp = 1.0 / sp, dm = 1.0 / sm;
                    const double db = A.kappa * (dp - dm);
                    const double w = 1.0 / (__ldg(S.r2 + j) + A.kappa * (dp * dp + dm * dm));
                    const double r = rdu_expr(__ldg(S.r2 + j), __ldg(S.rl + j), uu, hu[idx], db);
                    dbar[idx] = db; pinv[idx] = w; rdu[idx] = r;
                    ssd = fma(r, r, ssd);
                    pv = r * w;
                }
                Us[e] = pv;
            }
            for (int e = tid; e < xsp; e += NTHREADS) {
                const int rr = e / ld, k = e - rr * ld, sr = rr - 2;
                double pv = 0.0;
                if (sr >= 0 && sr < T && k < n) {
                    const size_t idx = (size_t)sr * n + k;
                    const bool last = (sr == T - 1);
                    const double r = rdx_expr(__ldg((last ? S.q2f : S.q2) + k), __ldg((last ? S.qfl : S.ql) + k), x[idx], hx[idx]);
                    rdx[idx] = r;
                    ssd = fma(r, r, ssd);
                    pv = r * __ldg((last ? S.qif : S.qi) + k);
                }
                Xs[e] = pv;
            }
            double ssp = 0.0;
            for (int e = tid; e < NB * n; e += NTHREADS) ssp = fma(rp[e], rp[e], ssp);
            const double tot_p = block_sum(ssp, red);
            const double tot_d = block_sum(ssd, red);
            const double nr0 = sqrt(tot_d + tot_p);
            // ---- early exit (inf_newton_solver.m:19-22) ----
            if (!isfinite(nr0)) { status = ST_NONFINITE; break; }
            if (nr0 <= A.tol_r && sqrt(tot_p) <= A.tol_p) { status = ST_EARLY_EXIT; break; }
            // ---- rhs of  Y dnu = -beta,  beta = -r_p + C inv(Phi) r_d  (:28-29) ----
            mma_apply_C(kc, rp, yv, true);
            __syncthreads();

            PB_T(0);
            // ---- D_t = B diag(w_t) B' for all stages: Dsc(t, pair) = G * W ----
            for (int e = tid; e < usp; e += NTHREADS) {
                const int t = e / ldp, j = e - t * ldp;
                Us[e] = (t < T && j < m) ? pinv[(size_t)t * m + j] : 0.0;
            }
            __syncthreads();
            for (int g = wid; g < Mp / 8; g += NWARPS) {
                for (int tt0 = 0; tt0 < G.ntt; tt0 += MAXTT) {
                    const int live = min(MAXTT, G.ntt - tt0);
                    double acc[MAXTT][2];
#pragma unroll
                    for (int tt = 0; tt < MAXTT; ++tt) acc[tt][0] = acc[tt][1] = 0.0;
                    mma_gA_sB(acc, S.G + (size_t)(8 * g + gq) * mp + q, true, mp, mp / 4, Us + (size_t)(8 * tt0 + gq) * ldp + q, ldp, live, q);
                    const int pair = 8 * g + gq;
#pragma unroll
                    for (int tt = 0; tt < MAXTT; ++tt) {
                        const int t = 8 * (tt0 + tt) + 2 * q;
                        if (tt < live) {
                            if (t < T) Dsc[(size_t)t * Mp + pair] = acc[tt][0];
                            if (t + 1 < T) Dsc[(size_t)(t + 1) * Mp + pair] = acc[tt][1];
                        }
                    }
                }
            }
            __syncthreads();
            PB_T(1);
            // zero the operand blocks (padding rows / columns must stay exactly zero)
            for (int e = tid; e < (int)(5 * G.blk()); e += NTHREADS) ops[e] = 0.0;
            __syncthreads();

            PB_T(2);
            // ---- band-2 block Cholesky of Y fused with the forward solve (:30-31) ----
            // five blocks: Linv (doubles as S scratch), L1 ring x2, L2 ring x2 (the older one doubles as U scratch)
            double *bLinv = ops, *bM1 = ops + G.blk(), *bL1p = ops + 2 * G.blk();
            double *bL2p = ops + 3 * G.blk(), *bL2pp = ops + 4 * G.blk();
            bool fail = false;
            const int nS = nt * (nt + 1) / 2;
            for (int i = 0; i < NB; ++i) {
                const bool has1 = (i + 1 < NB), has2 = (i + 2 < NB) && S.has_a2;
                const bool up1 = (i >= 1), up2 = (i >= 2) && S.has_a2;
                double *bS = bLinv;                          // inv(L_{i-1}) is dead: its block receives S_i
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
                        // the four loads are issued here and only consumed after the tile products (no add in between: an
                        // in-order warp would otherwise wait for L2 before its first DMMA)
                        double i0 = 0.0, i1 = 0.0, d0 = 0.0, d1 = 0.0;
                        if (r < n) {
                            if (cc <= r) { i0 = __ldg(Yd + r * n + cc); if (i < T) d0 = Dsc[(size_t)i * Mp + r * (r + 1) / 2 + cc]; }
                            if (cc + 1 <= r) { i1 = __ldg(Yd + r * n + cc + 1); if (i < T) d1 = Dsc[(size_t)i * Mp + r * (r + 1) / 2 + cc + 1]; }
                        }
                        double p0 = 0.0, p1 = 0.0;
                        if (up1) tile_nt(p0, p1, bL1p + (8 * rt + gq) * ld + q, bL1p + (8 * ct + gq) * ld + q, ks);
                        if (up2) tile_nt(p0, p1, bL2pp + (8 * rt + gq) * ld + q, bL2pp + (8 * ct + gq) * ld + q, ks);
                        if (r < n) {
                            if (cc <= r) bS[r * lds + cc] = (i0 + d0) - p0;
                            if (cc + 1 <= r) bS[r * lds + cc + 1] = (i1 + d1) - p1;
                        }
                    } else {
                        const int tk = task - nS, rt = tk / nt, ct = tk - rt * nt;
                        const int r = 8 * rt + gq, cc = 8 * ct + 2 * q;
                        double i0 = 0.0, i1 = 0.0;
                        if (r < n && y1 >= 0) {
                            if (cc < n) i0 = __ldg(S.ypool + (size_t)y1 * nn + r * n + cc);
                            if (cc + 1 < n) i1 = __ldg(S.ypool + (size_t)y1 * nn + r * n + cc + 1);
                        }
                        double p0 = 0.0, p1 = 0.0;
                        if (up1 && S.has_a2) tile_nt(p0, p1, bL2p + (8 * rt + gq) * ld + q, bL1p + (8 * ct + gq) * ld + q, ks);
                        if (cc < KP) bM1[r * ld + cc] = i0 - p0;
                        if (cc + 1 < KP) bM1[r * ld + cc + 1] = i1 - p1;
                    }
                }
                // rhs_i = yv_i - L1p y_{i-1} - L2pp y_{i-2}   (4 lanes per row)
                for (int r = tid >> 2; r < RP; r += NTHREADS / 4) {          // RP is a multiple of 8: uniform per warp
                    double s = 0.0;
                    if (up1) for (int k = q; k < KP; k += 4) s = fma(bL1p[r * ld + k], sm_y1[k], s);
                    if (up2) for (int k = q; k < KP; k += 4) s = fma(bL2pp[r * ld + k], sm_y2[k], s);
                    s += __shfl_xor_sync(0xffffffffu, s, 1);
                    s += __shfl_xor_sync(0xffffffffu, s, 2);
                    if (q == 0) sm_rhs[r] = (r < n) ? yv[i * n + r] - s : 0.0;
                }
                __syncthreads();
                PB_T(3);
                // -- phase 2: factor + invert the diagonal block (one warp, rotating over the SM sub-partitions) --
                if constexpr (NP > 32) {
                    const int info = cta_potrf_inverse(bS, bL2pp, ld, RP, gLi + (size_t)i * nn, n, tid, NTHREADS, ptmp, s_info2);
                    __syncthreads();
                    if (tid == 0) s_flag = info;
                } else {
                    if (wid == (i & (NWARPS - 1))) {
                        const int info = warp_potrf_inverse<(NP > 32 ? 32 : NP)>(bS, bL2pp, bLinv, ld, gLi + (size_t)i * nn, n, colbuf, rsv, lane);
                        if (lane == 0) s_flag = info;
                    }
                }
                __syncthreads();
                if (s_flag) { fail = true; break; }
                PB_T(4);
                // -- phase 3: L1_i = M1 inv(L)' (in place), L2_i = Y2 inv(L)' (into the dead L2pp block), y_i = inv(L) rhs --
                double *bM2 = bL2pp;
                // 2 nt tasks (product, row tile) instead of nt (row tile, both products): nt = 9 row tiles on 8 warps left seven warps waiting
                // for the one with two tasks
                for (int task = wid; task < 2 * nt; task += NWARPS) {
                    const int rt = task % nt, which = task / nt;
                    const int r = 8 * rt + gq;
                    if (which == 0 && has1) {
                        double af[KSMAX];
#pragma unroll
                        for (int k = 0; k < KSMAX; ++k) af[k] = (k < ks) ? bM1[r * ld + 4 * k + q] : 0.0;
                        __syncwarp();
                        for (int ct = 0; ct < nt; ++ct) {
                            double c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0;
                            const int kmax = min(ks, 2 * (ct + 1));          // inv(L) is lower triangular
                            const double *Bp = bLinv + (8 * ct + gq) * ld + q;
#pragma unroll
                            for (int k = 0; k < KSMAX; ++k) if (k < kmax) { if (k & 1) dmma(e0, e1, af[k], Bp[4 * k]); else dmma(c0, c1, af[k], Bp[4 * k]); }
                            c0 += e0; c1 += e1;
                            const int cc = 8 * ct + 2 * q;
                            if (cc < KP) bM1[r * ld + cc] = c0;
                            if (cc + 1 < KP) bM1[r * ld + cc + 1] = c1;
                            if (r < n) {
                                if (cc < n) gL1[(size_t)i * nn + r * n + cc] = c0;
                                if (cc + 1 < n) gL1[(size_t)i * nn + r * n + cc + 1] = c1;
                            }
                        }
                    }
                    if (which == 1 && has2) {
                        const int y2 = S.y2i[i];
                        double af[KSMAX];
#pragma unroll
                        for (int k = 0; k < KSMAX; ++k) {
                            const int kk = 4 * k + q;
                            af[k] = (y2 >= 0 && r < n && kk < n) ? __ldg(S.ypool + (size_t)y2 * nn + r * n + kk) : 0.0;
                        }
                        for (int ct = 0; ct < nt; ++ct) {
                            double c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0;
                            const int kmax = min(ks, 2 * (ct + 1));
                            const double *Bp = bLinv + (8 * ct + gq) * ld + q;
#pragma unroll
                            for (int k = 0; k < KSMAX; ++k) if (k < kmax) { if (k & 1) dmma(e0, e1, af[k], Bp[4 * k]); else dmma(c0, c1, af[k], Bp[4 * k]); }
                            c0 += e0; c1 += e1;
                            const int cc = 8 * ct + 2 * q;
                            if (cc < KP) bM2[r * ld + cc] = c0;
                            if (cc + 1 < KP) bM2[r * ld + cc + 1] = c1;
                            if (r < n) {
                                if (cc < n) gL2[(size_t)i * nn + r * n + cc] = c0;
                                if (cc + 1 < n) gL2[(size_t)i * nn + r * n + cc + 1] = c1;
                            }
                        }
                    } else if (which == 1) {
                        // the block held the U scratch of the factorization: restore the zero padding invariant
                        for (int ct = 0; ct < nt; ++ct) {
                            const int cc = 8 * ct + 2 * q;
                            if (cc < KP) bM2[r * ld + cc] = 0.0;
                            if (cc + 1 < KP) bM2[r * ld + cc + 1] = 0.0;
                        }
                    }
                }
                {
                    double sv[RMAX];
#pragma unroll
                    for (int u = 0; u < RMAX; ++u) {
                        const int r = (tid >> 2) + u * (NTHREADS / 4);
                        double s = 0.0;
                        if (r < RP) for (int k = q; k < KP; k += 4) s = fma(bLinv[r * ld + k], sm_rhs[k], s);
                        s += __shfl_xor_sync(0xffffffffu, s, 1);
                        s += __shfl_xor_sync(0xffffffffu, s, 2);
                        sv[u] = s;
                    }
                    __syncthreads();                               // everyone is done with sm_y1 / sm_y2 / the *p blocks of this stage
#pragma unroll
                    for (int u = 0; u < RMAX; ++u) {
                        const int r = (tid >> 2) + u * (NTHREADS / 4);
                        if (r < RP && q == 0) {
                            if (r < n) yv[i * n + r] = sv[u];
                            sm_y2[r] = sm_y1[r];
                            sm_y1[r] = (r < n) ? sv[u] : 0.0;
                        }
                    }
                }
                {   // rotate the rings: L1p <-> M1 ; L2pp (now L2_i) becomes L2p, old L2p becomes L2pp
                    double *t1 = bL1p; bL1p = bM1; bM1 = t1;
                    double *t2 = bL2p; bL2p = bL2pp; bL2pp = t2;
                }
                __syncthreads();
                PB_T(5);
            }
            if (fail) { status = ST_NOT_PD; break; }

            // ---- backward solve  dnu_i = inv(L_i)' (y_i - L1_i' dnu_{i+1} - L2_i' dnu_{i+2})  (:32) ----
            const bool bw_smem = (NP > 32) && (n % 2 == 0) && (6 * nn <= 5 * G.blk()) && (FMPC_BW_SMEM != 0);
            if (bw_smem) {
                // The factor blocks of a stage (3 n^2 doubles) come back from the scratch through cp.async into the operand
                // region, which is free now (6 buffers of n^2 doubles fit the five RP x ld blocks): stage i - 1 is in flight
                // while stage i is consumed, and the GEMVs read shared memory instead of waiting on L2 for every element.
                auto issue_stage = [&](int i, int buf) {
                    const int h2 = (int)nn / 2;                   // 16-byte pieces per block (n even)
                    for (int e = tid; e < 3 * h2; e += NTHREADS) {
                        const int which = e / h2, off = 2 * (e - which * h2);
                        const double *src = (which == 0 ? gL1 : (which == 1 ? gL2 : gLi)) + (size_t)i * nn + off;
                        const unsigned dst = (unsigned)__cvta_generic_to_shared(ops + (size_t)(buf * 3 + which) * nn + off);
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
                    }
                    asm volatile("cp.async.commit_group;" ::: "memory");
                };
                double *dn1 = sm_y1, *dn2 = sm_y2;               // dnu_{i+1}, dnu_{i+2}
                issue_stage(NB - 1, 0);
                for (int i = NB - 1; i >= 0; --i) {
                    const int buf = (NB - 1 - i) & 1;
                    const bool has1 = (i + 1 < NB), has2 = (i + 2 < NB) && S.has_a2;
                    if (i > 0) { issue_stage(i - 1, buf ^ 1); asm volatile("cp.async.wait_group 1;" ::: "memory"); }
                    else asm volatile("cp.async.wait_group 0;" ::: "memory");
                    __syncthreads();
                    const double *sL1 = ops + (size_t)(buf * 3) * nn, *sL2 = sL1 + nn, *sLi = sL2 + nn;
                    for (int k = tid >> 2; k < RP; k += NTHREADS / 4) {   // v[k] : 4 threads per column k, rows split by q
                        double s = 0.0;
                        if (k < n) {
                            if (has1) for (int r = q; r < n; r += 4) s = fma(sL1[r * n + k], dn1[r], s);
                            if (has2) for (int r = q; r < n; r += 4) s = fma(sL2[r * n + k], dn2[r], s);
                        }
                        s += __shfl_xor_sync(0xffffffffu, s, 1);
                        s += __shfl_xor_sync(0xffffffffu, s, 2);
                        if (k < n && q == 0) sm_rhs[k] = yv[i * n + k] - s;
                    }
                    __syncthreads();
                    for (int k = tid >> 2; k < RP; k += NTHREADS / 4) {
                        double s = 0.0;
                        if (k < n) for (int r = k + q; r < n; r += 4) s = fma(sLi[r * n + k], sm_rhs[r], s);
                        s += __shfl_xor_sync(0xffffffffu, s, 1);
                        s += __shfl_xor_sync(0xffffffffu, s, 2);
                        if (k < n && q == 0) { dnu[i * n + k] = s; dn2[k] = s; }     // dn2 is dead: it becomes dnu_i
                    }
                    { double *t2 = dn1; dn1 = dn2; dn2 = t2; }       // (dnu_{i+1}, dnu_{i+2}) <- (dnu_i, dnu_{i+1})
                    __syncthreads();
                }
            } else {
                for (int i = NB - 1; i >= 0; --i) {
                    const bool has1 = (i + 1 < NB), has2 = (i + 2 < NB) && S.has_a2;
                    for (int k = tid >> 2; k < RP; k += NTHREADS / 4) {   // v[k] : 4 threads per column k, rows split by q
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
                    for (int k = tid >> 2; k < RP; k += NTHREADS / 4) {
                        double s = 0.0;
                        if (k < n) for (int r = k + q; r < n; r += 4) s = fma(gLi[(size_t)i * nn + r * n + k], sm_rhs[r], s);
                        s += __shfl_xor_sync(0xffffffffu, s, 1);
                        s += __shfl_xor_sync(0xffffffffu, s, 2);
                        if (k < n && q == 0) dnu[i * n + k] = s;
                    }
                    __syncthreads();
                }
            }

            PB_T(6);
            // ---- dz = inv(Phi)(-r_d - C' dnu)  (:34-35) ----
            stage_rows(kc, dnu, T, 0);
            __syncthreads();
            mma_apply_Ct<1>(kc, dnu, hdu, hdx, rdu, rdx, pinv, du, dx);
            __syncthreads();

            PB_T(7);
            // ---- backtracking on ||[r_p; r_d]||, d frozen (backtracking_inf_newton.m:2-11) ----
            double t = 1.0;
            int nh = 0;
            for (;;) {
                // trial point staged for C z_t; dual residual with the SAME expressions / summation order as above
                double sst = 0.0;
                for (int e = tid; e < usp; e += NTHREADS) {
                    const int tt = e / ldp, j = e - tt * ldp;
                    double uv = 0.0;
                    if (tt < T && j < m) {
                        const size_t idx = (size_t)tt * m + j;
                        uv = __fma_rn(t, du[idx], u[idx]);
                        un[idx] = uv;
                        const double r = rdu_expr(__ldg(S.r2 + j), __ldg(S.rl + j), uv, __fma_rn(t, hdu[idx], hu[idx]), dbar[idx]);
                        sst = fma(r, r, sst);
                    }
                    Us[e] = uv;
                }
                for (int e = tid; e < xsp; e += NTHREADS) {
                    const int rr = e / ld, k = e - rr * ld, sr = rr - 2;
                    double xv = 0.0;
                    if (sr >= 0 && sr < T && k < n) {
                        const size_t idx = (size_t)sr * n + k;
                        const bool last = (sr == T - 1);
                        xv = __fma_rn(t, dx[idx], x[idx]);
                        xn[idx] = xv;
                        const double r = rdx_expr(__ldg((last ? S.q2f : S.q2) + k), __ldg((last ? S.qfl : S.ql) + k), xv,
                                                  __fma_rn(t, hdx[idx], hx[idx]));
                        sst = fma(r, r, sst);
                    }
                    Xs[e] = xv;
                }
                __syncthreads();
                mma_apply_C(kc, bv, rpt, false);
                __syncthreads();
                double sspt = 0.0;
                for (int e = tid; e < NB * n; e += NTHREADS) sspt = fma(rpt[e], rpt[e], sspt);
                const double tp = block_sum(sspt, red);
                const double td = block_sum(sst, red);
                const double nrt = sqrt(td + tp);
                if (!(nrt > (1.0 - A.alpha * t) * nr0)) break;
                if (t == 0.0) break;
                if (A.ls_max > 0 && nh >= A.ls_max) { status = ST_LS_MAX; break; }
                t *= A.beta;
                ++nh;
            }
            PB_T(8);
            // accept: swap iterate / r_p buffers, advance the dual images
            { double *s1 = u; u = un; un = s1; double *s2 = x; x = xn; xn = s2; double *s3 = rp; rp = rpt; rpt = s3; }
            for (int e = tid; e < T * m; e += NTHREADS) hu[e] = __fma_rn(t, hdu[e], hu[e]);
            for (int e = tid; e < T * n; e += NTHREADS) hx[e] = __fma_rn(t, hdx[e], hx[e]);
            for (int e = tid; e < NB * n; e += NTHREADS) nu[e] = __fma_rn(t, dnu[e], nu[e]);
            ++iters;
            __syncthreads();
        }
        __syncthreads();
        if (u != uo) {                                       // the iterate ended in the scratch buffers
            for (int e = tid; e < T * m; e += NTHREADS) uo[e] = u[e];
            for (int e = tid; e < T * n; e += NTHREADS) xo[e] = x[e];
        }
        if (tid == 0) {
            if (A.status) A.status[b] = status;
            if (A.iters) A.iters[b] = iters;
            atomicAdd(A.iters_total, (unsigned long long)iters);
            PB_PRINT;
        }
    }
}

// =============================================================================================
template <int NP>
static int config_np(const Geom &G, SolveLaunchCfg *cfg, int sms)
{
    const size_t smem = G.smem_doubles() * sizeof(double);
    if (cudaFuncSetAttribute(fmpc_solve_kernel_mma<NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -4;
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fmpc_solve_kernel_mma<NP>, KCfg<NP>::NTH, smem) != cudaSuccess || per_sm < 1)
        return -5;
    if (const char *e = getenv("FMPC_CTAS_PER_SM")) { const int v = atoi(e); if (v >= 1 && v < per_sm) per_sm = v; }   // experiments
    cfg->grid = sms * per_sm;
    cfg->block = KCfg<NP>::NTH;
    cfg->smem = smem;
    cfg->use_mma = 1;
    cfg->np = NP;
    return 0;
}

int fmpc_mma_config(const DevSys &S, int device, SolveLaunchCfg *cfg)
{
    if (S.n > 72) return -1;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return -2;
    const Geom G = Geom::make(S.n, S.m, S.T);
    if (G.smem_doubles() * sizeof(double) > (size_t)prop.sharedMemPerBlockOptin) return -3;
    if (S.n > 32) return config_np<72>(G, cfg, prop.multiProcessorCount);
    switch (G.KP) {
    case 8: return config_np<8>(G, cfg, prop.multiProcessorCount);
    case 16: return config_np<16>(G, cfg, prop.multiProcessorCount);
    case 28: return config_np<28>(G, cfg, prop.multiProcessorCount);
    default: return config_np<32>(G, cfg, prop.multiProcessorCount);
    }
}

void fmpc_launch_solve_mma(const DevSys &S, const StepArgs &A, const SolveLaunchCfg &cfg, void *stream)
{
    int grid = cfg.grid < A.nbatch ? cfg.grid : A.nbatch;
    if (grid < 1) grid = 1;
    cudaStream_t st = (cudaStream_t)stream;
    switch (cfg.np) {
    case 72: fmpc_solve_kernel_mma<72><<<grid, cfg.block, cfg.smem, st>>>(S, A); break;
    case 8: fmpc_solve_kernel_mma<8><<<grid, cfg.block, cfg.smem, st>>>(S, A); break;
    case 16: fmpc_solve_kernel_mma<16><<<grid, cfg.block, cfg.smem, st>>>(S, A); break;
    case 28: fmpc_solve_kernel_mma<28><<<grid, cfg.block, cfg.smem, st>>>(S, A); break;
    default: fmpc_solve_kernel_mma<32><<<grid, cfg.block, cfg.smem, st>>>(S, A); break;
    }
}
