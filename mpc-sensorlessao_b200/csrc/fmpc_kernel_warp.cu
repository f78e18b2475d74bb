// fastMPC batched Newton solve -- warp-per-instance DMMA kernel for n <= 32 (sm_100a, fp64).
//
// One WARP owns one MPC instance from the first residual to the last line search
// (inf_newton_solver.m:10-41 on the block structure of DESIGN.md section 2).  Nothing in the solve needs
// a CTA barrier: the warps of a CTA only share the problem constants (B, A1, A2, bounds, weights) that
// are staged once in shared memory.  Every dense contraction is an FP64 tensor-pipe instruction
// (mma.sync.aligned.m8n8k4.f64 = DMMA); on sm_100a DMMA and DFMA share one FP64 pipe per SM sub-partition
// (profiles/r01_ubench_fp64_pipes_latency.log), so scalar FP64 work is kept to the barrier terms and the
// 8 x 8 diagonal blocks.
//   * C z, C inv(Phi) r_d and the trial residuals are horizon GEMMs  [T x (m+2n)] * [(m+2n) x n]  with
//     the stage index as the M dimension; the barrier terms / trial point are computed by the thread
//     that owns the A-fragment element, in registers, straight from global memory (software pipelined);
//   * C' v is the transposed pair  [T x n] * [n x m],  [T x n] * [n x n];
//   * band-2 block Cholesky of Y = C inv(Phi) C' per stage i, entirely in registers:
//       S_i  = Yd_i + B diag(w_i) B' - L1_{i-1} L1_{i-1}' - L2_{i-2} L2_{i-2}'     (lower 8x8 tiles, accumulators)
//       L_i  = chol(S_i), inv(L_i): blocked right-looking on the accumulator tiles.  A tile in accumulator
//              layout {row gq, cols 2q, 2q+1} is at the same time a valid A operand (k = 2q+e) and a valid B
//              operand of X T' (n = gq, k = 2q+e), so panel solves, trailing updates and the triangular
//              inverse are DMMAs on registers; only the 8x8 diagonal blocks are factored/inverted with
//              shuffles.  The right-hand side rides along as one more row of S (row n), so the forward
//              substitution y_i = inv(L_i) rhs_i falls out of the panel operations.
//       L1_i = (Y1_i - L2_{i-1} L1_{i-1}') inv(L_i)' ,  L2_i = Y2 inv(L_i)'         (X T' with T = inv(L) tiles)
//   * the factor (inv(L_i), L1_i, L2_i) streams to a per-warp global scratch and comes back through a
//     cp.async ring for the backward substitution.
// Shared memory per warp: 3 history blocks (L1_{i-1}, L2_{i-1}, L2_{i-2}; row n of each carries y) + the
// staged w_i rows; per CTA: the constants.  8 warps (= 8 instances) per SM.
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
#define PROF_PARAMS , long long (&p_acc)[12], long long &p_last
#define PROF_ARGS , p_acc, p_last
#else
#define PROF_DECL
#define PROF_T(idx) do { } while (0)
#define PROF_PARAMS
#define PROF_ARGS
#endif

__device__ __forceinline__ void dmma(double (&c)[2], const double a, const double b)
{
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
// o += X T'   for 8 x 8 tiles X, T both in accumulator layout (k runs over {2q+e})
__device__ __forceinline__ void mma_xt(double (&o)[2], const double (&x)[2], const double (&t)[2])
{
    dmma(o, x[0], t[0]);
    dmma(o, x[1], t[1]);
}
__device__ __forceinline__ double dneg(const double x)
{
    return __hiloint2double(__double2hiint(x) ^ (int)0x80000000, __double2loint(x));
}
__device__ __forceinline__ double shfl_d(const double v, const int src) { return __shfl_sync(FULL, v, src); }
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
// 1/sqrt(x), x > 0 normal : MUFU.RSQ64H seed + one cubic step
__device__ __forceinline__ double rsqrt_nr(const double x)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    // cubic (Halley) step: y <- y (1 + e/2 + 3 e^2/8), e = 1 - x y^2 : the seed's relative error 2^-20 becomes ~2^-58, below the
    // rounding of the result.  (Round 1 added a quadratic polish step on top: 4 more FP64 operations on the pivot chain of every
    // column for nothing measurable -- all parity tests incl. iteration counts are unchanged without it, +1.3 % throughput.)
    const double e = fma(-x * y, y, 1.0);
    y = fma(y * e, fma(0.375, e, 0.5), y);
    return y;
}
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
// cp.async.cg: the scratch streams are read once and must not pass through L1 -- the 7 resident warps leave only ~29 KB of
// the unified array to L1, which has to keep the pool of iterate-independent Schur blocks (Yd / Y1, __ldg) and the operand
// rows the passes re-read.  Measured +10 % over .ca (profiles/r02_l1_bypass.log); bypassing L1 for the plain scratch loads as
// well costs 1.5 %.
__device__ __forceinline__ void cp_async16(double *smem_dst, const double *gsrc)
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
// DRAM -> L2 prefetch of `bytes` bytes at p (one 128-byte line per lane and step); no registers are tied up
__device__ __forceinline__ void l2_prefetch(const double *p, const int bytes, const int lane)
{
    const char *b = reinterpret_cast<const char *>(p);
    for (int o = lane * 128; o < bytes; o += 32 * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(b + o));
}
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
// Geometry (host + device).  NPOT = padded block size (multiple of 4 = k-steps of the block products),
// CT = 8-column tiles, RT = 8-row tiles including row n (the row that carries the forward substitution).
// Conventions that let the operand loads run without predicates:
//   * an operand row that does not exist (row > n of a block, row >= n of B / A1 / A2) only ever feeds an
//     output row or column that is never used, so it may read whatever FINITE data follows in shared memory
//     (all shared memory is zero-initialised and only finite values are ever stored);
//   * the k (contraction) index is always < NPOT and padded with exact zeros on at least one side.
// ---------------------------------------------------------------------------------------------
__host__ __device__ inline int npot_of(int n) { return n <= 8 ? 8 : (n <= 16 ? 16 : (n <= 24 ? 24 : (n <= 28 ? 28 : 32))); }
__host__ __device__ inline int ld_of(int npot) { return (npot % 8 == 4) ? npot : npot + 4; }

struct WGeom {
    int n, m, T, NPOT, CT, RT, KS, LD, BLK, WB, mpad, MK, MT8, npad, LDB, NTT, TP8, NN;
    size_t const_doubles, warp_doubles, tail_doubles;
    __host__ __device__ static WGeom make(int n, int m, int T)
    {
        WGeom g;
        g.n = n; g.m = m; g.T = T;
        g.NPOT = npot_of(n);
        g.CT = (g.NPOT + 7) / 8;
        g.RT = n / 8 + 1;
        g.KS = g.NPOT / 4;
        g.LD = ld_of(g.NPOT);
        g.BLK = ((n + 1) * g.LD + 1) & ~1;
        g.MT8 = (m + 7) / 8;
        g.mpad = 8 * g.MT8;
        g.MK = g.mpad / 4;
        g.npad = 8 * g.CT;
        g.LDB = (g.mpad % 8 == 4) ? g.mpad : g.mpad + 4;
        g.NTT = (T + 7) / 8;
        g.TP8 = 8 * g.NTT;
        g.NN = g.NPOT * g.NPOT;
        g.WB = 2 * g.mpad + 128;                       // two staged w rows + the stash of the paired last tile row (forward_sweep)
        if (g.WB < 160) g.WB = 160;                    // also holds the 5 x 32 vectors of the backward sweep
        //          B            A1, A2                umax umin r2 rl     q2 q2f ql qfl qi qif
        g.const_doubles = (size_t)n * g.LDB + 2 * ((size_t)n * g.LD) + 4 * (size_t)g.mpad + 6 * (size_t)g.npad;
        g.const_doubles = (g.const_doubles + 1) & ~(size_t)1;
        g.warp_doubles = 3 * (size_t)g.BLK + (size_t)g.WB;
        if (g.warp_doubles < 3 * (size_t)g.NN + 160) g.warp_doubles = 3 * (size_t)g.NN + 160;   // backward ring: 3 slots of NN + 5 x 32 vectors
        if (g.warp_doubles < 2400) g.warp_doubles = 2400;                                   // u-space stream ring of the passes
        g.tail_doubles = 8 * (size_t)g.LD + 8;   // overrun reads of the last rows stay inside the allocation
        return g;
    }
    __host__ __device__ size_t smem_doubles(int warps) const { return const_doubles + (size_t)warps * warp_doubles + tail_doubles; }
};

// per-warp global scratch (doubles); u- and x-space arrays have TP8 = 8 ceil(T/8) rows so that the 8-row tiles
// of the horizon GEMMs never need a row predicate (rows >= T hold zeros / finite junk that is never used)
struct WsW {
    size_t tu, tx, tb, bl, total;
    __host__ __device__ static WsW make(const WGeom &g)
    {
        WsW L;
        L.tu = (size_t)g.TP8 * g.mpad;                 // UC UT HU HUT HDU DU WV DB RDU
        L.tx = (size_t)g.TP8 * g.npad;                 // XC XT HX HXT HDX DX RDX
        L.tb = (size_t)(g.TP8 + 1) * g.npad;           // RP RPT YV DNU BV
        L.bl = (size_t)(g.T + 1) * g.NN;               // inv(L), L1  (row-major, leading dimension NPOT; L2 is rebuilt by the backward sweep)
        L.total = (9 * L.tu + 7 * L.tx + 5 * L.tb + 2 * L.bl + 31) & ~(size_t)31;
        return L;
    }
};

struct WCtx {
    int n, m, T, NB, a2, has_xf;
    int mpad, MK, MT8, LDB, NTT;
    int lane, gq, q;
    double kappa;
    // Everything below is addressed as base + index * stride, computed where it is used, so that a pass keeps
    // only the few pointers it needs in registers.
    double *smem;                       // CTA constants: B | A1 | A2 | umax umin r2 rl | q2[2] ql[2] qi[2]
    int oA1, oA2, oU, oQ, npad;
    double *wsm;                        // this warp's shared memory: 3 history blocks | w rows
    int BLK;
    double *ws;                         // this warp's global scratch (layout WsW)
    int tu, tx, tb, bl;
    int pp;                             // ping-pong parity of the iterate buffers
    const double *ypool;
    const int *ydi, *y1i, *y2i;
    // start of the iterate (K_RP pass): warm start arrays of this instance or NULL = midpoint cold start
    const double *U0, *X0, *xmin, *xmax;
    int sh;                             // 1: read the warm start shifted one stage (resident closed loop), else 0

    __device__ __forceinline__ const double *sB() const { return smem; }
    __device__ __forceinline__ const double *sA1() const { return smem + oA1; }
    __device__ __forceinline__ const double *sA2() const { return smem + oA2; }
    __device__ __forceinline__ const double *sUmax() const { return smem + oU; }
    __device__ __forceinline__ const double *sUmin() const { return smem + oU + mpad; }
    __device__ __forceinline__ const double *sR2() const { return smem + oU + 2 * mpad; }
    __device__ __forceinline__ const double *sRl() const { return smem + oU + 3 * mpad; }
    __device__ __forceinline__ const double *sQ2() const { return smem + oQ; }
    __device__ __forceinline__ const double *sQl() const { return smem + oQ + 2 * npad; }
    __device__ __forceinline__ const double *sQi() const { return smem + oQ + 4 * npad; }
    __device__ __forceinline__ double *blk(int k) const { return wsm + k * BLK; }
    __device__ __forceinline__ double *wbuf() const { return wsm + 3 * BLK; }
    __device__ __forceinline__ double *ua(int a) const { return ws + (size_t)a * tu; }
    __device__ __forceinline__ double *xa(int a) const { return ws + (size_t)9 * tu + (size_t)a * tx; }
    __device__ __forceinline__ double *ba(int a) const { return ws + (size_t)9 * tu + (size_t)7 * tx + (size_t)a * tb; }
    __device__ __forceinline__ double *fa(int a) const { return ws + (size_t)9 * tu + (size_t)7 * tx + (size_t)5 * tb + (size_t)a * bl; }
    __device__ __forceinline__ double *UC() const { return ua(pp); }
    __device__ __forceinline__ double *UT() const { return ua(pp ^ 1); }
    __device__ __forceinline__ double *HU() const { return ua(2 + pp); }
    __device__ __forceinline__ double *HUT() const { return ua(2 + (pp ^ 1)); }
    __device__ __forceinline__ double *HDU() const { return ua(4); }
    __device__ __forceinline__ double *DU() const { return ua(5); }
    __device__ __forceinline__ double *WV() const { return ua(6); }
    __device__ __forceinline__ double *DB() const { return ua(7); }
    __device__ __forceinline__ double *RDU() const { return ua(8); }
    __device__ __forceinline__ double *XC() const { return xa(pp); }
    __device__ __forceinline__ double *XT() const { return xa(pp ^ 1); }
    __device__ __forceinline__ double *HX() const { return xa(2 + pp); }
    __device__ __forceinline__ double *HXT() const { return xa(2 + (pp ^ 1)); }
    __device__ __forceinline__ double *HDX() const { return xa(4); }
    __device__ __forceinline__ double *DX() const { return xa(5); }
    __device__ __forceinline__ double *RDX() const { return xa(6); }
    __device__ __forceinline__ double *RP() const { return ba(pp); }
    __device__ __forceinline__ double *RPT() const { return ba(pp ^ 1); }
    __device__ __forceinline__ double *YV() const { return ba(2); }
    __device__ __forceinline__ double *DNU() const { return ba(3); }
    __device__ __forceinline__ double *BV() const { return ba(4); }
    __device__ __forceinline__ double *gLi() const { return fa(0); }
    __device__ __forceinline__ double *gL1() const { return fa(1); }
};

// K_NORM: the residual norms of K_NEWTON (same expressions, same summation order) without its GEMM and scratch stores.
// Not instantiated at present: the accepted K_TRIAL pass now returns the next iteration's residual itself (ss_f), which
// made the separate norm-only pass of v3.8 unnecessary; the variant is kept for callers that need the norms alone.
enum { K_RP = 0, K_NEWTON = 1, K_TRIAL = 2, K_NORM = 3 };

// ---------------------------------------------------------------------------------------------
// u-space streams of the passes: chunks of (8 TT rows) x (CW columns) of NA arrays go through a 4-slot cp.async ring in
// this warp's shared memory (free while no sweep is running).  No registers are tied up by data in flight, the ring
// runs 4 chunks ahead, and the consumer reads its fragment elements with conflict-free LDS.
// Slot layout [array][row 0..23][RS], RS chosen so that the fragment reads are conflict-free.
// ---------------------------------------------------------------------------------------------
constexpr int STREAM_RD = 4;
__device__ __forceinline__ void cp_async16_s(const unsigned saddr, const double *gsrc)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(gsrc) : "memory");
}
// per-lane piece table of one chunk shape (fixed over the chunks of a pass): 16-byte piece p = lane + 32 it covers
// row p / PPR, columns 2 (p % PPR) .. +1 of the chunk
template <int CW> struct StreamIdx {
    static constexpr int PPR = CW / 2, NIT = (8 * TTMAX * PPR + 31) / 32;
    int soff[NIT];
    unsigned doff[NIT];
    bool ok[NIT];
};
template <int CW, int RS>
__device__ __forceinline__ void stream_setup(StreamIdx<CW> &si, const WCtx &c, const int TT)
{
#pragma unroll
    for (int it = 0; it < StreamIdx<CW>::NIT; ++it) {
        const int p = c.lane + 32 * it, row = p / StreamIdx<CW>::PPR, pc = p - row * StreamIdx<CW>::PPR;
        si.ok[it] = p < 8 * TT * StreamIdx<CW>::PPR;
        si.soff[it] = row * c.mpad + 2 * pc;
        si.doff[it] = (unsigned)((row * RS + 2 * pc) * 8);
    }
}
template <int NA, int CW, int RS, int RD>
__device__ __forceinline__ void stream_issue(const WCtx &c, const StreamIdx<CW> &si, const double *const (&arr)[NA], const unsigned sbase,
                                             const int tt0, const int ch, const int nch)
{
    if (ch < nch) {
        const unsigned sl = sbase + (unsigned)(ch % RD) * (unsigned)(NA * 8 * TTMAX * RS * 8);
#pragma unroll
        for (int a = 0; a < NA; ++a) {
            const double *src = arr[a] + (size_t)(8 * tt0) * c.mpad + CW * ch;
#pragma unroll
            for (int it = 0; it < StreamIdx<CW>::NIT; ++it)
                if (si.ok[it]) cp_async16_s(sl + (unsigned)(a * 8 * TTMAX * RS * 8) + si.doff[it], src + si.soff[it]);
        }
    }
    cp_async_commit();
}

// ---------------------------------------------------------------------------------------------
// Horizon GEMM  acc[t][k] = (B v_u,t + A1 v_x,t-1 + A2 v_x,t-2)[k]  and its three uses
//   K_RP     : v = z0 (start)     ; UC, XC <- z0 ; RP  = x_t+1 - acc - b      (fast_mpc_init.m:12-26, r_p = C z - b)
//   K_NEWTON : v = inv(Phi) r_d   ; YV  = r_p - C v ; barrier terms, r_d, norms (inf_newton_KKT_H.m:3-13, :12, :28-29)
//   K_TRIAL  : v = z + ts dz      ; RPT = C v - b ; r_d(ts) with d frozen, norms (backtracking_inf_newton.m:4);
//              the dual images at the trial point go to HUT / HXT so that accepting the step is a pointer swap
// ss_d / ss_p return this lane's partial sums of squares (same accumulation order in NEWTON and TRIAL).
// Every global stream is software pipelined: the loads of step k+1 are issued before the arithmetic and the
// stores of step k (the scratch pointers may alias as far as the compiler knows, so it cannot do this itself).
// ---------------------------------------------------------------------------------------------
template <int KIND> struct UArr { static constexpr int N = (KIND == K_RP) ? 1 : ((KIND == K_NEWTON || KIND == K_NORM) ? 2 : 5); };

template <int NPOT, int KIND>
__device__ __forceinline__ void pass_Cv(const WCtx &c, const double ts, double &ss_d, double &ss_p, double &ss_f)
{
    constexpr int CT = (NPOT + 7) / 8, KS = NPOT / 4, LD = (NPOT % 8 == 4) ? NPOT : NPOT + 4, NA = UArr<KIND>::N, npad = 8 * CT;
    constexpr int XU = 4;
    const int n = c.n, m = c.m, T = c.T, mpad = c.mpad, gq = c.gq, q = c.q, lane = c.lane;
    ss_d = 0.0; ss_p = 0.0; ss_f = 0.0;
    // the u-space streams of this pass were written long ago (they have left L2 for DRAM): pull them back into L2 now,
    // the register pipeline below then only has to cover L2 latency
    {
        const int ub = T * mpad * 8;
        (void)ub;
        if (KIND == K_RP) { if (c.U0) l2_prefetch(c.U0, T * m * 8, lane); }
    }
    // ---- x-space elementwise (linear ownership, XU elements in flight per lane), then visible to the whole warp ----
    {
        const int tot = T * npad;
        double *pXC = c.XC(), *pHX = c.HX();
        for (int e0 = lane; e0 < tot; e0 += 32 * XU) {
            double a[XU], b[XU], d[XU], h[XU];
#pragma unroll
            for (int u = 0; u < XU; ++u) {
                const int e = e0 + 32 * u;
                a[u] = b[u] = d[u] = h[u] = 0.0;
                if (e < tot) {
                    if (KIND == K_RP) {
                        const int t = e / npad, k = e - t * npad;
                        if (k < n) a[u] = c.X0 ? c.X0[(size_t)min(t + c.sh, T - 1) * n + k] : (c.xmin[k] + c.xmax[k]) / 2;
                    } else if (KIND == K_NEWTON || KIND == K_NORM) {
                        a[u] = pXC[e]; b[u] = pHX[e];
                    } else {
                        a[u] = pXC[e]; b[u] = pHX[e]; d[u] = c.DX()[e]; h[u] = c.HDX()[e];
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < XU; ++u) {
                const int e = e0 + 32 * u;
                if (e < tot) {
                    const int t = e / npad, k = e - t * npad, st = (t == T - 1) ? npad : 0;
                    if (KIND == K_RP) {
                        pXC[e] = a[u];
                    } else if (KIND == K_NEWTON || KIND == K_NORM) {
                        const double r = rdx_expr(c.sQ2()[st + k], c.sQl()[st + k], a[u], b[u]);
                        ss_d = fma(r, r, ss_d);
                        if (KIND == K_NEWTON) {
                            c.RDX()[e] = r;
                            c.DX()[e] = r * c.sQi()[st + k];                    // p_x = inv(2Q) r_dx
                        }
                    } else {
                        const double xv = __fma_rn(ts, d[u], a[u]);
                        const double hv = __fma_rn(ts, h[u], b[u]);
                        c.XT()[e] = xv;
                        c.HXT()[e] = hv;
                        const double r = rdx_expr(c.sQ2()[st + k], c.sQl()[st + k], xv, hv);
                        ss_d = fma(r, r, ss_d);
                    }
                }
            }
        }
    }
    __syncwarp();
    // K_TRIAL also returns ss_f: the dual residual the NEXT iteration's K_NEWTON pass would find at this trial point (barrier
    // gradient re-evaluated there; same expressions, same summation order, so the sums are bit-identical).  If the trial is
    // accepted, the early-exit test of the next iteration (inf_newton_solver.m:19-22) is decided without another pass.
    if (KIND == K_TRIAL) ss_f = ss_d;
    const double *xsrc = (KIND == K_RP) ? c.XC() : (KIND == K_NEWTON ? c.DX() : c.XT());
    // operand row pointers (no predicates: see the conventions above WGeom)
    const double *pB[CT];
#pragma unroll
    for (int nt = 0; nt < CT; ++nt) pB[nt] = c.sB() + (size_t)(8 * nt + gq) * c.LDB + q;
    const double *pA1 = c.sA1() + gq * LD + q, *pA2 = c.sA2() + gq * LD + q;
    const double *pUmax = c.sUmax() + q, *pUmin = c.sUmin() + q, *pR2 = c.sR2() + q, *pRl = c.sRl() + q;
    // columns k >= n of the outputs must be stored as exact zeros (they pad the k index of later products)
    const bool kok0 = (8 * (CT - 1) + 2 * q) < n, kok1 = (8 * (CT - 1) + 2 * q + 1) < n;

    for (int tt0 = 0; tt0 < c.NTT; tt0 += TTMAX) {
        const int TT = min(TTMAX, c.NTT - tt0);
        double acc[TTMAX][CT][2];
#pragma unroll
        for (int tt = 0; tt < TTMAX; ++tt)
#pragma unroll
            for (int nt = 0; nt < CT; ++nt) acc[tt][nt][0] = acc[tt][nt][1] = 0.0;
        int row[TTMAX];                                      // element offset of (t, q) in a u-space array
        bool tok[TTMAX];
#pragma unroll
        for (int tt = 0; tt < TTMAX; ++tt) {
            const int t = 8 * (tt0 + tt) + gq;
            tok[tt] = (tt < TT) && (t < T);
            row[tt] = ((tt < TT) ? t : 0) * mpad + q;        // rows of dead tiles alias row 0 (finite data, unused result)
        }
        // ---- u part: A fragment element (t = 8 tt + gq, j = 4 kk + q) is produced by its owner ----
        if (KIND == K_RP) {
            // start of the iterate straight from the caller's arrays (leading dimension m): register pipeline, kk in pairs
            double cur[2][TTMAX], nxt[2][TTMAX];
            auto load = [&](double (&bf)[2][TTMAX], const int kp) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int j4 = 8 * kp + 4 * h;
#pragma unroll
                    for (int tt = 0; tt < TTMAX; ++tt) {
                        const int t = 8 * (tt0 + tt) + gq, j = j4 + q;
                        bf[h][tt] = (tok[tt] && j < m) ? (c.U0 ? c.U0[(size_t)min(t + c.sh, T - 1) * m + j] : (pUmin[j4] + pUmax[j4]) / 2) : 0.0;
                    }
                }
            };
            const int NKP = c.MK / 2;
            load(cur, 0);
            for (int kp = 0; kp < NKP; ++kp) {
                if (kp + 1 < NKP) load(nxt, kp + 1);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int j4 = 8 * kp + 4 * h;
                    double bf[CT];
#pragma unroll
                    for (int nt = 0; nt < CT; ++nt) bf[nt] = pB[nt][j4];
#pragma unroll
                    for (int tt = 0; tt < TTMAX; ++tt)
                        if (tt < TT) {
                            c.UC()[row[tt] + j4] = cur[h][tt];
#pragma unroll
                            for (int nt = 0; nt < CT; ++nt) dmma(acc[tt][nt], cur[h][tt], bf[nt]);
                        }
                }
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int tt = 0; tt < TTMAX; ++tt) cur[h][tt] = nxt[h][tt];
            }
        } else {
            // NEWTON: UC, HU in chunks of 2 kk ; TRIAL: UC, DU, HU, HDU, DB in chunks of 1 kk
            constexpr bool NWT = (KIND == K_NEWTON || KIND == K_NORM);
            constexpr int CK = NWT ? 2 : 1, CW = 4 * CK, RS = (CK == 1) ? 4 : 12;
            constexpr int RD = (NA * 8 * TTMAX * RS * 5 <= 2400) ? 5 : 4;         // ring depth: what fits the guaranteed 2400 doubles
            const double *arr[NA];
            if (NWT) { arr[0] = c.UC(); arr[NA > 1 ? 1 : 0] = c.HU(); }
            else { arr[0] = c.UC(); arr[NA > 1 ? 1 : 0] = c.DU(); arr[NA > 2 ? 2 : 0] = c.HU(); arr[NA > 3 ? 3 : 0] = c.HDU(); arr[NA > 4 ? 4 : 0] = c.DB(); }
            const int nch = c.MK / CK;
            StreamIdx<CW> si;
            stream_setup<CW, RS>(si, c, TT);
            const unsigned sbase = (unsigned)__cvta_generic_to_shared(c.wsm);
#pragma unroll
            for (int d = 0; d < RD; ++d) stream_issue<NA, CW, RS, RD>(c, si, arr, sbase, tt0, d, nch);
            double *pDB = c.DB(), *pWV = c.WV(), *pRDU = c.RDU(), *pUT = c.UT(), *pHUT = c.HUT();
            for (int ch = 0; ch < nch; ++ch) {
                cp_async_wait<RD - 1>();
                __syncwarp();
                double v[CK][TTMAX][NA];
                {
                    const double *slot = c.wsm + (ch % RD) * (NA * 8 * TTMAX * RS) + gq * RS + q;
#pragma unroll
                    for (int h = 0; h < CK; ++h)
#pragma unroll
                        for (int tt = 0; tt < TTMAX; ++tt)
#pragma unroll
                            for (int a = 0; a < NA; ++a) v[h][tt][a] = (tt < TT) ? slot[(a * 8 * TTMAX + 8 * tt) * RS + 4 * h] : 0.0;
                }
                __syncwarp();                                    // the slot is in registers: refill it
                stream_issue<NA, CW, RS, RD>(c, si, arr, sbase, tt0, ch + RD, nch);
#pragma unroll
                for (int h = 0; h < CK; ++h) {
                    const int j4 = CW * ch + 4 * h;
                    double a[TTMAX];
                    double bf[CT];
#pragma unroll
                    for (int nt = 0; nt < CT; ++nt) bf[nt] = pB[nt][j4];
                    const double r2 = pR2[j4], rl = pRl[j4];
#pragma unroll
                    for (int tt = 0; tt < TTMAX; ++tt) {
                        const int idx = row[tt] + j4;
                        if (NWT) {
                            const double uu = v[h][tt][0], hh = v[h][tt][NA > 1 ? 1 : 0];
                            const double sp = pUmax[j4] - uu, sm = uu - pUmin[j4];
                            const double dp = rcp_nr(sp), dm = rcp_nr(sm);
                            const double db = c.kappa * (dp - dm);
                            const double w = rcp_nr(fma(c.kappa, fma(dp, dp, dm * dm), r2));
                            const double r = rdu_expr(r2, rl, uu, hh, db);
                            if (KIND == K_NEWTON && tt < TT) { pDB[idx] = db; pWV[idx] = w; pRDU[idx] = r; }
                            const double rm = tok[tt] ? r : 0.0;
                            ss_d = fma(rm, rm, ss_d);
                            a[tt] = r * w;                                  // p_u = inv(Phi_u) r_du
                        } else {
                            const double uv = __fma_rn(ts, v[h][tt][NA > 1 ? 1 : 0], v[h][tt][0]);
                            const double hv = __fma_rn(ts, v[h][tt][NA > 3 ? 3 : 0], v[h][tt][NA > 2 ? 2 : 0]);
                            if (tt < TT) { pUT[idx] = uv; pHUT[idx] = hv; }
                            const double r = rdu_expr(r2, rl, uv, hv, v[h][tt][NA > 4 ? 4 : 0]);
                            const double rm = tok[tt] ? r : 0.0;
                            ss_d = fma(rm, rm, ss_d);
                            {   // the same element with the barrier gradient of the trial point (what K_NEWTON computes next)
                                const double sp = pUmax[j4] - uv, sm = uv - pUmin[j4];
                                const double dp = rcp_nr(sp), dm = rcp_nr(sm);
                                const double rf = rdu_expr(r2, rl, uv, hv, c.kappa * (dp - dm));
                                const double rfm = tok[tt] ? rf : 0.0;
                                ss_f = fma(rfm, rfm, ss_f);
                            }
                            a[tt] = uv;
                        }
                    }
                    if (KIND != K_NORM) {
#pragma unroll
                        for (int tt = 0; tt < TTMAX; ++tt)
                            if (tt < TT) {
#pragma unroll
                                for (int nt = 0; nt < CT; ++nt) dmma(acc[tt][nt], a[tt], bf[nt]);
                            }
                    }
                }
            }
            cp_async_wait<0>();
        }
        // ---- x part: A1 v_{t-1} + A2 v_{t-2} (shifted rows read from the scratch arrays, all loads up front) ----
        if (KIND != K_NORM) {
            double a1[KS][TTMAX], a2[KS][TTMAX];
#pragma unroll
            for (int tt = 0; tt < TTMAX; ++tt) {
                const int t = 8 * (tt0 + tt) + gq;
                const bool ok1 = tok[tt] && t >= 1, ok2 = tok[tt] && t >= 2;
                const double *p1 = xsrc + (size_t)(ok1 ? t - 1 : 0) * npad + q, *p2 = xsrc + (size_t)(ok2 ? t - 2 : 0) * npad + q;
#pragma unroll
                for (int kk = 0; kk < KS; ++kk) {
                    const double v1 = p1[4 * kk], v2 = p2[4 * kk];   // columns >= n of the scratch arrays are zero
                    a1[kk][tt] = ok1 ? v1 : 0.0;
                    a2[kk][tt] = ok2 ? v2 : 0.0;
                }
            }
#pragma unroll
            for (int kk = 0; kk < KS; ++kk) {
                double b1[CT], b2[CT];
#pragma unroll
                for (int nt = 0; nt < CT; ++nt) { b1[nt] = pA1[8 * nt * LD + 4 * kk]; b2[nt] = pA2[8 * nt * LD + 4 * kk]; }
#pragma unroll
                for (int tt = 0; tt < TTMAX; ++tt)
                    if (tt < TT) {
#pragma unroll
                        for (int nt = 0; nt < CT; ++nt) {
                            dmma(acc[tt][nt], a1[kk][tt], b1[nt]);
                            dmma(acc[tt][nt], a2[kk][tt], b2[nt]);          // A2 = 0 for a VAR(1) model
                        }
                    }
            }
        }
        // ---- epilogue in the accumulator layout: (t = 8 tt + gq, k = 8 nt + 2 q + e); loads first ----
        {
            double2 e1[TTMAX][CT], e2[TTMAX][CT];
#pragma unroll
            for (int tt = 0; tt < TTMAX; ++tt) {
                const int t = 8 * (tt0 + tt) + gq;
#pragma unroll
                for (int nt = 0; nt < CT; ++nt) {
                    e1[tt][nt] = e2[tt][nt] = make_double2(0.0, 0.0);
                    if (tok[tt]) {
                        const size_t idx = (size_t)t * npad + 8 * nt + 2 * q;
                        if (KIND == K_RP || KIND == K_TRIAL) {
                            e1[tt][nt] = *reinterpret_cast<const double2 *>(xsrc + idx);
                            e2[tt][nt] = *reinterpret_cast<const double2 *>(c.BV() + idx);
                        } else {
                            e1[tt][nt] = *reinterpret_cast<const double2 *>(c.RP() + idx);
                            if (KIND == K_NEWTON) e2[tt][nt] = *reinterpret_cast<const double2 *>(c.DX() + idx);
                        }
                    }
                }
            }
#pragma unroll
            for (int tt = 0; tt < TTMAX; ++tt) {
                const int t = 8 * (tt0 + tt) + gq;
                if (tok[tt]) {
#pragma unroll
                    for (int nt = 0; nt < CT; ++nt) {
                        const size_t idx = (size_t)t * npad + 8 * nt + 2 * q;
                        const bool k0 = (nt < CT - 1) || kok0, k1 = (nt < CT - 1) || kok1;
                        if (KIND == K_RP || KIND == K_TRIAL) {
                            double2 r;
                            r.x = k0 ? e1[tt][nt].x - acc[tt][nt][0] - e2[tt][nt].x : 0.0;
                            r.y = k1 ? e1[tt][nt].y - acc[tt][nt][1] - e2[tt][nt].y : 0.0;
                            *reinterpret_cast<double2 *>((KIND == K_RP ? c.RP() : c.RPT()) + idx) = r;
                            ss_p = fma(r.x, r.x, ss_p);
                            ss_p = fma(r.y, r.y, ss_p);
                        } else {
                            const double2 rp = e1[tt][nt], px = e2[tt][nt];
                            ss_p = fma(rp.x, rp.x, ss_p);
                            ss_p = fma(rp.y, rp.y, ss_p);
                            if (KIND == K_NEWTON) {
                                double2 y;
                                y.x = k0 ? rp.x - px.x + acc[tt][nt][0] : 0.0;   // -beta = r_p - C p
                                y.y = k1 ? rp.y - px.y + acc[tt][nt][1] : 0.0;
                                *reinterpret_cast<double2 *>(c.YV() + idx) = y;
                            }
                        }
                    }
                }
            }
        }
    }
    // ---- terminal row x_T = xf (fast_mpc_eq_const.m:67-71) ----
    if (c.has_xf && lane < n) {
        const size_t iT = (size_t)T * npad + lane, iL = (size_t)(T - 1) * npad + lane;
        if (KIND == K_RP || KIND == K_TRIAL) {
            const double r = xsrc[iL] - c.BV()[iT];
            (KIND == K_RP ? c.RP() : c.RPT())[iT] = r;
            ss_p = fma(r, r, ss_p);
        } else {
            const double rp = c.RP()[iT];
            ss_p = fma(rp, rp, ss_p);
            if (KIND == K_NEWTON) c.YV()[iT] = rp - c.DX()[iL];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// C' v for v = V = DNU buffer (rows t < NB, leading dimension npad, columns >= n zero):
//   hu_t = B' v_t ,  hx_t = v_t - A1' v_{t+1} - A2' v_{t+2} (+ v_T)
//   MODE 0 : HU, HX stored (images of the dual start)
//   MODE 1 : HDU, HDX stored and  du = -(r_du - hdu) w ,  dx = -(r_dx + hdx) inv(2Q)      (inf_newton_solver.m:34-35)
// ---------------------------------------------------------------------------------------------
template <int NPOT, int MODE>
__device__ __forceinline__ void pass_Ct(const WCtx &c)
{
    constexpr int CT = (NPOT + 7) / 8, KS = NPOT / 4, LD = (NPOT % 8 == 4) ? NPOT : NPOT + 4, npad = 8 * CT;
    const int n = c.n, T = c.T, mpad = c.mpad, gq = c.gq, q = c.q;
    const double *V = c.DNU();
    const double *pBk[KS];                                   // B rows k = 4 kk + q, column gq
#pragma unroll
    for (int kk = 0; kk < KS; ++kk) pBk[kk] = c.sB() + (size_t)(4 * kk + q) * c.LDB + gq;
    const double *pA1 = c.sA1() + q * LD + gq, *pA2 = c.sA2() + q * LD + gq;
    const bool kok0 = (8 * (CT - 1) + 2 * q) < n, kok1 = (8 * (CT - 1) + 2 * q + 1) < n;
    for (int tt0 = 0; tt0 < c.NTT; tt0 += TTMAX) {
        const int TT = min(TTMAX, c.NTT - tt0);
        bool tok[TTMAX];
#pragma unroll
        for (int tt = 0; tt < TTMAX; ++tt) tok[tt] = (tt < TT) && (8 * (tt0 + tt) + gq < T);
        {   // ---- u part ----
            double av[TTMAX][KS];
#pragma unroll
            for (int tt = 0; tt < TTMAX; ++tt) {
                const int t = 8 * (tt0 + tt) + gq;
                const double *pv = V + (size_t)(tok[tt] ? t : 0) * npad + q;
#pragma unroll
                for (int kk = 0; kk < KS; ++kk) { const double v = pv[4 * kk]; av[tt][kk] = tok[tt] ? v : 0.0; }
            }
            // MODE 1: r_du and w (8 columns per step) come through the cp.async ring
            const double *arr[2] = {c.RDU(), c.WV()};
            StreamIdx<8> si;
            const unsigned sbase = (unsigned)__cvta_generic_to_shared(c.wsm);
            if (MODE == 1) {
                stream_setup<8, 8>(si, c, TT);
#pragma unroll
                for (int d = 0; d < STREAM_RD; ++d) stream_issue<2, 8, 8, STREAM_RD>(c, si, arr, sbase, tt0, d, c.MT8);
            }
            for (int jt = 0; jt < c.MT8; ++jt) {
                double2 rc[TTMAX], wc[TTMAX];
                if (MODE == 1) {
                    cp_async_wait<STREAM_RD - 1>();
                    __syncwarp();
                    const double *slot = c.wsm + (jt % STREAM_RD) * (2 * 8 * TTMAX * 8) + gq * 8 + 2 * q;
#pragma unroll
                    for (int tt = 0; tt < TTMAX; ++tt) {
                        rc[tt] = wc[tt] = make_double2(0.0, 0.0);
                        if (tt < TT) {
                            rc[tt] = *reinterpret_cast<const double2 *>(slot + 8 * tt * 8);
                            wc[tt] = *reinterpret_cast<const double2 *>(slot + (8 * TTMAX + 8 * tt) * 8);
                        }
                    }
                    __syncwarp();
                    stream_issue<2, 8, 8, STREAM_RD>(c, si, arr, sbase, tt0, jt + STREAM_RD, c.MT8);
                }
                double acc[TTMAX][2];
#pragma unroll
                for (int tt = 0; tt < TTMAX; ++tt) acc[tt][0] = acc[tt][1] = 0.0;
#pragma unroll
                for (int kk = 0; kk < KS; ++kk) {
                    const double bf = pBk[kk][8 * jt];               // rows k >= n meet av = 0
#pragma unroll
                    for (int tt = 0; tt < TTMAX; ++tt)
                        if (tt < TT) dmma(acc[tt], av[tt][kk], bf);
                }
#pragma unroll
                for (int tt = 0; tt < TTMAX; ++tt) {
                    const int t = 8 * (tt0 + tt) + gq;
                    if (tok[tt]) {
                        const size_t idx = (size_t)t * mpad + 8 * jt + 2 * q;
                        const double2 h = make_double2(acc[tt][0], acc[tt][1]);
                        if (MODE == 0) {
                            *reinterpret_cast<double2 *>(c.HU() + idx) = h;
                        } else {
                            *reinterpret_cast<double2 *>(c.HDU() + idx) = h;
                            double2 d;
                            d.x = -(rc[tt].x - h.x) * wc[tt].x;
                            d.y = -(rc[tt].y - h.y) * wc[tt].y;
                            *reinterpret_cast<double2 *>(c.DU() + idx) = d;
                        }
                    }
                }
            }
            if (MODE == 1) cp_async_wait<0>();
        }
        {   // ---- x part (global loads up front; the epilogue operands two k-steps before the end) ----
            double a1[KS][TTMAX], a2[KS][TTMAX];
#pragma unroll
            for (int tt = 0; tt < TTMAX; ++tt) {
                const int t = 8 * (tt0 + tt) + gq;
                const bool ok1 = (tt < TT) && (t + 1 < T), ok2 = (tt < TT) && (t + 2 < T);
                const double *p1 = V + (size_t)(ok1 ? t + 1 : 0) * npad + q, *p2 = V + (size_t)(ok2 ? t + 2 : 0) * npad + q;
#pragma unroll
                for (int kk = 0; kk < KS; ++kk) {
                    const double v1 = p1[4 * kk], v2 = p2[4 * kk];
                    a1[kk][tt] = ok1 ? v1 : 0.0;
                    a2[kk][tt] = ok2 ? v2 : 0.0;
                }
            }
            double2 vown[TTMAX][CT], vterm[CT], rdx[TTMAX][CT];
            double acc[TTMAX][CT][2];
#pragma unroll
            for (int tt = 0; tt < TTMAX; ++tt)
#pragma unroll
                for (int nt = 0; nt < CT; ++nt) acc[tt][nt][0] = acc[tt][nt][1] = 0.0;
#pragma unroll
            for (int kk = 0; kk < KS; ++kk) {
                if (kk == (KS >= 2 ? KS - 2 : 0)) {
#pragma unroll
                    for (int nt = 0; nt < CT; ++nt)
                        vterm[nt] = c.has_xf ? *reinterpret_cast<const double2 *>(V + (size_t)T * npad + 8 * nt + 2 * q) : make_double2(0.0, 0.0);
#pragma unroll
                    for (int tt = 0; tt < TTMAX; ++tt) {
                        const int t = 8 * (tt0 + tt) + gq;
#pragma unroll
                        for (int nt = 0; nt < CT; ++nt) {
                            vown[tt][nt] = rdx[tt][nt] = make_double2(0.0, 0.0);
                            if (tok[tt]) {
                                vown[tt][nt] = *reinterpret_cast<const double2 *>(V + (size_t)t * npad + 8 * nt + 2 * q);
                                if (MODE == 1) rdx[tt][nt] = *reinterpret_cast<const double2 *>(c.RDX() + (size_t)t * npad + 8 * nt + 2 * q);
                            }
                        }
                    }
                }
                double b1[CT], b2[CT];
#pragma unroll
                for (int nt = 0; nt < CT; ++nt) { b1[nt] = pA1[4 * kk * LD + 8 * nt]; b2[nt] = pA2[4 * kk * LD + 8 * nt]; }
#pragma unroll
                for (int tt = 0; tt < TTMAX; ++tt)
                    if (tt < TT) {
#pragma unroll
                        for (int nt = 0; nt < CT; ++nt) {
                            dmma(acc[tt][nt], a1[kk][tt], b1[nt]);
                            dmma(acc[tt][nt], a2[kk][tt], b2[nt]);
                        }
                    }
            }
#pragma unroll
            for (int tt = 0; tt < TTMAX; ++tt) {
                const int t = 8 * (tt0 + tt) + gq;
                if (tok[tt]) {
                    const int st = (t == T - 1) ? npad : 0;
#pragma unroll
                    for (int nt = 0; nt < CT; ++nt) {
                        const int k0 = 8 * nt + 2 * q;
                        const bool ok0 = (nt < CT - 1) || kok0, ok1 = (nt < CT - 1) || kok1;
                        double h0 = vown[tt][nt].x - acc[tt][nt][0], h1 = vown[tt][nt].y - acc[tt][nt][1];
                        if (t == T - 1) { h0 += vterm[nt].x; h1 += vterm[nt].y; }
                        h0 = ok0 ? h0 : 0.0;                                   // columns >= n stay exact zeros
                        h1 = ok1 ? h1 : 0.0;
                        const size_t idx = (size_t)t * npad + k0;
                        if (MODE == 0) {
                            *reinterpret_cast<double2 *>(c.HX() + idx) = make_double2(h0, h1);
                        } else {
                            *reinterpret_cast<double2 *>(c.HDX() + idx) = make_double2(h0, h1);
                            double2 d;
                            d.x = -(rdx[tt][nt].x + h0) * c.sQi()[st + k0];
                            d.y = -(rdx[tt][nt].y + h1) * c.sQi()[st + k0 + 1];
                            *reinterpret_cast<double2 *>(c.DX() + idx) = d;
                        }
                    }
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// 8 x 8 diagonal block in accumulator layout (thread (gq,q) holds D[gq][2q], D[gq][2q+1]): right-looking Cholesky
// with the inverse of the factor formed alongside (forward elimination applied to the identity).  Only the first
// nv columns are pivots; rows below them (e.g. the rhs row) are carried as ordinary rows.
// On exit d = L (lower triangle; the rest is junk), x = inv(L) (lower triangular).  Returns 0 or failing column + 1.
// ---------------------------------------------------------------------------------------------
template <int NVMAX, bool LAST>
__device__ __forceinline__ int diag8(double (&d)[2], double (&x)[2], const int nv, const int gq, const int q)
{
    x[0] = (2 * q == gq) ? 1.0 : 0.0;
    x[1] = (2 * q + 1 == gq) ? 1.0 : 0.0;
    int info = 0;
    // Straight-line code (no warp-divergent control flow around the shuffles): a column >= nv is an identity
    // column (unit pivot, zero column), for which every update below is an exact no-op.
#pragma unroll
    for (int cidx = 0; cidx < NVMAX; ++cidx) {
        const bool live = LAST ? (cidx < nv) : true;          // only the last diagonal tile can have fewer than 8 pivots
        const int qc = cidx >> 1, ec = cidx & 1;
        double p = shfl_d(d[ec], 4 * cidx + qc);                     // pivot D[c][c]
        p = live ? p : 1.0;
        if (!(p > 0.0) || !(p < 1.0e300)) { if (!info) info = cidx + 1; }
        const double rs = rsqrt_nr(p);
        const double mine = live ? d[ec] * rs : 0.0;                 // column c scaled (held where q == qc)
        const double lr = shfl_d(mine, 4 * gq + qc);                 // L[gq][c]
        const double lc0 = shfl_d(mine, 4 * (2 * q) + qc);           // L[2q][c]
        const double lc1 = shfl_d(mine, 4 * (2 * q + 1) + qc);       // L[2q+1][c]
        d[ec] = (live && q == qc) ? lr : d[ec];
        d[0] = (2 * q > cidx) ? fma(-lr, lc0, d[0]) : d[0];
        d[1] = (2 * q + 1 > cidx) ? fma(-lr, lc1, d[1]) : d[1];
        const double e0 = shfl_d(x[0], 4 * cidx + q) * rs, e1 = shfl_d(x[1], 4 * cidx + q) * rs;
        x[0] = (gq == cidx) ? e0 : ((gq > cidx) ? fma(-lr, e0, x[0]) : x[0]);
        x[1] = (gq == cidx) ? e1 : ((gq > cidx) ? fma(-lr, e1, x[1]) : x[1]);
    }
    return info;
}
// transpose of an 8 x 8 tile in accumulator layout
__device__ __forceinline__ void tile_T(double (&o)[2], const double (&t)[2], const int gq, const int q)
{
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        const int src = 4 * (2 * q + e) + (gq >> 1);
        const double v0 = shfl_d(t[0], src), v1 = shfl_d(t[1], src);
        o[e] = (gq & 1) ? v1 : v0;
    }
}

// ---------------------------------------------------------------------------------------------
// B diag(w) B' into the lower tiles of s, software-pipelined operand loads.
//   MODE 0 : every lower tile, one w row.
//   MODE 1 : NPOT = 28 (n = 25..28): the last tile row has at most 4 real rows (24..27); its fragment rows 4..7 would multiply
//            padding.  They take the rows 24..27 of the NEXT stage instead (A operand = B[24 + (gq & 3)] scaled by w_i for
//            gq < 4 and by w_{i+1} for gq >= 4), so the last tile row of two consecutive stages costs one set of DMMAs.
//            The column operand of tile (CT-1, CT-1) is B[24 + (gq & 3)] too: columns 28..31 only ever hold padding.
//   MODE 2 : the stage whose last tile row was computed by its predecessor: tile rows 0 .. CT-2 only.
// ---------------------------------------------------------------------------------------------
template <int RT, int CT, int MODE>
__device__ __forceinline__ void bwb_loop(double (&s)[RT][CT][2], const double *const (&pB)[CT], const double *pBl, const double *wb,
                                         const double *wl, const int MK)
{
    constexpr int NR = (MODE == 2) ? CT - 1 : CT;          // tile rows computed
    double fr[CT], af[CT];
#pragma unroll
    for (int rt = 0; rt < NR; ++rt) fr[rt] = (MODE == 1 && rt == CT - 1) ? pBl[0] : pB[rt][0];
    {
        const double wv = wb[0], wlv = (MODE == 1) ? wl[0] : 0.0;
#pragma unroll
        for (int rt = 0; rt < NR; ++rt) af[rt] = fr[rt] * ((MODE == 1 && rt == CT - 1) ? wlv : wv);
    }
#pragma unroll 4
    for (int kk = 0; kk < MK; ++kk) {
        // operands of the next step (the last step re-reads its own: harmless)
        const int kn = (kk + 1 < MK) ? kk + 1 : kk;
        const double wn = wb[4 * kn], wln = (MODE == 1) ? wl[4 * kn] : 0.0;
        double frn[CT], afn[CT];
#pragma unroll
        for (int rt = 0; rt < NR; ++rt) frn[rt] = (MODE == 1 && rt == CT - 1) ? pBl[4 * kn] : pB[rt][4 * kn];
#pragma unroll
        for (int rt = 0; rt < NR; ++rt) afn[rt] = frn[rt] * ((MODE == 1 && rt == CT - 1) ? wln : wn);
#pragma unroll
        for (int rt = 0; rt < NR; ++rt) {
#pragma unroll
            for (int ct = 0; ct <= rt; ++ct) dmma(s[rt][ct], af[rt], fr[ct]);
        }
#pragma unroll
        for (int rt = 0; rt < NR; ++rt) { fr[rt] = frn[rt]; af[rt] = afn[rt]; }
    }
}

// ---------------------------------------------------------------------------------------------
// Band-2 block Cholesky of Y fused with the forward substitution (inf_newton_solver.m:27-31).
// On return YV holds y = inv(L) (-beta).  Returns 0 or the failing stage + 1.
// ---------------------------------------------------------------------------------------------
template <int NPOT, int RT>
__device__ __forceinline__ int forward_sweep(const WCtx &c PROF_PARAMS)
{
    constexpr int CT = (NPOT + 7) / 8, KS = NPOT / 4, LD = (NPOT % 8 == 4) ? NPOT : NPOT + 4, npad = 8 * CT;
    const int n = c.n, T = c.T, NB = c.NB, mpad = c.mpad, gq = c.gq, q = c.q, lane = c.lane;
    const bool yrow = (gq == (n & 7));                       // row n lives in tile RT-1, fragment row n % 8
    const int nn = n * n;
    double *bL1 = c.blk(0), *bL2p = c.blk(1), *bL2pp = c.blk(2);
    const bool cok0 = (8 * (CT - 1) + 2 * q) < NPOT, cok1 = (8 * (CT - 1) + 2 * q + 1) < NPOT;     // block columns < NPOT
    const bool nok0 = (8 * (CT - 1) + 2 * q) < n, nok1 = (8 * (CT - 1) + 2 * q + 1) < n;          // real columns < n

    // stage the first two w rows: stage i reads slot i & 1 (a paired stage also the other one) and refills it with row i + 2
    constexpr bool PAIR = (NPOT % 8 == 4) && (RT == CT);      // NPOT = 28: the last tile row is at most half full
    for (int ch = lane; ch < mpad / 2; ch += 32) cp_async16(c.wbuf() + 2 * ch, c.WV() + 2 * ch);
    if (T > 1) for (int ch = lane; ch < mpad / 2; ch += 32) cp_async16(c.wbuf() + mpad + 2 * ch, c.WV() + mpad + 2 * ch);
    cp_async_commit();
    double2 *stash = reinterpret_cast<double2 *>(c.wbuf() + 2 * mpad);     // [CT][16 lanes]: rows 24..27 of the next stage

    for (int i = 0; i < NB; ++i) {
        const bool has1 = (i + 1 < NB), has2 = (i + 2 < NB) && c.a2;
        const bool up1 = (i >= 1), up2 = (i >= 2) && c.a2;
        // ================= phase A: S_i (lower tiles) + rhs row =================
        double sacc[RT][CT][2];
#pragma unroll
        for (int rt = 0; rt < RT; ++rt)
#pragma unroll
            for (int ct = 0; ct < CT; ++ct) sacc[rt][ct][0] = sacc[rt][ct][1] = 0.0;
        // iterate-independent part of Y[i,i] and -beta_i: requested now, consumed after the B diag(w) B' loop
        double ydv[CT][CT][2];
        double2 yrv[CT];
        {
            const int yd = c.ydi[i];
            const double *Yd = c.ypool + (size_t)(yd >= 0 ? yd : 0) * nn;
#pragma unroll
            for (int rt = 0; rt < CT; ++rt) {
                const int r = 8 * rt + gq;
#pragma unroll
                for (int ct = 0; ct <= rt; ++ct) {
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int cc = 8 * ct + 2 * q + e;
                        ydv[rt][ct][e] = (yd >= 0 && r < n && cc <= r) ? __ldg(Yd + r * n + cc) : 0.0;
                    }
                }
            }
#pragma unroll
            for (int ct = 0; ct < CT; ++ct)
                yrv[ct] = yrow ? *reinterpret_cast<const double2 *>(c.YV() + (size_t)i * npad + 8 * ct + 2 * q) : make_double2(0.0, 0.0);
        }
        if (i < T) {
            cp_async_wait<0>();
            __syncwarp();
            const double *wb = c.wbuf() + (i & 1) * mpad + q;
            // B diag(w_i) B'  (rows >= n of B read finite junk that only reaches unused rows / columns of S)
            const double *pB[CT];
#pragma unroll
            for (int rt = 0; rt < CT; ++rt) pB[rt] = c.sB() + (size_t)(8 * rt + gq) * c.LDB + q;
            const bool pair_now = PAIR && !(i & 1) && (i + 1 < T), pair_done = PAIR && (i & 1);
            if (pair_now) {
                const double *pBl = c.sB() + (size_t)(8 * (CT - 1) + (gq & 3)) * c.LDB + q;
                const double *wl = c.wbuf() + (((gq < 4) ? i : i + 1) & 1) * mpad + q;
                bwb_loop<RT, CT, 1>(sacc, pB, pBl, wb, wl, c.MK);
                if (gq >= 4) {                                   // rows 24..27 of stage i + 1: kept for the next stage
#pragma unroll
                    for (int ct = 0; ct < CT; ++ct) stash[16 * ct + 4 * (gq - 4) + q] = make_double2(sacc[CT - 1][ct][0], sacc[CT - 1][ct][1]);
                }
            } else if (pair_done) {
                bwb_loop<RT, CT, 2>(sacc, pB, nullptr, wb, nullptr, c.MK);
#pragma unroll
                for (int ct = 0; ct < CT; ++ct) {
                    const double2 v = stash[16 * ct + 4 * (gq & 3) + q];
                    sacc[CT - 1][ct][0] = (gq < 4) ? v.x : 0.0;
                    sacc[CT - 1][ct][1] = (gq < 4) ? v.y : 0.0;
                }
            } else {
                bwb_loop<RT, CT, 0>(sacc, pB, nullptr, wb, nullptr, c.MK);
            }
            // slot i & 1 is free again: fetch row i + 2 underneath the rest of this stage
            __syncwarp();
            if (i + 2 < T) {
                const double *src = c.WV() + (size_t)(i + 2) * mpad;
                double *dst = c.wbuf() + (i & 1) * mpad;
                for (int ch = lane; ch < mpad / 2; ch += 32) cp_async16(dst + 2 * ch, src + 2 * ch);
            }
            cp_async_commit();
        }
        {   // + iterate-independent part of Y[i,i]; rhs row <- -beta_i
#pragma unroll
            for (int rt = 0; rt < CT; ++rt)
#pragma unroll
                for (int ct = 0; ct <= rt; ++ct) { sacc[rt][ct][0] += ydv[rt][ct][0]; sacc[rt][ct][1] += ydv[rt][ct][1]; }
            if (yrow) {
#pragma unroll
                for (int ct = 0; ct < CT; ++ct) { sacc[RT - 1][ct][0] = yrv[ct].x; sacc[RT - 1][ct][1] = yrv[ct].y; }
            }
        }
        if (up1) {
            const double *p = bL1 + gq * LD + q;
#pragma unroll
            for (int kk = 0; kk < KS; ++kk) {
                double fr[RT];
#pragma unroll
                for (int rt = 0; rt < RT; ++rt) fr[rt] = p[8 * rt * LD + 4 * kk];
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
            const double *p = bL2pp + gq * LD + q;
#pragma unroll
            for (int kk = 0; kk < KS; ++kk) {
                double fr[RT];
#pragma unroll
                for (int rt = 0; rt < RT; ++rt) fr[rt] = p[8 * rt * LD + 4 * kk];
#pragma unroll
                for (int rt = 0; rt < RT; ++rt) {
                    const double a = dneg(fr[rt]);
#pragma unroll
                    for (int ct = 0; ct < CT; ++ct)
                        if (ct <= rt) dmma(sacc[rt][ct], a, fr[ct]);
                }
            }
        }
        PROF_T(8);
        // ================= phase B: L_i and inv(L_i) on the accumulator tiles =================
        double linv[CT][CT][2];                              // inv(L_i) tiles (rt, ct <= rt)
        int info = 0;
#pragma unroll
        for (int kb = 0; kb < CT; ++kb) {
            const int nv = min(8, n - 8 * kb);
            if (nv < 8) {                                    // columns >= n of the last diagonal tile: identity
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int cg = 2 * q + e;
                    if (cg >= nv) sacc[kb][kb][e] = (gq == cg) ? 1.0 : 0.0;
                }
            }
            // columns of this tile that can be pivots: 8, except NPOT - 8 (CT - 1) for the last tile
            const int inf = (kb < CT - 1) ? diag8<8, false>(sacc[kb][kb], linv[kb][kb], nv, gq, q)
                                          : diag8<NPOT - 8 * (CT - 1), true>(sacc[kb][kb], linv[kb][kb], nv, gq, q);
            if (inf && !info) info = 8 * kb + inf;
            if (nv < 8 && gq >= nv) { linv[kb][kb][0] = 0.0; linv[kb][kb][1] = 0.0; }
#pragma unroll
            for (int rt = kb + 1; rt < RT; ++rt) {           // panel: L[rt][kb] = S[rt][kb] inv(L_kk)'
                double o[2] = {0.0, 0.0};
                mma_xt(o, sacc[rt][kb], linv[kb][kb]);
                sacc[rt][kb][0] = o[0]; sacc[rt][kb][1] = o[1];
            }
#pragma unroll
            for (int rt = kb + 1; rt < RT; ++rt) {           // trailing update
                const double nx[2] = {dneg(sacc[rt][kb][0]), dneg(sacc[rt][kb][1])};
#pragma unroll
                for (int ct = kb + 1; ct < CT; ++ct)
                    if (ct <= rt) mma_xt(sacc[rt][ct], nx, sacc[ct][kb]);
            }
        }
        if (info) return i + 1;
        // the iterate-independent part of Y[i+1,i]: requested here (the diagonal blocks are done, registers are free again),
        // consumed in phase C after the triangular inverse
        double macc[CT][CT][2];
        {
            const int y1 = c.y1i[i];
            const double *Y1 = c.ypool + (size_t)(y1 >= 0 ? y1 : 0) * nn;
            const bool ld1 = has1 && (y1 >= 0);
#pragma unroll
            for (int rt = 0; rt < CT; ++rt) {
                const int r = 8 * rt + gq;
#pragma unroll
                for (int ct = 0; ct < CT; ++ct) {
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int cc = 8 * ct + 2 * q + e;
                        macc[rt][ct][e] = (ld1 && r < n && cc < n) ? __ldg(Y1 + r * n + cc) : 0.0;
                    }
                }
            }
        }
        {   // inv(L) off-diagonal tiles, row by row from the diagonal leftwards:
            //   inv(L)[rt][ct] = -(sum_{j=ct+1..rt} inv(L)[rt][j] L[j][ct]) inv(L_ct,ct)
            double lt[CT][CT][2], xdt[CT][2];
#pragma unroll
            for (int j = 1; j < CT; ++j)
#pragma unroll
                for (int ct = 0; ct < j; ++ct) tile_T(lt[j][ct], sacc[j][ct], gq, q);
#pragma unroll
            for (int ct = 0; ct + 1 < CT; ++ct) tile_T(xdt[ct], linv[ct][ct], gq, q);
#pragma unroll
            for (int rt = 1; rt < CT; ++rt) {
#pragma unroll
                for (int ct = rt - 1; ct >= 0; --ct) {
                    double z[2] = {0.0, 0.0};
#pragma unroll
                    for (int j = ct + 1; j <= rt; ++j) mma_xt(z, linv[rt][j], lt[j][ct]);
                    const double nz[2] = {dneg(z[0]), dneg(z[1])};
                    linv[rt][ct][0] = linv[rt][ct][1] = 0.0;
                    mma_xt(linv[rt][ct], nz, xdt[ct]);
                }
            }
        }
        // y_i: row n of the L tiles (columns >= n are not part of y)
        double yv[CT][2];
#pragma unroll
        for (int ct = 0; ct < CT; ++ct) {
            yv[ct][0] = ((ct < CT - 1) || nok0) ? sacc[RT - 1][ct][0] : 0.0;
            yv[ct][1] = ((ct < CT - 1) || nok1) ? sacc[RT - 1][ct][1] : 0.0;
        }
        // inv(L_i) -> global scratch (row-major, leading dimension NPOT; tiles above the diagonal are never written)
        {
            double *g = c.gLi() + (size_t)i * (NPOT * NPOT) + gq * NPOT + 2 * q;
#pragma unroll
            for (int rt = 0; rt < CT; ++rt)
#pragma unroll
                for (int ct = 0; ct <= rt; ++ct) {
                    const bool rok = (rt < CT - 1) || (8 * rt + gq < NPOT);
                    const bool ok = rok && ((ct < CT - 1) || cok0);
                    if (ok) *reinterpret_cast<double2 *>(g + 8 * rt * NPOT + 8 * ct) = make_double2(linv[rt][ct][0], linv[rt][ct][1]);
                }
            if (yrow) {
#pragma unroll
                for (int ct = 0; ct < CT; ++ct) *reinterpret_cast<double2 *>(c.YV() + (size_t)i * npad + 8 * ct + 2 * q) = make_double2(yv[ct][0], yv[ct][1]);
            }
        }
        PROF_T(9);
        // ================= phase C: L1_i = (Y1 - L2_{i-1} L1_{i-1}') inv(L_i)' ,  L2_i = Y2 inv(L_i)' =================
        if (has1) {
            if (up1 && c.a2) {
                const double *pa = bL2p + gq * LD + q, *pb = bL1 + gq * LD + q;
#pragma unroll
                for (int kk = 0; kk < KS; ++kk) {
                    double fa[CT], fb[CT];
#pragma unroll
                    for (int rt = 0; rt < CT; ++rt) { fa[rt] = dneg(pa[8 * rt * LD + 4 * kk]); fb[rt] = pb[8 * rt * LD + 4 * kk]; }
#pragma unroll
                    for (int rt = 0; rt < CT; ++rt)
#pragma unroll
                        for (int ct = 0; ct < CT; ++ct) dmma(macc[rt][ct], fa[rt], fb[ct]);
                }
            }
            __syncwarp();                                    // every read of L1_{i-1} is done: its block receives L1_i
            double *g = c.gL1() + (size_t)i * (NPOT * NPOT) + gq * NPOT + 2 * q;
            double *s = bL1 + gq * LD + 2 * q;
#pragma unroll
            for (int rt = 0; rt < CT; ++rt) {
                const bool rok = (rt < CT - 1) || (8 * rt + gq < n);
#pragma unroll
                for (int ct = 0; ct < CT; ++ct) {
                    double o[2] = {0.0, 0.0};
#pragma unroll
                    for (int jt = 0; jt <= ct; ++jt) mma_xt(o, macc[rt][jt], linv[ct][jt]);
                    const bool c0 = (ct < CT - 1) || cok0, c1 = (ct < CT - 1) || cok1;
                    if (rok && c0) s[8 * rt * LD + 8 * ct] = o[0];
                    if (rok && c1) s[8 * rt * LD + 8 * ct + 1] = o[1];
                    if (rok && c0) *reinterpret_cast<double2 *>(g + 8 * rt * NPOT + 8 * ct) = make_double2(o[0], o[1]);
                }
            }
        } else {
            __syncwarp();
        }
        if (yrow) {                                          // row n of the L1 block carries y_i
            double *s = bL1 + n * LD + 2 * q;
#pragma unroll
            for (int ct = 0; ct < CT; ++ct) {
                if ((ct < CT - 1) || cok0) s[8 * ct] = yv[ct][0];
                if ((ct < CT - 1) || cok1) s[8 * ct + 1] = yv[ct][1];
            }
        }
        if (has2) {
            const bool y2ok = (c.y2i[i] >= 0);
            double *s = bL2pp + gq * LD + 2 * q;             // L2_i stays on chip: the backward sweep rebuilds L2_i' v from inv(L_i) and Y2
            double nqi[CT][2];                                // -inv(2Q)(k) for k = 8 jt + 2 q + e  (0 for k >= n)
#pragma unroll
            for (int jt = 0; jt < CT; ++jt) {
                nqi[jt][0] = y2ok ? -c.sQi()[8 * jt + 2 * q] : 0.0;
                nqi[jt][1] = y2ok ? -c.sQi()[8 * jt + 2 * q + 1] : 0.0;
            }
            const double *pa2 = c.sA2() + gq * LD + 2 * q;
#pragma unroll
            for (int rt = 0; rt < CT; ++rt) {
                const bool rok = (rt < CT - 1) || (8 * rt + gq < n);
                double af[CT][2];                             // Y2 = -A2 inv(2Q) : rows of A2, k in pair order
#pragma unroll
                for (int jt = 0; jt < CT; ++jt) {
                    af[jt][0] = pa2[8 * rt * LD + 8 * jt] * nqi[jt][0];
                    af[jt][1] = pa2[8 * rt * LD + 8 * jt + 1] * nqi[jt][1];
                }
#pragma unroll
                for (int ct = 0; ct < CT; ++ct) {
                    double o[2] = {0.0, 0.0};
#pragma unroll
                    for (int jt = 0; jt <= ct; ++jt) mma_xt(o, af[jt], linv[ct][jt]);
                    const bool c0 = (ct < CT - 1) || cok0, c1 = (ct < CT - 1) || cok1;
                    if (rok && c0) s[8 * rt * LD + 8 * ct] = o[0];
                    if (rok && c1) s[8 * rt * LD + 8 * ct + 1] = o[1];
                }
            }
            if (yrow) {                                       // row n of the L2 block carries y_i as well
                double *sy = bL2pp + n * LD + 2 * q;
#pragma unroll
                for (int ct = 0; ct < CT; ++ct) {
                    if ((ct < CT - 1) || cok0) sy[8 * ct] = yv[ct][0];
                    if ((ct < CT - 1) || cok1) sy[8 * ct + 1] = yv[ct][1];
                }
            }
        }
        { double *t2 = bL2p; bL2p = bL2pp; bL2pp = t2; }
        __syncwarp();
        PROF_T(10);
    }
    cp_async_wait<0>();
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Backward substitution  dnu_i = inv(L_i)' (y_i - L1_i' dnu_{i+1} - L2_i' dnu_{i+2})  (inf_newton_solver.m:32).
// The factor comes back from the global scratch through a 3-slot cp.async ring (3 entries per stage);
// lane k owns component k: its column of each matrix is read into registers, then 4 independent FMA chains.
// Rows / columns >= n of the stored matrices are zero (never written), so nothing is predicated.
// ---------------------------------------------------------------------------------------------
template <int NPOT>
__device__ __forceinline__ void ring_issue(const WCtx &c, const int e, const int nent)
{
    if (e < nent) {
        const int i = c.NB - 1 - e / 2, kind = e % 2;        // per stage: L1_i (kind 0), inv(L_i) (kind 1)
        const bool on = (kind == 1) || (i + 1 < c.NB);
        if (on) {
            const double *src = (kind == 0 ? c.gL1() : c.gLi()) + (size_t)i * (NPOT * NPOT);
            double *dst = c.wsm + (e % 3) * (NPOT * NPOT);
            for (int ch = c.lane; ch < NPOT * NPOT / 2; ch += 32) cp_async16(dst + 2 * ch, src + 2 * ch);
        }
    }
    cp_async_commit();
}

// L2_i = Y2 inv(L_i)' is not read back (nor stored: a third of the factor scratch traffic): L2_i' v = inv(L_i) (Y2' v) with
// Y2 = -A2 inv(2Q) from the constants in shared memory and inv(L_i), which the stage needs anyway.
template <int NPOT>
__device__ __forceinline__ void backward_sweep(const WCtx &c)
{
    constexpr int npad = 8 * ((NPOT + 7) / 8), LD = (NPOT % 8 == 4) ? NPOT : NPOT + 4;
    const int n = c.n, NB = c.NB, lane = c.lane;
    double *vec = c.wsm + 3 * (NPOT * NPOT);                 // [3][32] dnu ring + [32] tmp + [32] Y2' dnu (behind the 3 ring slots)
    const int nent = 2 * NB;
    const bool act = lane < n;
    __syncwarp();
    vec[lane] = 0.0; vec[32 + lane] = 0.0; vec[64 + lane] = 0.0; vec[96 + lane] = 0.0; vec[128 + lane] = 0.0;
    for (int e = 0; e < 3; ++e) ring_issue<NPOT>(c, e, nent);
    const double nqi = act ? -c.sQi()[lane] : 0.0;           // Y2[r][k] = -A2[r][k] inv(2Q)[k]
    const double *a2c = c.sA2() + (lane < NPOT ? lane : 0);  // column `lane` of A2
    double acc = 0.0, yi = 0.0;
    for (int e = 0; e < nent; ++e) {
        const int i = NB - 1 - e / 2, kind = e % 2;
        if (kind == 0) yi = act ? c.YV()[(size_t)i * npad + lane] : 0.0;       // consumed one entry later
        cp_async_wait<2>();
        __syncwarp();
        const double *M = c.wsm + (e % 3) * (NPOT * NPOT) + (lane < NPOT ? lane : 0);
        if (kind == 0) {
            acc = 0.0;
            if (i + 1 < NB) {                                 // L1_i' dnu_{i+1}
                const double *dv = vec + ((i + 1) % 3) * 32;
                double mv[NPOT];
#pragma unroll
                for (int r = 0; r < NPOT; ++r) mv[r] = M[r * NPOT];
                double s[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
                for (int r = 0; r < NPOT; ++r) s[r & 3] = fma(mv[r], dv[r], s[r & 3]);
                acc += (s[0] + s[1]) + (s[2] + s[3]);
            }
            if (i + 2 < NB && c.a2 && c.y2i[i] >= 0) {        // w = Y2' dnu_{i+2} (shared-memory constants), consumed with inv(L_i)
                const double *dv = vec + ((i + 2) % 3) * 32;
                double s[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
                for (int r = 0; r < NPOT; ++r) s[r & 3] = fma(a2c[r * LD], dv[r], s[r & 3]);
                vec[128 + lane] = act ? nqi * ((s[0] + s[1]) + (s[2] + s[3])) : 0.0;
            } else {
                vec[128 + lane] = 0.0;
            }
        } else {
            double *tmp = vec + 96;
            const double *w2 = vec + 128;
            const double *Mr = c.wsm + (e % 3) * (NPOT * NPOT) + (lane < NPOT ? lane : 0) * NPOT;      // row `lane` of inv(L_i)
            double s2[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
            for (int r = 0; r < NPOT; ++r) s2[r & 3] = fma(Mr[r], w2[r], s2[r & 3]);                    // (inv(L_i) w)[lane]
            acc += (s2[0] + s2[1]) + (s2[2] + s2[3]);
            double mv[NPOT];
#pragma unroll
            for (int r = 0; r < NPOT; ++r) mv[r] = M[r * NPOT];
            if (act) tmp[lane] = yi - acc;
            __syncwarp();
            double s[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
            for (int r = 0; r < NPOT; ++r) s[r & 3] = fma(mv[r], tmp[r], s[r & 3]);
            const double d = (s[0] + s[1]) + (s[2] + s[3]);
            if (act) {
                vec[(i % 3) * 32 + lane] = d;
                c.DNU()[(size_t)i * npad + lane] = d;
            }
        }
        __syncwarp();                                        // slot consumed by every lane
        ring_issue<NPOT>(c, e + 3, nent);
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
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nwarps = blockDim.x >> 5;
    constexpr int LD = (NPOT % 8 == 4) ? NPOT : NPOT + 4, npad = 8 * ((NPOT + 7) / 8);

    // ---- shared memory: zero everything (overrun reads must see finite values), then the constants ----
    {
        const size_t tot = G.smem_doubles(nwarps);
        for (size_t e = tid; e < tot; e += blockDim.x) smem[e] = 0.0;
    }
    __syncthreads();
    double *sB = smem;
    double *sA1 = sB + (size_t)n * G.LDB;
    double *sA2 = sA1 + (size_t)n * LD;
    double *sUmax = sA2 + (size_t)n * LD;
    double *sUmin = sUmax + G.mpad, *sR2 = sUmin + G.mpad, *sRl = sR2 + G.mpad;
    double *sQ2 = sRl + G.mpad, *sQl = sQ2 + 2 * npad, *sQi = sQl + 2 * npad;
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
    for (int k = tid; k < npad; k += blockDim.x) {
        const bool ok = k < n;
        sQ2[k] = ok ? S.q2[k] : 0.0;  sQ2[npad + k] = ok ? S.q2f[k] : 0.0;
        sQl[k] = ok ? S.ql[k] : 0.0;  sQl[npad + k] = ok ? S.qfl[k] : 0.0;
        sQi[k] = ok ? S.qi[k] : 0.0;  sQi[npad + k] = ok ? S.qif[k] : 0.0;
    }
    __syncthreads();

    WCtx c;
    c.n = n; c.m = m; c.T = T; c.NB = T + (A.has_xf ? 1 : 0); c.a2 = S.has_a2; c.has_xf = A.has_xf;
    c.mpad = G.mpad; c.MK = G.MK; c.MT8 = G.MT8; c.npad = npad; c.LDB = G.LDB; c.NTT = G.NTT;
    c.lane = lane; c.gq = lane >> 2; c.q = lane & 3;
    c.kappa = A.kappa;
    c.smem = smem; c.oA1 = (int)(sA1 - smem); c.oA2 = (int)(sA2 - smem); c.oU = (int)(sUmax - smem); c.oQ = (int)(sQ2 - smem);
    c.wsm = smem + G.const_doubles + (size_t)wid * G.warp_doubles; c.BLK = G.BLK;
    // scratch slot: slot_base + blockIdx.x, or the SM id when slot_base < 0 (branch-free: the kernel sits at the register
    // limit and an extra branch in the prologue measurably perturbs the allocation of the hot loops)
    unsigned smid;
    asm("mov.u32 %0, %%smid;" : "=r"(smid));
    const unsigned bysm = (unsigned)(A.slot_base >> 31);            // all ones iff slot_base < 0
    const unsigned slot = (smid & bysm) | ((unsigned)(A.slot_base + (int)blockIdx.x) & ~bysm);
    double *ws = A.ws + ((size_t)slot * nwarps + wid) * A.ws_stride;
    c.ws = ws; c.tu = G.TP8 * G.mpad; c.tx = G.TP8 * npad; c.tb = (G.TP8 + 1) * npad; c.bl = (T + 1) * G.NN; c.pp = 0;
    c.ypool = S.ypool; c.ydi = S.ydi; c.y1i = S.y1i; c.y2i = S.y2i;
    c.xmin = S.xmin; c.xmax = S.xmax;
    c.sh = A.warm_shift ? 1 : 0;
    const int NB = c.NB, mpad = G.mpad;

    for (;;) {
        int b = 0;
        if (lane == 0) b = (int)atomicAdd(A.counter, 1u);
        b = __shfl_sync(FULL, b, 0);
        if (b >= A.nbatch) break;
        PROF_DECL
        c.pp = 0;
        c.U0 = A.cold ? nullptr : A.U0 + (size_t)b * m * T;
        c.X0 = A.cold ? nullptr : A.X0 + (size_t)b * n * T;

        // ---- b (fast_mpc_eq_const.m:39,44,47,68): lane k owns row k; x0 / x0_pre broadcast by shuffles;
        //      the dual start nu goes to the (column padded) DNU buffer ----
        {
            const double x0v = (lane < n) ? A.x0[(size_t)b * n + lane] : 0.0;
            const double x0pv = (lane < n && A.x0_pre) ? A.x0_pre[(size_t)b * n + lane] : 0.0;
            double s0 = 0.0, s1 = 0.0, s2 = 0.0;
            const double *r1 = sA1 + (size_t)(lane < n ? lane : 0) * LD, *r2 = sA2 + (size_t)(lane < n ? lane : 0) * LD;
#pragma unroll
            for (int kc = 0; kc < NPOT; ++kc) {
                const double xv = __shfl_sync(FULL, x0v, kc), xpv = __shfl_sync(FULL, x0pv, kc);   // 0 for kc >= n
                const double a1 = r1[kc], a2v = r2[kc];
                s0 = fma(a1, xv, s0);
                s2 = fma(a2v, xpv, s2);
                s1 = fma(a2v, xv, s1);
            }
            s0 += s2;
            const double *nu0 = A.nu0 + (size_t)b * NB * n;
            for (int i0 = 0; i0 < NB; i0 += 4) {
                double wv[4], nv[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int i = i0 + u;
                    wv[u] = nv[u] = 0.0;
                    if (lane < n && i < NB) {
                        if (i < T) wv[u] = A.w ? A.w[(size_t)b * T * n + (size_t)i * n + lane] : 0.0;
                        else wv[u] = A.xf[(size_t)b * n + lane];
                        nv[u] = nu0[(size_t)i * n + lane];
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int i = i0 + u;
                    if (lane < n && i < NB) {
                        double v = wv[u];
                        if (i < T) { if (i == 0) v += s0; else if (i == 1) v += s1; }
                        c.BV()[(size_t)i * npad + lane] = v;
                        c.DNU()[(size_t)i * npad + lane] = nv[u];
                    }
                }
            }
        }
        __syncwarp();
        double ssd, ssp, ssf;
        pass_Cv<NPOT, K_RP>(c, 0.0, ssd, ssp, ssf);                         // z <- start ; r_p = C z - b
        pass_Ct<NPOT, 0>(c);                                           // images of the dual start nu
        __syncwarp();
        PROF_T(0);

        int status = ST_OK, iters = 0;
        int next_exit = 0;                     // decided by the accepted trial pass: 1 = early exit, 2 = non-finite residual
        for (int it = 0; it < A.niters; ++it) {
            if (next_exit) { status = (next_exit == 1) ? ST_EARLY_EXIT : ST_NONFINITE; break; }
            pass_Cv<NPOT, K_NEWTON>(c, 0.0, ssd, ssp, ssf);
            const double tot_p = warp_sum(ssp), tot_d = warp_sum(ssd);
            const double nr0 = sqrt(tot_d + tot_p);
            PROF_T(1);
            // ---- early exit (inf_newton_solver.m:19-22) ----
            if (!isfinite(nr0)) { status = ST_NONFINITE; break; }
            if (nr0 <= A.tol_r && sqrt(tot_p) <= A.tol_p) { status = ST_EARLY_EXIT; break; }
            __syncwarp();
            const int fail = forward_sweep<NPOT, RT>(c PROF_ARGS);
            PROF_T(2);
            if (fail) { status = ST_NOT_PD; break; }
            backward_sweep<NPOT>(c);
            PROF_T(3);
            pass_Ct<NPOT, 1>(c);                                       // dz = inv(Phi)(-r_d - C' dnu)
            __syncwarp();
            PROF_T(4);
            // ---- backtracking on ||[r_p; r_d]||, d frozen (backtracking_inf_newton.m:2-11) ----
            double t = 1.0;
            int nh = 0;
            for (;;) {
                pass_Cv<NPOT, K_TRIAL>(c, t, ssd, ssp, ssf);
                const double tp = warp_sum(ssp), td = warp_sum(ssd);
                const double nrt = sqrt(td + tp);
                __syncwarp();
                {   // residual of the next iteration at this trial point (valid if this trial is the accepted one)
                    const double nrn = sqrt(warp_sum(ssf) + tp);
                    next_exit = !isfinite(nrn) ? 2 : ((nrn <= A.tol_r && sqrt(tp) <= A.tol_p) ? 1 : 0);
                }
                if (!(nrt > (1.0 - A.alpha * t) * nr0)) break;
                if (t == 0.0) break;
                if (A.ls_max > 0 && nh >= A.ls_max) { status = ST_LS_MAX; break; }
                t *= A.beta;
                ++nh;
            }
            PROF_T(5);
            // accept: the trial pass left z + t dz, C(z + t dz) - b and the advanced dual images in the *T buffers
            c.pp ^= 1;
            ++iters;
        }
        __syncwarp();
        {   // iterate -> outputs: rows of the padded scratch arrays to the dense output rows, 4 rows in flight per lane
            double *uo = A.U + (size_t)b * m * T, *xo = A.X + (size_t)b * n * T;
            const double *pu = c.UC(), *px = c.XC();
            for (int j0 = 0; j0 < m; j0 += 32) {
                const int j = j0 + lane;
                for (int t0 = 0; t0 < T; t0 += 4) {
                    double v[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) v[u] = (j < m && t0 + u < T) ? pu[(size_t)(t0 + u) * mpad + j] : 0.0;
#pragma unroll
                    for (int u = 0; u < 4; ++u) if (j < m && t0 + u < T) uo[(size_t)(t0 + u) * m + j] = v[u];
                    if (t0 == 0 && A.u_first && j < m) A.u_first[(size_t)b * m + j] = v[0];
                }
            }
            for (int t0 = 0; t0 < T; t0 += 8) {
                double v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] = (lane < n && t0 + u < T) ? px[(size_t)(t0 + u) * npad + lane] : 0.0;
#pragma unroll
                for (int u = 0; u < 8; ++u) if (lane < n && t0 + u < T) xo[(size_t)(t0 + u) * n + lane] = v[u];
            }
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
    if (G.smem_doubles(1) * 8 > avail) return -3;
    int warps = 8;
    while (warps > 1 && G.smem_doubles(warps) * 8 > avail) --warps;
    // measured on B200 (profiles/r01_v3_warp_sweep.log): filling the last ~16 KB of the unified L1/shared array with an
    // eighth warp costs more (no L1 left for the constant pool and the scratch streams) than the extra instance gains
    if (warps == 8 && G.smem_doubles(warps) * 8 + 16384 > avail) warps = 7;
    if (const char *e = getenv("FMPC_WARPS_PER_CTA")) { const int v = atoi(e); if (v >= 1 && G.smem_doubles(v) * 8 <= avail && v <= 8) warps = v; }
    const size_t smem = G.smem_doubles(warps) * 8;
    if (cudaFuncSetAttribute(fmpc_solve_kernel_warp<NPOT, RT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -4;
    cfg->grid = prop.multiProcessorCount;
    cfg->block = 32 * warps;
    cfg->smem = smem;
    cfg->use_mma = 2;
    cfg->slots = cfg->grid * warps;
    cfg->ws_stride = WsW::make(G).total;
    return 0;
}

#ifdef FMPC_FAST_BUILD          /* compile-time experiments: two instantiations only */
#define WARP_DISPATCH(G, CALL)                                                                         \
    switch ((G).NPOT * 8 + (G).RT) {                                                                   \
    case 8 * 8 + 1: CALL(8, 1); break;                                                                 \
    case 28 * 8 + 4: CALL(28, 4); break;                                                               \
    default: break;                                                                                    \
    }
#else
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
#endif

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

namespace { __global__ void fmpc_nsmid_kernel(unsigned *out) { unsigned v; asm("mov.u32 %0, %%nsmid;" : "=r"(v)); *out = v; } }

// Scratch slots may be indexed by the SM id when SM ids are dense below the slot count and two CTAs can never share an SM
// (each needs more than half of its shared memory).
int fmpc_warp_smid_slots_ok(const SolveLaunchCfg &cfg, int device)
{
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return 0;
    if (2 * (cfg.smem + 1024) <= (size_t)prop.sharedMemPerMultiprocessor) return 0;
    unsigned *d = nullptr, nsmid = 0;
    if (cudaMalloc(&d, 4) != cudaSuccess) return 0;
    fmpc_nsmid_kernel<<<1, 1>>>(d);
    const bool ok = cudaMemcpy(&nsmid, d, 4, cudaMemcpyDeviceToHost) == cudaSuccess;
    cudaFree(d);
    return ok && nsmid > 0 && (int)nsmid <= cfg.grid;
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
