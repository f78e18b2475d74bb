// Internal (non-ABI) declarations shared by the host layer and the CUDA kernels.
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

// Problem-constant device data of one handle.  All n x n "blocks" are ROW-major on the device
// (the host layer transposes MATLAB's column-major input once, at fmpc_create).
struct DevSys {
    int n, m, T;
    int has_a2;
    const double *B;     // n x m, column-major:  B[k + n*j]      (rows k contiguous)
    const double *Bt;    // m x n, column-major:  Bt[j + m*k] = B(k,j)
    const double *A1;    // n x n, column-major:  A1[k + n*k']
    const double *A1t;   // transpose
    const double *A2, *A2t;
    const double *r2;    // m : 2*R_jj
    const double *rl;    // m : linear cost r
    const double *q2;    // n : 2*Q_kk          (stages 1..T-1)
    const double *q2f;   // n : 2*Qf_kk         (stage T)
    const double *qi;    // n : 1/(2 Q_kk)
    const double *qif;   // n : 1/(2 Qf_kk)
    const double *ql;    // n : linear cost q
    const double *qfl;   // n : linear cost qf
    const double *umin, *umax, *xmin, *xmax;
    // iterate-independent Schur blocks (row-major n x n each), de-duplicated pool + index tables (T+1 entries)
    const double *ypool;
    const int *ydi, *y1i, *y2i;   // block index per stage; -1 = zero block
    // DMMA path: pair-product matrix G[p][j] = B(r,j) B(c,j), p = r(r+1)/2 + c (r >= c), row-major Mp x mp,
    // so that  B diag(w_t) B'  for ALL stages is one GEMM  G (Mp x mp) * W (mp x T)
    const double *G;
    int npairs, Mp, mp;
    const double *ypk;   // the pool blocks again, lower triangles packed by pair index: [block][Mp]
};

struct StepArgs {
    int nbatch, has_xf, cold;
    double kappa;
    int niters, ls_max;
    double alpha, beta, tol_r, tol_p;
    const double *x0, *x0_pre, *u_prev, *w, *xf, *X0, *U0, *nu0;
    double *X, *U;
    int *status, *iters;
    unsigned int *counter;              // dynamic instance counter (zeroed before launch)
    unsigned long long *iters_total;    // sum of Newton iterations (zeroed before launch)
    double *ws;                         // per-slot scratch (slot = CTA for the CTA kernels, warp for the warp kernel)
    size_t ws_stride;                   // doubles per slot
    int slot_base;                      // first CTA-slot of this launch; -1: the CTA's scratch slot is its SM id (%smid), so that
                                        // any number of concurrent launches of one handle share the per-SM scratch without
                                        // collisions (at most one CTA of this kernel fits an SM)
    long long *prof;                    // optional phase-cycle accumulators (builds with -DFMPC_PROF), else NULL
    // resident closed loop (fmpc_step_r), honoured by the warp kernel only (the host layer runs separate kernels otherwise):
    int warm_shift;                     // 1: the warm start is (X0, U0) shifted one stage (stage t <- t + 1, last stage repeated)
    double *u_first;                    // m x nbatch | NULL: U(:,0) of every instance, compact
};

// Per-CTA scratch layout (in doubles), computed identically on host and device.
struct WsLayout {
    size_t nu, dnu, yv, rp, rpt, bv;            // (T+1) n each
    size_t hx, hdx, dx, xt, rdx;                // T n each
    size_t hu, hdu, du, ut, dbar, pinv, rdu;    // T m each
    size_t Lf, L1, L2;                          // (T+1) n n each
    size_t Dsc;                                 // T * Mp : packed lower triangles of B diag(w_t) B'
    size_t total;
    __host__ __device__ static WsLayout make(int n, int m, int T)
    {
        WsLayout L;
        size_t nb = (size_t)(T + 1) * n, tn = (size_t)T * n, tm = (size_t)T * m, bl = (size_t)(T + 1) * n * n;
        size_t o = 0;
        L.nu = o; o += nb; L.dnu = o; o += nb; L.yv = o; o += nb; L.rp = o; o += nb; L.rpt = o; o += nb; L.bv = o; o += nb;
        L.hx = o; o += tn; L.hdx = o; o += tn; L.dx = o; o += tn; L.xt = o; o += tn; L.rdx = o; o += tn;
        L.hu = o; o += tm; L.hdu = o; o += tm; L.du = o; o += tm; L.ut = o; o += tm; L.dbar = o; o += tm;
        L.pinv = o; o += tm; L.rdu = o; o += tm;
        L.Lf = o; o += bl; L.L1 = o; o += bl; L.L2 = o; o += bl;
        { size_t np = (size_t)n * (n + 1) / 2, Mp = (np + 7) & ~(size_t)7; L.Dsc = o; o += (size_t)T * Mp; }
        L.total = (o + 15) & ~(size_t)15;
        return L;
    }
};

// Device tables of the general-structure kernel (fmpc_kernel_gen.cu): C as the reference builds it (one dense row
// window per block row), the u-column blocks of every block row, the iterate-independent x part of the Schur
// complement, dense 2Q / inv(2Q), ramp-rate bounds.
struct GenSys {
    int n, m, T, N, ramp, ldyx, panel_rows;
    const double *cw;                       // pool of window matrices, row-major n x cw_len[i]
    const int *cw_ptr, *cw_off, *cw_len;    // per block row (T+1): offset into cw, first column, width
    const double *cu, *cut;                 // pool of n x m blocks C[i, u_t] (row-major) and their transposes (m x n)
    const int *ue_cnt, *ue_t, *ue_ptr;      // per block row: number of u blocks (<= 4), their stages, offsets into cu
    const int *ue_lo, *ue_hi;               // per (block row, u block): the columns [lo, hi) of the block that are not all zero (multiples of 4)
    const double *Yx;                       // ((T+1) n)^2 row-major: C_x inv(Phi_xx) C_x'
    const double *Yxd;                      // same layout: per n x n block, transpose minus the block (what a mirrored Schur tile adds)
    const double *Q2, *Q2f, *Qi, *Qif;      // n x n row-major: 2Q, 2Qf and their inverses
    const double *dumin, *dumax;            // m
    int dense_r;                            // 1: R is a dense SPD matrix (box rows only): Phi_uu of a stage is dense m x m
    const double *R2;                       // m x m row-major: R + R' (= 2R for symmetric R), dense_r only
    int qdiag;                              // 1: Q and Qf are diagonal (the x part of 2 H z and inv(Phi_xx) need no dot products)
    const int *sch; int nsch;               // Schur assembly tasks (pair, tile row, tile-column group), most expensive first
    int cw_main, ls_stage;                  // line search: offset (into cw) of the window most block rows share; 1: it and the trial point fit the panel area
    int cu_main;                            // offset (into cu) of the u block most block rows share: staged in shared memory for the Schur assembly
};

// status words (mirror include/fmpc.h)
enum { ST_OK = 0, ST_EARLY_EXIT = 1, ST_NOT_PD = 2, ST_LS_MAX = 3, ST_NONFINITE = 4 };

// use_mma: 0 = generic CTA kernel (any n), 1 = CTA-per-instance DMMA kernel (n <= 72; the default for 32 < n <= 72), 2 = warp-per-instance DMMA kernel (n <= 32),
//          3 = general-structure kernel (ramp rows / literal VAR_1 columns / dense Q)
struct SolveLaunchCfg { int grid, block; size_t smem; int use_mma; int slots; size_t ws_stride; int np; /* block-size class of the CTA DMMA kernel */ };

// kernels.cu
int  fmpc_solve_config(const DevSys &S, int device, SolveLaunchCfg *cfg);        // 0 ok
void fmpc_launch_solve(const DevSys &S, const StepArgs &A, const SolveLaunchCfg &cfg, void *stream);
void fmpc_launch_state_update(const DevSys &S, int nbatch, const double *x, const double *xpre, const double *u,
                              const double *w, double *xnext, void *stream);
void fmpc_launch_shift_warm(const DevSys &S, int nbatch, const double *a_k, int a_stride, double *X, double *U,
                            double *x0, double *x0_pre, double *u_prev, int first, void *stream);

void fmpc_launch_shift_inplace(int n, int m, int T, int nbatch, double *X, double *U, void *stream);
void fmpc_launch_extract_first(int m, int T, int nbatch, const double *U, double *u_first, void *stream);
void fmpc_launch_log_step(int n, int m, int T, int nbatch, int K, int k, const double *U, const double *x0, const int *iters,
                          double *Uacc, double *Xacc, int *itacc, void *stream);

// MATLAB default stream MT19937 on the device: appends `count` doubles of the stream at `out`; *idx = next unread word of the
// stored block (host-side bookkeeping), `raw` = staging for the raw words (>= 2 count + 1248)
void fmpc_launch_mt_fill(unsigned *state, unsigned *raw, int *idx, double *out, unsigned long long count, void *stream);

// kernel_mma.cu : CTA-per-instance DMMA path (n <= 72)
int  fmpc_mma_config(const DevSys &S, int device, SolveLaunchCfg *cfg);          // 0 ok, <0 not applicable
void fmpc_launch_solve_mma(const DevSys &S, const StepArgs &A, const SolveLaunchCfg &cfg, void *stream);

// kernel_warp.cu : warp-per-instance DMMA path (n <= 32), the default
int  fmpc_warp_config(const DevSys &S, int device, SolveLaunchCfg *cfg);         // 0 ok, <0 not applicable
void fmpc_launch_solve_warp(const DevSys &S, const StepArgs &A, const SolveLaunchCfg &cfg, void *stream);
int  fmpc_warp_smid_slots_ok(const SolveLaunchCfg &cfg, int device);            // 1 if scratch slots may be indexed by %smid

// kernel_gen.cu : general-structure path (VAR_1 ramp rows, literal VAR_1 column placement, dense Q / Qf)
struct fmpc_sys;
int  fmpc_gen_create(const fmpc_sys *s, int device, GenSys *out, std::vector<void *> &allocs, SolveLaunchCfg *cfg);   // FMPC_* code
void fmpc_launch_solve_gen(const DevSys &S, const GenSys &G, const StepArgs &A, const SolveLaunchCfg &cfg, void *stream);
