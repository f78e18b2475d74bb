/* fmpc.h -- C-ABI boundary of the B200-native fastMPC hot path.
 *
 * The reference (jinsungkim96/MPC-SensorlessAO) exposes this path as a MATLAB value class,
 * not as a C/MEX interface.  Every entry point below cites the reference
 * interface it replaces; a MEX shim (mpc-sensorlessao_b200/matlab/fmpc_step_mex.c), a MATLAB
 * wrapper class and a Python ctypes mirror bind exactly these symbols.
 *
 * Conventions
 *   - plain C, no torch / CUDA types in any signature; `void *stream` is a cudaStream_t
 *     (NULL = the handle's own stream).
 *   - all matrices COLUMN-MAJOR double, exactly as MATLAB stores them; NULL = MATLAB [].
 *   - pointers are HOST memory unless the function name ends in `_d` (device memory on the
 *     handle's GPU).
 *   - per-instance arrays hold one column per instance: x0 is n x nb, w is (T n) x nb,
 *     X is n x T x nb (stage-major inside an instance), U is m x T x nb.
 *     The reference's interleaved z = [u_0; x_1; ...; u_{T-1}; x_T] (fast_mpc_init.m:22-25)
 *     is the de-interleaving README.md:558-570 of (U, X).
 *   - return value: 0 = ok, negative = argument / runtime error (fmpc_strerror); one code per
 *     reference error() string.  Per-instance numerical trouble is reported in status[], the
 *     batch is never aborted.
 *   - there is NO CPU fallback: every compute entry point fails with FMPC_ERR_CUDA when no
 *     sm_100 device is usable.
 */
#ifndef FMPC_H
#define FMPC_H

#ifdef __cplusplus
extern "C" {
#endif

#define FMPC_VERSION 100

/* ---- error codes (negative returns) --------------------------------------------------- */
enum {
    FMPC_OK                 =   0,
    FMPC_ERR_NULL           =  -1,  /* required pointer missing */
    FMPC_ERR_DIM            =  -2,  /* n, m, T, nbatch out of range */
    FMPC_ERR_Q_NOT_SQUARE   =  -3,  /* fast_mpc_objective.m:17-19  'State stage cost must a square matrix' */
    FMPC_ERR_R_NOT_SQUARE   =  -4,  /* fast_mpc_objective.m:20-21  'Control stage cost must a square matrix' */
    FMPC_ERR_LIN_COST_SIZE  =  -5,  /* fast_mpc_objective.m:26-47  linear cost size */
    FMPC_ERR_X_BOUND_SIZE   =  -6,  /* fast_mpc_ineq_const.m:4-6   'Check the state inequality constraints dimensions' */
    FMPC_ERR_U_BOUND_SIZE   =  -7,  /* fast_mpc_ineq_const.m:7-9   'Check cotrol iequality constraint dimension' */
    FMPC_ERR_NO_A           =  -8,  /* fast_mpc_eq_const.m:19-22   'Define the state dynamics/equality constrained matrix' */
    FMPC_ERR_NO_B           =  -9,  /* fast_mpc_eq_const.m:23-24   'Define the control dynamics/equality constrained matrix' */
    FMPC_ERR_A_SIZE         = -10,  /* fast_mpc_eq_const.m:27-30   'The equality state dynamics matrix size does not match' */
    FMPC_ERR_B_SIZE         = -11,  /* fast_mpc_eq_const.m:31-32   'The equality control dynamics matrix size does not match' */
    FMPC_ERR_INIT_SIZE      = -12,  /* fast_mpc_init.m:13-14       'Initialization size mismatch (T*(n+m))' */
    FMPC_ERR_NOT_PD         = -13,  /* chol() failure on a problem-constant block (Q, Qf, R) */
    FMPC_ERR_UNSUPPORTED    = -14,  /* valid reference input this build does not cover (non-diagonal R with ramp rows) */
    FMPC_ERR_BATCH          = -15,  /* nbatch > max_batch of the handle */
    FMPC_ERR_CUDA           = -16,  /* no usable sm_100 device / CUDA runtime error (no CPU fallback) */
    FMPC_ERR_PARAM          = -17   /* kappa <= 0, niters < 0, beta not in (0,1) ... */
};

/* ---- per-instance status words --------------------------------------------------------- */
enum {
    FMPC_ST_OK         = 0,   /* ran all `niters` Newton steps */
    FMPC_ST_EARLY_EXIT = 1,   /* inf_newton_solver.m:19-22 residual test passed before a step */
    FMPC_ST_NOT_PD     = 2,   /* chol(Schur) would have thrown (inf_newton_solver.m:30): z of the last good step returned */
    FMPC_ST_LS_MAX     = 3,   /* line search hit ls_max halvings (reference: unbounded) */
    FMPC_ST_NONFINITE  = 4    /* residual norm became NaN/Inf */
};

/* ---- problem data shared by all instances of a handle ---------------------------------- */
/* Replaces the problem-constant arguments of the Fast_MPC2 constructor
 * (VAR_2/Fast_MPC2.m:28-55, VAR_1/Fast_MPC2.m:26-51). */
typedef struct fmpc_sys {
    int n, m, T;            /* states, inputs, horizon */
    int var_order;          /* 1 (VAR_1: A2 ignored) or 2 */
    const double *A1;       /* n x n */
    const double *A2;       /* n x n, NULL iff var_order == 1 */
    const double *B;        /* n x m */
    const double *Q;        /* n x n */
    const double *R;        /* m x m */
    const double *Qf;       /* n x n */
    const double *q, *r, *qf;       /* optional linear costs (NULL = zeros, fast_mpc_objective.m:26-47) */
    const double *x_min, *x_max;    /* n; used for the cold-start midpoint only (fast_mpc_init.m:19) */
    const double *u_min, *u_max;    /* m; box rows fast_mpc_ineq_const.m:42-56 */
    const double *du_min, *du_max;  /* m; VAR_1 ramp rows (VAR_1/fast_mpc_ineq_const.m:58-79); may be NULL if !ramp_rows */
    int ramp_rows;          /* 0: box only (VAR_2 semantics); 1: VAR_1 ramp rows (needs u_prev at every step) */
    int var1_literal_bug;   /* var_order == 1 only.  1: write the second block row of C at columns n : 3n+m-1 exactly as
                             * VAR_1/fast_mpc_eq_const.m:34-37 does -- bit-for-bit the reference's VAR_1;
                             * 0: the corrected placement m+1 : 2(n+m) (= VAR_2 code with A2 = 0) */
} fmpc_sys;

/* Solver parameters: arguments (nw, k) of mpc_fixed_log_newton (VAR_2/Fast_MPC2.m:124) plus the
 * constants hard-coded in inf_newton_solver.m:9,19,36-37.  fmpc_default_params fills the
 * reference values. */
typedef struct fmpc_params {
    double kappa;       /* barrier weight k */
    int    niters;      /* nw: Newton steps (reference [] => 1000) */
    int    ls_max;      /* cap on halvings per line search; 0 = unbounded like the reference */
    double alpha;       /* 1e-4 */
    double beta;        /* 0.5  */
    double tol_r;       /* 1e-6 : ||[r_d; r_p]|| */
    double tol_p;       /* 1e-8 : ||r_p||         */
} fmpc_params;

typedef struct fmpc_handle fmpc_handle;

/* Fill `p` with the reference constants (kappa = 0.01, niters = 5: test_fast_mpc.m:53,59). */
void fmpc_default_params(fmpc_params *p);

/* Number of usable sm_100 devices (0 if none / no driver). */
int fmpc_device_count(void);

/* Validate `sys` (reference error() checks), upload the shared matrices to GPU `device`,
 * precompute the iterate-independent Schur blocks, allocate workspaces for up to `max_batch`
 * instances.  Replaces: Fast_MPC2 ctor + fast_mpc_objective / _ineq_const / _eq_const assembly. */
int fmpc_create(fmpc_handle **out, const fmpc_sys *sys, int max_batch, int device);
void fmpc_destroy(fmpc_handle *h);

/* One batched `mpc_fixed_log_newton(niters, kappa)` (VAR_2/Fast_MPC2.m:124-130 ->
 * inf_newton_solver.m:1-43) for `nbatch` independent instances, HOST buffers.
 *   x0      n x nb           current state            (ctor arg x0)
 *   x0_pre  n x nb | NULL    previous state           (ctor arg x0_pre; required iff var_order == 2)
 *   u_prev  m x nb | NULL    previous input           (ctor arg u_prev; ramp rows only)
 *   w       (T n) x nb|NULL  per-stage offsets        (ctor arg w; NULL = zeros)
 *   xf      n x nb | NULL    terminal state x_T = xf  (ctor arg xf; NULL = no terminal row)
 *   X0, U0  n x T x nb, m x T x nb | both NULL => cold start (fast_mpc_init.m:19-25)
 *                            (ctor arg x_init, de-interleaved)
 *   nu0     (T n [+ n if xf]) x nb | NULL => MATLAB default stream MT19937(5489), consumed
 *                            instance after instance (inf_newton_solver.m:2)
 *   X, U    outputs, same shapes as X0, U0 (x_opt de-interleaved, README.md:558-570)
 *   status  nb ints | NULL   FMPC_ST_*
 *   iters   nb ints | NULL   Newton steps actually taken
 *   telapsed  seconds of device time for the solve kernels (CUDA events) | NULL
 */
int fmpc_step(fmpc_handle *h, const fmpc_params *p, int nbatch,
              const double *x0, const double *x0_pre, const double *u_prev,
              const double *w, const double *xf,
              const double *X0, const double *U0, const double *nu0,
              double *X, double *U, int *status, int *iters, double *telapsed);

/* Same solve on DEVICE buffers, asynchronous on `stream` (no host sync, no allocation).
 * X/U may alias X0/U0 (in-place warm start).  nu0 must be given (device).
 * ONE solve in flight per handle: every launch of a handle shares its instance counter and scratch slots, so a second
 * fmpc_step_d / fmpc_step_r_d on the same handle must be issued on the same stream (or after the first has completed);
 * use one handle per stream for concurrent solves.  The host-buffer entry points block, so they are safe as they are. */
int fmpc_step_d(fmpc_handle *h, const fmpc_params *p, int nbatch,
                const double *x0, const double *x0_pre, const double *u_prev,
                const double *w, const double *xf,
                const double *X0, const double *U0, const double *nu0,
                double *X, double *U, int *status, int *iters, void *stream);

/* Resident closed-loop step: what the loop body of README.md:444-626 needs from the solver and nothing more.  The handle
 * keeps the solution (X, U) of its previous fmpc_step_r call on the device and warm-starts from it shifted one stage
 * (stage t <- t + 1, last stage repeated); it also keeps the previous x0 (= x0_pre of this step, README.md:483-488) and the
 * input it returned last (= u_prev of the VAR_1 ramp rows, README.md:447-452).  Per step the caller sends the estimator
 * output x0 and receives the input the loop applies, U(:,0) (README.md:589); nothing of the horizon crosses the host
 * link unless X / U are requested.  HOST buffers, blocking.
 *   flags   FMPC_R_RESET: first step of a loop -- cold start (fast_mpc_init.m:19-25), x0_pre = 0, u_prev = 0
 *           (README.md:483-485, 446-448).  Required on the first call and whenever nbatch changes.
 *   x0_pre  NULL => x0 of the previous call on this handle        u_prev  NULL => U(:,0) of the previous call (ramp rows only)
 *   w, xf   as in fmpc_step (NULL = zeros / no terminal row)       nu0  NULL => the handle's MATLAB stream, as in fmpc_step
 *   u0      m x nb  output: U(:,0) per instance
 *   X, U    both NULL, or full horizons as in fmpc_step (n x T x nb, m x T x nb)
 * Equals fmpc_step called with X0 / U0 = the previous outputs shifted one stage (tests/test_gpu_fmpc.py). */
enum { FMPC_R_RESET = 1 };
int fmpc_step_r(fmpc_handle *h, const fmpc_params *p, int nbatch, int flags,
                const double *x0, const double *x0_pre, const double *u_prev,
                const double *w, const double *xf, const double *nu0,
                double *u0, double *X, double *U, int *status, int *iters, double *telapsed);

/* The same step on DEVICE buffers, asynchronous on `stream` (NULL = the handle's own stream): for drivers that keep the
 * estimator output and the applied input on the GPU too.  Calls on one handle must be issued in order on one stream. */
int fmpc_step_r_d(fmpc_handle *h, const fmpc_params *p, int nbatch, int flags,
                  const double *x0, const double *x0_pre, const double *u_prev,
                  const double *w, const double *xf, const double *nu0,
                  double *u0, double *X, double *U, int *status, int *iters, void *stream);

/* Interleaved convenience wrapper with the reference's own I/O shape: z0 / z are
 * (T (n+m)) x nb in the layout of fast_mpc_init.m:22-25; z0 NULL => cold start. */
int fmpc_step_z(fmpc_handle *h, const fmpc_params *p, int nbatch,
                const double *x0, const double *x0_pre, const double *u_prev,
                const double *w, const double *xf, const double *z0, const double *nu0,
                double *z, int *status, int *iters, double *telapsed);

/* kappa-continuation front-ends (VAR_2/Fast_MPC2.m:88-144).  mode:
 *   FMPC_FE_FIXED_LOG     mpc_fixed_log(k)        : one solve, niters = 1000
 *   FMPC_FE_FIXED_NEWTON  mpc_fixed_newton(nw)    : k = 1, 0.1, ... while k*N >= 1e-2, nw steps each
 *   FMPC_FE_SOLVE_FULL    mpc_solve_full          : same schedule, niters = 1000 each
 *   FMPC_FE_SOLVE_CHECK   mpc_solve_check(kmin,kmax): 5 linearly spaced k from kmax down to kmin
 * nu0 (if given) holds one dual start PER OUTER kappa: (NBn) x nb x n_outer; NULL = MATLAB stream. */
enum { FMPC_FE_FIXED_LOG = 1, FMPC_FE_FIXED_NEWTON = 2, FMPC_FE_SOLVE_FULL = 3, FMPC_FE_SOLVE_CHECK = 4 };
int fmpc_frontend(fmpc_handle *h, int mode, const fmpc_params *p, double k_min, double k_max, int nbatch,
                  const double *x0, const double *x0_pre, const double *u_prev,
                  const double *w, const double *xf, const double *X0, const double *U0,
                  const double *nu0, double *X, double *U, int *status, int *iters, double *telapsed);
/* Number of outer kappa values `fmpc_frontend` visits for this handle/mode. */
int fmpc_frontend_nouter(const fmpc_handle *h, int mode);

/* Batched state update, north-star item (d) = the equality row of
 * VAR_2/fast_mpc_eq_const.m:39-47 used as a recurrence:
 *   x_next[:,b] = A1 x[:,b] + A2 x_pre[:,b] + B u[:,b] (+ w[:,b])     HOST buffers. */
int fmpc_state_update(fmpc_handle *h, int nbatch, const double *x, const double *x_pre,
                      const double *u, const double *w, double *x_next);
int fmpc_state_update_d(fmpc_handle *h, int nbatch, const double *x, const double *x_pre,
                        const double *u, const double *w, double *x_next, void *stream);

/* K closed-loop steps entirely on the device (README.md:444-626 restricted to the synthetic
 * modal loop of SURVEY.md 3.4 / 8d):  per step k and instance b
 *     x0  = a[:,k,b] + B u_prev          (perfect estimator: residual aberration)
 *     solve (warm start = previous solution shifted one stage, cold at k = 0)
 *     u_prev <- U(:,0);  logs U_acc[:,k,b] = u_prev, X_acc[:,k,b] = x0
 *   a      n x K x nb   open-loop aberration sequence (host)
 *   nu0    (T n) x nb x K | NULL (MATLAB stream: one draw per solve, instance-major within a step)
 *   U_acc  m x K x nb,  X_acc  n x K x nb   (host outputs), iters_acc K x nb | NULL
 */
int fmpc_closed_loop(fmpc_handle *h, const fmpc_params *p, int nbatch, int K,
                     const double *a, const double *nu0,
                     double *U_acc, double *X_acc, int *iters_acc, double *telapsed);

/* Re-seeds the handle's MT19937 stream (the source of nu0 == NULL dual starts); 5489 = MATLAB's `rng default`. */
int fmpc_seed_stream(fmpc_handle *h, unsigned seed);

/* ---- all GPUs of the box behind ONE blocking call ------------------------------------------------------
 * One handle, one stream set and one host thread per device, in a single process: the batch is cut into contiguous shards
 * of ceil(nbatch / G) instances, every device solves its shard (no inter-GPU traffic: instances are independent), the
 * call returns when all shards are on the host.  Same arguments and results as fmpc_step / fmpc_step_r -- with explicit
 * nu0 bit-identical to a single-device call (tests/test_gpu_multi.py).  nu0 == NULL draws from one MT19937 stream PER
 * DEVICE (device 0: MATLAB's default stream, device g: seed 5489 + g): one sequential stream cannot feed several GPUs.
 *   ngpus   <= 0: every usable sm_100 device          devices  NULL: 0 .. ngpus-1 */
typedef struct fmpc_multi fmpc_multi;
typedef struct fmpc_multi_stats {          /* one 64-byte record per device, all doubles (it travels over ncclAllGather) */
    double device, n_solves, device_seconds, newton_iters, status_hist[4];   /* FMPC_ST_OK, _EARLY_EXIT, _NOT_PD, _LS_MAX + _NONFINITE */
} fmpc_multi_stats;
int  fmpc_multi_create(fmpc_multi **out, const fmpc_sys *sys, int max_batch, int ngpus, const int *devices);
void fmpc_multi_destroy(fmpc_multi *M);
int  fmpc_multi_ngpus(const fmpc_multi *M);
/* shard of device slot g for a batch of nbatch instances: first instance and count */
int  fmpc_multi_shard(const fmpc_multi *M, int nbatch, int g, int *first, int *count);
fmpc_handle *fmpc_multi_handle(fmpc_multi *M, int g);          /* the per-device handle (e.g. for fmpc_kernel_kind) */
int  fmpc_multi_step(fmpc_multi *M, const fmpc_params *p, int nbatch,
                     const double *x0, const double *x0_pre, const double *u_prev,
                     const double *w, const double *xf,
                     const double *X0, const double *U0, const double *nu0,
                     double *X, double *U, int *status, int *iters, double *telapsed /* max over devices */);
int  fmpc_multi_step_r(fmpc_multi *M, const fmpc_params *p, int nbatch, int flags,
                       const double *x0, const double *x0_pre, const double *u_prev,
                       const double *w, const double *xf, const double *nu0,
                       double *u0, double *X, double *U, int *status, int *iters, double *telapsed);
/* Per-device statistics of the last step (solves, device seconds, Newton iterations, status histogram).  use_nccl != 0: gathered with ncclAllGather over NVLink (ncclCommInitAll in this process, libnccl.so.2
 * loaded with dlopen) -- the only collective of the path; else, or if NCCL is not available, read by the host.
 * Returns the number of records written to out[0 .. G-1]; *used_nccl tells which way they came. */
int  fmpc_multi_last_stats(fmpc_multi *M, fmpc_multi_stats *out, int use_nccl, int *used_nccl);

/* Dimensions the handle was created with (any pointer may be NULL). */
int fmpc_get_dims(const fmpc_handle *h, int *n, int *m, int *T);

/* Device-resident workspace access for benchmarks / pipelines: size in bytes the handle holds. */
long long fmpc_workspace_bytes(const fmpc_handle *h);
/* Kernels launched by this handle since creation (bench.py's gpu_launches). */
long long fmpc_launch_count(const fmpc_handle *h);
/* Total Newton iterations executed by the last fmpc_step* call on this handle (sum over instances);
 * requires a device sync, so call it outside timed regions. */
long long fmpc_last_newton_iters(fmpc_handle *h);
/* Which solve kernel the handle selected: 2 = warp-per-instance DMMA kernel (n <= 32), 1 = CTA-per-instance
 * DMMA kernel (32 < n <= 72), 0 = scalar reference kernel (only when forced), 3 = general-structure kernel (VAR_1 ramp rows, literal
 * VAR_1 columns, dense Q / Qf, dense R, n > 72). */
int fmpc_kernel_kind(const fmpc_handle *h);
/* Phase cycle counters of the last solve launch, summed over warps (all zero unless the library was
 * built with -DFMPC_PROF): warp kernel: init, newton pass, forward sweep, backward sweep, C' pass, line search,
 * accept, copy-out, 4 spare; general-structure kernel (thread 0 of every CTA): init, barrier + residuals, inv(Phi_uu), rhs,
 * Schur assembly, potrf + staging, panel, trailing update, backward substitution, dz, line search, copy-out. */
int fmpc_last_profile(fmpc_handle *h, long long *out12);

const char *fmpc_strerror(int code);

/* FP64 pipe micro-benchmarks used as roofline denominators (MEASURED_PEAKS.json has no FP64 entry).
 * kind 0: dependent-free DFMA streams; kind 1: mma.sync m8n8k4 f64 (DMMA).  Returns TFLOP/s (<0 on error). */
double fmpc_fp64_peak(int device, int kind, int iters);

/* ======================================================================================= */
/* zernmodfit: masked least-squares projection of nL x nL phase frames onto the Zernike basis */
/* Replaces zernmodfit.m:195-213 + zernfun.m:140-192 called per frame by README.md:88-93.   */
typedef struct zmf_handle zmf_handle;

/* Builds, once, the pupil grid of README.md:78-84, the basis Z (npix_in x nmodes, modes ordered
 * n = 0..N, m = -n:2:n) and its least-squares operator W = pinv(Z) in fp64 on the host, and
 * uploads W scattered to the full nL x nL frame (zeros outside the pupil). */
int zmf_create(zmf_handle **out, int nL, int N, int max_frames, int device);
/* zernmodfit on an ARBITRARY sample set (zernmodfit.m:154-213 takes any vectors r, theta, data): the basis is evaluated at
 * the npts given samples (0 <= r <= 1 else FMPC_ERR_DIM, zernmodfit.m:182-184), W = pinv(Z) once; zmf_fit then takes
 * `frames` = npts x nf (one data vector per column, in the order of r / theta) and returns coef nmodes x nf.
 * FMPC_ERR_NOT_PD if the samples do not determine the modes (Z rank deficient); zmf_synth is not available. */
int zmf_create_samples(zmf_handle **out, int npts, const double *r, const double *theta, int N, int max_frames, int device);
void zmf_destroy(zmf_handle *h);
int zmf_nmodes(const zmf_handle *h);
int zmf_npix_in(const zmf_handle *h);
/* frames: nL x nL x nf column-major (MATLAB phase(:,:,j)); pixels outside the pupil are ignored
 * (may be NaN, zernmodfit.m:30).  coef: nmodes x nf (column j = ad(:,1) of frame j). HOST buffers. */
int zmf_fit(zmf_handle *h, int nframes, const double *frames, double *coef, double *telapsed);
int zmf_fit_d(zmf_handle *h, int nframes, const double *frames, double *coef, void *stream);
/* Synthesis, the step after the path (README.md:592-598  phase_cor = sum_j ad_cor(j) * Z_j):
 * frames(:,:,f) = sum_j coef(j,f) Z_j inside the pupil, 0 outside.  coef: nmodes x nf, frames: nL x nL x nf (column-major). */
int zmf_synth(zmf_handle *h, int nframes, const double *coef, double *frames, double *telapsed);   /* HOST buffers */
int zmf_synth_d(zmf_handle *h, int nframes, const double *coef, double *frames, void *stream);     /* DEVICE buffers */
/* Copies the host-side basis (npix_in x nmodes, col-major) / mask (nL*nL bytes, col-major) out, for tests. */
int zmf_get_basis(const zmf_handle *h, double *Z);
int zmf_get_mask(const zmf_handle *h, unsigned char *mask);
long long zmf_launch_count(const zmf_handle *h);

/* ======================================================================================= */
/* Estimator step in front of the solve (README.md:478):                                      */
/*     ad_est = lsqminnorm((A_s'*A_s), ((A_s)'*(Y_M - b_s)));                                 */
/* with A_s (npix x nmodes, column-major; the caller removes the piston column like           */
/* README.md:289-290) and b_s (npix) from model_approx.mat.  Batched: y is npix x nb (one      */
/* measurement vector Y_M per column), x_hat is nmodes x nb.  b_s may be NULL (zeros).         */
/* Rank-deficient A_s: the minimum-norm solution, like lsqminnorm (eigenvalues of A_s'A_s below max(size) * eps(norm) dropped). */
typedef struct zmf_handle est_handle;
int est_create(est_handle **out, int npix, int nmodes, const double *A_s, const double *b_s, int max_batch, int device);
void est_destroy(est_handle *h);
int est_apply(est_handle *h, int nb, const double *y, double *x_hat, double *telapsed);          /* HOST buffers */
int est_apply_d(est_handle *h, int nb, const double *y, double *x_hat, void *stream);            /* DEVICE buffers */
long long est_launch_count(const est_handle *h);

/* ======================================================================================= */
/* VAR(p) identification of the coefficient time series (README.md:116-130), batched:         */
/*     AA(i-PN, n(j-1)+1:n j) = ad_acc(i-j,:), BB(i-PN,:) = ad_acc(i,:), i = PN+1..num_train   */
/*     PARA = (AA'*AA)\AA'*BB;  A_j = PARA(n(j-1)+1:n j, :)'                                   */
/*   ad   K x n x nseq   training series, column-major like ad_acc(1:num_train,:) (HOST)       */
/*   A    n x n x order x nseq   A(:,:,j,s) = A_j of sequence s (HOST, output)                 */
/*   info nseq | NULL    0, or failing column + 1 of chol(AA'AA) for that sequence             */
int var_identify(int nseq, int K, int n, int order, const double *ad, double *A, int *info, int device, double *telapsed);

#ifdef __cplusplus
}
#endif
#endif /* FMPC_H */
