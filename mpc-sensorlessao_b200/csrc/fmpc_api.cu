// C-ABI layer of the fastMPC hot path (include/fmpc.h): argument validation with the reference's
// error() semantics, host-side precompute of the iterate-independent Schur blocks, device buffers,
// the MATLAB default random stream, and the batched entry points.  No CPU fallback: every compute
// entry point fails with FMPC_ERR_CUDA when no sm_100 device is usable.
#include <cuda_runtime.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <mutex>
#include <new>
#include <thread>
#include <utility>
#include <vector>
#include "../../include/fmpc.h"
#include "fmpc_internal.h"

namespace {

// ---------------------------------------------------------------------------------------------
// MATLAB's default global stream: MT19937 seeded with 5489, doubles from genrand_res53
// (`rand` in inf_newton_solver.m:2).  The host only seeds the state; the stream itself is
// generated on the device (fmpc_mt_fill_kernel).
// ---------------------------------------------------------------------------------------------
struct MT19937 {
    uint32_t mt[624];
    int idx;
    explicit MT19937(uint32_t seed = 5489u) { reseed(seed); }
    void reseed(uint32_t seed)
    {
        mt[0] = seed;
        for (int i = 1; i < 624; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
        idx = 624;
    }
};

#define CU_OK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { last_cuda_error = e_; return FMPC_ERR_CUDA; } } while (0)
thread_local cudaError_t last_cuda_error = cudaSuccess;

struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
    int ensure(size_t need)
    {
        if (need <= bytes) return 0;
        if (p) cudaFree(p);
        p = nullptr; bytes = 0;
        if (cudaMalloc(&p, need) != cudaSuccess) return -1;
        bytes = need;
        return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
    template <class T> T *as() const { return (T *)p; }
};

// ---------------------------------------------------------------------------------------------
// Pageable host buffers (what mxGetPr hands the MEX gateway): cudaMemcpyAsync from / to them is a synchronous staged
// copy inside the driver, which serialises the copy-in / solve / copy-out pipeline of fmpc_step.  Instead the library stages
// them itself: a few worker threads memcpy each chunk between the caller's buffer and a ring of pinned slots, and the DMA
// engines only ever see pinned memory.
// ---------------------------------------------------------------------------------------------
struct CopyPool {
    struct Job { char *dst; const char *src; size_t bytes; };
    std::vector<std::thread> th;
    std::mutex mu;
    std::condition_variable cv_work, cv_done;
    std::vector<Job> jobs;
    size_t next = 0, pending = 0;
    bool stop = false;
    void start(int nthreads)
    {
        for (int i = 0; i < nthreads; ++i) th.emplace_back([this] { run(); });
    }
    void run()
    {
        std::unique_lock<std::mutex> lk(mu);
        for (;;) {
            cv_work.wait(lk, [this] { return stop || next < jobs.size(); });
            if (stop) return;
            const Job j = jobs[next++];
            lk.unlock();
            std::memcpy(j.dst, j.src, j.bytes);
            lk.lock();
            if (--pending == 0) cv_done.notify_all();
        }
    }
    // copies every (dst, src, bytes) triple, cut into pieces of <= 1 MiB, on the workers and the calling thread
    void copy(const std::vector<Job> &list)
    {
        std::unique_lock<std::mutex> lk(mu);
        jobs.clear(); next = 0;
        const size_t piece = (size_t)1 << 20;
        for (const Job &j : list)
            for (size_t o = 0; o < j.bytes; o += piece) jobs.push_back({j.dst + o, j.src + o, (j.bytes - o < piece) ? j.bytes - o : piece});
        pending = jobs.size();
        if (!pending) return;
        cv_work.notify_all();
        while (next < jobs.size()) {          // the caller works too
            const Job j = jobs[next++];
            lk.unlock();
            std::memcpy(j.dst, j.src, j.bytes);
            lk.lock();
            --pending;
        }
        cv_done.wait(lk, [this] { return pending == 0; });
    }
    ~CopyPool()
    {
        { std::lock_guard<std::mutex> lk(mu); stop = true; }
        cv_work.notify_all();
        for (auto &t : th) t.join();
    }
};

struct PinBuf {
    void *p = nullptr;
    size_t bytes = 0;
    int ensure(size_t need)
    {
        if (need <= bytes) return 0;
        if (p) cudaFreeHost(p);
        p = nullptr; bytes = 0;
        if (cudaMallocHost(&p, need) != cudaSuccess) return -1;
        bytes = need;
        return 0;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; bytes = 0; }
};

// true if the DMA engines can reach `p` directly (pinned / registered host memory, device or managed memory)
bool dma_reachable(const void *p)
{
    if (!p) return true;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type != cudaMemoryTypeUnregistered;
}

bool is_diag(const double *A, int n)
{
    for (int j = 0; j < n; ++j)
        for (int i = 0; i < n; ++i)
            if (i != j && A[(size_t)j * n + i] != 0.0) return false;
    return true;
}

} // namespace

struct fmpc_handle {
    int device = 0;
    int n = 0, m = 0, T = 0, var_order = 2, max_batch = 0;
    DevSys S{};
    GenSys G{};
    int ramp = 0;
    SolveLaunchCfg cfg{};
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // host-buffer entry points: copy-in / solve / copy-out of successive instance chunks overlap on three streams
    static constexpr int MAX_CHUNKS = 32, MAX_PART = 4;
    bool other_blocks_dirty = false;      // counter blocks 1.. hold totals of a multi-chunk call
    int smid_slots = 0;                   // warp kernel: scratch slots indexed by SM id -> chunk kernels of one call may overlap freely
    cudaStream_t s_in = nullptr, s_out = nullptr;
    cudaStream_t s_k[MAX_PART] = {};      // solve streams: partition p of the SMs / scratch slots runs chunks p, p + NP, ...
    cudaEvent_t ev_in[MAX_CHUNKS] = {}, ev_k[MAX_CHUNKS] = {};
    std::vector<void *> sys_allocs;       // problem-constant device arrays
    DevBuf ws, counters;                  // per-CTA scratch; {counter u32 (pad), iters_total u64}
    // staging for the host-pointer entry points
    DevBuf d_x0, d_x0pre, d_uprev, d_w, d_xf, d_X, d_U, d_nu0, d_status, d_iters;
    // closed-loop state
    DevBuf d_a, d_Uacc, d_Xacc, d_itacc;
    long long launches = 0;
    size_t ws_stride = 0;
    // MATLAB's default stream on the device (nu0 == NULL): MT19937 state (624 words + read index) and two buffers of stream
    // doubles.  nu_buf[nu_cur][0 .. nu_have) are the next unread doubles, generated on s_gen ahead of the call that uses them.
    DevBuf d_mt, d_mt_raw, nu_buf[2];
    int mt_idx = 624;                     // next unread word of the block stored in d_mt (624: nothing unread)
    int nu_cur = 0;
    size_t nu_have = 0, nu_cap = 0;
    cudaStream_t s_gen = nullptr;
    cudaEvent_t ev_nu_ready[2] = {}, ev_nu_free[2] = {};
    bool nu_free_pending[2] = {false, false};
    bool nu_stream_active = false;        // the last call drew its dual starts from the stream (and prefetched the next ones)
    // staging of pageable host buffers (fmpc_step): NSTAGE pinned slots per direction, worker threads created on first use
    static constexpr int NSTAGE = 3;
    PinBuf stg_in[NSTAGE], stg_out[NSTAGE];
    cudaEvent_t ev_stg_in[NSTAGE] = {}, ev_stg_out[NSTAGE] = {};
    CopyPool *pool = nullptr;
    // resident closed-loop state of fmpc_step_r: the previous solution (r_X, r_U), x0 of the previous call, U(:,0)
    DevBuf r_X, r_U, r_x0, r_x0pre, r_u0;
    int r_nb = 0;                         // instances the resident state is valid for (0 = none: the next call must reset)
};

namespace {

template <class T> T *upload(fmpc_handle *h, const std::vector<T> &v)
{
    void *p = nullptr;
    if (cudaMalloc(&p, v.size() * sizeof(T) + 16) != cudaSuccess) return nullptr;
    if (!v.empty() && cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice) != cudaSuccess) { cudaFree(p); return nullptr; }
    h->sys_allocs.push_back(p);
    return (T *)p;
}

// Builds the iterate-independent blocks of Y = C inv(Phi) C' (row-major n x n) for diagonal Q/Qf:
//   Y[i,i]   = B Rt_i^-1 B' (iterate dependent, added on the device)
//              + Qi_{i+1} + A1 Qi_i A1' [i>=1] + A2 Qi_{i-1} A2' [i>=2]         ; Y[T,T] = Qi_T (xf row)
//   Y[i+1,i] = -A1 Qi_{i+1} + A2 Qi_i A1' [i>=1]   (i+1 <= T-1)                 ; Y[T,T-1] = Qi_T
//   Y[i+2,i] = -A2 Qi_{i+1}                         (i+2 <= T-1)                 ; Y[T,T-2] = 0
// with Qi_j = inv(2 Q_j), Q_T = Qf  (C rows: VAR_2/fast_mpc_eq_const.m:38-49,67-71; H: fast_mpc_objective.m:50-55).
struct YTables {
    std::vector<double> pool;
    std::vector<int> ydi, y1i, y2i;
};

int intern_block(YTables &Y, const std::vector<double> &blk, size_t nn)
{
    bool zero = true;
    for (double v : blk) if (v != 0.0) { zero = false; break; }
    if (zero) return -1;
    const size_t nblk = Y.pool.size() / nn;
    for (size_t b = 0; b < nblk; ++b)
        if (std::memcmp(Y.pool.data() + b * nn, blk.data(), nn * sizeof(double)) == 0) return (int)b;
    Y.pool.insert(Y.pool.end(), blk.begin(), blk.end());
    return (int)nblk;
}

void build_y_tables(const fmpc_sys *s, const std::vector<double> &qi, const std::vector<double> &qif, YTables &Y)
{
    const int n = s->n, T = s->T;
    const size_t nn = (size_t)n * n;
    const bool a2 = (s->var_order == 2);
    auto A1 = [&](int r, int c) { return s->A1[(size_t)c * n + r]; };
    auto A2 = [&](int r, int c) { return s->A2[(size_t)c * n + r]; };
    auto Qi = [&](int j, int k) { return (j == T) ? qif[k] : qi[k]; };
    Y.ydi.assign(T + 1, -1); Y.y1i.assign(T + 1, -1); Y.y2i.assign(T + 1, -1);
    std::vector<double> blk(nn);
    for (int i = 0; i <= T; ++i) {
        // diagonal block
        std::fill(blk.begin(), blk.end(), 0.0);
        if (i == T) {
            for (int k = 0; k < n; ++k) blk[(size_t)k * n + k] = Qi(T, k);
        } else {
            for (int r = 0; r < n; ++r)
                for (int c = 0; c < n; ++c) {
                    double v = (r == c) ? Qi(i + 1, r) : 0.0;
                    if (i >= 1) for (int k = 0; k < n; ++k) v += A1(r, k) * Qi(i, k) * A1(c, k);
                    if (i >= 2 && a2) for (int k = 0; k < n; ++k) v += A2(r, k) * Qi(i - 1, k) * A2(c, k);
                    blk[(size_t)r * n + c] = v;
                }
        }
        Y.ydi[i] = intern_block(Y, blk, nn);
        // first sub-diagonal Y[i+1,i]
        std::fill(blk.begin(), blk.end(), 0.0);
        if (i + 1 <= T - 1) {
            for (int r = 0; r < n; ++r)
                for (int c = 0; c < n; ++c) {
                    double v = -A1(r, c) * Qi(i + 1, c);
                    if (i >= 1 && a2) for (int k = 0; k < n; ++k) v += A2(r, k) * Qi(i, k) * A1(c, k);
                    blk[(size_t)r * n + c] = v;
                }
        } else if (i + 1 == T) {
            for (int k = 0; k < n; ++k) blk[(size_t)k * n + k] = Qi(T, k);
        }
        Y.y1i[i] = intern_block(Y, blk, nn);
        // second sub-diagonal Y[i+2,i]
        std::fill(blk.begin(), blk.end(), 0.0);
        if (i + 2 <= T - 1 && a2)
            for (int r = 0; r < n; ++r)
                for (int c = 0; c < n; ++c) blk[(size_t)r * n + c] = -A2(r, c) * Qi(i + 1, c);
        Y.y2i[i] = intern_block(Y, blk, nn);
    }
    if (Y.pool.empty()) Y.pool.assign(nn, 0.0);
}

int validate_sys(const fmpc_sys *s)
{
    if (!s) return FMPC_ERR_NULL;
    if (s->n < 1 || s->m < 1 || s->T < 1 || s->n > 512 || s->m > 4096 || s->T > 4096) return FMPC_ERR_DIM;
    if (s->var_order != 1 && s->var_order != 2) return FMPC_ERR_DIM;
    if (!s->Q || !s->Qf) return FMPC_ERR_Q_NOT_SQUARE;
    if (!s->R) return FMPC_ERR_R_NOT_SQUARE;
    if (!s->x_min || !s->x_max) return FMPC_ERR_X_BOUND_SIZE;
    if (!s->u_min || !s->u_max) return FMPC_ERR_U_BOUND_SIZE;
    if (!s->A1) return FMPC_ERR_NO_A;
    if (s->var_order == 2 && !s->A2) return FMPC_ERR_NO_A;
    if (!s->B) return FMPC_ERR_NO_B;
    if (s->ramp_rows && (!s->du_min || !s->du_max)) return FMPC_ERR_U_BOUND_SIZE;
    return FMPC_OK;
}

int validate_params(const fmpc_params *p)
{
    if (!p) return FMPC_ERR_NULL;
    if (!(p->kappa > 0.0) || p->niters < 0 || p->ls_max < 0) return FMPC_ERR_PARAM;
    if (!(p->beta > 0.0 && p->beta < 1.0) || !(p->alpha >= 0.0 && p->alpha < 1.0)) return FMPC_ERR_PARAM;
    return FMPC_OK;
}

int ensure_device(int device)
{
    int cnt = 0;
    if (cudaGetDeviceCount(&cnt) != cudaSuccess || device < 0 || device >= cnt) return FMPC_ERR_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return FMPC_ERR_CUDA;
    if (prop.major != 10) return FMPC_ERR_CUDA;       // built for sm_100a only
    if (cudaSetDevice(device) != cudaSuccess) return FMPC_ERR_CUDA;
    return FMPC_OK;
}

} // namespace

extern "C" {

void fmpc_default_params(fmpc_params *p)
{
    if (!p) return;
    p->kappa = 0.01;     // test_fast_mpc.m:53, README.md:551
    p->niters = 5;       // test_fast_mpc.m:59
    p->ls_max = 0;
    p->alpha = 1e-4;     // inf_newton_solver.m:36
    p->beta = 0.5;       // inf_newton_solver.m:37
    p->tol_r = 1e-6;     // inf_newton_solver.m:9
    p->tol_p = 1e-8;     // inf_newton_solver.m:19
}

int fmpc_device_count(void)
{
    int cnt = 0;
    if (cudaGetDeviceCount(&cnt) != cudaSuccess) return 0;
    int ok = 0;
    for (int d = 0; d < cnt; ++d) {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, d) == cudaSuccess && prop.major == 10) ++ok;
    }
    return ok;
}

const char *fmpc_strerror(int code)
{
    switch (code) {
    case FMPC_OK: return "ok";
    case FMPC_ERR_NULL: return "required pointer is NULL";
    case FMPC_ERR_DIM: return "dimension out of range";
    case FMPC_ERR_Q_NOT_SQUARE: return "State stage cost must a square matrix";
    case FMPC_ERR_R_NOT_SQUARE: return "Control stage cost must a square matrix";
    case FMPC_ERR_LIN_COST_SIZE: return "Linear state cost needs to be a vector of size n";
    case FMPC_ERR_X_BOUND_SIZE: return "Check the state inequality constraints dimensions";
    case FMPC_ERR_U_BOUND_SIZE: return "Check cotrol iequality constraint dimension";
    case FMPC_ERR_NO_A: return "Define the state dynamics/equality constrained matrix";
    case FMPC_ERR_NO_B: return "Define the control dynamics/equality constrained matrix";
    case FMPC_ERR_A_SIZE: return "The equality state dynamics matrix size does not match";
    case FMPC_ERR_B_SIZE: return "The equality control dynamics matrix size does not match";
    case FMPC_ERR_INIT_SIZE: return "Initialization size mismatch (T*(n+m))";
    case FMPC_ERR_NOT_PD: return "cost matrix is not positive definite";
    case FMPC_ERR_UNSUPPORTED: return "input not covered by this build (non-diagonal R together with ramp rows, a state block too large for shared memory, or a literal VAR_1 C that MATLAB itself would reject; see DESIGN.md)";
    case FMPC_ERR_BATCH: return "nbatch exceeds the handle's max_batch";
    case FMPC_ERR_CUDA: return "no usable sm_100 CUDA device or CUDA runtime error (there is no CPU fallback)";
    case FMPC_ERR_PARAM: return "invalid solver parameter";
    default: return "unknown fmpc error";
    }
}

void fmpc_destroy(fmpc_handle *h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    for (void *p : h->sys_allocs) cudaFree(p);
    DevBuf *bufs[] = {&h->ws, &h->counters, &h->d_x0, &h->d_x0pre, &h->d_uprev, &h->d_w, &h->d_xf, &h->d_X, &h->d_U,
                      &h->d_nu0, &h->d_status, &h->d_iters, &h->d_a, &h->d_Uacc, &h->d_Xacc, &h->d_itacc,
                      &h->d_mt, &h->d_mt_raw, &h->nu_buf[0], &h->nu_buf[1], &h->r_X, &h->r_U, &h->r_x0, &h->r_x0pre, &h->r_u0};
    for (DevBuf *b : bufs) b->release();
    for (int i = 0; i < 2; ++i) { if (h->ev_nu_ready[i]) cudaEventDestroy(h->ev_nu_ready[i]); if (h->ev_nu_free[i]) cudaEventDestroy(h->ev_nu_free[i]); }
    if (h->s_gen) cudaStreamDestroy(h->s_gen);
    for (int i = 0; i < fmpc_handle::NSTAGE; ++i) {
        h->stg_in[i].release(); h->stg_out[i].release();
        if (h->ev_stg_in[i]) cudaEventDestroy(h->ev_stg_in[i]);
        if (h->ev_stg_out[i]) cudaEventDestroy(h->ev_stg_out[i]);
    }
    delete h->pool;
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    for (int i = 0; i < fmpc_handle::MAX_CHUNKS; ++i) { if (h->ev_in[i]) cudaEventDestroy(h->ev_in[i]); if (h->ev_k[i]) cudaEventDestroy(h->ev_k[i]); }
    for (int i = 0; i < fmpc_handle::MAX_PART; ++i) if (h->s_k[i]) cudaStreamDestroy(h->s_k[i]);
    if (h->s_in) cudaStreamDestroy(h->s_in);
    if (h->s_out) cudaStreamDestroy(h->s_out);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

int fmpc_create(fmpc_handle **out, const fmpc_sys *s, int max_batch, int device)
{
    if (!out) return FMPC_ERR_NULL;
    *out = nullptr;
    int rc = validate_sys(s);
    if (rc) return rc;
    if (max_batch < 1) return FMPC_ERR_DIM;
    const int n = s->n, m = s->m, T = s->T;
    // general-structure kernel: VAR_1 ramp rows, the literal VAR_1 column placement, dense Q / Qf, dense R
    const bool lit = (s->var_order == 1) && s->var1_literal_bug;
    const bool dense_r = !is_diag(s->R, m);
    // n > 72: the stage blocks of the block-banded kernels no longer fit shared memory; the general kernel (dense Schur complement,
    // n x n diagonal block + row panel in shared memory) takes over -- the reference has no size limit (fast_mpc_eq_const.m:14)
    const bool need_gen = s->ramp_rows || lit || dense_r || !is_diag(s->Q, n) || !is_diag(s->Qf, n) || n > 72;
    if (lit && T < 3) return FMPC_ERR_UNSUPPORTED;           // fast_mpc_eq_const.m:55 then rewrites the mis-placed row itself
    for (int k = 0; k < n; ++k)
        if (!(s->Q[(size_t)k * n + k] > 0.0) || !(s->Qf[(size_t)k * n + k] > 0.0)) return FMPC_ERR_NOT_PD;
    for (int j = 0; j < m; ++j)
        if (!(s->R[(size_t)j * m + j] > 0.0)) return FMPC_ERR_NOT_PD;
    rc = ensure_device(device);
    if (rc) return rc;

    fmpc_handle *h = new (std::nothrow) fmpc_handle();
    if (!h) return FMPC_ERR_CUDA;
    h->device = device; h->n = n; h->m = m; h->T = T; h->var_order = s->var_order; h->max_batch = max_batch;
    h->ramp = s->ramp_rows ? 1 : 0;
    const bool a2 = (s->var_order == 2);

    std::vector<double> B(s->B, s->B + (size_t)n * m), Bt((size_t)n * m);
    for (int k = 0; k < n; ++k) for (int j = 0; j < m; ++j) Bt[(size_t)k * m + j] = B[(size_t)j * n + k];
    std::vector<double> A1(s->A1, s->A1 + (size_t)n * n), A1t((size_t)n * n), A2((size_t)n * n, 0.0), A2t((size_t)n * n, 0.0);
    if (a2) A2.assign(s->A2, s->A2 + (size_t)n * n);
    for (int r = 0; r < n; ++r) for (int c = 0; c < n; ++c) { A1t[(size_t)r * n + c] = A1[(size_t)c * n + r]; A2t[(size_t)r * n + c] = A2[(size_t)c * n + r]; }
    std::vector<double> r2(m), rl(m, 0.0), q2(n), q2f(n), qi(n), qif(n), ql(n, 0.0), qfl(n, 0.0);
    for (int j = 0; j < m; ++j) { r2[j] = 2.0 * s->R[(size_t)j * m + j]; if (s->r) rl[j] = s->r[j]; }
    for (int k = 0; k < n; ++k) {
        q2[k] = 2.0 * s->Q[(size_t)k * n + k]; q2f[k] = 2.0 * s->Qf[(size_t)k * n + k];
        qi[k] = 1.0 / q2[k]; qif[k] = 1.0 / q2f[k];
        if (s->q) ql[k] = s->q[k];
        if (s->qf) qfl[k] = s->qf[k];
    }
    YTables Y;
    if (!need_gen) build_y_tables(s, qi, qif, Y);
    else { Y.pool.assign((size_t)n * n, 0.0); Y.ydi.assign(T + 1, -1); Y.y1i.assign(T + 1, -1); Y.y2i.assign(T + 1, -1); }

    DevSys &S = h->S;
    S.n = n; S.m = m; S.T = T; S.has_a2 = a2 ? 1 : 0;
    bool ok = true;
#define UP(field, vec) do { S.field = upload(h, vec); if (!S.field) ok = false; } while (0)
    UP(B, B); UP(Bt, Bt); UP(A1, A1); UP(A1t, A1t); UP(A2, A2); UP(A2t, A2t);
    UP(r2, r2); UP(rl, rl); UP(q2, q2); UP(q2f, q2f); UP(qi, qi); UP(qif, qif); UP(ql, ql); UP(qfl, qfl);
    std::vector<double> umin(s->u_min, s->u_min + m), umax(s->u_max, s->u_max + m), xmin(s->x_min, s->x_min + n), xmax(s->x_max, s->x_max + n);
    UP(umin, umin); UP(umax, umax); UP(xmin, xmin); UP(xmax, xmax);
    UP(ypool, Y.pool); UP(ydi, Y.ydi); UP(y1i, Y.y1i); UP(y2i, Y.y2i);
    if (need_gen) { S.G = nullptr; S.ypk = nullptr; S.npairs = S.Mp = S.mp = 0; }
    else {   // pair-product matrix for the DMMA path
        const int np = n * (n + 1) / 2, Mp = (np + 7) & ~7, mp = (m + 3) & ~3;
        std::vector<double> G((size_t)Mp * mp, 0.0);
        for (int r = 0; r < n; ++r)
            for (int c = 0; c <= r; ++c) {
                double *row = &G[(size_t)(r * (r + 1) / 2 + c) * mp];
                for (int j = 0; j < m; ++j) row[j] = B[(size_t)j * n + r] * B[(size_t)j * n + c];
            }
        S.npairs = np; S.Mp = Mp; S.mp = mp;
        UP(G, G);
        const size_t nblk = Y.pool.size() / ((size_t)n * n);
        std::vector<double> ypk(nblk * Mp, 0.0);
        for (size_t bidx = 0; bidx < nblk; ++bidx)
            for (int r = 0; r < n; ++r)
                for (int c = 0; c <= r; ++c) ypk[bidx * Mp + r * (r + 1) / 2 + c] = Y.pool[bidx * n * n + (size_t)r * n + c];
        UP(ypk, ypk);
    }
#undef UP
    if (ok && need_gen) {
        const int rcg = fmpc_gen_create(s, device, &h->G, h->sys_allocs, &h->cfg);
        if (rcg != FMPC_OK) { fmpc_destroy(h); return rcg; }
        if (h->cfg.slots > max_batch) h->cfg.slots = max_batch;
    } else if (ok) {
        // kernel selection: warp-per-instance DMMA kernel (n <= 32) > CTA DMMA kernel (experiments only) > generic
        const char *force = getenv("FMPC_FORCE_KERNEL");       // "v1" | "v2" : A/B experiments
        const bool want_v1 = force && force[0] == 'v' && force[1] == '1';
        const bool want_v2 = force && force[0] == 'v' && force[1] == '2';
        const WsLayout L = WsLayout::make(n, m, T);
        int rc2 = -1;
        if (!want_v1 && !want_v2) rc2 = fmpc_warp_config(S, device, &h->cfg);
        if (rc2 != 0 && !want_v1) {
            rc2 = fmpc_mma_config(S, device, &h->cfg);
            if (rc2 == 0) { h->cfg.slots = h->cfg.grid; h->cfg.ws_stride = L.total; }
        }
        if (rc2 != 0) {
            const int rc1 = fmpc_solve_config(S, device, &h->cfg);
            if (rc1 == -2) { fmpc_destroy(h); return FMPC_ERR_UNSUPPORTED; }     // state blocks too large for shared memory (n > 72)
            if (rc1 != 0) ok = false;
            else { h->cfg.slots = h->cfg.grid; h->cfg.ws_stride = L.total; }
        }
    }
    if (ok) {
        h->ws_stride = h->cfg.ws_stride;
        const size_t wsb = (size_t)h->cfg.slots * h->cfg.ws_stride * sizeof(double);
        if (h->ws.ensure(wsb) || h->counters.ensure(256 * fmpc_handle::MAX_CHUNKS)) ok = false;
        if (ok && cudaMemset(h->counters.p, 0, 256 * fmpc_handle::MAX_CHUNKS) != cudaSuccess) ok = false;
        // SM-id indexed scratch slots are OPT-IN (FMPC_SMID_SLOTS=1): %smid is not stable under compute preemption / MPS
        // time-slicing / a debugger, where a migrated CTA would share a slot with a newly scheduled one.  The default keeps the
        // disjoint slot_base partitions of step_device.
        if (ok && h->cfg.use_mma == 2 && getenv("FMPC_SMID_SLOTS") && atoi(getenv("FMPC_SMID_SLOTS")) > 0)
            h->smid_slots = fmpc_warp_smid_slots_ok(h->cfg, device);
        // the warp kernel relies on never-written padding columns of its scratch staying zero
        if (ok && cudaMemset(h->ws.p, 0, wsb) != cudaSuccess) ok = false;
    }
    if (ok && cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) ok = false;
    if (ok && (cudaEventCreate(&h->ev0) != cudaSuccess || cudaEventCreate(&h->ev1) != cudaSuccess)) ok = false;
    if (ok && (cudaStreamCreateWithFlags(&h->s_in, cudaStreamNonBlocking) != cudaSuccess ||
               cudaStreamCreateWithFlags(&h->s_out, cudaStreamNonBlocking) != cudaSuccess)) ok = false;
    for (int i = 0; ok && i < fmpc_handle::MAX_PART; ++i)
        if (cudaStreamCreateWithFlags(&h->s_k[i], cudaStreamNonBlocking) != cudaSuccess) ok = false;
    for (int i = 0; ok && i < fmpc_handle::MAX_CHUNKS; ++i)
        if (cudaEventCreateWithFlags(&h->ev_in[i], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&h->ev_k[i], cudaEventDisableTiming) != cudaSuccess) ok = false;
    if (ok && cudaStreamCreateWithFlags(&h->s_gen, cudaStreamNonBlocking) != cudaSuccess) ok = false;
    for (int i = 0; ok && i < 2; ++i)
        if (cudaEventCreateWithFlags(&h->ev_nu_ready[i], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&h->ev_nu_free[i], cudaEventDisableTiming) != cudaSuccess) ok = false;
    for (int i = 0; ok && i < fmpc_handle::NSTAGE; ++i)
        if (cudaEventCreateWithFlags(&h->ev_stg_in[i], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&h->ev_stg_out[i], cudaEventDisableTiming) != cudaSuccess) ok = false;
    if (ok) {   // rand's generator as a fresh MATLAB session has it: MT19937 seeded with 5489, nothing drawn yet
        MT19937 g(5489u);
        if (h->d_mt.ensure(sizeof(g.mt)) || cudaMemcpy(h->d_mt.p, g.mt, sizeof(g.mt), cudaMemcpyHostToDevice) != cudaSuccess) ok = false;
        h->mt_idx = 624;
    }
    if (!ok) { fmpc_destroy(h); return FMPC_ERR_CUDA; }
    *out = h;
    return FMPC_OK;
}

int fmpc_seed_stream(fmpc_handle *h, unsigned seed)
{
    if (!h) return FMPC_ERR_NULL;
    CU_OK(cudaSetDevice(h->device));
    CU_OK(cudaStreamSynchronize(h->s_gen));         // doubles generated ahead from the old state are dropped
    MT19937 g(seed);
    CU_OK(cudaMemcpy(h->d_mt.p, g.mt, sizeof(g.mt), cudaMemcpyHostToDevice));
    h->mt_idx = 624;
    h->nu_have = 0;
    return FMPC_OK;
}

int fmpc_get_dims(const fmpc_handle *h, int *n, int *m, int *T)
{
    if (!h) return FMPC_ERR_NULL;
    if (n) *n = h->n;
    if (m) *m = h->m;
    if (T) *T = h->T;
    return FMPC_OK;
}

long long fmpc_workspace_bytes(const fmpc_handle *h) { return h ? (long long)h->ws.bytes : 0; }
long long fmpc_launch_count(const fmpc_handle *h) { return h ? h->launches : 0; }

int fmpc_kernel_kind(const fmpc_handle *h) { return h ? h->cfg.use_mma : -1; }

int fmpc_last_profile(fmpc_handle *h, long long *out12)
{
    if (!h || !out12) return FMPC_ERR_NULL;
    cudaSetDevice(h->device);
    if (cudaStreamSynchronize(h->stream) != cudaSuccess) return FMPC_ERR_CUDA;
    if (cudaMemcpy(out12, h->counters.as<char>() + 64, 12 * 8, cudaMemcpyDeviceToHost) != cudaSuccess) return FMPC_ERR_CUDA;
    return FMPC_OK;
}

long long fmpc_last_newton_iters(fmpc_handle *h)
{
    if (!h) return -1;
    cudaSetDevice(h->device);
    unsigned long long v = 0;
    if (cudaStreamSynchronize(h->stream) != cudaSuccess) return -1;
    for (int p = 0; p < fmpc_handle::MAX_PART; ++p)
        if (cudaStreamSynchronize(h->s_k[p]) != cudaSuccess) return -1;
    std::vector<char> blk(256 * fmpc_handle::MAX_CHUNKS);     // one counter block per (possibly concurrent) chunk launch: one copy
    if (cudaMemcpy(blk.data(), h->counters.p, blk.size(), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    for (int p = 0; p < fmpc_handle::MAX_CHUNKS; ++p) {
        unsigned long long vp = 0;
        std::memcpy(&vp, blk.data() + 256 * p + 8, 8);
        v += vp;
    }
    return (long long)v;
}

// The next `need` doubles of the handle's MATLAB stream, in device memory, readable by work enqueued on `consumer` after this
// call.  The doubles for the FOLLOWING call (assumed to need as many) are generated right away on s_gen into the other buffer,
// underneath the solve that consumes this one; whatever a call leaves unread is carried over, so the stream is consumed
// strictly in order whatever the batch sizes.  The consumer calls nu_release(which) once its last reader is enqueued.
static int nu_take(fmpc_handle *h, size_t need, cudaStream_t consumer, const double **ptr, int *which)
{
    const size_t cap = (size_t)h->max_batch * (size_t)(h->T + 1) * (size_t)h->n;
    if (need > cap) return FMPC_ERR_BATCH;
    if (h->nu_cap < cap) {
        if (h->nu_buf[0].ensure(cap * 8) || h->nu_buf[1].ensure(cap * 8) || h->d_mt_raw.ensure((2 * cap + 1248) * 4)) return FMPC_ERR_CUDA;
        h->nu_cap = cap;
    }
    const int cur = h->nu_cur, oth = cur ^ 1;
    double *B = h->nu_buf[cur].as<double>(), *O = h->nu_buf[oth].as<double>();
    if (h->nu_have < need) {
        if (h->nu_free_pending[cur]) { CU_OK(cudaStreamWaitEvent(h->s_gen, h->ev_nu_free[cur], 0)); h->nu_free_pending[cur] = false; }
        fmpc_launch_mt_fill(h->d_mt.as<unsigned>(), h->d_mt_raw.as<unsigned>(), &h->mt_idx, B + h->nu_have, need - h->nu_have, h->s_gen);
        CU_OK(cudaGetLastError());
        h->launches += 2;
        h->nu_have = need;
    }
    CU_OK(cudaEventRecord(h->ev_nu_ready[cur], h->s_gen));
    CU_OK(cudaStreamWaitEvent(consumer, h->ev_nu_ready[cur], 0));
    *ptr = B;
    *which = cur;
    const size_t left = h->nu_have - need;
    if (h->nu_free_pending[oth]) { CU_OK(cudaStreamWaitEvent(h->s_gen, h->ev_nu_free[oth], 0)); h->nu_free_pending[oth] = false; }
    if (left) CU_OK(cudaMemcpyAsync(O, B + need, left * 8, cudaMemcpyDeviceToDevice, h->s_gen));
    size_t have = left;
    if (left < need) {
        fmpc_launch_mt_fill(h->d_mt.as<unsigned>(), h->d_mt_raw.as<unsigned>(), &h->mt_idx, O + left, need - left, h->s_gen);
        CU_OK(cudaGetLastError());
        h->launches += 2;
        have = need;
    }
    h->nu_cur = oth;
    h->nu_have = have;
    h->nu_stream_active = true;
    return FMPC_OK;
}
static int nu_release(fmpc_handle *h, int which, cudaStream_t consumer)
{
    CU_OK(cudaEventRecord(h->ev_nu_free[which], consumer));
    h->nu_free_pending[which] = true;
    return FMPC_OK;
}

static int step_device(fmpc_handle *h, const fmpc_params *p, int nbatch, const double *x0, const double *x0_pre,
                       const double *u_prev, const double *w, const double *xf, const double *X0, const double *U0, const double *nu0,
                       double *X, double *U, int *status, int *iters, cudaStream_t st, bool keep_totals = false,
                       int part = 0, int nparts = 1, bool by_smid = false, int warm_shift = 0, double *u_first = nullptr)
{
    StepArgs A{};
    A.warm_shift = warm_shift; A.u_first = u_first;
    A.nbatch = nbatch; A.has_xf = xf ? 1 : 0; A.cold = (X0 == nullptr || U0 == nullptr) ? 1 : 0;
    A.kappa = p->kappa; A.niters = p->niters; A.ls_max = p->ls_max;
    A.alpha = p->alpha; A.beta = p->beta; A.tol_r = p->tol_r; A.tol_p = p->tol_p;
    A.x0 = x0; A.x0_pre = x0_pre; A.u_prev = u_prev; A.w = w; A.xf = xf; A.X0 = X0; A.U0 = U0; A.nu0 = nu0;
    A.X = X; A.U = U; A.status = status; A.iters = iters;
    // partition `part` of `nparts`: its own counter block, a disjoint range of CTA slots (scratch) and a grid of that size
    char *cblk = h->counters.as<char>() + 256 * part;
    A.counter = (unsigned int *)cblk;
    A.iters_total = (unsigned long long *)(cblk + 8);
    SolveLaunchCfg cfg = h->cfg;
    if (by_smid) A.slot_base = -1;                              // `part` only selects the counter block
    else if (nparts > 1) { cfg.grid = h->cfg.grid / nparts; A.slot_base = part * cfg.grid; }
    // While the handle's MATLAB stream is in use its generator (one CTA, fmpc_mt_fill_kernel) runs one call ahead, underneath
    // this solve.  The persistent solve kernels fill every SM, and a generator CTA that does not fit beside one of their CTAs
    // would wait for the whole solve (and the NEXT solve for it): leave it one SM.
    if (h->nu_stream_active && nparts == 1 && !by_smid && cfg.grid > 1 && (h->cfg.use_mma == 2 || h->cfg.use_mma == 1)) cfg.grid -= 1;
    A.ws = h->ws.as<double>(); A.ws_stride = h->ws_stride;
    A.prof = (long long *)(h->counters.as<char>() + 64);
    CU_OK(cudaMemsetAsync(cblk, 0, keep_totals ? 4 : 256, st));     // instance counter [+ iteration total, phase counters]
    if (!keep_totals && nparts == 1 && part == 0 && h->other_blocks_dirty) {   // a whole-grid launch starts a new total
        CU_OK(cudaMemsetAsync(h->counters.as<char>() + 256, 0, 256 * (fmpc_handle::MAX_CHUNKS - 1), st));
        h->other_blocks_dirty = false;
    }
    if (part > 0) h->other_blocks_dirty = true;
    if (h->cfg.use_mma == 3) {
        if (nbatch > h->max_batch) return FMPC_ERR_BATCH;     // its scratch is sized by max_batch
        fmpc_launch_solve_gen(h->S, h->G, A, h->cfg, st);
    } else if (h->cfg.use_mma == 2) fmpc_launch_solve_warp(h->S, A, cfg, st);
    else if (h->cfg.use_mma == 1) fmpc_launch_solve_mma(h->S, A, cfg, st);
    else fmpc_launch_solve(h->S, A, h->cfg, st);
    CU_OK(cudaGetLastError());
    h->launches += 1;
    return FMPC_OK;
}

int fmpc_step_d(fmpc_handle *h, const fmpc_params *p, int nbatch, const double *x0, const double *x0_pre,
                const double *u_prev, const double *w, const double *xf, const double *X0, const double *U0,
                const double *nu0, double *X, double *U, int *status, int *iters, void *stream)
{
    if (!h || !x0 || !X || !U || !nu0) return FMPC_ERR_NULL;
    if (h->ramp && !u_prev) return FMPC_ERR_U_BOUND_SIZE;
    int rc = validate_params(p);
    if (rc) return rc;
    if (nbatch < 0) return FMPC_ERR_DIM;
    if (h->var_order == 2 && !x0_pre) return FMPC_ERR_A_SIZE;
    if ((X0 == nullptr) != (U0 == nullptr)) return FMPC_ERR_INIT_SIZE;
    if (nbatch == 0) return FMPC_OK;
    CU_OK(cudaSetDevice(h->device));
    return step_device(h, p, nbatch, x0, x0_pre, u_prev, w, xf, X0, U0, nu0, X, U, status, iters,
                       stream ? (cudaStream_t)stream : h->stream);
}

int fmpc_step(fmpc_handle *h, const fmpc_params *p, int nbatch, const double *x0, const double *x0_pre,
              const double *u_prev, const double *w, const double *xf, const double *X0, const double *U0,
              const double *nu0, double *X, double *U, int *status, int *iters, double *telapsed)
{
    if (!h || !x0 || !X || !U) return FMPC_ERR_NULL;
    if (h->ramp && !u_prev) return FMPC_ERR_U_BOUND_SIZE;
    int rc = validate_params(p);
    if (rc) return rc;
    if (nbatch < 0) return FMPC_ERR_DIM;
    if (nbatch > h->max_batch) return FMPC_ERR_BATCH;
    if (h->var_order == 2 && !x0_pre) return FMPC_ERR_A_SIZE;
    if ((X0 == nullptr) != (U0 == nullptr)) return FMPC_ERR_INIT_SIZE;
    if (telapsed) *telapsed = 0.0;
    if (nbatch == 0) return FMPC_OK;
    CU_OK(cudaSetDevice(h->device));
    const int n = h->n, m = h->m, T = h->T;
    const size_t nb = (size_t)nbatch, NBn = (size_t)(T + (xf ? 1 : 0)) * n;
    cudaStream_t st = h->stream;
    if (h->d_x0.ensure(nb * n * 8) || h->d_x0pre.ensure(nb * n * 8) || h->d_w.ensure(nb * T * n * 8) || h->d_xf.ensure(nb * n * 8) ||
        h->d_uprev.ensure(nb * m * 8) || h->d_X.ensure(nb * n * T * 8) || h->d_U.ensure(nb * m * T * 8) || h->d_nu0.ensure(nb * NBn * 8) ||
        h->d_status.ensure(nb * 4) || h->d_iters.ensure(nb * 4))
        return FMPC_ERR_CUDA;
    // nu0 == NULL: MATLAB default stream, one rand(length(b),1) per instance, instance after instance -- generated on the device
    const double *nu_dev = nullptr;
    int nu_which = -1;
    // Instances are independent: the batch is cut into chunks, and copy-in (s_in), solve and copy-out (s_out) of successive
    // chunks overlap.  The warp kernel additionally splits the SMs (and its per-slot scratch) into NP partitions, each with its
    // own solve stream: chunk c runs as one wave on partition c % NP, so NP chunk kernels are resident side by side and the
    // pipeline fill / drain is a 1/NP-wave chunk instead of a whole wave.  With pageable host memory the copies degrade to
    // staged synchronous ones; the results are the same.
    // SM-id scratch slots: chunk kernels need no SM partitions -- each chunk is launched on the next of MAX_PART streams with
    // its own counter block and a grid of its own size, and the hardware packs the CTAs of overlapping chunks onto whatever
    // SMs are free (no idle tail between the chunks of a partition).
    const bool smid_mode = h->smid_slots && nb > (size_t)h->cfg.slots / 2 && !getenv("FMPC_STEP_PARTS");
    int NP = 1;
    // measured (profiles/r01_v3_7_e2e_partitions.log): 2 partitions beat 1 and 4 -- the host->device copies of a step
    // (133 MB at C2) take nearly as long as its solves, so finer chunks only add per-copy overhead
    if (h->cfg.use_mma == 2 && h->cfg.grid >= 4 * fmpc_handle::MAX_PART && nb > (size_t)h->cfg.slots / 2) NP = 2;
    if (const char *e = getenv("FMPC_STEP_PARTS")) { const int v = atoi(e); if (v >= 1 && v <= fmpc_handle::MAX_PART && h->cfg.use_mma == 2) NP = v; }
    const size_t wave = (NP > 1) ? (size_t)(h->cfg.grid / NP) * (size_t)(h->cfg.block / 32) : (size_t)h->cfg.slots;
    size_t per = wave;
    if (smid_mode) {
        per = (size_t)h->cfg.slots / 2;
        if (const char *e = getenv("FMPC_STEP_CHUNK")) { const long v = atol(e); if (v >= 32) per = (size_t)v; }
    }
    int nch = (int)((nb + per - 1) / per);
    if (NP == 1) { nch = (int)((nb + wave / 2) / wave); if (nch < 1) nch = 1; }
    while (nch > fmpc_handle::MAX_CHUNKS) { per += smid_mode ? (size_t)h->cfg.slots / 2 : wave; nch = (int)((nb + per - 1) / per); }
    if (NP == 1) per = (nb + nch - 1) / nch;
    cudaStream_t si = h->s_in, so = h->s_out;
    const size_t Tn = (size_t)T * n, Tm = (size_t)T * m;
    while (nch > 1 && (size_t)(nch - 1) * per >= nb) --nch;           // no empty trailing chunk
    if (smid_mode) CU_OK(cudaMemsetAsync(h->counters.p, 0, 256 * fmpc_handle::MAX_CHUNKS, si));   // every chunk's counter block, ahead of all chunks
    if (!nu0) { rc = nu_take(h, nb * NBn, si, &nu_dev, &nu_which); if (rc) return rc; }
    else h->nu_stream_active = false;
    // the per-instance vectors are small: one copy each for the whole batch, ahead of the chunked arrays
    CU_OK(cudaMemcpyAsync(h->d_x0.p, x0, nb * n * 8, cudaMemcpyHostToDevice, si));
    if (x0_pre) CU_OK(cudaMemcpyAsync(h->d_x0pre.p, x0_pre, nb * n * 8, cudaMemcpyHostToDevice, si));
    if (u_prev && h->ramp) CU_OK(cudaMemcpyAsync(h->d_uprev.p, u_prev, nb * m * 8, cudaMemcpyHostToDevice, si));
    if (xf) CU_OK(cudaMemcpyAsync(h->d_xf.p, xf, nb * n * 8, cudaMemcpyHostToDevice, si));
    // pageable caller buffers go through the pinned staging ring (see CopyPool); pinned / registered ones are copied directly
    const bool stage_in = !(dma_reachable(w) && dma_reachable(X0) && dma_reachable(U0) && dma_reachable(nu0));
    const bool stage_out = !(dma_reachable(X) && dma_reachable(U));
    constexpr int NSTG = fmpc_handle::NSTAGE;
    if (stage_in || stage_out) {
        if (!h->pool) {
            h->pool = new (std::nothrow) CopyPool();
            if (!h->pool) return FMPC_ERR_CUDA;
            int nt = (int)std::thread::hardware_concurrency() - 1;
            if (const char *e = getenv("FMPC_COPY_THREADS")) nt = atoi(e);
            h->pool->start(nt < 0 ? 0 : (nt > 7 ? 7 : nt));
        }
        const size_t in_b = per * ((w ? Tn : 0) + (X0 ? Tn + Tm : 0) + (nu0 ? NBn : 0)) * 8, out_b = per * (Tn + Tm) * 8;
        for (int i = 0; i < NSTG; ++i)
            if ((stage_in && h->stg_in[i].ensure(in_b)) || (stage_out && h->stg_out[i].ensure(out_b))) return FMPC_ERR_CUDA;
    }
    auto drain = [&](int cj) -> int {         // chunk cj: pinned out-slot -> the caller's X, U
        const size_t b0 = (size_t)cj * per, b1 = (b0 + per < nb) ? b0 + per : nb, cb = b1 - b0;
        CU_OK(cudaEventSynchronize(h->ev_stg_out[cj % NSTG]));
        char *sp = (char *)h->stg_out[cj % NSTG].p;
        h->pool->copy({{(char *)(X + b0 * Tn), sp, cb * Tn * 8}, {(char *)(U + b0 * Tm), sp + cb * Tn * 8, cb * Tm * 8}});
        return FMPC_OK;
    };
    // enqueue order per chunk: copy-in(ci), solve(ci), copy-out(ci-1) -- the host-side staging copies of one chunk run
    // underneath the solve of the previous one
    for (int ci = 0; ci <= nch; ++ci) {
        if (ci < nch) {
            const size_t b0 = (size_t)ci * per, b1 = (b0 + per < nb) ? b0 + per : nb, cb = b1 - b0;
            const int part = smid_mode ? ci : ci % NP;
            cudaStream_t sk = smid_mode ? h->s_k[ci % fmpc_handle::MAX_PART] : ((NP > 1) ? h->s_k[part] : st);
            const double *sw = w ? w + b0 * Tn : nullptr, *sX0 = X0 ? X0 + b0 * Tn : nullptr, *sU0 = U0 ? U0 + b0 * Tm : nullptr,
                         *snu = nu0 ? nu0 + b0 * NBn : nullptr;
            if (stage_in) {
                const int slot = ci % NSTG;
                if (ci >= NSTG) CU_OK(cudaEventSynchronize(h->ev_stg_in[slot]));      // its previous DMA is done
                char *sp = (char *)h->stg_in[slot].p;
                std::vector<CopyPool::Job> jobs;
                auto put = [&](const double *&src, size_t bytes) {
                    if (!src) return;
                    jobs.push_back({sp, (const char *)src, bytes});
                    src = (const double *)sp;
                    sp += bytes;
                };
                put(sw, cb * Tn * 8); put(sX0, cb * Tn * 8); put(sU0, cb * Tm * 8); put(snu, cb * NBn * 8);
                h->pool->copy(jobs);
            }
            if (w) CU_OK(cudaMemcpyAsync(h->d_w.as<double>() + b0 * Tn, sw, cb * Tn * 8, cudaMemcpyHostToDevice, si));
            if (X0) {
                CU_OK(cudaMemcpyAsync(h->d_X.as<double>() + b0 * Tn, sX0, cb * Tn * 8, cudaMemcpyHostToDevice, si));
                CU_OK(cudaMemcpyAsync(h->d_U.as<double>() + b0 * Tm, sU0, cb * Tm * 8, cudaMemcpyHostToDevice, si));
            }
            if (nu0) CU_OK(cudaMemcpyAsync(h->d_nu0.as<double>() + b0 * NBn, snu, cb * NBn * 8, cudaMemcpyHostToDevice, si));
            if (stage_in) CU_OK(cudaEventRecord(h->ev_stg_in[ci % NSTG], si));
            CU_OK(cudaEventRecord(h->ev_in[ci], si));
            CU_OK(cudaStreamWaitEvent(sk, h->ev_in[ci], 0));
            if (ci == 0) CU_OK(cudaEventRecord(h->ev0, sk));
            rc = step_device(h, p, (int)cb, h->d_x0.as<double>() + b0 * n, x0_pre ? h->d_x0pre.as<double>() + b0 * n : nullptr,
                             (u_prev && h->ramp) ? h->d_uprev.as<double>() + b0 * m : nullptr, w ? h->d_w.as<double>() + b0 * Tn : nullptr, xf ? h->d_xf.as<double>() + b0 * n : nullptr,
                             X0 ? h->d_X.as<double>() + b0 * Tn : nullptr, X0 ? h->d_U.as<double>() + b0 * Tm : nullptr,
                             (nu0 ? h->d_nu0.as<double>() : nu_dev) + b0 * NBn, h->d_X.as<double>() + b0 * Tn, h->d_U.as<double>() + b0 * Tm,
                             h->d_status.as<int>() + b0, h->d_iters.as<int>() + b0, sk, smid_mode ? true : ci >= NP, part, NP, smid_mode);
            if (rc) return rc;
            CU_OK(cudaEventRecord(h->ev_k[ci], sk));
            if (NP > 1 || smid_mode) CU_OK(cudaStreamWaitEvent(st, h->ev_k[ci], 0));      // st collects every chunk (ev1, final sync)
        }
        if (ci >= 1) {
            const int cj = ci - 1;
            const size_t b0 = (size_t)cj * per, b1 = (b0 + per < nb) ? b0 + per : nb, cb = b1 - b0;
            CU_OK(cudaStreamWaitEvent(so, h->ev_k[cj], 0));
            double *dX = X + b0 * Tn, *dU = U + b0 * Tm;
            if (stage_out) { dX = (double *)h->stg_out[cj % NSTG].p; dU = dX + cb * Tn; }
            CU_OK(cudaMemcpyAsync(dX, h->d_X.as<double>() + b0 * Tn, cb * Tn * 8, cudaMemcpyDeviceToHost, so));
            CU_OK(cudaMemcpyAsync(dU, h->d_U.as<double>() + b0 * Tm, cb * Tm * 8, cudaMemcpyDeviceToHost, so));
            if (stage_out) {
                CU_OK(cudaEventRecord(h->ev_stg_out[cj % NSTG], so));
                if (cj >= 1) { rc = drain(cj - 1); if (rc) return rc; }
            }
        }
    }
    if (stage_out) { rc = drain(nch - 1); if (rc) return rc; }
    if (status) CU_OK(cudaMemcpyAsync(status, h->d_status.p, nb * 4, cudaMemcpyDeviceToHost, so));
    if (iters) CU_OK(cudaMemcpyAsync(iters, h->d_iters.p, nb * 4, cudaMemcpyDeviceToHost, so));
    CU_OK(cudaEventRecord(h->ev1, st));
    if (nu_which >= 0) { rc = nu_release(h, nu_which, st); if (rc) return rc; }
    CU_OK(cudaStreamSynchronize(so));
    CU_OK(cudaStreamSynchronize(st));
    CU_OK(cudaStreamSynchronize(si));
    if (telapsed) { float ms = 0.f; CU_OK(cudaEventElapsedTime(&ms, h->ev0, h->ev1)); *telapsed = ms * 1e-3; }
    return FMPC_OK;
}

// Shared body of fmpc_step_r (host buffers, blocking) and fmpc_step_r_d (device buffers, asynchronous on `st`): every copy is
// cudaMemcpyDefault, so the same code moves host or device operands.
static int step_resident(fmpc_handle *h, const fmpc_params *p, int nbatch, int flags, const double *x0, const double *x0_pre,
                         const double *u_prev, const double *w, const double *xf, const double *nu0,
                         double *u0, double *X, double *U, int *status, int *iters, double *telapsed, cudaStream_t st, bool blocking)
{
    if (!h || !x0 || !u0) return FMPC_ERR_NULL;
    int rc = validate_params(p);
    if (rc) return rc;
    if (nbatch < 0) return FMPC_ERR_DIM;
    if (nbatch > h->max_batch) return FMPC_ERR_BATCH;
    if ((X == nullptr) != (U == nullptr)) return FMPC_ERR_NULL;
    const bool reset = (flags & FMPC_R_RESET) != 0;
    if (!reset && h->r_nb != nbatch) return FMPC_ERR_INIT_SIZE;      // no resident solution of that size to warm-start from
    if (telapsed) *telapsed = 0.0;
    if (nbatch == 0) return FMPC_OK;
    CU_OK(cudaSetDevice(h->device));
    const int n = h->n, m = h->m, T = h->T;
    const size_t nb = (size_t)nbatch, Tn = (size_t)T * n, Tm = (size_t)T * m, NBn = (size_t)(T + (xf ? 1 : 0)) * n;
    if (h->r_X.ensure(nb * Tn * 8) || h->r_U.ensure(nb * Tm * 8) || h->r_x0.ensure(nb * n * 8) || h->r_x0pre.ensure(nb * n * 8) ||
        h->r_u0.ensure(nb * m * 8) || h->d_status.ensure(nb * 4) || h->d_iters.ensure(nb * 4) ||
        (w && h->d_w.ensure(nb * Tn * 8)) || (xf && h->d_xf.ensure(nb * n * 8)) || (u_prev && h->d_uprev.ensure(nb * m * 8)) ||
        (nu0 && h->d_nu0.ensure(nb * NBn * 8)))
        return FMPC_ERR_CUDA;
    // x0_pre = the x0 of the previous call (README.md:483-488: zeros at the first step) unless the caller passes one
    std::swap(h->r_x0, h->r_x0pre);
    if (reset) {
        h->r_nb = 0;
        CU_OK(cudaMemsetAsync(h->r_x0pre.p, 0, nb * n * 8, st));
        CU_OK(cudaMemsetAsync(h->r_u0.p, 0, nb * m * 8, st));        // u_prev of the first step (README.md:447-452: no ramp shift yet)
    }
    CU_OK(cudaMemcpyAsync(h->r_x0.p, x0, nb * n * 8, cudaMemcpyDefault, st));
    if (x0_pre) CU_OK(cudaMemcpyAsync(h->r_x0pre.p, x0_pre, nb * n * 8, cudaMemcpyDefault, st));
    const double *d_up = nullptr;
    if (h->ramp) {       // ramp rows need the input applied last: the caller's, else U(:,0) of the previous call
        if (u_prev) {
            CU_OK(cudaMemcpyAsync(h->d_uprev.p, u_prev, nb * m * 8, cudaMemcpyDefault, st));
            d_up = h->d_uprev.as<double>();
        } else {
            if (h->d_uprev.ensure(nb * m * 8)) return FMPC_ERR_CUDA;
            CU_OK(cudaMemcpyAsync(h->d_uprev.p, h->r_u0.p, nb * m * 8, cudaMemcpyDeviceToDevice, st));
            d_up = h->d_uprev.as<double>();
        }
    }
    if (w) CU_OK(cudaMemcpyAsync(h->d_w.p, w, nb * Tn * 8, cudaMemcpyDefault, st));
    if (xf) CU_OK(cudaMemcpyAsync(h->d_xf.p, xf, nb * n * 8, cudaMemcpyDefault, st));
    const double *nu_dev = nullptr;
    int nu_which = -1;
    if (nu0) {
        h->nu_stream_active = false;
        CU_OK(cudaMemcpyAsync(h->d_nu0.p, nu0, nb * NBn * 8, cudaMemcpyDefault, st));
        nu_dev = h->d_nu0.as<double>();
    } else {
        rc = nu_take(h, nb * NBn, st, &nu_dev, &nu_which);
        if (rc) return rc;
    }
    const bool fused = (h->cfg.use_mma == 2);       // the warp kernel reads the shifted warm start and writes U(:,0) itself
    if (!reset && !fused) {
        fmpc_launch_shift_inplace(n, m, T, nbatch, h->r_X.as<double>(), h->r_U.as<double>(), st);
        CU_OK(cudaGetLastError());
        h->launches += 1;
    }
    if (blocking) CU_OK(cudaEventRecord(h->ev0, st));
    rc = step_device(h, p, nbatch, h->r_x0.as<double>(), h->var_order == 2 ? h->r_x0pre.as<double>() : nullptr, d_up,
                     w ? h->d_w.as<double>() : nullptr, xf ? h->d_xf.as<double>() : nullptr,
                     reset ? nullptr : h->r_X.as<double>(), reset ? nullptr : h->r_U.as<double>(), nu_dev,
                     h->r_X.as<double>(), h->r_U.as<double>(), h->d_status.as<int>(), h->d_iters.as<int>(), st,
                     false, 0, 1, false, (fused && !reset) ? 1 : 0, fused ? h->r_u0.as<double>() : nullptr);
    if (rc) return rc;
    if (blocking) CU_OK(cudaEventRecord(h->ev1, st));
    if (nu_which >= 0) { rc = nu_release(h, nu_which, st); if (rc) return rc; }
    if (!fused) {
        fmpc_launch_extract_first(m, T, nbatch, h->r_U.as<double>(), h->r_u0.as<double>(), st);
        CU_OK(cudaGetLastError());
        h->launches += 1;
    }
    CU_OK(cudaMemcpyAsync(u0, h->r_u0.p, nb * m * 8, cudaMemcpyDefault, st));
    if (status) CU_OK(cudaMemcpyAsync(status, h->d_status.p, nb * 4, cudaMemcpyDefault, st));
    if (iters) CU_OK(cudaMemcpyAsync(iters, h->d_iters.p, nb * 4, cudaMemcpyDefault, st));
    if (X) {
        CU_OK(cudaMemcpyAsync(X, h->r_X.p, nb * Tn * 8, cudaMemcpyDefault, st));
        CU_OK(cudaMemcpyAsync(U, h->r_U.p, nb * Tm * 8, cudaMemcpyDefault, st));
    }
    h->r_nb = nbatch;
    if (!blocking) return FMPC_OK;
    CU_OK(cudaStreamSynchronize(st));
    if (telapsed) { float ms = 0.f; CU_OK(cudaEventElapsedTime(&ms, h->ev0, h->ev1)); *telapsed = ms * 1e-3; }
    return FMPC_OK;
}

int fmpc_step_r(fmpc_handle *h, const fmpc_params *p, int nbatch, int flags, const double *x0, const double *x0_pre,
                const double *u_prev, const double *w, const double *xf, const double *nu0,
                double *u0, double *X, double *U, int *status, int *iters, double *telapsed)
{
    if (!h) return FMPC_ERR_NULL;
    return step_resident(h, p, nbatch, flags, x0, x0_pre, u_prev, w, xf, nu0, u0, X, U, status, iters, telapsed, h->stream, true);
}

int fmpc_step_r_d(fmpc_handle *h, const fmpc_params *p, int nbatch, int flags, const double *x0, const double *x0_pre,
                  const double *u_prev, const double *w, const double *xf, const double *nu0,
                  double *u0, double *X, double *U, int *status, int *iters, void *stream)
{
    if (!h) return FMPC_ERR_NULL;
    return step_resident(h, p, nbatch, flags, x0, x0_pre, u_prev, w, xf, nu0, u0, X, U, status, iters, nullptr,
                         stream ? (cudaStream_t)stream : h->stream, false);
}

int fmpc_step_z(fmpc_handle *h, const fmpc_params *p, int nbatch, const double *x0, const double *x0_pre,
                const double *u_prev, const double *w, const double *xf, const double *z0, const double *nu0,
                double *z, int *status, int *iters, double *telapsed)
{
    if (!h || !z) return FMPC_ERR_NULL;
    if (nbatch < 0) return FMPC_ERR_DIM;
    const int n = h->n, m = h->m, T = h->T;
    const size_t nb = (size_t)nbatch;
    std::vector<double> X(nb * n * T), U(nb * m * T), X0, U0;
    if (z0) {       // de-interleave, README.md:558-570
        X0.resize(nb * n * T); U0.resize(nb * m * T);
        for (size_t b = 0; b < nb; ++b)
            for (int t = 0; t < T; ++t) {
                const double *zs = z0 + b * (size_t)T * (n + m) + (size_t)t * (n + m);
                std::memcpy(&U0[b * m * T + (size_t)t * m], zs, m * 8);
                std::memcpy(&X0[b * n * T + (size_t)t * n], zs + m, n * 8);
            }
    }
    int rc = fmpc_step(h, p, nbatch, x0, x0_pre, u_prev, w, xf, z0 ? X0.data() : nullptr, z0 ? U0.data() : nullptr, nu0,
                       X.data(), U.data(), status, iters, telapsed);
    if (rc) return rc;
    for (size_t b = 0; b < nb; ++b)
        for (int t = 0; t < T; ++t) {
            double *zs = z + b * (size_t)T * (n + m) + (size_t)t * (n + m);
            std::memcpy(zs, &U[b * m * T + (size_t)t * m], m * 8);
            std::memcpy(zs + m, &X[b * n * T + (size_t)t * n], n * 8);
        }
    return FMPC_OK;
}

int fmpc_frontend_nouter(const fmpc_handle *h, int mode)
{
    if (!h) return FMPC_ERR_NULL;
    switch (mode) {
    case FMPC_FE_FIXED_LOG: return 1;
    case FMPC_FE_SOLVE_CHECK: return 5;                      // linspace(k_max,k_min,5), Fast_MPC2.m:89
    case FMPC_FE_FIXED_NEWTON:
    case FMPC_FE_SOLVE_FULL: {                               // k = 1; while k*length(z) >= 10e-3: k = k/10
        const double N = (double)h->T * (h->n + h->m);
        int cnt = 0;
        double k = 1.0;
        const double mu = 1.0 / 10;
        while (k * N >= 10e-3) { ++cnt; k = mu * k; }
        return cnt;
    }
    default: return FMPC_ERR_PARAM;
    }
}

int fmpc_frontend(fmpc_handle *h, int mode, const fmpc_params *p, double k_min, double k_max, int nbatch,
                  const double *x0, const double *x0_pre, const double *u_prev, const double *w, const double *xf,
                  const double *X0, const double *U0, const double *nu0, double *X, double *U, int *status, int *iters,
                  double *telapsed)
{
    if (!h || !p) return FMPC_ERR_NULL;
    const int nouter = fmpc_frontend_nouter(h, mode);
    if (nouter < 0) return nouter;
    const size_t NBn = (size_t)(h->T + (xf ? 1 : 0)) * h->n;
    fmpc_params q = *p;
    if (mode == FMPC_FE_FIXED_LOG || mode == FMPC_FE_SOLVE_FULL || mode == FMPC_FE_SOLVE_CHECK) q.niters = 1000;   // nw = []
    double k = (mode == FMPC_FE_FIXED_LOG) ? p->kappa : 1.0;
    const double mu = 1.0 / 10;
    double tsum = 0.0;
    std::vector<int> it_acc;
    if (iters) it_acc.assign((size_t)nbatch, 0);
    const double *Xs = X0, *Us = U0;
    for (int o = 0; o < nouter; ++o) {
        if (mode == FMPC_FE_SOLVE_CHECK) k = k_max + (k_min - k_max) * (double)o / 4.0;   // linspace(k_max,k_min,5)
        q.kappa = k;
        double te = 0.0;
        int rc = fmpc_step(h, &q, nbatch, x0, x0_pre, u_prev, w, xf, Xs, Us, nu0 ? nu0 + (size_t)o * NBn * nbatch : nullptr,
                           X, U, status, iters, &te);
        if (rc) return rc;
        tsum += te;
        if (iters) for (int b = 0; b < nbatch; ++b) it_acc[b] += iters[b];
        Xs = X; Us = U;                                       // z = x_opt
        if (mode == FMPC_FE_FIXED_NEWTON || mode == FMPC_FE_SOLVE_FULL) k = mu * k;
    }
    if (iters) for (int b = 0; b < nbatch; ++b) iters[b] = it_acc[b];
    if (telapsed) *telapsed = tsum;
    return FMPC_OK;
}

int fmpc_state_update_d(fmpc_handle *h, int nbatch, const double *x, const double *x_pre, const double *u, const double *w,
                        double *x_next, void *stream)
{
    if (!h || !x || !u || !x_next) return FMPC_ERR_NULL;
    if (h->var_order == 2 && !x_pre) return FMPC_ERR_A_SIZE;
    if (nbatch <= 0) return nbatch < 0 ? FMPC_ERR_DIM : FMPC_OK;
    CU_OK(cudaSetDevice(h->device));
    fmpc_launch_state_update(h->S, nbatch, x, x_pre, u, w, x_next, stream ? stream : (void *)h->stream);
    CU_OK(cudaGetLastError());
    h->launches += 1;
    return FMPC_OK;
}

int fmpc_state_update(fmpc_handle *h, int nbatch, const double *x, const double *x_pre, const double *u, const double *w,
                      double *x_next)
{
    if (!h || !x || !u || !x_next) return FMPC_ERR_NULL;
    if (h->var_order == 2 && !x_pre) return FMPC_ERR_A_SIZE;
    if (nbatch <= 0) return nbatch < 0 ? FMPC_ERR_DIM : FMPC_OK;
    CU_OK(cudaSetDevice(h->device));
    const size_t nb = (size_t)nbatch, n = h->n, m = h->m;
    // staging: x -> d_x0, x_pre -> d_x0pre, u -> d_uprev, w -> d_xf, out -> d_w
    if (h->d_x0.ensure(nb * n * 8) || h->d_x0pre.ensure(nb * n * 8) || h->d_uprev.ensure(nb * m * 8) || h->d_xf.ensure(nb * n * 8) ||
        h->d_w.ensure(nb * n * 8))
        return FMPC_ERR_CUDA;
    cudaStream_t st = h->stream;
    CU_OK(cudaMemcpyAsync(h->d_x0.p, x, nb * n * 8, cudaMemcpyHostToDevice, st));
    if (x_pre) CU_OK(cudaMemcpyAsync(h->d_x0pre.p, x_pre, nb * n * 8, cudaMemcpyHostToDevice, st));
    CU_OK(cudaMemcpyAsync(h->d_uprev.p, u, nb * m * 8, cudaMemcpyHostToDevice, st));
    if (w) CU_OK(cudaMemcpyAsync(h->d_xf.p, w, nb * n * 8, cudaMemcpyHostToDevice, st));
    int rc = fmpc_state_update_d(h, nbatch, h->d_x0.as<double>(), x_pre ? h->d_x0pre.as<double>() : nullptr, h->d_uprev.as<double>(),
                                 w ? h->d_xf.as<double>() : nullptr, h->d_w.as<double>(), st);
    if (rc) return rc;
    CU_OK(cudaMemcpyAsync(x_next, h->d_w.p, nb * n * 8, cudaMemcpyDeviceToHost, st));
    CU_OK(cudaStreamSynchronize(st));
    return FMPC_OK;
}

int fmpc_closed_loop(fmpc_handle *h, const fmpc_params *p, int nbatch, int K, const double *a, const double *nu0,
                     double *U_acc, double *X_acc, int *iters_acc, double *telapsed)
{
    if (!h || !a || !U_acc || !X_acc) return FMPC_ERR_NULL;
    int rc = validate_params(p);
    if (rc) return rc;
    if (nbatch < 0 || K < 0) return FMPC_ERR_DIM;
    if (nbatch > h->max_batch) return FMPC_ERR_BATCH;
    if (telapsed) *telapsed = 0.0;
    if (nbatch == 0 || K == 0) return FMPC_OK;
    CU_OK(cudaSetDevice(h->device));
    const size_t nb = (size_t)nbatch, n = h->n, m = h->m, T = h->T, NBn = T * n;
    cudaStream_t st = h->stream;
    if (h->d_x0.ensure(nb * n * 8) || h->d_x0pre.ensure(nb * n * 8) || h->d_uprev.ensure(nb * m * 8) ||
        h->d_X.ensure(nb * n * T * 8) || h->d_U.ensure(nb * m * T * 8) || h->d_nu0.ensure(2 * nb * NBn * 8) ||
        h->d_status.ensure(nb * 4) || h->d_iters.ensure(nb * 4) || h->d_a.ensure(nb * n * K * 8) ||
        h->d_Uacc.ensure(nb * m * K * 8) || h->d_Xacc.ensure(nb * n * K * 8) || h->d_itacc.ensure(nb * K * 4))
        return FMPC_ERR_CUDA;
    cudaStream_t si = h->s_in;
    CU_OK(cudaMemcpyAsync(h->d_a.p, a, nb * n * K * 8, cudaMemcpyHostToDevice, st));
    CU_OK(cudaMemsetAsync(h->d_x0.p, 0, nb * n * 8, st));
    CU_OK(cudaEventRecord(h->ev0, st));
    // The dual start of step k + 1 is uploaded on the copy stream (two device buffers) while the solve of step k runs: with
    // pageable host memory that copy blocks the calling thread, not the GPU.  ev_in[b]: buffer b filled; ev_k[b]: buffer b consumed.
    auto upload_nu = [&](int k) -> int {        // explicit dual starts: step k's block goes up while step k - 1 solves
        const int buf = k & 1;
        if (k >= 2) CU_OK(cudaStreamWaitEvent(si, h->ev_k[buf], 0));       // the solve of step k - 2 has read this buffer
        CU_OK(cudaMemcpyAsync(h->d_nu0.as<double>() + (size_t)buf * nb * NBn, nu0 + (size_t)k * nb * NBn, nb * NBn * 8, cudaMemcpyHostToDevice, si));
        CU_OK(cudaEventRecord(h->ev_in[buf], si));
        return FMPC_OK;
    };
    if (nu0) { h->nu_stream_active = false; rc = upload_nu(0); if (rc) return rc; }
    for (int k = 0; k < K; ++k) {
        const int buf = k & 1;
        // x0 = a[:,k,b] + B u_prev ; x0_pre <- previous x0 ; warm start shifted one stage
        fmpc_launch_shift_warm(h->S, nbatch, h->d_a.as<double>() + (size_t)k * n, (int)(n * K), h->d_X.as<double>(),
                               h->d_U.as<double>(), h->d_x0.as<double>(), h->d_x0pre.as<double>(), h->d_uprev.as<double>(),
                               k == 0, st);
        CU_OK(cudaGetLastError());
        h->launches += 1;
        const double *nu_k = nullptr;
        int nu_which = -1;
        if (nu0) { CU_OK(cudaStreamWaitEvent(st, h->ev_in[buf], 0)); nu_k = h->d_nu0.as<double>() + (size_t)buf * nb * NBn; }
        else {      // MATLAB default stream, one rand(length(b),1) per instance and step: the device generator runs one step ahead
            rc = nu_take(h, nb * NBn, st, &nu_k, &nu_which);
            if (rc) return rc;
        }
        rc = step_device(h, p, nbatch, h->d_x0.as<double>(), h->d_x0pre.as<double>(), h->d_uprev.as<double>(), nullptr, nullptr,
                         k == 0 ? nullptr : h->d_X.as<double>(), k == 0 ? nullptr : h->d_U.as<double>(), nu_k,
                         h->d_X.as<double>(), h->d_U.as<double>(), h->d_status.as<int>(), h->d_iters.as<int>(), st);
        if (rc) return rc;
        if (nu0) CU_OK(cudaEventRecord(h->ev_k[buf], st));
        else { rc = nu_release(h, nu_which, st); if (rc) return rc; }
        // logs: U_acc[:,k,b] = U(:,0,b), X_acc[:,k,b] = x0, iters
        fmpc_launch_log_step((int)n, (int)m, (int)T, nbatch, K, k, h->d_U.as<double>(), h->d_x0.as<double>(), h->d_iters.as<int>(),
                             h->d_Uacc.as<double>(), h->d_Xacc.as<double>(), h->d_itacc.as<int>(), st);
        CU_OK(cudaGetLastError());
        h->launches += 1;
        if (nu0 && k + 1 < K) { rc = upload_nu(k + 1); if (rc) return rc; }
    }
    CU_OK(cudaEventRecord(h->ev1, st));
    CU_OK(cudaMemcpyAsync(U_acc, h->d_Uacc.p, nb * m * K * 8, cudaMemcpyDeviceToHost, st));
    CU_OK(cudaMemcpyAsync(X_acc, h->d_Xacc.p, nb * n * K * 8, cudaMemcpyDeviceToHost, st));
    if (iters_acc) CU_OK(cudaMemcpyAsync(iters_acc, h->d_itacc.p, nb * K * 4, cudaMemcpyDeviceToHost, st));
    CU_OK(cudaStreamSynchronize(st));
    if (telapsed) { float ms = 0.f; CU_OK(cudaEventElapsedTime(&ms, h->ev0, h->ev1)); *telapsed = ms * 1e-3; }
    return FMPC_OK;
}

} // extern "C"
