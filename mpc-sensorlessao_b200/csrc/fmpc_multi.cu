// Multi-GPU inside the boundary: ONE blocking call drives every B200 of the box, so the MATLAB call site
// (README.md:548-555, one synchronous call per control step) reaches all of them without any host-side change.
// One fmpc_handle + one persistent host thread per device in a single process; the batch is cut into contiguous shards of
// ceil(nb / G) instances; the solve needs NO inter-GPU traffic (instances are independent).  The only collective is the
// optional gather of a small per-device statistics record over NCCL (ncclCommInitAll + ncclAllGather), loaded with dlopen so
// that the library itself does not depend on NCCL.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <functional>
#include <mutex>
#include <new>
#include <thread>
#include <vector>
#include "../../include/fmpc.h"

namespace {

struct Worker {
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    std::function<int()> job;
    bool has_job = false, done = true, stop = false;
    int rc = 0;
    void start()
    {
        th = std::thread([this] {
            std::unique_lock<std::mutex> lk(mu);
            for (;;) {
                cv.wait(lk, [this] { return stop || has_job; });
                if (stop) return;
                std::function<int()> j = std::move(job);
                has_job = false;
                lk.unlock();
                const int r = j();
                lk.lock();
                rc = r;
                done = true;
                cv.notify_all();
            }
        });
    }
    void submit(std::function<int()> j)
    {
        std::lock_guard<std::mutex> lk(mu);
        job = std::move(j);
        has_job = true;
        done = false;
        cv.notify_all();
    }
    int wait()
    {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [this] { return done; });
        return rc;
    }
    void shutdown()
    {
        { std::lock_guard<std::mutex> lk(mu); stop = true; }
        cv.notify_all();
        if (th.joinable()) th.join();
    }
};

// the five NCCL entry points the statistics gather needs, resolved at run time
struct Nccl {
    void *lib = nullptr;
    int (*CommInitAll)(void **, int, const int *) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, void *, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    bool load()
    {
        if (lib) return true;
        lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
        if (!lib) return false;
        CommInitAll = (int (*)(void **, int, const int *))dlsym(lib, "ncclCommInitAll");
        AllGather = (int (*)(const void *, void *, size_t, int, void *, cudaStream_t))dlsym(lib, "ncclAllGather");
        GroupStart = (int (*)())dlsym(lib, "ncclGroupStart");
        GroupEnd = (int (*)())dlsym(lib, "ncclGroupEnd");
        CommDestroy = (int (*)(void *))dlsym(lib, "ncclCommDestroy");
        return CommInitAll && AllGather && GroupStart && GroupEnd && CommDestroy;
    }
};

} // namespace

struct fmpc_multi {
    int G = 0, n = 0, m = 0, T = 0, max_batch = 0, per = 0;
    std::vector<int> dev;
    std::vector<fmpc_handle *> h;
    std::vector<Worker *> w;
    std::vector<fmpc_multi_stats> last;         // per device, of the last step
    Nccl nccl;
    std::vector<void *> comms;
    std::vector<double *> d_send, d_recv;
    std::vector<cudaStream_t> cs;
};

extern "C" {

void fmpc_multi_destroy(fmpc_multi *M)
{
    if (!M) return;
    for (Worker *w : M->w) { if (w) { w->shutdown(); delete w; } }
    for (size_t g = 0; g < M->comms.size(); ++g) {
        cudaSetDevice(M->dev[g]);
        if (M->comms[g] && M->nccl.CommDestroy) M->nccl.CommDestroy(M->comms[g]);
        if (g < M->d_send.size() && M->d_send[g]) cudaFree(M->d_send[g]);
        if (g < M->d_recv.size() && M->d_recv[g]) cudaFree(M->d_recv[g]);
        if (g < M->cs.size() && M->cs[g]) cudaStreamDestroy(M->cs[g]);
    }
    for (fmpc_handle *h : M->h) fmpc_destroy(h);
    delete M;
}

int fmpc_multi_create(fmpc_multi **out, const fmpc_sys *sys, int max_batch, int ngpus, const int *devices)
{
    if (!out) return FMPC_ERR_NULL;
    *out = nullptr;
    if (max_batch < 1) return FMPC_ERR_DIM;
    const int avail = fmpc_device_count();
    if (avail < 1) return FMPC_ERR_CUDA;
    if (ngpus <= 0) ngpus = avail;
    if (!devices && ngpus > avail) return FMPC_ERR_CUDA;
    fmpc_multi *M = new (std::nothrow) fmpc_multi();
    if (!M) return FMPC_ERR_CUDA;
    M->G = ngpus; M->max_batch = max_batch;
    M->per = (max_batch + ngpus - 1) / ngpus;
    for (int g = 0; g < ngpus; ++g) M->dev.push_back(devices ? devices[g] : g);
    M->h.assign(ngpus, nullptr);
    M->last.assign(ngpus, fmpc_multi_stats{});
    // the handles are created in parallel too: each uploads the constants and allocates its scratch on its own device
    for (int g = 0; g < ngpus; ++g) { Worker *w = new (std::nothrow) Worker(); if (!w) { fmpc_multi_destroy(M); return FMPC_ERR_CUDA; } w->start(); M->w.push_back(w); }
    for (int g = 0; g < ngpus; ++g) {
        fmpc_handle **hp = &M->h[g];
        const int d = M->dev[g], mb = M->per;
        M->w[g]->submit([hp, sys, mb, d, g] {
            int rc = fmpc_create(hp, sys, mb, d);
            // nu0 == NULL draws from one MT19937 stream PER DEVICE: a single sequential stream cannot feed several GPUs at
            // their solve rate.  Device 0 keeps MATLAB's default stream (seed 5489), device g is seeded 5489 + g.
            if (rc == FMPC_OK && g > 0) rc = fmpc_seed_stream(*hp, 5489u + (unsigned)g);
            return rc;
        });
    }
    int rc = FMPC_OK;
    for (int g = 0; g < ngpus; ++g) { const int r = M->w[g]->wait(); if (r && !rc) rc = r; }
    if (rc) { fmpc_multi_destroy(M); return rc; }
    fmpc_get_dims(M->h[0], &M->n, &M->m, &M->T);
    *out = M;
    return FMPC_OK;
}

int fmpc_multi_ngpus(const fmpc_multi *M) { return M ? M->G : 0; }
int fmpc_multi_shard(const fmpc_multi *M, int nbatch, int g, int *first, int *count)
{
    if (!M || g < 0 || g >= M->G) return FMPC_ERR_DIM;
    const int per = (nbatch + M->G - 1) / M->G;
    int b0 = g * per, b1 = b0 + per;
    if (b0 > nbatch) b0 = nbatch;
    if (b1 > nbatch) b1 = nbatch;
    if (first) *first = b0;
    if (count) *count = b1 - b0;
    return FMPC_OK;
}
fmpc_handle *fmpc_multi_handle(fmpc_multi *M, int g) { return (M && g >= 0 && g < M->G) ? M->h[g] : nullptr; }

static void fill_stats(fmpc_multi *M, int g, int cb, const int *status, double te, const int *iters = nullptr)
{
    fmpc_multi_stats &s = M->last[g];
    std::memset(&s, 0, sizeof s);
    s.device = M->dev[g];
    s.n_solves = cb;
    s.device_seconds = te;
    if (iters) { double t = 0.0; for (int b = 0; b < cb; ++b) t += iters[b]; s.newton_iters = t; }     // no device round trip
    else s.newton_iters = (double)fmpc_last_newton_iters(M->h[g]);
    if (status) for (int b = 0; b < cb; ++b) { const int v = status[b]; if (v >= 0 && v < 5) s.status_hist[v < 3 ? v : 3] += 1.0; }
}

int fmpc_multi_step(fmpc_multi *M, const fmpc_params *p, int nbatch, const double *x0, const double *x0_pre, const double *u_prev,
                    const double *w, const double *xf, const double *X0, const double *U0, const double *nu0,
                    double *X, double *U, int *status, int *iters, double *telapsed)
{
    if (!M || !x0 || !X || !U) return FMPC_ERR_NULL;
    if (nbatch < 0) return FMPC_ERR_DIM;
    if (nbatch > M->max_batch) return FMPC_ERR_BATCH;
    if (telapsed) *telapsed = 0.0;
    const size_t n = M->n, m = M->m, T = M->T, NBn = (T + (xf ? 1 : 0)) * n;
    std::vector<double> te(M->G, 0.0);
    for (int g = 0; g < M->G; ++g) {
        int b0 = 0, cb = 0;
        fmpc_multi_shard(M, nbatch, g, &b0, &cb);
        double *tep = &te[g];
        M->w[g]->submit([=] {
            if (cb == 0) { fill_stats(M, g, 0, nullptr, 0.0); return (int)FMPC_OK; }
            auto at = [b0](const double *a, size_t stride) { return a ? a + (size_t)b0 * stride : nullptr; };
            const int rc = fmpc_step(M->h[g], p, cb, at(x0, n), at(x0_pre, n), at(u_prev, m), at(w, T * n), at(xf, n), at(X0, T * n),
                                     at(U0, T * m), at(nu0, NBn), X + (size_t)b0 * T * n, U + (size_t)b0 * T * m,
                                     status ? status + b0 : nullptr, iters ? iters + b0 : nullptr, tep);
            if (rc == FMPC_OK) fill_stats(M, g, cb, status ? status + b0 : nullptr, *tep, iters ? iters + b0 : nullptr);
            return rc;
        });
    }
    int rc = FMPC_OK;
    for (int g = 0; g < M->G; ++g) { const int r = M->w[g]->wait(); if (r && !rc) rc = r; }
    if (telapsed) for (int g = 0; g < M->G; ++g) if (te[g] > *telapsed) *telapsed = te[g];
    return rc;
}

int fmpc_multi_step_r(fmpc_multi *M, const fmpc_params *p, int nbatch, int flags, const double *x0, const double *x0_pre,
                      const double *u_prev, const double *w, const double *xf, const double *nu0,
                      double *u0, double *X, double *U, int *status, int *iters, double *telapsed)
{
    if (!M || !x0 || !u0) return FMPC_ERR_NULL;
    if (nbatch < 0) return FMPC_ERR_DIM;
    if (nbatch > M->max_batch) return FMPC_ERR_BATCH;
    if (telapsed) *telapsed = 0.0;
    const size_t n = M->n, m = M->m, T = M->T, NBn = (T + (xf ? 1 : 0)) * n;
    std::vector<double> te(M->G, 0.0);
    for (int g = 0; g < M->G; ++g) {
        int b0 = 0, cb = 0;
        fmpc_multi_shard(M, nbatch, g, &b0, &cb);
        double *tep = &te[g];
        M->w[g]->submit([=] {
            if (cb == 0) { fill_stats(M, g, 0, nullptr, 0.0); return (int)FMPC_OK; }
            auto at = [b0](const double *a, size_t stride) { return a ? a + (size_t)b0 * stride : nullptr; };
            const int rc = fmpc_step_r(M->h[g], p, cb, flags, at(x0, n), at(x0_pre, n), at(u_prev, m), at(w, T * n), at(xf, n),
                                       at(nu0, NBn), u0 + (size_t)b0 * m, X ? X + (size_t)b0 * T * n : nullptr,
                                       U ? U + (size_t)b0 * T * m : nullptr, status ? status + b0 : nullptr,
                                       iters ? iters + b0 : nullptr, tep);
            if (rc == FMPC_OK) fill_stats(M, g, cb, status ? status + b0 : nullptr, *tep, iters ? iters + b0 : nullptr);
            return rc;
        });
    }
    int rc = FMPC_OK;
    for (int g = 0; g < M->G; ++g) { const int r = M->w[g]->wait(); if (r && !rc) rc = r; }
    if (telapsed) for (int g = 0; g < M->G; ++g) if (te[g] > *telapsed) *telapsed = te[g];
    return rc;
}

// Statistics of the last step, one record per device.  use_nccl != 0: the records travel over NCCL (ncclAllGather of one
// 64-byte record per GPU -- the only inter-GPU traffic of the whole path); otherwise, or when libnccl.so.2 cannot
// be loaded, the host reads them directly.  Returns the number of records written, or a negative error.
int fmpc_multi_last_stats(fmpc_multi *M, fmpc_multi_stats *out, int use_nccl, int *used_nccl)
{
    if (!M || !out) return FMPC_ERR_NULL;
    if (used_nccl) *used_nccl = 0;
    const int G = M->G;
    const size_t rec = sizeof(fmpc_multi_stats) / sizeof(double);
    if (use_nccl && G > 1 && M->nccl.load()) {
        bool ok = true;
        if (M->comms.empty()) {
            M->comms.assign(G, nullptr); M->d_send.assign(G, nullptr); M->d_recv.assign(G, nullptr); M->cs.assign(G, nullptr);
            if (M->nccl.CommInitAll(M->comms.data(), G, M->dev.data()) != 0) { M->comms.clear(); ok = false; }
            for (int g = 0; ok && g < G; ++g) {
                if (cudaSetDevice(M->dev[g]) != cudaSuccess || cudaMalloc(&M->d_send[g], rec * 8) != cudaSuccess ||
                    cudaMalloc(&M->d_recv[g], rec * 8 * G) != cudaSuccess || cudaStreamCreate(&M->cs[g]) != cudaSuccess) ok = false;
            }
        }
        for (int g = 0; ok && g < G; ++g) {
            if (cudaSetDevice(M->dev[g]) != cudaSuccess ||
                cudaMemcpyAsync(M->d_send[g], &M->last[g], rec * 8, cudaMemcpyHostToDevice, M->cs[g]) != cudaSuccess) ok = false;
        }
        if (ok) {
            M->nccl.GroupStart();
            for (int g = 0; g < G; ++g)
                if (M->nccl.AllGather(M->d_send[g], M->d_recv[g], rec, /*ncclFloat64*/ 8, M->comms[g], M->cs[g]) != 0) ok = false;
            if (M->nccl.GroupEnd() != 0) ok = false;
        }
        if (ok) {
            cudaSetDevice(M->dev[0]);
            if (cudaMemcpyAsync(out, M->d_recv[0], rec * 8 * G, cudaMemcpyDeviceToHost, M->cs[0]) != cudaSuccess ||
                cudaStreamSynchronize(M->cs[0]) != cudaSuccess) ok = false;
            for (int g = 1; ok && g < G; ++g) { cudaSetDevice(M->dev[g]); if (cudaStreamSynchronize(M->cs[g]) != cudaSuccess) ok = false; }
        }
        if (ok) { if (used_nccl) *used_nccl = 1; return G; }
    }
    for (int g = 0; g < G; ++g) out[g] = M->last[g];
    return G;
}

} // extern "C"
