// VAR(p) identification of the Zernike-coefficient time series (README.md:116-130), batched over training sequences:
//
//     for i = PN+1 : num_train,  AA(i-PN, n(j-1)+1 : n j) = ad_acc(i-j, :)  (j = 1..PN);  BB(i-PN, :) = ad_acc(i, :);
//     PARA = (AA'*AA) \ AA'*BB;      A_j = PARA(n(j-1)+1 : n j, :)'
//
// One CTA per sequence.  The Gram blocks AA'AA (pn x pn, pn = PN n) and AA'BB (pn x n) are sums over time of outer
// products of the stacked vector v_i = [a_{i-1}; ...; a_{i-PN}; a_i]: the series streams through shared memory in time
// chunks, every thread owns a fixed set of (row, column) pairs and accumulates them in double-double (TwoSum / FMA
// error terms), because the normal equations square the conditioning of AA (cond(AA'AA) ~ 1e4..1e5 for AR(2) Zernike
// data) and the result must agree with any other correct fp64 evaluation to 1e-9.  Then Cholesky of AA'AA in shared
// memory (MATLAB's mldivide takes the same route for a symmetric positive definite matrix) and one forward / backward
// substitution per right-hand side column.  A one-off on-ramp computation, not a per-step kernel: sized for clarity.
#include <cuda_runtime.h>
#include <cmath>
#include <new>
#include <vector>
#include "../../include/fmpc.h"
#include "fmpc_device.cuh"

namespace {

constexpr int VT = 256;          // threads per CTA
constexpr int TCH = 32;          // time steps per shared-memory chunk

__device__ __forceinline__ void dd_add_prod(double &hi, double &lo, const double a, const double b)
{
    // (hi, lo) += a * b  with the rounding errors of the product and of the sum kept in lo
    const double p = a * b;
    const double pe = fma(a, b, -p);
    const double s = hi + p;
    const double bb = s - hi;
    const double se = (hi - (s - bb)) + (p - bb);
    hi = s;
    lo += se + pe;
}

__global__ void __launch_bounds__(VT) var_identify_kernel(const double *__restrict__ ad, int K, int n, int PN, double *__restrict__ Aout,
                                                          int *__restrict__ info, int max_pairs)
{
    extern __shared__ double sm[];
    const int pn = PN * n, nv = pn + n;                 // stacked vector length
    const int ldg = pn | 1;
    double *sG = sm;                                    // pn x ldg : AA'AA, then its Cholesky factor
    double *sH = sG + (size_t)pn * ldg;                 // pn x n   : AA'BB, then PARA
    double *sa = sH + (size_t)pn * n;                   // (TCH + PN) x n : time chunk of the series (rows = time)
    __shared__ int s_info;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const double *a = ad + (size_t)blockIdx.x * K * n;  // column-major K x n : a(k, j) = a[k + K j]
    const int npairs = pn * nv;                          // (r, c): r < pn, c < nv  (c < pn -> G, else H)

    // every thread owns pairs e = tid, tid + VT, ... (at most max_pairs of them), accumulators in registers
    constexpr int MAXP = 24;
    double hi[MAXP], lo[MAXP];
#pragma unroll
    for (int u = 0; u < MAXP; ++u) hi[u] = lo[u] = 0.0;
    (void)max_pairs;
    for (int t0 = PN; t0 < K; t0 += TCH) {
        const int tc = min(TCH, K - t0);
        __syncthreads();
        // rows t0 - PN .. t0 + tc - 1 of the series
        for (int e = tid; e < (tc + PN) * n; e += VT) {
            const int r = e / n, j = e - r * n;
            sa[e] = a[(size_t)(t0 - PN + r) + (size_t)K * j];
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < MAXP; ++u) {
            const int e = tid + u * VT;
            if (e < npairs) {
                const int r = e / nv, c = e - r * nv;
                // v_i[r] = a(i - 1 - r / n, r % n)   (lag block r / n);   v_i[c] likewise for c < pn, a(i, c - pn) for c >= pn
                const int lr = r / n + 1, jr = r - (lr - 1) * n;
                const int lc = (c < pn) ? c / n + 1 : 0, jc = (c < pn) ? c - (lc - 1) * n : c - pn;
                double h = hi[u], l = lo[u];
                for (int i = 0; i < tc; ++i) dd_add_prod(h, l, sa[(size_t)(i + PN - lr) * n + jr], sa[(size_t)(i + PN - lc) * n + jc]);
                hi[u] = h; lo[u] = l;
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < MAXP; ++u) {
        const int e = tid + u * VT;
        if (e < npairs) {
            const int r = e / nv, c = e - r * nv;
            const double v = hi[u] + lo[u];
            if (c < pn) sG[r * ldg + c] = v; else sH[r * n + (c - pn)] = v;
        }
    }
    __syncthreads();
    // (AA'AA) \ (AA'BB): Cholesky + two triangular solves per column
    if (wid == 0) {
        const int f = fmpc_dev::warp_potrf(sG, pn, ldg, lane);
        if (lane == 0) s_info = f;
    }
    __syncthreads();
    if (s_info) { if (tid == 0) info[blockIdx.x] = s_info; return; }
    for (int c = tid; c < n; c += VT) {
        for (int j = 0; j < pn; ++j) {
            double s = sH[j * n + c];
            for (int k = 0; k < j; ++k) s = fma(-sG[j * ldg + k], sH[k * n + c], s);
            sH[j * n + c] = s / sG[j * ldg + j];
        }
        for (int j = pn - 1; j >= 0; --j) {
            double s = sH[j * n + c];
            for (int k = j + 1; k < pn; ++k) s = fma(-sG[k * ldg + j], sH[k * n + c], s);
            sH[j * n + c] = s / sG[j * ldg + j];
        }
    }
    __syncthreads();
    // A_l = PARA(n (l-1) + 1 : n l, :)'  ->  Aout[s][l][r + n c] = PARA[(l n + c), r]   (column-major n x n per lag)
    double *Ao = Aout + (size_t)blockIdx.x * PN * n * n;
    for (int e = tid; e < PN * n * n; e += VT) {
        const int l = e / (n * n), rem = e - l * n * n, c = rem / n, r = rem - c * n;
        Ao[e] = sH[(size_t)(l * n + c) * n + r];
    }
    if (tid == 0) info[blockIdx.x] = 0;
}

} // namespace

extern "C" int var_identify(int nseq, int K, int n, int order, const double *ad, double *A, int *info, int device, double *telapsed)
{
    if (!ad || !A) return FMPC_ERR_NULL;
    if (telapsed) *telapsed = 0.0;
    if (nseq < 0 || n < 1 || order < 1 || order > 4 || K <= order) return FMPC_ERR_DIM;
    const int pn = order * n, nv = pn + n;
    if ((long long)pn * nv > 24LL * VT) return FMPC_ERR_DIM;           // register accumulators: (order n) (order + 1) n <= 6144
    if (K - order < pn) return FMPC_ERR_DIM;                           // fewer equations than unknowns: AA'AA singular
    if (nseq == 0) return FMPC_OK;
    int cnt = 0;
    if (cudaGetDeviceCount(&cnt) != cudaSuccess || device < 0 || device >= cnt) return FMPC_ERR_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major != 10) return FMPC_ERR_CUDA;
    if (cudaSetDevice(device) != cudaSuccess) return FMPC_ERR_CUDA;
    const size_t smem = ((size_t)pn * (pn | 1) + (size_t)pn * n + (size_t)(TCH + order) * n) * 8;
    if (smem > (size_t)prop.sharedMemPerBlockOptin) return FMPC_ERR_DIM;
    if (cudaFuncSetAttribute(var_identify_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return FMPC_ERR_CUDA;
    double *d_ad = nullptr, *d_A = nullptr;
    int *d_info = nullptr;
    const size_t nin = (size_t)nseq * K * n, nout = (size_t)nseq * order * n * n;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    bool ok = cudaMalloc(&d_ad, nin * 8) == cudaSuccess && cudaMalloc(&d_A, nout * 8) == cudaSuccess &&
              cudaMalloc(&d_info, (size_t)nseq * 4) == cudaSuccess && cudaEventCreate(&e0) == cudaSuccess && cudaEventCreate(&e1) == cudaSuccess;
    ok = ok && cudaMemcpy(d_ad, ad, nin * 8, cudaMemcpyHostToDevice) == cudaSuccess;
    int rc = FMPC_OK;
    if (ok) {
        cudaEventRecord(e0);
        var_identify_kernel<<<nseq, VT, smem>>>(d_ad, K, n, order, d_A, d_info, 24);
        cudaEventRecord(e1);
        ok = cudaGetLastError() == cudaSuccess && cudaMemcpy(A, d_A, nout * 8, cudaMemcpyDeviceToHost) == cudaSuccess;
        std::vector<int> hinfo((size_t)nseq, 0);
        ok = ok && cudaMemcpy(hinfo.data(), d_info, (size_t)nseq * 4, cudaMemcpyDeviceToHost) == cudaSuccess;
        if (ok) {
            for (int s = 0; s < nseq; ++s) { if (info) info[s] = hinfo[s]; if (hinfo[s] && !info) rc = FMPC_ERR_NOT_PD; }
            if (telapsed) { float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1); *telapsed = ms * 1e-3; }
        }
    }
    if (d_ad) cudaFree(d_ad);
    if (d_A) cudaFree(d_A);
    if (d_info) cudaFree(d_info);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    return ok ? rc : FMPC_ERR_CUDA;
}
