// Micro-benchmark: single-warp Cholesky + in-place inverse of an n x n block in shared memory (the serial section of the
// general-structure kernel's blocked factorisation).  Prints cycles per phase for W concurrent warps (each on its own block).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o potrf potrf.cu ; run: ./potrf [n] [W]
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

__host__ __device__ inline int gen_ld(int n) { int l = (n + 3) & ~3; if ((l & 7) != 4) l += 4; return l; }

__device__ int potrf_v1(double *Sm, int n, int ld, int lane, double *rdiag, long long *t)
{
    long long t0 = clock64();
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
        if (!(d > 0.0)) return k + 1;
        const double rs = rsqrt(d);
        __syncwarp();
        for (int r = k + lane; r < n; r += 32) Sm[(size_t)r * ld + k] *= rs;
        if (lane == 0) rdiag[k] = rs;
        __syncwarp();
    }
    t[0] = clock64() - t0;
    return 0;
}
__device__ void inv_v1(double *Sm, int n, int ld, int lane, const double *rdiag, double *tmp, long long *t)
{
    long long t0 = clock64();
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
    t[1] = clock64() - t0;
}

// v2: right-looking Cholesky, a lane per row, the row in shared memory but the column-k multipliers broadcast by shuffles;
// one shared-memory round trip per column instead of three
__device__ int potrf_v2(double *Sm, int n, int ld, int lane, double *rdiag, long long *t)
{
    long long t0 = clock64();
    double *rr = Sm + (size_t)min(lane, n - 1) * ld;
    for (int k = 0; k < n; ++k) {
        const double akk = __shfl_sync(0xffffffffu, rr[k], k);
        if (!(akk > 0.0)) return k + 1;
        const double rs = rsqrt(akk);
        const double l = rr[k] * rs;                 // L[r][k] for r >= k (own row only)
        if (lane >= k && lane < n) rr[k] = l;
        if (lane == 0) rdiag[k] = rs;
        for (int c = k + 1; c < n; ++c) {
            const double lc = __shfl_sync(0xffffffffu, l, c);
            if (lane >= c && lane < n) rr[c] = fma(-l, lc, rr[c]);
        }
    }
    __syncwarp();
    t[0] = clock64() - t0;
    return 0;
}
// v2 inverse: a lane per COLUMN j of X, right-looking forward substitution; L[r][k] read from shared memory (broadcast), X column in
// a second shared block (column j = row j of Xt, i.e. the transpose of inv(L))
__device__ void inv_v2(const double *Sm, double *Xt, int n, int ld, int lane, const double *rdiag, long long *t)
{
    long long t0 = clock64();
    double *x = Xt + (size_t)min(lane, n - 1) * ld;
    const bool on = lane < n;
    if (on) for (int r = 0; r < n; ++r) x[r] = (r == lane) ? 1.0 : 0.0;
    for (int k = 0; k < n; ++k) {
        const double xk = x[k] * rdiag[k];
        if (on) x[k] = xk;
        for (int r = k + 1; r < n; ++r) if (on) x[r] = fma(-Sm[(size_t)r * ld + k], xk, x[r]);
    }
    __syncwarp();
    t[1] = clock64() - t0;
}

__global__ void bench(const double *A, double *out, long long *tm, int n, int ver)
{
    extern __shared__ double sm[];
    const int ld = gen_ld(n), lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    double *S = sm + (size_t)wid * (2 * 32 * ld + 64), *X = S + 32 * ld, *rd = X + 32 * ld, *tmp = rd + 32;
    for (int e = lane; e < 32 * ld; e += 32) { const int r = e / ld, c = e % ld; S[e] = (r < n && c <= r) ? A[r * n + c] : 0.0; X[e] = 0.0; }
    __syncwarp();
    long long t[2] = {0, 0};
    if (ver == 1) { potrf_v1(S, n, ld, lane, rd, t); inv_v1(S, n, ld, lane, rd, tmp, t); }
    else { potrf_v2(S, n, ld, lane, rd, t); inv_v2(S, X, n, ld, lane, rd, t); }
    __syncwarp();
    if (wid == 0) {
        for (int e = lane; e < n * n; e += 32) { const int r = e / n, c = e % n; out[e] = (ver == 1) ? S[r * ld + c] : X[c * ld + r]; }
        if (lane == 0) { tm[0] = t[0]; tm[1] = t[1]; }
    }
}

int main(int argc, char **argv)
{
    const int n = argc > 1 ? atoi(argv[1]) : 27;
    std::vector<double> B(n * n), A(n * n), Lr(n * n, 0.0), Xr(n * n, 0.0);
    srand(1);
    for (auto &v : B) v = rand() / (double)RAND_MAX - 0.5;
    for (int r = 0; r < n; ++r) for (int c = 0; c < n; ++c) { double s = (r == c) ? 1.0 : 0.0; for (int k = 0; k < n; ++k) s += B[r * n + k] * B[c * n + k]; A[r * n + c] = s; }
    for (int k = 0; k < n; ++k) {
        double d = A[k * n + k]; for (int j = 0; j < k; ++j) d -= Lr[k * n + j] * Lr[k * n + j];
        Lr[k * n + k] = sqrt(d);
        for (int r = k + 1; r < n; ++r) { double s = A[r * n + k]; for (int j = 0; j < k; ++j) s -= Lr[r * n + j] * Lr[k * n + j]; Lr[r * n + k] = s / Lr[k * n + k]; }
    }
    for (int c = 0; c < n; ++c) { Xr[c * n + c] = 1.0 / Lr[c * n + c]; for (int r = c + 1; r < n; ++r) { double s = 0; for (int k = c; k < r; ++k) s -= Lr[r * n + k] * Xr[k * n + c]; Xr[r * n + c] = s / Lr[r * n + r]; } }
    double *dA, *dO; long long *dT;
    cudaMalloc(&dA, n * n * 8); cudaMalloc(&dO, n * n * 8); cudaMalloc(&dT, 16);
    cudaMemcpy(dA, A.data(), n * n * 8, cudaMemcpyHostToDevice);
    const int ld = gen_ld(n);
    for (int ver = 1; ver <= 2; ++ver)
        for (int W : {1, 8}) {
            const size_t smem = (size_t)W * (2 * 32 * ld + 64) * 8;
            cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            for (int rep = 0; rep < 2; ++rep) bench<<<1, 32 * W, smem>>>(dA, dO, dT, n, ver);
            if (cudaDeviceSynchronize() != cudaSuccess) { printf("cuda error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
            std::vector<double> O(n * n); long long T[2];
            cudaMemcpy(O.data(), dO, n * n * 8, cudaMemcpyDeviceToHost); cudaMemcpy(T, dT, 16, cudaMemcpyDeviceToHost);
            double err = 0; for (int r = 0; r < n; ++r) for (int c = 0; c <= r; ++c) err = fmax(err, fabs(O[r * n + c] - Xr[r * n + c]));
            printf("n=%d v%d warps=%d: potrf %lld cycles, inverse %lld cycles, max |inv(L) - ref| = %.2e\n", n, ver, W, T[0], T[1], err);
        }
    return 0;
}
