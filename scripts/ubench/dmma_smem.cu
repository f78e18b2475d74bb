// Micro-benchmark: FP64 tensor-pipe (DMMA m8n8k4) throughput when the fragments come from shared memory, for the tile blockings the
// solve kernels use.  Prints cycles per DMMA per SM sub-partition (16 = the issue limit) for W warps per SM.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_smem dmma_smem.cu ; run: ./dmma_smem
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double &c0, double &c1, const double a, const double b)
{
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NR, int NC, int UNR>
__global__ void bench(double *out, long long *cyc, int reps, int ks, int ld)
{
    extern __shared__ double sm[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, gq = lane >> 2, q = lane & 3;
    for (int e = threadIdx.x; e < 80 * ld; e += blockDim.x) sm[e] = 1.0 / (1 + (e % 97));
    __syncthreads();
    double acc[NR][NC][2];
    for (int r = 0; r < NR; ++r) for (int u = 0; u < NC; ++u) acc[r][u][0] = acc[r][u][1] = 0.0;
    const long long t0 = clock64();
    for (int rep = 0; rep < reps; ++rep) {
        const int rt0 = (rep + wid) % 7, ct0 = (rep * 3 + wid) % 5;          // operands move so that nothing is hoisted out of the loop
        const double *A[NR], *B[NC];
        for (int r = 0; r < NR; ++r) A[r] = sm + (8 * (rt0 + r) + gq) * ld + q;
        for (int u = 0; u < NC; ++u) B[u] = sm + (8 * (ct0 + u) + gq) * ld + q;
#pragma unroll UNR
        for (int k = 0; k < ks; ++k) {
            double a[NR];
#pragma unroll
            for (int r = 0; r < NR; ++r) a[r] = A[r][4 * k];
#pragma unroll
            for (int u = 0; u < NC; ++u) {
                const double b = B[u][4 * k];
#pragma unroll
                for (int r = 0; r < NR; ++r) dmma(acc[r][u][0], acc[r][u][1], a[r], b);
            }
        }
    }
    const long long t1 = clock64();
    double s = 0.0;
    for (int r = 0; r < NR; ++r) for (int u = 0; u < NC; ++u) s += acc[r][u][0] + acc[r][u][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int NR, int NC, int UNR> void run(const char *name, int W, double *out, long long *cyc)
{
    const int reps = 200, ks = 18, ld = 76;
    const size_t smem = 80 * ld * 8;
    cudaFuncSetAttribute(bench<NR, NC, UNR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    bench<NR, NC, UNR><<<148, 32 * W, smem>>>(out, cyc, reps, ks, ld);
    bench<NR, NC, UNR><<<148, 32 * W, smem>>>(out, cyc, reps, ks, ld);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const double dm = (double)reps * ks * NR * NC * W / 4.0;           // DMMAs per sub-partition
    printf("%-28s warps/SM %2d : %6.1f cycles per DMMA per sub-partition (%.0f %% of the pipe)\n", name, W, c / dm, 1600.0 / (c / dm));
}

int main()
{
    double *out; long long *cyc;
    cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 8);
    for (int W : {4, 8, 16}) {
        run<1, 1, 2>("1 x 1 tile, unroll 2", W, out, cyc);
        run<1, 4, 2>("1 x 4 tiles, unroll 2", W, out, cyc);
        run<2, 4, 2>("2 x 4 tiles, unroll 2", W, out, cyc);
        run<2, 4, 6>("2 x 4 tiles, unroll 6", W, out, cyc);
        run<2, 4, 18>("2 x 4 tiles, unroll 18", W, out, cyc);
    }
    return 0;
}
