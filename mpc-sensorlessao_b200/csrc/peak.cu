// FP64 pipe micro-benchmarks: the roofline denominators for the fastMPC solve.
// MEASURED_PEAKS.json (driver-written) carries HBM and bf16 peaks only, so the FP64 peak is
// measured here, live, with the two instruction forms a kernel could use on sm_100a:
//   kind 0 : DFMA   -- 16 independent fused-multiply-add chains per thread
//   kind 1 : DMMA   -- mma.sync.aligned.m8n8k4.row.col.f64, 8 independent accumulator tiles per warp
#include <cuda_runtime.h>
#include "../../include/fmpc.h"

namespace {

constexpr int DFMA_CHAINS = 16;
constexpr int DFMA_INNER = 32;

__global__ void __launch_bounds__(256) dfma_peak_kernel(double *out, int iters, double a, double b)
{
    double acc[DFMA_CHAINS];
#pragma unroll
    for (int i = 0; i < DFMA_CHAINS; ++i) acc[i] = (double)(threadIdx.x + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < DFMA_INNER; ++r)
#pragma unroll
            for (int i = 0; i < DFMA_CHAINS; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < DFMA_CHAINS; ++i) s += acc[i];
    if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;     // keep the chains alive
}

constexpr int DMMA_TILES = 8;
constexpr int DMMA_INNER = 16;

__global__ void __launch_bounds__(256) dmma_peak_kernel(double *out, int iters, double a, double b)
{
    double c0[DMMA_TILES], c1[DMMA_TILES];
#pragma unroll
    for (int i = 0; i < DMMA_TILES; ++i) { c0[i] = threadIdx.x + i; c1[i] = i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < DMMA_INNER; ++r)
#pragma unroll
            for (int i = 0; i < DMMA_TILES; ++i)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < DMMA_TILES; ++i) s += c0[i] + c1[i];
    if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

} // namespace

extern "C" double fmpc_fp64_peak(int device, int kind, int iters)
{
    int cnt = 0;
    if (cudaGetDeviceCount(&cnt) != cudaSuccess || device < 0 || device >= cnt) return -1.0;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major != 10) return -1.0;
    if (cudaSetDevice(device) != cudaSuccess) return -1.0;
    if (iters < 1) iters = 1;
    const int block = 256, grid = prop.multiProcessorCount * 8;
    double *d_out = nullptr;
    if (cudaMalloc(&d_out, (size_t)grid * block * 8) != cudaSuccess) return -1.0;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {       // rep 0 = warm-up
        cudaEventRecord(e0);
        if (kind == 0) dfma_peak_kernel<<<grid, block>>>(d_out, iters, 0.999999, 1e-7);
        else dmma_peak_kernel<<<grid, block>>>(d_out, iters, 0.5, 0.25);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { best = -1.0; break; }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        double flops;
        if (kind == 0) flops = 2.0 * (double)grid * block * DFMA_CHAINS * DFMA_INNER * (double)iters;
        else flops = 2.0 * 8 * 8 * 4 * (double)grid * (block / 32) * DMMA_TILES * DMMA_INNER * (double)iters;
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d_out);
    return best;
}
