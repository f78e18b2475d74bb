// Do DMMA (mma.sync m8n8k4 f64) and DFMA share one FP64 pipe on sm_100a?  8 warps per CTA, 1 CTA per SM:
//   mode 0: 8 warps DMMA        mode 1: 8 warps DFMA        mode 2: warps 0-3 DMMA + warps 4-7 DFMA (one of each per SMSP)
//   mode 3: 4 warps DMMA only   mode 4: 4 warps DFMA only   mode 5: 1 DMMA warp per SMSP with NCH independent chains (latency/issue)
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int NCH> __device__ double run_dmma(int iters, double a, double b) {
    double c0[NCH], c1[NCH];
#pragma unroll
    for (int i = 0; i < NCH; ++i) { c0[i] = i; c1[i] = -i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < NCH; ++i) dmma(c0[i], c1[i], a, b);
    }
    double s = 0; for (int i = 0; i < NCH; ++i) s += c0[i] + c1[i];
    return s;
}
__device__ double run_dfma(int iters, double a, double b) {
    double x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = i + threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x[i]) : "d"(a), "d"(b));
    }
    double s = 0; for (int i = 0; i < 16; ++i) s += x[i];
    return s;
}
__global__ void k(double *out, int mode, int iters, int nch, double a, double b) {
    const int w = threadIdx.x >> 5;
    double s = 0;
    bool do_dmma = (mode == 0) || (mode == 2 && w < 4) || mode == 3 || mode == 5;
    bool do_dfma = (mode == 1) || (mode == 2 && w >= 4) || mode == 4;
    if (mode == 5) {
        if (nch == 1) s = run_dmma<1>(iters, a, b); else if (nch == 2) s = run_dmma<2>(iters, a, b);
        else if (nch == 3) s = run_dmma<3>(iters, a, b); else if (nch == 4) s = run_dmma<4>(iters, a, b); else s = run_dmma<8>(iters, a, b);
    } else if (do_dmma) s = run_dmma<8>(iters, a, b);
    else if (do_dfma) s = run_dfma(iters, a, b);
    if (s == 123.456) out[threadIdx.x] = s;
}
int main() {
    double *out; cudaMalloc(&out, 8192);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000, sms = 148;
    auto run = [&](int mode, int threads, int nch) {
        k<<<sms, threads>>>(out, mode, 100, nch, 0.5, 0.25); cudaDeviceSynchronize();
        cudaEventRecord(e0); k<<<sms, threads>>>(out, mode, iters, nch, 0.5, 0.25); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); return ms * 1e-3;
    };
    const double dmma_fl = 8.0 * 8 * 512 * iters;         // per warp: 8 reps x 8 chains
    const double dfma_fl = 4.0 * 16 * 64 * iters;         // per warp
    double t;
    t = run(0, 256, 8); printf("mode0 8 DMMA warps/SM : %.3f ms  DMMA %.2f TF\n", t * 1e3, sms * 8 * dmma_fl / t / 1e12);
    t = run(1, 256, 8); printf("mode1 8 DFMA warps/SM : %.3f ms  DFMA %.2f TF\n", t * 1e3, sms * 8 * dfma_fl / t / 1e12);
    t = run(3, 128, 8); printf("mode3 4 DMMA warps/SM : %.3f ms  DMMA %.2f TF\n", t * 1e3, sms * 4 * dmma_fl / t / 1e12);
    t = run(4, 128, 8); printf("mode4 4 DFMA warps/SM : %.3f ms  DFMA %.2f TF\n", t * 1e3, sms * 4 * dfma_fl / t / 1e12);
    t = run(2, 256, 8); printf("mode2 4 DMMA + 4 DFMA : %.3f ms  DMMA %.2f TF + DFMA %.2f TF (if time ~ max(mode3,mode4): separate pipes; if ~ sum: shared)\n",
                               t * 1e3, sms * 4 * dmma_fl / t / 1e12, sms * 4 * dfma_fl / t / 1e12);
    for (int nch : {1, 2, 3, 4, 8}) {
        t = run(5, 128, nch);
        const double per = t * 1.965e9 / (8.0 * nch * iters);   // cycles per DMMA per warp at 1965 MHz
        printf("mode5 1 warp/SMSP, %d chains: %.3f ms, %.1f cycles per DMMA issue (@1965 MHz)\n", nch, t * 1e3, per);
    }
    return 0;
}
