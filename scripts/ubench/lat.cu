// Dependent-chain latencies on sm_100a (one warp, clock64 around N dependent ops).
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__global__ void k(double *out, long long *cyc, double a, double b, double *sm_dummy) {
    __shared__ double sh[1024];
    for (int i = threadIdx.x; i < 1024; i += 32) sh[i] = i * 1e-3;
    __syncwarp();
    const int N = 256;
    long long t0, t1;
    // 1. dependent DMMA chain
    double c0 = 1, c1 = 2;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) dmma(c0, c1, a, b);
    t1 = clock64(); if (threadIdx.x == 0) cyc[0] = (t1 - t0);
    // 2. two interleaved chains
    double d0 = 1, d1 = 2;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) { dmma(c0, c1, a, b); dmma(d0, d1, a, b); }
    t1 = clock64(); if (threadIdx.x == 0) cyc[1] = (t1 - t0);
    // 3. four interleaved chains
    double e0 = 1, e1 = 2, f0 = 3, f1 = 4;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) { dmma(c0, c1, a, b); dmma(d0, d1, a, b); dmma(e0, e1, a, b); dmma(f0, f1, a, b); }
    t1 = clock64(); if (threadIdx.x == 0) cyc[2] = (t1 - t0);
    // 4. dependent DFMA chain
    double x = a;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x) : "d"(a), "d"(b));
    t1 = clock64(); if (threadIdx.x == 0) cyc[3] = (t1 - t0);
    // 5. dependent shfl chain (64-bit = 2 SHFL)
    double y = a + threadIdx.x;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) y = __shfl_sync(0xffffffffu, y, (i + 1) & 31);
    t1 = clock64(); if (threadIdx.x == 0) cyc[4] = (t1 - t0);
    // 6. dependent rcp.approx.f64 chain
    double z = a + 1.5;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) asm volatile("rcp.approx.ftz.f64 %0, %0;" : "+d"(z));
    t1 = clock64(); if (threadIdx.x == 0) cyc[5] = (t1 - t0);
    // 7. dependent LDS chain (pointer chasing in smem, 64-bit)
    int idx = threadIdx.x;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) { double v = sh[idx]; idx = (int)(v * 0.0) + ((idx + 33) & 1023); }
    t1 = clock64(); if (threadIdx.x == 0) cyc[6] = (t1 - t0);
    // 8. LDS -> DMMA -> dependent (operand from smem each time)
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) { double av = sh[(threadIdx.x + 4 * i) & 1023]; dmma(c0, c1, av, b); }
    t1 = clock64(); if (threadIdx.x == 0) cyc[7] = (t1 - t0);
    // 9. full IEEE division chain
    double w = a + 2.5;
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; ++i) w = 1.0 / (w + 1.0);
    t1 = clock64(); if (threadIdx.x == 0) cyc[8] = (t1 - t0);
    // 10. st.global then ld.global same address (L1) round trip chain
    double *g = sm_dummy + threadIdx.x;
    double gv = a;
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; ++i) { *(volatile double *)g = gv; gv = *(volatile double *)g + 1.0; }
    t1 = clock64(); if (threadIdx.x == 0) cyc[9] = (t1 - t0);
    // 11. __syncthreads cost with 4 warps is measured in k2
    out[threadIdx.x] = c0 + c1 + d0 + d1 + e0 + e1 + f0 + f1 + x + y + z + idx + w + gv;
}
__global__ void k2(long long *cyc) {
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < 256; ++i) __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[10] = t1 - t0;
}
// L2 load latency: pointer chase in a 64 MB buffer
__global__ void k3(const int *chain, long long *cyc, int *sink) {
    int idx = 0;
    long long t0 = clock64();
    for (int i = 0; i < 256; ++i) idx = __ldcg(chain + idx);
    long long t1 = clock64();
    cyc[11] = t1 - t0; *sink = idx;
}
int main() {
    double *out, *gd; long long *cyc; cudaMalloc(&out, 4096); cudaMalloc(&gd, 4096); cudaMalloc(&cyc, 128); cudaMemset(cyc, 0, 128);
    for (int rep = 0; rep < 2; ++rep) { k<<<1, 32>>>(out, cyc, 0.5, 0.25, gd); k2<<<1, 128>>>(cyc); }
    // chase buffer: stride 4 KB over 32 MB (L2 resident after first pass)
    const int NEL = 8 << 20; int *h = new int[NEL]; for (int i = 0; i < NEL; ++i) h[i] = (i + 1024 * 17) % NEL;
    int *chain, *sink; cudaMalloc(&chain, NEL * 4); cudaMalloc(&sink, 4); cudaMemcpy(chain, h, NEL * 4, cudaMemcpyHostToDevice);
    k3<<<1, 1>>>(chain, cyc, sink); k3<<<1, 1>>>(chain, cyc, sink);
    long long hc[16]; cudaMemcpy(hc, cyc, 128, cudaMemcpyDeviceToHost);
    const char *nm[] = {"DMMA dependent", "DMMA 2 chains (per pair)", "DMMA 4 chains (per quad)", "DFMA dependent", "SHFL f64 dependent", "RCP64H dependent",
                        "LDS dependent", "LDS->DMMA dependent", "1/x IEEE div dependent", "STG->LDG round trip", "__syncthreads (4 warps)", "L2 load (ldcg chase)"};
    for (int i = 0; i < 12; ++i) printf("%-28s %8.1f cycles\n", nm[i], hc[i] / 256.0);
    return 0;
}
