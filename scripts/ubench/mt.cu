// micro-benchmark of the MT19937 block regeneration: where do the ~850 cycles per 624-word block go?
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned mt_f(unsigned a, unsigned b)
{
    const unsigned y = (a & 0x80000000u) | (b & 0x7fffffffu);
    return (y >> 1) ^ ((0u - (y & 1u)) & 0x9908b0dfu);
}
__device__ __forceinline__ unsigned mt_temper(unsigned y)
{
    y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
    return y;
}
__device__ __forceinline__ unsigned mt_next_word(const unsigned *o, const int e)
{
    if (e < 227) return o[e + 397] ^ mt_f(o[e], o[e + 1]);
    if (e < 454) return o[e + 170] ^ mt_f(o[e - 227], o[e - 226]) ^ mt_f(o[e], o[e + 1]);
    if (e < 623) return o[e - 57] ^ mt_f(o[e - 454], o[e - 453]) ^ mt_f(o[e - 227], o[e - 226]) ^ mt_f(o[e], o[e + 1]);
    const unsigned n0 = o[397] ^ mt_f(o[0], o[1]);
    const unsigned n396 = o[566] ^ mt_f(o[169], o[170]) ^ mt_f(o[396], o[397]);
    return n396 ^ mt_f(o[623], n0);
}
// MODE bit0: twist, bit1: temper+convert, bit2: global store
template <int THREADS, int TWIST, int MODE>
__global__ void __launch_bounds__(THREADS, 1) k(unsigned *state, double *out, int nblocks, long long *cyc)
{
    constexpr int OUTT = THREADS - TWIST, NTW = (624 + TWIST - 1) / TWIST, NOUT = (312 + OUTT - 1) / OUTT;
    __shared__ __align__(16) unsigned mt[2][624];
    const int tid = threadIdx.x;
    for (int i = tid; i < 624; i += THREADS) { mt[0][i] = state[i]; mt[1][i] = state[i] * 3u; }
    __syncthreads();
    int cur = 0;
    double acc = 0.0;
    const long long t0 = clock64();
    for (int b = 0; b < nblocks; ++b) {
        const unsigned *o = mt[cur];
        if (tid < TWIST) {
            if (MODE & 1) {
                unsigned *w = mt[cur ^ 1];
                unsigned v[NTW];
#pragma unroll
                for (int j = 0; j < NTW; ++j) { const int e = tid + TWIST * j; if (e < 624) v[j] = mt_next_word(o, e); }
#pragma unroll
                for (int j = 0; j < NTW; ++j) { const int e = tid + TWIST * j; if (e < 624) w[e] = v[j]; }
            }
        } else if (MODE & 2) {
#pragma unroll
            for (int j = 0; j < NOUT; ++j) {
                const int t = tid - TWIST + OUTT * j;
                if (t < 312) {
                    const uint2 y = *reinterpret_cast<const uint2 *>(o + 2 * t);
                    const unsigned a = mt_temper(y.x) >> 5, bb = mt_temper(y.y) >> 6;
                    const double da = __longlong_as_double(0x4330000000000000ll | (long long)a) - 4503599627370496.0;
                    const double db = __longlong_as_double(0x4330000000000000ll | (long long)bb) - 4503599627370496.0;
                    const double r = (da * 67108864.0 + db) * (1.0 / 9007199254740992.0);
                    if (MODE & 4) out[(size_t)b * 312 + t] = r; else acc += r;
                }
            }
        }
        __syncthreads();
        cur ^= 1;
    }
    const long long t1 = clock64();
    if (tid == 0) cyc[0] = t1 - t0;
    if (acc == 12345.678) out[0] = acc;
    for (int i = tid; i < 624; i += THREADS) state[i] = mt[cur][i];
}

// ---- variant B: one word per thread, branch-free (7 loads + 3 twists, masks decide what enters), raw block stored to global ----
template <int STORE>
__global__ void __launch_bounds__(640, 1) kb(unsigned *state, unsigned *raw, int nblocks, long long *cyc)
{
    __shared__ __align__(16) unsigned mt[2][624];
    const int e = threadIdx.x;
    for (int i = e; i < 624; i += 640) { mt[0][i] = state[i]; mt[1][i] = 0u; }
    // loop-invariant operand indices of word e
    int ia0, ia1, ib0, ib1, ic0, ic1, ix;
    unsigned mb = 0u, mc = 0u;
    const bool act = e < 624, last = (e == 623);
    if (e < 227) { ix = e + 397; ia0 = e; ia1 = e + 1; ib0 = ib1 = ic0 = ic1 = 0; }
    else if (e < 454) { ix = e + 170; ia0 = e; ia1 = e + 1; ib0 = e - 227; ib1 = e - 226; mb = ~0u; ic0 = ic1 = 0; }
    else { ix = e - 57; ia0 = e; ia1 = min(e + 1, 623); ib0 = e - 227; ib1 = e - 226; mb = ~0u; ic0 = e - 454; ic1 = e - 453; mc = ~0u; }
    if (!act) { ix = ia0 = ia1 = ib0 = ib1 = ic0 = ic1 = 0; }
    __syncthreads();
    int cur = 0;
    const long long t0 = clock64();
    for (int b = 0; b < nblocks; ++b) {
        const unsigned *o = mt[cur];
        unsigned *w = mt[cur ^ 1];
        unsigned v;
        if (!last) {
            const unsigned x = o[ix], a0 = o[ia0], a1 = o[ia1], b0 = o[ib0], b1 = o[ib1], c0 = o[ic0], c1 = o[ic1];
            v = x ^ mt_f(a0, a1) ^ (mt_f(b0, b1) & mb) ^ (mt_f(c0, c1) & mc);
        } else {
            const unsigned n0 = o[397] ^ mt_f(o[0], o[1]);
            const unsigned n396 = o[566] ^ mt_f(o[169], o[170]) ^ mt_f(o[396], o[397]);
            v = n396 ^ mt_f(o[623], n0);
        }
        if (act) { w[e] = v; if (STORE) raw[(size_t)b * 624 + e] = v; }
        __syncthreads();
        cur ^= 1;
    }
    const long long t1 = clock64();
    if (e == 0) cyc[0] = t1 - t0;
    if (act) state[e] = mt[cur][e];
}
template <int STORE> void runb(const char *name, unsigned *st, unsigned *raw, long long *cyc, int nb)
{
    kb<STORE><<<1, 640>>>(st, raw, nb, cyc);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    kb<STORE><<<1, 640>>>(st, raw, nb, cyc);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-44s threads  640: %.3f ms, %.0f cycles / block (%s)\n", name, ms, (double)c / nb, cudaGetErrorString(cudaGetLastError()));
}

template <int THREADS, int TWIST, int MODE> void run(const char *name, unsigned *st, double *out, long long *cyc, int nb)
{
    k<THREADS, TWIST, MODE><<<1, THREADS>>>(st, out, nb, cyc);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<THREADS, TWIST, MODE><<<1, THREADS>>>(st, out, nb, cyc);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-44s threads %4d twist %4d: %.3f ms, %.0f cycles / block (%s)\n", name, THREADS, TWIST, ms, (double)c / nb, cudaGetErrorString(cudaGetLastError()));
}
int main()
{
    const int nb = 7351;
    unsigned *st; double *out; long long *cyc;
    cudaMalloc(&st, 625 * 4); cudaMalloc(&out, (size_t)nb * 312 * 8); cudaMalloc(&cyc, 8);
    unsigned h[625]; h[0] = 5489u; for (int i = 1; i < 624; ++i) h[i] = 1812433253u * (h[i - 1] ^ (h[i - 1] >> 30)) + i; h[624] = 624;
    cudaMemcpy(st, h, sizeof h, cudaMemcpyHostToDevice);
    run<320, 192, 0>("barrier only", st, out, cyc, nb);
    run<320, 192, 1>("twist only", st, out, cyc, nb);
    run<320, 192, 2>("temper+convert only", st, out, cyc, nb);
    run<320, 192, 6>("temper+convert+store", st, out, cyc, nb);
    run<320, 192, 3>("twist + temper", st, out, cyc, nb);
    run<320, 192, 7>("all", st, out, cyc, nb);
    run<1024, 640, 7>("all", st, out, cyc, nb);
    run<1024, 640, 1>("twist only", st, out, cyc, nb);
    run<256, 128, 7>("all", st, out, cyc, nb);
    run<128, 96, 7>("all", st, out, cyc, nb);
    unsigned *raw; cudaMalloc(&raw, (size_t)nb * 624 * 4);
    runb<0>("B: branch-free twist, one word/thread", st, raw, cyc, nb);
    runb<1>("B: same + raw block to global", st, raw, cyc, nb);
    return 0;
}
