// zernmodfit on the GPU: batched masked least-squares projection of nL x nL phase frames onto
// the Zernike basis (zernmodfit.m:195-213 + zernfun.m:140-192, driver loop README.md:78-93).
//
// The reference rebuilds Z (npix_in x nmodes) and runs a QR per frame although Z depends only on
// the fixed pupil grid.  Here the host builds Z once in fp64, forms the least-squares operator
// W = inv(Z'Z) Z' (cond(Z) = 3.8 at N = 6, 5.0 at N = 10: normal equations are safe at 1e-10),
// scatters it to the full frame (zeros outside the pupil) and the device computes
// coef = W * frames, an HBM-streaming fp64 GEMM with M = nmodes, K = nL^2, N = nframes.
//
// The GEMM is balanced between the two rooflines (2 nmodes npix_in flop per 8 nL^2 bytes = 5.4 flop/B at N = 6, the
// FP64 ridge of a B200 is ~5.7), so the product kernel (zmf_fit_dmma_kernel) runs it on the FP64 tensor pipe:
//   * K is cut into 8-pixel groups; groups wholly outside the pupil are dropped (their bytes are never read),
//     partially covered ones carry a pixel mask so that NaNs outside the pupil (zernmodfit.m:30) never reach a DMMA;
//   * the frames (A operand, M = 8 frames per tile) go HBM -> registers directly: every element is needed by exactly one
//     lane, each lane loads 16 B (2 k-steps), 4 lanes cover one 64-byte run of a frame, a 4-group register ring keeps
//     ~4 KB per warp in flight;
//   * W (B operand) is pre-permuted on the host into DMMA fragment order, so a chunk of groups is ONE contiguous block:
//     it is brought into shared memory by the bulk-copy engine (cp.async.bulk + mbarrier, 3 stages) and read back with
//     conflict-free 128-bit loads; MT frame tiles per warp share every B fragment;
//   * split-K over blockIdx.y (deterministic: partial sums to a scratch buffer, then a small reduction kernel) fills
//     the 148 SMs when there are few frames and removes the wave-quantisation tail when there are many.
#include <cuda_runtime.h>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>
#include "../../include/fmpc.h"

namespace {

double prod2(int k) { double p = 1.0; for (int j = 2; j <= k; ++j) p *= j; return p; }   // MATLAB prod(2:k)

// zernfun.m:161-192 for one (n, m) and one sample
double zern_eval(int n, int m, double r, double theta)
{
    const int ma = m < 0 ? -m : m;
    double z = 0.0;
    const int smax = (n - ma) / 2;
    for (int s = smax; s >= 0; --s) {                    // :165  k = length(s):-1:1
        const double p = (1 - 2 * (s % 2)) * prod2(n - s) / prod2(s) / prod2((n - ma) / 2 - s) / prod2((n + ma) / 2 - s);
        const int pw = n - 2 * s;
        z += p * (pw == 0 ? 1.0 : std::pow(r, (double)pw));
    }
    if (m > 0) z *= std::cos(theta * ma);
    else if (m < 0) z *= std::sin(theta * ma);
    return z;
}

} // namespace

struct zmf_handle {
    int device = 0, nL = 0, N = 0, nmodes = 0, npix = 0, npix_in = 0;
    std::vector<double> Z;                 // npix_in x nmodes, column-major
    std::vector<unsigned char> mask;       // nL*nL, column-major linear index
    double *d_W = nullptr;                 // nmodes x npix (row j contiguous over pixels), zeros outside pupil
    unsigned char *d_mask = nullptr;
    double *d_frames = nullptr, *d_coef = nullptr;
    size_t cap_frames = 0;
    // DMMA path
    double *d_Wf = nullptr;                // [group][ntile][lane][2] fragment-ordered W
    unsigned *d_ginfo = nullptr;           // per group: (first pixel / 8) | (pixel mask << 24)
    int ngroups = 0, ntiles = 0;
    double *d_Zdev = nullptr;              // synthesis, generic kernel: Z (npix_in x nmodes, column-major) on the device
    int *d_pin = nullptr;                  // synthesis, generic kernel: index of every frame pixel among the in-pupil samples (-1 outside)
    double *d_Zf = nullptr;                // synthesis: [group of 8 pixels][k-step][lane] fragment-ordered Z' (zeros outside the pupil)
    unsigned char *d_gmask = nullptr;      // synthesis: pupil mask byte of every 8-pixel group
    int syn_ks = 0;                        // k-steps (4 modes each) of the synthesis kernel instantiation: 7 or 18
    double *d_off = nullptr;               // nmodes: constant subtracted from every output column (estimator: W b_s), or NULL
    double *d_part = nullptr;              // split-K partial sums [ksplit][nframes][8 ntiles]
    size_t part_doubles = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    long long launches = 0;
};

// ---------------------------------------------------------------------------------------------
// v1 kernel: FT frames per CTA; W streamed from L2 once per CTA and reused across the FT frames.
// Thread t owns pixels p = t, t+NT, ... ; accumulates nmodes x FT partial sums in registers
// (modes processed in chunks of MC), CTA-reduces them at the end.
// ---------------------------------------------------------------------------------------------
template <int MC, int FT, int NT>
__global__ void __launch_bounds__(NT) zmf_fit_kernel(const double *__restrict__ W, const unsigned char *__restrict__ mask,
                                                     const double *__restrict__ frames, double *__restrict__ coef,
                                                     int npix, int nmodes, int nframes, const double *__restrict__ off)
{
    __shared__ double red[NT / 32][MC * FT];
    const int f0 = blockIdx.x * FT;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (int j0 = 0; j0 < nmodes; j0 += MC) {
        double acc[MC][FT];
#pragma unroll
        for (int a = 0; a < MC; ++a)
#pragma unroll
            for (int b = 0; b < FT; ++b) acc[a][b] = 0.0;
        for (int p = tid; p < npix; p += NT) {
            if (!mask[p]) continue;                       // outside the pupil: may hold NaN (zernmodfit.m:30)
            double fv[FT];
#pragma unroll
            for (int b = 0; b < FT; ++b) fv[b] = (f0 + b < nframes) ? __ldcs(frames + (size_t)(f0 + b) * npix + p) : 0.0;
#pragma unroll
            for (int a = 0; a < MC; ++a) {
                const double w = (j0 + a < nmodes) ? __ldg(W + (size_t)(j0 + a) * npix + p) : 0.0;
#pragma unroll
                for (int b = 0; b < FT; ++b) acc[a][b] = fma(w, fv[b], acc[a][b]);
            }
        }
#pragma unroll
        for (int a = 0; a < MC; ++a)
#pragma unroll
            for (int b = 0; b < FT; ++b) {
                double v = acc[a][b];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane == 0) red[wid][a * FT + b] = v;
            }
        __syncthreads();
        if (tid < MC * FT) {
            double v = 0.0;
#pragma unroll
            for (int w = 0; w < NT / 32; ++w) v += red[w][tid];
            const int a = tid / FT, b = tid % FT;
            if (j0 + a < nmodes && f0 + b < nframes) coef[(size_t)(f0 + b) * nmodes + j0 + a] = off ? v - off[j0 + a] : v;
        }
        __syncthreads();
    }
}


// ---------------------------------------------------------------------------------------------
// DMMA kernel (see the header comment).  Grid = (frame blocks of NW * 8 MT frames, ksplit).
// ---------------------------------------------------------------------------------------------
namespace {

#ifndef ZSYN_STAGES
#define ZSYN_STAGES 2      // measured (profiles/r01_zmf_occupancy.log): 2 stages (3 resident CTAs) beat 3 stages and 8-group chunks
#endif
#ifndef ZSYN_CH
#define ZSYN_CH 16
#endif
template <int NT> struct ZChunk { static constexpr int CH = (NT <= 4) ? 16 : 8; };     // groups per shared-memory stage
#ifndef ZMF_PF
#define ZMF_PF 16
#endif
constexpr int ZPF = ZMF_PF;                                                           // L2 prefetch distance (groups), 0 = off
#ifndef ZMF_RING1
#define ZMF_RING1 8
#endif
template <int MT> struct ZRingDepth { static constexpr int V = (MT == 1) ? ZMF_RING1 : 4; };   // groups in flight per warp

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void zdmma(double (&c)[2], const double a, const double b)
{
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
__device__ __forceinline__ bool mbar_try_wait(unsigned bar, unsigned parity)
{
    unsigned ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}

// ST = 3 stages, or ST = 2 with the register budget of three resident CTAs per SM (80 registers, 24 warps): measured faster for
// large frame counts (32 768 frames: 1.15 vs 1.24 ms), slower for small ones
template <int NT, int MT, int NW, bool VEC, int ST>
__global__ void __launch_bounds__(NW * 32, ST == 2 ? 3 : 1) zmf_fit_dmma_kernel(const double *__restrict__ Wf, const unsigned *__restrict__ ginfo, int ngroups,
                                                               int groups_per_split, const double *__restrict__ frames,
                                                               double *__restrict__ out, int out_ld, size_t out_split_stride,
                                                               int npix, int nmodes, int nframes, const double *__restrict__ offv)
{
    constexpr int CH = ZChunk<NT>::CH, GD = NT * 64;            // doubles of W per group
    constexpr int ZRING = ZRingDepth<MT>::V;
    static_assert(CH % ZRING == 0, "ring depth must divide the chunk");
    extern __shared__ __align__(128) double wbuf[];             // ST x CH x GD
    __shared__ __align__(8) unsigned long long bars[ST];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, gq = lane >> 2, q = lane & 3;
    const int g_begin = blockIdx.y * groups_per_split;
    const int g_end = min(ngroups, g_begin + groups_per_split);
    const int nch = (g_end - g_begin + CH - 1) / CH;
    const int f0 = (blockIdx.x * NW + warp) * 8 * MT;

    if (tid == 0) {
        for (int s = 0; s < ST; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[s])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](int c) {       // thread 0: chunk c of this split -> stage c % ST
        const int s = c % ST, g0 = g_begin + c * CH, ng = min(CH, g_end - g0);
        const unsigned bytes = (unsigned)ng * GD * 8u, bar = smem_u32(&bars[s]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(wbuf + (size_t)s * CH * GD)), "l"(Wf + (size_t)g0 * GD), "r"(bytes), "r"(bar) : "memory");
    };
    if (tid == 0) for (int c = 0; c < ST && c < nch; ++c) issue(c);

    // frame rows of this lane (rows past the last frame alias the last one: loaded, never stored)
    const double *fr[MT];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) fr[mt] = frames + (size_t)min(f0 + 8 * mt + gq, nframes - 1) * npix + 2 * q;
    double acc[MT][NT][2];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;

    // register ring: frame data of the next ZRING groups, group descriptors of the next 2 ZRING groups (so that the
    // address of a frame load never waits on the descriptor load in front of it)
    double2 ring[ZRING][MT];
    unsigned rinfo[ZRING], ninfo[ZRING];
    auto load_info = [&](int g) { return __ldg(ginfo + min(g, ngroups - 1)); };
    auto load_frames = [&](int d, unsigned info) {
        const size_t off = (size_t)(info & 0xFFFFFFu) * 8;
        if (VEC) {
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) ring[d][mt] = __ldcs(reinterpret_cast<const double2 *>(fr[mt] + off));
        } else {        // rows that are not 16-byte aligned (odd row length): two scalar loads, never past the end of a row
            const unsigned mk = info >> 24;
            const bool p0 = (mk >> (2 * q)) & 1u, p1 = (mk >> (2 * q + 1)) & 1u;
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
                ring[d][mt].x = p0 ? __ldcs(fr[mt] + off) : 0.0;
                ring[d][mt].y = p1 ? __ldcs(fr[mt] + off + 1) : 0.0;
            }
        }
    };
#pragma unroll
    for (int d = 0; d < ZRING; ++d) { rinfo[d] = load_info(g_begin + d); ninfo[d] = load_info(g_begin + ZRING + d); }
#pragma unroll
    for (int d = 0; d < ZRING; ++d) load_frames(d, rinfo[d]);

    for (int c = 0; c < nch; ++c) {
        const int s = c % ST;
        const unsigned bar = smem_u32(&bars[s]), parity = (unsigned)((c / ST) & 1);
        while (!mbar_try_wait(bar, parity)) { }
        const double *wb = wbuf + (size_t)s * CH * GD + 2 * lane;
        const int g0 = g_begin + c * CH;
#pragma unroll 1
        for (int gi = 0; gi < CH; gi += ZRING) {
            // DRAM -> L2 prefetch, ZPF groups ahead: lane l < 8 MT owns frame row f0 + l and pulls the 128-byte lines of
            // the next 4 groups of that frame in one burst (neighbouring lines of one DRAM page arrive together)
            if (ZPF > 0 && lane < 8 * MT) {
                const double *row = frames + (size_t)min(f0 + lane, nframes - 1) * npix;
#pragma unroll
                for (int d = 0; d < ZRING; d += 2) {
                    const unsigned pi = __ldg(ginfo + min(g0 + gi + ZPF + d, ngroups - 1));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(row + (size_t)(pi & 0xFFFFFFu) * 8));
                }
            }
#pragma unroll
            for (int d = 0; d < ZRING; ++d) {
                const int g = g0 + gi + d;
                if (g < g_end) {
                    double2 av[MT];
                    const unsigned mk = rinfo[d] >> 24;
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) av[mt] = ring[d][mt];
                    if (mk != 0xFFu) {                   // pupil edge: pixels outside may hold NaN -> exact zeros
#pragma unroll
                        for (int mt = 0; mt < MT; ++mt) {
                            if (!((mk >> (2 * q)) & 1u)) av[mt].x = 0.0;
                            if (!((mk >> (2 * q + 1)) & 1u)) av[mt].y = 0.0;
                        }
                    }
                    rinfo[d] = ninfo[d];
                    load_frames(d, rinfo[d]);            // group g + ZRING
                    ninfo[d] = load_info(g + 2 * ZRING);
                    const double *wg = wb + (size_t)(gi + d) * GD;
                    double2 b[NT];
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) b[nt] = *reinterpret_cast<const double2 *>(wg + nt * 64);
                    // all first k-steps, then all second k-steps: NT * MT independent DMMAs between dependent ones
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                        for (int mt = 0; mt < MT; ++mt) zdmma(acc[mt][nt], av[mt].x, b[nt].x);
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                        for (int mt = 0; mt < MT; ++mt) zdmma(acc[mt][nt], av[mt].y, b[nt].y);
                }
            }
        }
        __syncthreads();                                 // every warp is done with stage s
        if (tid == 0 && c + ST < nch) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            issue(c + ST);
        }
    }
    double *o = out + (size_t)blockIdx.y * out_split_stride;
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
        const int f = f0 + 8 * mt + gq;
        if (f < nframes) {
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int j = 8 * nt + 2 * q + e;
                    if (j < out_ld && (out_ld != nmodes || j < nmodes))
                        o[(size_t)f * out_ld + j] = (offv && j < nmodes) ? acc[mt][nt][e] - offv[j] : acc[mt][nt][e];
                }
        }
    }
}

// coef[f][j] = sum over splits of part[split][f][j]   (fixed order: deterministic)
__global__ void zmf_reduce_kernel(const double *__restrict__ part, int ksplit, size_t split_stride, int ld, double *__restrict__ coef,
                                  int nmodes, int nframes, const double *__restrict__ offv)
{
    const size_t tot = (size_t)nframes * nmodes;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += (size_t)gridDim.x * blockDim.x) {
        const size_t f = e / nmodes;
        const int j = (int)(e - f * nmodes);
        double s = 0.0;
        for (int k = 0; k < ksplit; ++k) s += part[(size_t)k * split_stride + f * ld + j];
        coef[e] = offv ? s - offv[j] : s;
    }
}

template <int NT, int MT, int NW, bool VEC, int ST>
cudaError_t zmf_launch_dmma(const zmf_handle *h, int nframes, const double *frames, double *out, int out_ld, size_t split_stride,
                            int ksplit, int gps, cudaStream_t st)
{
    constexpr int CH = ZChunk<NT>::CH;
    const size_t smem = (size_t)ST * CH * NT * 64 * 8;
    static bool attr_done_dev[64] = {};          // the attribute is per device
    bool &attr_done = attr_done_dev[h->device & 63];
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(zmf_fit_dmma_kernel<NT, MT, NW, VEC, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr_done = true;
    }
    dim3 grid((nframes + NW * 8 * MT - 1) / (NW * 8 * MT), ksplit);
    zmf_fit_dmma_kernel<NT, MT, NW, VEC, ST><<<grid, NW * 32, smem, st>>>(h->d_Wf, h->d_ginfo, h->ngroups, gps, frames, out, out_ld, split_stride,
                                                                      h->npix, h->nmodes, nframes, ksplit > 1 ? nullptr : h->d_off);
    return cudaGetLastError();
}

template <int NT>
cudaError_t zmf_launch_nt(const zmf_handle *h, int mt, int nframes, const double *frames, double *out, int out_ld, size_t split_stride,
                          int ksplit, int gps, cudaStream_t st)
{
    if (mt < 0) return zmf_launch_dmma<NT, 1, 8, false, 3>(h, nframes, frames, out, out_ld, split_stride, ksplit, gps, st);
    if (mt == 3) return zmf_launch_dmma<NT, 2, 8, true, 2>(h, nframes, frames, out, out_ld, split_stride, ksplit, gps, st);   // large batches
    if (mt == 2) return zmf_launch_dmma<NT, 2, 8, true, 3>(h, nframes, frames, out, out_ld, split_stride, ksplit, gps, st);
    return zmf_launch_dmma<NT, 1, 8, true, 3>(h, nframes, frames, out, out_ld, split_stride, ksplit, gps, st);
}

} // namespace


// ---------------------------------------------------------------------------------------------
// Synthesis (the step after the path, README.md:592-598:  phase_cor = sum_j ad_cor(j) Z_j):
//   frames[f][p] = sum_j coef[f][j] Z[p][j]   inside the pupil, 0 outside.
// HBM-write bound (8 nL^2 bytes per frame out, 8 nmodes in) and, at 2 nmodes flop per 8 bytes, again next to the FP64
// ridge: M = 8 frames, N = 8 pixels, K = nmodes on the FP64 tensor pipe.  The coefficient fragments of a warp's frame
// tiles stay in registers; Z' comes pre-permuted in fragment order through the same bulk-copy stages as W in the fit
// kernel; every lane stores 16 bytes, 4 lanes one 64-byte run of a frame (streaming stores).
// ---------------------------------------------------------------------------------------------
namespace {

template <int KS, int MT, int NW>
__global__ void __launch_bounds__(NW * 32) zmf_synth_dmma_kernel(const double *__restrict__ Zf, const unsigned char *__restrict__ gmask, int ngroups,
                                                                 int groups_per_split, const double *__restrict__ coef,
                                                                 double *__restrict__ frames, int npix, int nmodes, int nframes)
{
    constexpr int CH = (KS <= 9) ? ZSYN_CH : 8, GD = KS * 32;
    extern __shared__ __align__(128) double zbuf[];             // ZSYN_STAGES x CH x GD
    __shared__ __align__(8) unsigned long long bars[ZSYN_STAGES];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, gq = lane >> 2, q = lane & 3;
    const int g_begin = blockIdx.y * groups_per_split;
    const int g_end = min(ngroups, g_begin + groups_per_split);
    const int nch = (g_end - g_begin + CH - 1) / CH;
    const int f0 = (blockIdx.x * NW + warp) * 8 * MT;
    if (tid == 0) {
        for (int s = 0; s < ZSYN_STAGES; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[s])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](int c) {
        const int s = c % ZSYN_STAGES, g0 = g_begin + c * CH, ng = min(CH, g_end - g0);
        const unsigned bytes = (unsigned)ng * GD * 8u, bar = smem_u32(&bars[s]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(zbuf + (size_t)s * CH * GD)), "l"(Zf + (size_t)g0 * GD), "r"(bytes), "r"(bar) : "memory");
    };
    if (tid == 0) for (int c = 0; c < ZSYN_STAGES && c < nch; ++c) issue(c);

    // A fragments: coef[frame f0 + 8 mt + gq][mode 4 s + q], zero beyond nmodes / nframes
    double a[MT][KS];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
        const int f = f0 + 8 * mt + gq;
#pragma unroll
        for (int s = 0; s < KS; ++s) {
            const int j = 4 * s + q;
            a[mt][s] = (f < nframes && j < nmodes) ? __ldg(coef + (size_t)f * nmodes + j) : 0.0;
        }
    }
    double *orow[MT];
    bool ok[MT];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
        const int f = f0 + 8 * mt + gq;
        ok[mt] = f < nframes;
        orow[mt] = frames + (size_t)min(f, nframes - 1) * npix + 2 * q;
    }
    for (int c = 0; c < nch; ++c) {
        const int s = c % ZSYN_STAGES;
        const unsigned bar = smem_u32(&bars[s]), parity = (unsigned)((c / ZSYN_STAGES) & 1);
        while (!mbar_try_wait(bar, parity)) { }
        const double *zb = zbuf + (size_t)s * CH * GD + lane;
        const int g0 = g_begin + c * CH;
#pragma unroll 2
        for (int gi = 0; gi < CH; ++gi) {
            const int g = g0 + gi;
            if (g < g_end) {
                double acc[MT][2];
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) acc[mt][0] = acc[mt][1] = 0.0;
                if (__ldg(gmask + g)) {                  // groups wholly outside the pupil are plain zero stores
                    const double *zg = zb + (size_t)gi * GD;
#pragma unroll
                    for (int ks = 0; ks < KS; ++ks) {
                        const double b = zg[32 * ks];
#pragma unroll
                        for (int mt = 0; mt < MT; ++mt) zdmma(acc[mt], a[mt][ks], b);
                    }
                }
#pragma unroll
                for (int mt = 0; mt < MT; ++mt)
                    if (ok[mt]) __stcs(reinterpret_cast<double2 *>(orow[mt] + (size_t)g * 8), make_double2(acc[mt][0], acc[mt][1]));
            }
        }
        __syncthreads();
        if (tid == 0 && c + ZSYN_STAGES < nch) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            issue(c + ZSYN_STAGES);
        }
    }
}

// generic synthesis (frame length not a multiple of 8, or > 72 modes): one thread per pixel
__global__ void zmf_synth_kernel(const double *__restrict__ Z, const unsigned char *__restrict__ mask, const int *__restrict__ pin,
                                 const double *__restrict__ coef, double *__restrict__ frames, int npix, int npix_in, int nmodes, int nframes)
{
    const size_t tot = (size_t)nframes * npix;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += (size_t)gridDim.x * blockDim.x) {
        const size_t f = e / npix;
        const int p = (int)(e - f * npix);
        double s = 0.0;
        if (mask[p]) {
            const int pi = pin[p];
            for (int j = 0; j < nmodes; ++j) s = fma(__ldg(coef + f * nmodes + j), __ldg(Z + (size_t)j * npix_in + pi), s);
        }
        frames[e] = s;
    }
}

template <int KS, int MT>
cudaError_t zmf_launch_synth(const zmf_handle *h, int nframes, const double *coef, double *frames, int ksplit, int gps, cudaStream_t st)
{
    constexpr int CH = (KS <= 9) ? ZSYN_CH : 8, NW = 8;
    const size_t smem = (size_t)ZSYN_STAGES * CH * KS * 32 * 8;
    static bool attr_done_dev[64] = {};          // the attribute is per device
    bool &attr_done = attr_done_dev[h->device & 63];
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(zmf_synth_dmma_kernel<KS, MT, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr_done = true;
    }
    dim3 grid((nframes + NW * 8 * MT - 1) / (NW * 8 * MT), ksplit);
    zmf_synth_dmma_kernel<KS, MT, NW><<<grid, NW * 32, smem, st>>>(h->d_Zf, h->d_gmask, h->npix / 8, gps, coef, frames, h->npix, h->nmodes, nframes);
    return cudaGetLastError();
}

} // namespace

// Uploads the least-squares operator W (nmodes x npix, row j contiguous over pixels, exact zeros where mask == 0) and
// builds the DMMA tables; allocates staging for max_frames.  Shared by zmf_create and est_create.
// W (M x ldw, row j contiguous over samples; sample p lands in column pidx[p], or p if pidx == NULL) such that W y is the
// least-squares solution of Z c = y, Z = P x M column-major:  W = inv(Z'Z) Z' by a Cholesky of the Gram matrix in extended
// precision.  If the Gram matrix is numerically rank deficient: FMPC_ERR_NOT_PD, unless `minnorm`, in which case
// W = pinv(Z'Z) Z' -- what lsqminnorm(Z'Z, Z'y) returns (README.md:478): eigen-decomposition of the Gram matrix (cyclic
// Jacobi, extended precision), eigenvalues below lsqminnorm's default tolerance max(size) * eps(norm) dropped.
// `off` (if given) receives W b for the estimator's constant term.
static int lsq_operator(const std::vector<double> &Z, int P, int M, const int *pidx, size_t ldw, bool minnorm, const double *b,
                        std::vector<double> &W, std::vector<double> *off)
{
    std::vector<long double> G((size_t)M * M, 0.0L);
    for (int a = 0; a < M; ++a)
        for (int c = a; c < M; ++c) {
            long double sacc = 0.0L;
            const double *za = &Z[(size_t)a * P], *zc = &Z[(size_t)c * P];
            for (int p = 0; p < P; ++p) sacc += (long double)za[p] * zc[p];
            G[(size_t)a * M + c] = G[(size_t)c * M + a] = sacc;
        }
    std::vector<long double> Gi;                // pinv of the Gram matrix, only in the minimum-norm branch
    std::vector<long double> L = G;
    bool pd = true;
    for (int j = 0; j < M && pd; ++j) {         // in-place lower Cholesky, row-major
        const long double g0 = G[(size_t)j * M + j];
        long double d = L[(size_t)j * M + j];
        for (int k = 0; k < j; ++k) d -= L[(size_t)j * M + k] * L[(size_t)j * M + k];
        if (!(d > 1e-13L * g0) || !(g0 > 0.0L)) { pd = false; break; }
        d = sqrtl(d);
        L[(size_t)j * M + j] = d;
        for (int i = j + 1; i < M; ++i) {
            long double sacc = L[(size_t)i * M + j];
            for (int k = 0; k < j; ++k) sacc -= L[(size_t)i * M + k] * L[(size_t)j * M + k];
            L[(size_t)i * M + j] = sacc / d;
        }
    }
    if (!pd) {
        if (!minnorm) return FMPC_ERR_NOT_PD;
        std::vector<long double> A = G, V((size_t)M * M, 0.0L);
        for (int i = 0; i < M; ++i) V[(size_t)i * M + i] = 1.0L;
        for (int sweep = 0; sweep < 60; ++sweep) {
            long double offn = 0.0L, dn = 0.0L;
            for (int i = 0; i < M; ++i) for (int j = 0; j < M; ++j) (i == j ? dn : offn) += A[(size_t)i * M + j] * A[(size_t)i * M + j];
            if (offn <= 1e-38L * dn) break;
            for (int pp = 0; pp < M - 1; ++pp)
                for (int q = pp + 1; q < M; ++q) {
                    const long double apq = A[(size_t)pp * M + q];
                    if (apq == 0.0L) continue;
                    const long double th = (A[(size_t)q * M + q] - A[(size_t)pp * M + pp]) / (2.0L * apq);
                    const long double t = (th >= 0 ? 1.0L : -1.0L) / (fabsl(th) + sqrtl(th * th + 1.0L));
                    const long double c = 1.0L / sqrtl(t * t + 1.0L), sn = t * c;
                    for (int k = 0; k < M; ++k) {
                        const long double akp = A[(size_t)k * M + pp], akq = A[(size_t)k * M + q];
                        A[(size_t)k * M + pp] = c * akp - sn * akq; A[(size_t)k * M + q] = sn * akp + c * akq;
                    }
                    for (int k = 0; k < M; ++k) {
                        const long double apk = A[(size_t)pp * M + k], aqk = A[(size_t)q * M + k];
                        A[(size_t)pp * M + k] = c * apk - sn * aqk; A[(size_t)q * M + k] = sn * apk + c * aqk;
                    }
                    for (int k = 0; k < M; ++k) {
                        const long double vkp = V[(size_t)k * M + pp], vkq = V[(size_t)k * M + q];
                        V[(size_t)k * M + pp] = c * vkp - sn * vkq; V[(size_t)k * M + q] = sn * vkp + c * vkq;
                    }
                }
        }
        long double lmax = 0.0L;
        for (int i = 0; i < M; ++i) lmax = fmaxl(lmax, fabsl(A[(size_t)i * M + i]));
        const long double tol = (long double)M * 2.220446049250313e-16L * lmax;       // max(size(G)) * eps(norm(G))
        Gi.assign((size_t)M * M, 0.0L);
        for (int e = 0; e < M; ++e) {
            const long double lam = A[(size_t)e * M + e];
            if (!(lam > tol)) continue;
            for (int i = 0; i < M; ++i)
                for (int j = 0; j < M; ++j) Gi[(size_t)i * M + j] += V[(size_t)i * M + e] * V[(size_t)j * M + e] / lam;
        }
    }
    W.assign((size_t)M * ldw, 0.0);
    std::vector<long double> col(M), res(M), offl(M, 0.0L);
    for (int p = 0; p < P; ++p) {       // column p of W: solve (Z'Z) w = Z(p,:)'
        for (int j = 0; j < M; ++j) col[j] = Z[(size_t)j * P + p];
        if (pd) {
            for (int j = 0; j < M; ++j) { long double sacc = col[j]; for (int k = 0; k < j; ++k) sacc -= L[(size_t)j * M + k] * col[k]; col[j] = sacc / L[(size_t)j * M + j]; }
            for (int j = M - 1; j >= 0; --j) { long double sacc = col[j]; for (int k = j + 1; k < M; ++k) sacc -= L[(size_t)k * M + j] * col[k]; col[j] = sacc / L[(size_t)j * M + j]; }
            res = col;
        } else {
            for (int j = 0; j < M; ++j) { long double sacc = 0.0L; for (int k = 0; k < M; ++k) sacc += Gi[(size_t)j * M + k] * col[k]; res[j] = sacc; }
        }
        const size_t cidx = pidx ? (size_t)pidx[p] : (size_t)p;
        for (int j = 0; j < M; ++j) { W[(size_t)j * ldw + cidx] = (double)res[j]; if (b) offl[j] += res[j] * (long double)b[p]; }
    }
    if (off) { off->assign(M, 0.0); for (int j = 0; j < M; ++j) (*off)[j] = (double)offl[j]; }
    return FMPC_OK;
}

static bool linfit_upload(zmf_handle *h, const std::vector<double> &W, int sm_count, int max_frames)
{
    const int M = h->nmodes;
    bool ok = cudaMalloc(&h->d_W, W.size() * 8) == cudaSuccess && cudaMalloc(&h->d_mask, h->npix) == cudaSuccess;
    ok = ok && cudaMemcpy(h->d_W, W.data(), W.size() * 8, cudaMemcpyHostToDevice) == cudaSuccess;
    ok = ok && cudaMemcpy(h->d_mask, h->mask.data(), h->npix, cudaMemcpyHostToDevice) == cudaSuccess;
    {   // DMMA path tables (<= 72 modes)
        h->sm_count = sm_count;
        h->ntiles = (M + 7) / 8;
        if (h->ntiles <= 9) {
            std::vector<unsigned> ginfo;
            for (int g = 0; g < (h->npix + 7) / 8; ++g) {
                unsigned mk = 0;
                for (int e = 0; e < 8; ++e) if ((size_t)8 * g + e < (size_t)h->npix && h->mask[(size_t)8 * g + e]) mk |= 1u << e;
                if (mk) ginfo.push_back((unsigned)g | (mk << 24));
            }
            h->ngroups = (int)ginfo.size();
            const int NTl = h->ntiles;
            std::vector<double> Wf((size_t)h->ngroups * NTl * 64, 0.0);
            for (int gi = 0; gi < h->ngroups; ++gi) {
                const size_t p0 = (size_t)(ginfo[gi] & 0xFFFFFFu) * 8;
                for (int nt = 0; nt < NTl; ++nt)
                    for (int ln = 0; ln < 32; ++ln)
                        for (int e = 0; e < 2; ++e) {
                            const int j = 8 * nt + (ln >> 2);
                            const size_t px = p0 + 2 * (ln & 3) + e;
                            Wf[(((size_t)gi * NTl + nt) * 32 + ln) * 2 + e] = (j < M && px < (size_t)h->npix && h->mask[px]) ? W[(size_t)j * h->npix + px] : 0.0;
                        }
            }
            ok = ok && cudaMalloc(&h->d_Wf, Wf.size() * 8) == cudaSuccess && cudaMalloc(&h->d_ginfo, ginfo.size() * 4 + 16) == cudaSuccess;
            ok = ok && cudaMemcpy(h->d_Wf, Wf.data(), Wf.size() * 8, cudaMemcpyHostToDevice) == cudaSuccess;
            ok = ok && cudaMemcpy(h->d_ginfo, ginfo.data(), ginfo.size() * 4, cudaMemcpyHostToDevice) == cudaSuccess;
        }
    }
    h->cap_frames = (size_t)max_frames;
    ok = ok && cudaMalloc(&h->d_frames, h->cap_frames * h->npix * 8) == cudaSuccess;
    ok = ok && cudaMalloc(&h->d_coef, h->cap_frames * M * 8) == cudaSuccess;
    ok = ok && cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaEventCreate(&h->ev0) == cudaSuccess && cudaEventCreate(&h->ev1) == cudaSuccess;
    return ok;
}

extern "C" {

int zmf_create(zmf_handle **out, int nL, int N, int max_frames, int device)
{
    if (!out) return FMPC_ERR_NULL;
    *out = nullptr;
    if (nL < 2 || nL > 4096 || N < 0 || N > 40 || max_frames < 1) return FMPC_ERR_DIM;
    int cnt = 0;
    if (cudaGetDeviceCount(&cnt) != cudaSuccess || device < 0 || device >= cnt) return FMPC_ERR_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major != 10) return FMPC_ERR_CUDA;
    if (cudaSetDevice(device) != cudaSuccess) return FMPC_ERR_CUDA;
    zmf_handle *h = new (std::nothrow) zmf_handle();
    if (!h) return FMPC_ERR_CUDA;
    h->device = device; h->nL = nL; h->N = N; h->npix = nL * nL;
    h->nmodes = (N + 1) * (N + 2) / 2;
    // mode index vectors, zernmodfit.m:195-198: n = 0..N, m = -n:2:n
    std::vector<int> nn, mm;
    for (int x = 0; x <= N; ++x) for (int m = -x; m <= x; m += 2) { nn.push_back(x); mm.push_back(m); }
    // pupil grid, README.md:78-84; column-major linear index p = row + nL*col, X varies with col
    std::vector<double> xs(nL);
    for (int i = 0; i < nL; ++i) xs[i] = (double)(-(nL - 1) + 2 * i) / (double)(nL - 1);
    double xmax = 0.0;
    for (double v : xs) xmax = std::fmax(xmax, std::fabs(v));
    h->mask.assign(h->npix, 0);
    std::vector<double> rr, tt;
    std::vector<int> pidx;
    for (int col = 0; col < nL; ++col)
        for (int row = 0; row < nL; ++row) {
            const double X = xs[col], Y = xs[row];
            const double r = std::hypot(X, Y), th = std::atan2(Y, X);
            if (r <= xmax) { h->mask[(size_t)col * nL + row] = 1; rr.push_back(r); tt.push_back(th); pidx.push_back(col * nL + row); }
        }
    h->npix_in = (int)rr.size();
    const int P = h->npix_in, M = h->nmodes;
    if (P < M) { delete h; return FMPC_ERR_DIM; }
    h->Z.assign((size_t)P * M, 0.0);
    for (int j = 0; j < M; ++j)
        for (int p = 0; p < P; ++p) h->Z[(size_t)j * P + p] = zern_eval(nn[j], mm[j], rr[p], tt[p]);
    std::vector<double> W;
    {
        const int rcw = lsq_operator(h->Z, P, M, pidx.data(), (size_t)h->npix, false, nullptr, W, nullptr);
        if (rcw) { delete h; return rcw; }
    }
    const bool ok = linfit_upload(h, W, prop.multiProcessorCount, max_frames);
    if (!ok) { zmf_destroy(h); return FMPC_ERR_CUDA; }
    *out = h;
    return FMPC_OK;
}

int zmf_create_samples(zmf_handle **out, int npts, const double *r, const double *theta, int N, int max_frames, int device)
{
    if (!out) return FMPC_ERR_NULL;
    *out = nullptr;
    if (!r || !theta) return FMPC_ERR_NULL;
    if (npts < 1 || npts > (1 << 26) || N < 0 || N > 40 || max_frames < 1) return FMPC_ERR_DIM;
    for (int p = 0; p < npts; ++p) if (!(r[p] >= 0.0 && r[p] <= 1.0)) return FMPC_ERR_DIM;       // zernmodfit.m:182-184
    int cnt = 0;
    if (cudaGetDeviceCount(&cnt) != cudaSuccess || device < 0 || device >= cnt) return FMPC_ERR_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major != 10) return FMPC_ERR_CUDA;
    if (cudaSetDevice(device) != cudaSuccess) return FMPC_ERR_CUDA;
    zmf_handle *h = new (std::nothrow) zmf_handle();
    if (!h) return FMPC_ERR_CUDA;
    h->device = device; h->nL = 0; h->N = N; h->npix = npts; h->npix_in = npts;
    h->nmodes = (N + 1) * (N + 2) / 2;
    h->mask.assign(npts, 1);
    const int P = npts, M = h->nmodes;
    if (P < M) { delete h; return FMPC_ERR_DIM; }
    h->Z.assign((size_t)P * M, 0.0);
    int j = 0;
    for (int x = 0; x <= N; ++x)
        for (int m = -x; m <= x; m += 2, ++j)                      // zernmodfit.m:195-198
            for (int p = 0; p < P; ++p) h->Z[(size_t)j * P + p] = zern_eval(x, m, r[p], theta[p]);
    std::vector<double> W;
    const int rcw = lsq_operator(h->Z, P, M, nullptr, (size_t)P, false, nullptr, W, nullptr);
    if (rcw) { delete h; return rcw; }
    if (!linfit_upload(h, W, prop.multiProcessorCount, max_frames)) { zmf_destroy(h); return FMPC_ERR_CUDA; }
    *out = h;
    return FMPC_OK;
}

void zmf_destroy(zmf_handle *h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->d_W) cudaFree(h->d_W);
    if (h->d_mask) cudaFree(h->d_mask);
    if (h->d_Wf) cudaFree(h->d_Wf);
    if (h->d_ginfo) cudaFree(h->d_ginfo);
    if (h->d_part) cudaFree(h->d_part);
    if (h->d_off) cudaFree(h->d_off);
    if (h->d_Zf) cudaFree(h->d_Zf);
    if (h->d_gmask) cudaFree(h->d_gmask);
    if (h->d_Zdev) cudaFree(h->d_Zdev);
    if (h->d_pin) cudaFree(h->d_pin);
    if (h->d_frames) cudaFree(h->d_frames);
    if (h->d_coef) cudaFree(h->d_coef);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

int zmf_nmodes(const zmf_handle *h) { return h ? h->nmodes : FMPC_ERR_NULL; }
int zmf_npix_in(const zmf_handle *h) { return h ? h->npix_in : FMPC_ERR_NULL; }
long long zmf_launch_count(const zmf_handle *h) { return h ? h->launches : 0; }

int zmf_get_basis(const zmf_handle *h, double *Z)
{
    if (!h || !Z) return FMPC_ERR_NULL;
    std::memcpy(Z, h->Z.data(), h->Z.size() * 8);
    return FMPC_OK;
}
int zmf_get_mask(const zmf_handle *h, unsigned char *mask)
{
    if (!h || !mask) return FMPC_ERR_NULL;
    std::memcpy(mask, h->mask.data(), h->mask.size());
    return FMPC_OK;
}

int zmf_fit_d(zmf_handle *h, int nframes, const double *frames, double *coef, void *stream)
{
    if (!h || !frames || !coef) return FMPC_ERR_NULL;
    if (nframes <= 0) return nframes < 0 ? FMPC_ERR_DIM : FMPC_OK;
    if (cudaSetDevice(h->device) != cudaSuccess) return FMPC_ERR_CUDA;
    cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
    if (h->d_Wf) {
        // ---- DMMA path: frame-block size and split-K chosen so that the grid holds >= ~6 units per SM ----
        const bool vec = (h->npix % 8 == 0) && ((uintptr_t)frames & 15) == 0;      // 16-byte aligned rows of whole 8-pixel groups
        int mt = (nframes > 64) ? 2 : 1;                 // two frame tiles per warp share every W fragment
        if (const char *e = getenv("ZMF_MT")) { const int v = atoi(e); if (v == 1 || v == 2) mt = v; }      // experiments
        if (!vec) mt = 1;
        const int fblocks = (nframes + 64 * mt - 1) / (64 * mt);
        constexpr int CHmin = 8;
        // measured (scripts/zmf_sweep.sh, profiles/r01_zmf_sweep.log): ~14 work units per SM is the optimum at 2000 and at
        // 32768 frames (deterministic split-K: partial sums + reduction kernel)
        int ksplit = (14 * h->sm_count + fblocks - 1) / fblocks;
        if (ksplit > 16) ksplit = 16;
        if (const char *e = getenv("ZMF_KSPLIT")) { const int v = atoi(e); if (v >= 1 && v <= 64) ksplit = v; }   // experiments
        const int maxsplit = (h->ngroups + 4 * CHmin - 1) / (4 * CHmin);      // keep >= 4 chunks per split
        if (ksplit > maxsplit) ksplit = maxsplit;
        if (ksplit < 1) ksplit = 1;
        int gps = (h->ngroups + ksplit - 1) / ksplit;
        gps = (gps + 15) & ~15;                                               // whole shared-memory chunks
        ksplit = (h->ngroups + gps - 1) / gps;
        const int ld = 8 * h->ntiles;
        double *out = coef;
        int out_ld = h->nmodes;
        size_t stride = 0;
        if (ksplit > 1) {
            stride = (size_t)nframes * ld;
            const size_t need = stride * ksplit;
            if (need > h->part_doubles) {
                if (h->d_part) cudaFree(h->d_part);
                h->d_part = nullptr; h->part_doubles = 0;
                if (cudaMalloc(&h->d_part, need * 8) != cudaSuccess) return FMPC_ERR_CUDA;
                h->part_doubles = need;
            }
            out = h->d_part; out_ld = ld;
        }
        cudaError_t e = cudaErrorInvalidValue;
        int mtsel = vec ? mt : -1;
        if (mtsel == 2 && nframes > 8192 && !getenv("ZMF_NO_ST2")) mtsel = 3;     // 2-stage / 3-CTA variant
        switch (h->ntiles) {
        case 1: e = zmf_launch_nt<1>(h, mtsel, nframes, frames, out, out_ld, stride, ksplit, gps, st); break;
        case 2: e = zmf_launch_nt<2>(h, mtsel, nframes, frames, out, out_ld, stride, ksplit, gps, st); break;
        case 3: e = zmf_launch_nt<3>(h, mtsel, nframes, frames, out, out_ld, stride, ksplit, gps, st); break;
        case 4: e = zmf_launch_nt<4>(h, mtsel, nframes, frames, out, out_ld, stride, ksplit, gps, st); break;
        case 5: e = zmf_launch_nt<5>(h, mtsel, nframes, frames, out, out_ld, stride, ksplit, gps, st); break;
        case 6: e = zmf_launch_nt<6>(h, mtsel, nframes, frames, out, out_ld, stride, ksplit, gps, st); break;
        case 7: e = zmf_launch_nt<7>(h, mtsel, nframes, frames, out, out_ld, stride, ksplit, gps, st); break;
        case 8: e = zmf_launch_nt<8>(h, mtsel, nframes, frames, out, out_ld, stride, ksplit, gps, st); break;
        case 9: e = zmf_launch_nt<9>(h, mtsel, nframes, frames, out, out_ld, stride, ksplit, gps, st); break;
        default: break;
        }
        if (e != cudaSuccess) return FMPC_ERR_CUDA;
        h->launches += 1;
        if (ksplit > 1) {
            const size_t tot = (size_t)nframes * h->nmodes;
            int rg = (int)((tot + 255) / 256);
            if (rg > h->sm_count * 8) rg = h->sm_count * 8;
            zmf_reduce_kernel<<<rg, 256, 0, st>>>(h->d_part, ksplit, stride, ld, coef, h->nmodes, nframes, h->d_off);
            if (cudaGetLastError() != cudaSuccess) return FMPC_ERR_CUDA;
            h->launches += 1;
        }
        return FMPC_OK;
    }
    // generic path (odd frame sizes, > 72 modes): scalar FMA kernel
    constexpr int MC = 7, FT = 4, NT = 256;
    const int grid = (nframes + FT - 1) / FT;
    zmf_fit_kernel<MC, FT, NT><<<grid, NT, 0, st>>>(h->d_W, h->d_mask, frames, coef, h->npix, h->nmodes, nframes, h->d_off);
    if (cudaGetLastError() != cudaSuccess) return FMPC_ERR_CUDA;
    h->launches += 1;
    return FMPC_OK;
}

int zmf_fit(zmf_handle *h, int nframes, const double *frames, double *coef, double *telapsed)
{
    if (!h || !frames || !coef) return FMPC_ERR_NULL;
    if (telapsed) *telapsed = 0.0;
    if (nframes <= 0) return nframes < 0 ? FMPC_ERR_DIM : FMPC_OK;
    if ((size_t)nframes > h->cap_frames) return FMPC_ERR_BATCH;
    if (cudaSetDevice(h->device) != cudaSuccess) return FMPC_ERR_CUDA;
    cudaStream_t st = h->stream;
    if (cudaMemcpyAsync(h->d_frames, frames, (size_t)nframes * h->npix * 8, cudaMemcpyHostToDevice, st) != cudaSuccess) return FMPC_ERR_CUDA;
    cudaEventRecord(h->ev0, st);
    int rc = zmf_fit_d(h, nframes, h->d_frames, h->d_coef, st);
    if (rc) return rc;
    cudaEventRecord(h->ev1, st);
    if (cudaMemcpyAsync(coef, h->d_coef, (size_t)nframes * h->nmodes * 8, cudaMemcpyDeviceToHost, st) != cudaSuccess) return FMPC_ERR_CUDA;
    if (cudaStreamSynchronize(st) != cudaSuccess) return FMPC_ERR_CUDA;
    if (telapsed) { float ms = 0.f; cudaEventElapsedTime(&ms, h->ev0, h->ev1); *telapsed = ms * 1e-3; }
    return FMPC_OK;
}


/* ---- synthesis: frames = Z coef inside the pupil, 0 outside (README.md:592-598) ---- */
static int zmf_synth_prepare(zmf_handle *h)
{
    if (h->d_Zf || h->d_Zdev) return FMPC_OK;
    const int P = h->npix_in, M = h->nmodes;
    std::vector<int> pin(h->npix, -1);
    { int k = 0; for (int p = 0; p < h->npix; ++p) if (h->mask[p]) pin[p] = k++; }
    if (h->npix % 8 == 0 && M <= 72) {
        const int KS = (M <= 28) ? 7 : 18, ng = h->npix / 8;
        std::vector<double> Zf((size_t)ng * KS * 32, 0.0);
        std::vector<unsigned char> gm(ng, 0);
        for (int g = 0; g < ng; ++g)
            for (int ks = 0; ks < KS; ++ks)
                for (int ln = 0; ln < 32; ++ln) {
                    const int px = 8 * g + (ln >> 2), j = 4 * ks + (ln & 3);
                    if (h->mask[px]) { gm[g] = 1; if (j < M) Zf[((size_t)g * KS + ks) * 32 + ln] = h->Z[(size_t)j * P + pin[px]]; }
                }
        if (cudaMalloc(&h->d_Zf, Zf.size() * 8) != cudaSuccess || cudaMalloc(&h->d_gmask, gm.size() + 16) != cudaSuccess) return FMPC_ERR_CUDA;
        if (cudaMemcpy(h->d_Zf, Zf.data(), Zf.size() * 8, cudaMemcpyHostToDevice) != cudaSuccess ||
            cudaMemcpy(h->d_gmask, gm.data(), gm.size(), cudaMemcpyHostToDevice) != cudaSuccess) return FMPC_ERR_CUDA;
        h->syn_ks = KS;
    } else {
        if (cudaMalloc(&h->d_Zdev, h->Z.size() * 8) != cudaSuccess || cudaMalloc(&h->d_pin, (size_t)h->npix * 4) != cudaSuccess) return FMPC_ERR_CUDA;
        if (cudaMemcpy(h->d_Zdev, h->Z.data(), h->Z.size() * 8, cudaMemcpyHostToDevice) != cudaSuccess ||
            cudaMemcpy(h->d_pin, pin.data(), (size_t)h->npix * 4, cudaMemcpyHostToDevice) != cudaSuccess) return FMPC_ERR_CUDA;
    }
    return FMPC_OK;
}

int zmf_synth_d(zmf_handle *h, int nframes, const double *coef, double *frames, void *stream)
{
    if (!h || !frames || !coef) return FMPC_ERR_NULL;
    if (nframes <= 0) return nframes < 0 ? FMPC_ERR_DIM : FMPC_OK;
    if (cudaSetDevice(h->device) != cudaSuccess) return FMPC_ERR_CUDA;
    int rc = zmf_synth_prepare(h);
    if (rc) return rc;
    cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
    if (h->d_Zf && ((uintptr_t)frames & 15) == 0) {
        const int mt = (nframes > 4096) ? 2 : 1;
        const int fblocks = (nframes + 64 * mt - 1) / (64 * mt), ng = h->npix / 8;
        int ksplit = (6 * h->sm_count + fblocks - 1) / fblocks;      // split over pixel groups: independent outputs, no reduction
        if (ksplit > 32) ksplit = 32;
        if (ksplit > (ng + 63) / 64) ksplit = (ng + 63) / 64;
        if (ksplit < 1) ksplit = 1;
        int gps = (ng + ksplit - 1) / ksplit;
        gps = (gps + 15) & ~15;
        ksplit = (ng + gps - 1) / gps;
        cudaError_t e;
        if (h->syn_ks == 7) e = (mt == 2) ? zmf_launch_synth<7, 2>(h, nframes, coef, frames, ksplit, gps, st) : zmf_launch_synth<7, 1>(h, nframes, coef, frames, ksplit, gps, st);
        else e = (mt == 2) ? zmf_launch_synth<18, 2>(h, nframes, coef, frames, ksplit, gps, st) : zmf_launch_synth<18, 1>(h, nframes, coef, frames, ksplit, gps, st);
        if (e != cudaSuccess) return FMPC_ERR_CUDA;
    } else {
        if (!h->d_Zdev) {       // DMMA tables exist but the output is not 16-byte aligned: build the generic tables too
            std::vector<int> pin(h->npix, -1);
            { int k = 0; for (int p = 0; p < h->npix; ++p) if (h->mask[p]) pin[p] = k++; }
            if (cudaMalloc(&h->d_Zdev, h->Z.size() * 8) != cudaSuccess || cudaMalloc(&h->d_pin, (size_t)h->npix * 4) != cudaSuccess) return FMPC_ERR_CUDA;
            if (cudaMemcpy(h->d_Zdev, h->Z.data(), h->Z.size() * 8, cudaMemcpyHostToDevice) != cudaSuccess ||
                cudaMemcpy(h->d_pin, pin.data(), (size_t)h->npix * 4, cudaMemcpyHostToDevice) != cudaSuccess) return FMPC_ERR_CUDA;
        }
        const size_t tot = (size_t)nframes * h->npix;
        int grid = (int)((tot + 255) / 256);
        if (grid > h->sm_count * 16) grid = h->sm_count * 16;
        zmf_synth_kernel<<<grid, 256, 0, st>>>(h->d_Zdev, h->d_mask, h->d_pin, coef, frames, h->npix, h->npix_in, h->nmodes, nframes);
        if (cudaGetLastError() != cudaSuccess) return FMPC_ERR_CUDA;
    }
    h->launches += 1;
    return FMPC_OK;
}

int zmf_synth(zmf_handle *h, int nframes, const double *coef, double *frames, double *telapsed)
{
    if (!h || !frames || !coef) return FMPC_ERR_NULL;
    if (telapsed) *telapsed = 0.0;
    if (nframes <= 0) return nframes < 0 ? FMPC_ERR_DIM : FMPC_OK;
    if ((size_t)nframes > h->cap_frames) return FMPC_ERR_BATCH;
    if (cudaSetDevice(h->device) != cudaSuccess) return FMPC_ERR_CUDA;
    cudaStream_t st = h->stream;
    if (cudaMemcpyAsync(h->d_coef, coef, (size_t)nframes * h->nmodes * 8, cudaMemcpyHostToDevice, st) != cudaSuccess) return FMPC_ERR_CUDA;
    cudaEventRecord(h->ev0, st);
    int rc = zmf_synth_d(h, nframes, h->d_coef, h->d_frames, st);
    if (rc) return rc;
    cudaEventRecord(h->ev1, st);
    if (cudaMemcpyAsync(frames, h->d_frames, (size_t)nframes * h->npix * 8, cudaMemcpyDeviceToHost, st) != cudaSuccess) return FMPC_ERR_CUDA;
    if (cudaStreamSynchronize(st) != cudaSuccess) return FMPC_ERR_CUDA;
    if (telapsed) { float ms = 0.f; cudaEventElapsedTime(&ms, h->ev0, h->ev1); *telapsed = ms * 1e-3; }
    return FMPC_OK;
}

// =============================================================================================
// Estimator step of the closed loop (README.md:478):  x_hat = lsqminnorm(A_s'A_s, A_s'(y - b_s)).
// A_s'A_s is a fixed nmodes x nmodes SPD matrix (cond(A_s) = 9.3 for the reference's model_approx.mat), so
// x_hat = W y - W b_s with W = inv(A_s'A_s) A_s' built once in extended precision on the host; the device part
// is the same streaming GEMM as zernmodfit (all pixels inside the "pupil", no 16-byte row alignment assumed).
// =============================================================================================
int est_create(est_handle **out, int npix, int nmodes, const double *A_s, const double *b_s, int max_batch, int device)
{
    if (!out) return FMPC_ERR_NULL;
    *out = nullptr;
    if (!A_s) return FMPC_ERR_NULL;
    if (npix < 1 || nmodes < 1 || nmodes > npix || npix > (1 << 26) || max_batch < 1) return FMPC_ERR_DIM;
    int cnt = 0;
    if (cudaGetDeviceCount(&cnt) != cudaSuccess || device < 0 || device >= cnt) return FMPC_ERR_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major != 10) return FMPC_ERR_CUDA;
    if (cudaSetDevice(device) != cudaSuccess) return FMPC_ERR_CUDA;
    zmf_handle *h = new (std::nothrow) zmf_handle();
    if (!h) return FMPC_ERR_CUDA;
    h->device = device; h->nL = 0; h->N = -1; h->npix = npix; h->npix_in = npix; h->nmodes = nmodes;
    h->mask.assign(npix, 1);
    const int P = npix, M = nmodes;
    h->Z.assign(A_s, A_s + (size_t)P * M);                 // column-major npix x nmodes, like the Zernike basis
    // rank-deficient A_s: lsqminnorm's minimum-norm branch (README.md:478)
    std::vector<double> W, off;
    {
        const int rcw = lsq_operator(h->Z, P, M, nullptr, (size_t)P, true, b_s, W, &off);
        if (rcw) { delete h; return rcw; }
    }
    bool ok = linfit_upload(h, W, prop.multiProcessorCount, max_batch);
    if (ok && b_s) {
        ok = cudaMalloc(&h->d_off, (size_t)M * 8) == cudaSuccess && cudaMemcpy(h->d_off, off.data(), (size_t)M * 8, cudaMemcpyHostToDevice) == cudaSuccess;
    }
    if (!ok) { zmf_destroy(h); return FMPC_ERR_CUDA; }
    *out = h;
    return FMPC_OK;
}

void est_destroy(est_handle *h) { zmf_destroy(h); }
int est_apply(est_handle *h, int nb, const double *y, double *x_hat, double *telapsed) { return zmf_fit(h, nb, y, x_hat, telapsed); }
int est_apply_d(est_handle *h, int nb, const double *y, double *x_hat, void *stream) { return zmf_fit_d(h, nb, y, x_hat, stream); }
long long est_launch_count(const est_handle *h) { return zmf_launch_count(h); }

} // extern "C"
