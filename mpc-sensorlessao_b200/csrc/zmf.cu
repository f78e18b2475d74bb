// zernmodfit on the GPU: batched masked least-squares projection of nL x nL phase frames onto
// the Zernike basis (zernmodfit.m:195-213 + zernfun.m:140-192, driver loop README.md:78-93).
//
// The reference rebuilds Z (npix_in x nmodes) and runs a QR per frame although Z depends only on
// the fixed pupil grid.  Here the host builds Z once in fp64, forms the least-squares operator
// W = inv(Z'Z) Z' (cond(Z) = 3.8 at N = 6, 5.0 at N = 10: normal equations are safe at 1e-10),
// scatters it to the full frame (zeros outside the pupil) and the device computes
// coef = W * frames, an HBM-streaming fp64 GEMM with M = nmodes, K = nL^2, N = nframes.
#include <cuda_runtime.h>
#include <cmath>
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
                                                     int npix, int nmodes, int nframes)
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
            if (j0 + a < nmodes && f0 + b < nframes) coef[(size_t)(f0 + b) * nmodes + j0 + a] = v;
        }
        __syncthreads();
    }
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
    // W = inv(Z'Z) Z'  via Cholesky of the Gram matrix (long double accumulation)
    std::vector<long double> G((size_t)M * M, 0.0L);
    for (int a = 0; a < M; ++a)
        for (int b = a; b < M; ++b) {
            long double s = 0.0L;
            const double *za = &h->Z[(size_t)a * P], *zb = &h->Z[(size_t)b * P];
            for (int p = 0; p < P; ++p) s += (long double)za[p] * zb[p];
            G[(size_t)a * M + b] = G[(size_t)b * M + a] = s;
        }
    for (int j = 0; j < M; ++j) {       // in-place lower Cholesky, row-major G[r*M+c]
        long double d = G[(size_t)j * M + j];
        for (int k = 0; k < j; ++k) d -= G[(size_t)j * M + k] * G[(size_t)j * M + k];
        if (!(d > 0.0L)) { delete h; return FMPC_ERR_NOT_PD; }
        d = sqrtl(d);
        G[(size_t)j * M + j] = d;
        for (int i = j + 1; i < M; ++i) {
            long double s = G[(size_t)i * M + j];
            for (int k = 0; k < j; ++k) s -= G[(size_t)i * M + k] * G[(size_t)j * M + k];
            G[(size_t)i * M + j] = s / d;
        }
    }
    std::vector<double> W((size_t)M * h->npix, 0.0);
    std::vector<long double> col(M);
    for (int p = 0; p < P; ++p) {       // solve G w = Z(p,:)'
        for (int j = 0; j < M; ++j) col[j] = h->Z[(size_t)j * P + p];
        for (int j = 0; j < M; ++j) { long double s = col[j]; for (int k = 0; k < j; ++k) s -= G[(size_t)j * M + k] * col[k]; col[j] = s / G[(size_t)j * M + j]; }
        for (int j = M - 1; j >= 0; --j) { long double s = col[j]; for (int k = j + 1; k < M; ++k) s -= G[(size_t)k * M + j] * col[k]; col[j] = s / G[(size_t)j * M + j]; }
        for (int j = 0; j < M; ++j) W[(size_t)j * h->npix + pidx[p]] = (double)col[j];
    }
    bool ok = cudaMalloc(&h->d_W, W.size() * 8) == cudaSuccess && cudaMalloc(&h->d_mask, h->npix) == cudaSuccess;
    ok = ok && cudaMemcpy(h->d_W, W.data(), W.size() * 8, cudaMemcpyHostToDevice) == cudaSuccess;
    ok = ok && cudaMemcpy(h->d_mask, h->mask.data(), h->npix, cudaMemcpyHostToDevice) == cudaSuccess;
    h->cap_frames = (size_t)max_frames;
    ok = ok && cudaMalloc(&h->d_frames, h->cap_frames * h->npix * 8) == cudaSuccess;
    ok = ok && cudaMalloc(&h->d_coef, h->cap_frames * M * 8) == cudaSuccess;
    ok = ok && cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaEventCreate(&h->ev0) == cudaSuccess && cudaEventCreate(&h->ev1) == cudaSuccess;
    if (!ok) { zmf_destroy(h); return FMPC_ERR_CUDA; }
    *out = h;
    return FMPC_OK;
}

void zmf_destroy(zmf_handle *h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->d_W) cudaFree(h->d_W);
    if (h->d_mask) cudaFree(h->d_mask);
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
    constexpr int MC = 7, FT = 4, NT = 256;
    const int grid = (nframes + FT - 1) / FT;
    zmf_fit_kernel<MC, FT, NT><<<grid, NT, 0, stream ? (cudaStream_t)stream : h->stream>>>(h->d_W, h->d_mask, frames, coef, h->npix,
                                                                                          h->nmodes, nframes);
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

} // extern "C"
