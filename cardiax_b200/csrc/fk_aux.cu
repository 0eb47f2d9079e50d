// fk_aux.cu -- sm_100a kernels around the Euler hot path: snapshot resize (cardiax/io.py:118-124),
// the array side of the Dormand-Prince integrator (cardiax/solve.py:114-124 -> jax.experimental.ode) and
// metrics.electrogram (cardiax/metrics.py:13-22).  All HBM-bound elementwise / reduction work, no tensor cores.
#include <cuda_runtime.h>

#include <string.h>

#include <algorithm>
#include <vector>

#include "../../include/fk.h"
#include "fk_aux.cuh"

namespace fk {
namespace {

#define AUX_CUDA(call)                                              \
    do {                                                            \
        cudaError_t e_ = (call);                                    \
        if (e_ != cudaSuccess) return api_cuda_fail((int)e_, #call); \
    } while (0)

// ------------------------------------------------------------------------------------------------ resize
// One thread per output pixel, 32 x 8 pixels per CTA.  The CTA's weights (Kw taps of its 32 columns, Kh taps of its 8
// rows) are staged in shared memory tap-major: the inner loop is one global load (L1-resident: neighbouring pixels'
// windows overlap) + one conflict-free shared load + one FMA per tap.
template <bool SMEM>
__global__ void __launch_bounds__(256) fk_resize_kernel(const __grid_constant__ ResizeArgs A) {
    extern __shared__ float sw[];   // [Kw][32] then [Kh][8]
    const int j0 = blockIdx.x * 32, i0 = blockIdx.y * 8, p = blockIdx.z;
    const int j = j0 + threadIdx.x, i = i0 + threadIdx.y;
    const int tid = threadIdx.y * 32 + threadIdx.x;
    if (SMEM) {
        for (int e = tid; e < A.Kw * 32; e += 256) {
            const int q = e >> 5, c = j0 + (e & 31);
            sw[e] = c < A.Wo ? __ldg(A.wt_w + (size_t)q * A.Wo + c) : 0.0f;
        }
        float* sh = sw + A.Kw * 32;
        for (int e = tid; e < A.Kh * 8; e += 256) {
            const int r = e >> 3, c = i0 + (e & 7);
            sh[e] = c < A.Ho ? __ldg(A.wt_h + (size_t)r * A.Ho + c) : 0.0f;
        }
        __syncthreads();
    }
    if (i >= A.Ho || j >= A.Wo) return;
    const float* in = A.n_planes <= RESIZE_BY_VALUE ? A.plane[p] : A.planes[p];
    float v;
    if (SMEM) v = resize_pixel(in, A.W, __ldg(A.lo_h + i), sw + A.Kw * 32 + threadIdx.y, 8, A.Kh, __ldg(A.lo_w + j), sw + threadIdx.x, 32, A.Kw);
    else v = resize_pixel(in, A.W, __ldg(A.lo_h + i), A.wt_h + i, A.Ho, A.Kh, __ldg(A.lo_w + j), A.wt_w + j, A.Wo, A.Kw);
    A.out[((size_t)p * A.Ho + i) * A.Wo + j] = v;
}

// ------------------------------------------------------------------------------------------------ reductions
// NS running sums per thread -> one double per (sum, block) in partial[s * nslots + slot]; fixed order, no atomics
template <int NS>
__device__ __forceinline__ void block_sums(double (&v)[NS], double* partial, int slot, int nslots) {
    __shared__ double sm[NS][ODE_THREADS / 32];
#pragma unroll
    for (int s = 0; s < NS; ++s) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[s] += __shfl_down_sync(0xffffffffu, v[s], o);
        if ((threadIdx.x & 31) == 0) sm[s][threadIdx.x >> 5] = v[s];
    }
    __syncthreads();
    if (threadIdx.x < NS) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < ODE_THREADS / 32; ++w) t += sm[threadIdx.x][w];
        partial[(size_t)threadIdx.x * nslots + slot] = t;
    }
}

__global__ void __launch_bounds__(ODE_THREADS) fk_ode_reduce_kernel(const double* __restrict__ partial, int nslots, int ns,
                                                                    double* __restrict__ result) {
    __shared__ double sm[ODE_THREADS];
    for (int s = 0; s < ns; ++s) {
        double t = 0.0;
        for (int i = threadIdx.x; i < nslots; i += ODE_THREADS) t += partial[(size_t)s * nslots + i];
        sm[threadIdx.x] = t;
        __syncthreads();
        for (int o = ODE_THREADS / 2; o > 0; o >>= 1) {
            if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
            __syncthreads();
        }
        if (threadIdx.x == 0) result[s] = sm[0];
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------ Dopri5 kernels
// blockIdx.y = which array of the State; grid-stride over its n elements
template <bool EXACT>
__global__ void __launch_bounds__(ODE_THREADS) fk_ode_axpy_kernel(P3 y, float h, P3 f, P3 out, long long n) {
    typedef Num<EXACT> N;
    const float* __restrict__ yy = y.a[blockIdx.y];
    const float* __restrict__ ff = f.a[blockIdx.y];
    float* __restrict__ oo = out.a[blockIdx.y];
    for (long long e = blockIdx.x * (long long)ODE_THREADS + threadIdx.x; e < n; e += (long long)gridDim.x * ODE_THREADS)
        oo[e] = N::add(yy[e], N::mul(h, ff[e]));
}

template <bool EXACT>
__global__ void __launch_bounds__(ODE_THREADS) fk_ode_init_norms_kernel(P3 y, P3 f, float rtol, float atol, long long n,
                                                                        double* partial) {
    typedef Num<EXACT> N;
    const float* __restrict__ yy = y.a[blockIdx.y];
    const float* __restrict__ ff = f.a[blockIdx.y];
    double acc[2] = {0.0, 0.0};
    for (long long e = blockIdx.x * (long long)ODE_THREADS + threadIdx.x; e < n; e += (long long)gridDim.x * ODE_THREADS) {
        float qy, qf;
        ode_scaled<EXACT>(yy[e], ff[e], rtol, atol, qy, qf);
        acc[0] += (double)N::mul(qy, qy);
        acc[1] += (double)N::mul(qf, qf);
    }
    block_sums<2>(acc, partial, blockIdx.y * gridDim.x + blockIdx.x, 3 * gridDim.x);
}

template <bool EXACT>
__global__ void __launch_bounds__(ODE_THREADS) fk_ode_diff_norm_kernel(P3 f1, P3 f0, P3 y, float rtol, float atol, long long n,
                                                                       double* partial) {
    typedef Num<EXACT> N;
    const float* __restrict__ a = f1.a[blockIdx.y];
    const float* __restrict__ b = f0.a[blockIdx.y];
    const float* __restrict__ yy = y.a[blockIdx.y];
    double acc[1] = {0.0};
    for (long long e = blockIdx.x * (long long)ODE_THREADS + threadIdx.x; e < n; e += (long long)gridDim.x * ODE_THREADS) {
        const float scale = N::add(atol, N::mul(fabsf(yy[e]), rtol));
        const float q = ode_div<EXACT>(N::sub(a[e], b[e]), scale);
        acc[0] += (double)N::mul(q, q);
    }
    block_sums<1>(acc, partial, blockIdx.y * gridDim.x + blockIdx.x, 3 * gridDim.x);
}

struct K7 {
    const float* k[7];
};

template <bool EXACT>
__global__ void __launch_bounds__(ODE_THREADS) fk_ode_stage_kernel(const __grid_constant__ Dopri T, int i, P3 y, K7 kv, K7 kw, K7 ku,
                                                                   float dt, P3 ys, long long n) {
    const K7& K = blockIdx.y == 0 ? kv : (blockIdx.y == 1 ? kw : ku);
    const float* __restrict__ yy = y.a[blockIdx.y];
    float* __restrict__ oo = ys.a[blockIdx.y];
    const float* beta = T.beta[i - 1];
    for (long long e = blockIdx.x * (long long)ODE_THREADS + threadIdx.x; e < n; e += (long long)gridDim.x * ODE_THREADS) {
        float k[6];
#pragma unroll
        for (int s = 0; s < 6; ++s) k[s] = (s < i && beta[s] != 0.0f) ? K.k[s][e] : 0.0f;
        oo[e] = ode_stage<EXACT>(yy[e], beta, k, i, dt);
    }
}

struct C5 {
    float* c[5];
};

template <bool EXACT>
__global__ void __launch_bounds__(ODE_THREADS) fk_ode_finish_kernel(const __grid_constant__ Dopri T, P3 y, K7 kv, K7 kw, K7 ku, float dt,
                                                                    float rtol, float atol, P3 yn, C5 cv, C5 cw, C5 cu, long long n,
                                                                    double* partial) {
    const K7& K = blockIdx.y == 0 ? kv : (blockIdx.y == 1 ? kw : ku);
    const C5& C = blockIdx.y == 0 ? cv : (blockIdx.y == 1 ? cw : cu);
    const float* __restrict__ yy = y.a[blockIdx.y];
    float* __restrict__ oo = yn.a[blockIdx.y];
    double acc[1] = {0.0};
    for (long long e = blockIdx.x * (long long)ODE_THREADS + threadIdx.x; e < n; e += (long long)gridDim.x * ODE_THREADS) {
        float k[7];
#pragma unroll
        for (int s = 0; s < 7; ++s) k[s] = s == 1 ? 0.0f : K.k[s][e];   // the second stage has zero weight everywhere
        float y1, r2, coef[5];
        ode_finish<EXACT>(T, yy[e], k, dt, rtol, atol, y1, r2, coef);
        oo[e] = y1;
#pragma unroll
        for (int s = 0; s < 5; ++s) C.c[s][e] = coef[s];
        acc[0] += (double)r2;
    }
    block_sums<1>(acc, partial, blockIdx.y * gridDim.x + blockIdx.x, 3 * gridDim.x);
}

template <bool EXACT>
__global__ void __launch_bounds__(ODE_THREADS) fk_ode_interp_kernel(C5 cv, C5 cw, C5 cu, float r, P3 out, long long off, long long n) {
    const C5& C = blockIdx.y == 0 ? cv : (blockIdx.y == 1 ? cw : cu);
    float* __restrict__ oo = out.a[blockIdx.y] + off;
    for (long long e = blockIdx.x * (long long)ODE_THREADS + threadIdx.x; e < n; e += (long long)gridDim.x * ODE_THREADS) {
        float coef[5];
#pragma unroll
        for (int s = 0; s < 5; ++s) coef[s] = C.c[s][e];
        oo[e] = ode_interp<EXACT>(coef, r);
    }
}

// ------------------------------------------------------------------------------------------------ electrogram
__global__ void __launch_bounds__(1024) fk_electrogram_kernel(const float* __restrict__ x, int H, int W, float p0, float p1,
                                                              float* __restrict__ out) {
    const float* frame = x + (size_t)blockIdx.x * H * W;
    double acc = 0.0;
    for (int i = threadIdx.y; i < H; i += blockDim.y)
        for (int j = threadIdx.x; j < W; j += blockDim.x) acc += (double)__fmul_rn(frame[(size_t)i * W + j], egm_weight(i, j, p0, p1));
    __shared__ double sm[32];
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if ((tid & 31) == 0) sm[tid >> 5] = acc;
    __syncthreads();
    if (tid == 0) {
        double t = 0.0;
        for (int w = 0; w < 32; ++w) t += sm[w];
        out[blockIdx.x] = (float)t;
    }
}

dim3 ode_grid(long long n) {
    const long long b = (n + ODE_THREADS - 1) / ODE_THREADS;
    return dim3((unsigned)std::min<long long>(std::max<long long>(b, 1), ODE_BLOCKS), 3);
}

K7 k7(const P3* k, int which) {
    K7 r;
    for (int s = 0; s < 7; ++s) r.k[s] = k[s].a[which];
    return r;
}
C5 c5(const P3* c, int which) {
    C5 r;
    for (int s = 0; s < 5; ++s) r.c[s] = c[s].a[which];
    return r;
}

int fetch_sums(const OdeScratch& S, int nslots, int ns, double* host, cudaStream_t st) {
    double* result = S.partial + 2 * 3 * ODE_BLOCKS;
    fk_ode_reduce_kernel<<<1, ODE_THREADS, 0, st>>>(S.partial, nslots, ns, result);
    api_count_launch(1);
    AUX_CUDA(cudaGetLastError());
    AUX_CUDA(cudaMemcpyAsync(host, result, sizeof(double) * ns, cudaMemcpyDeviceToHost, st));
    AUX_CUDA(cudaStreamSynchronize(st));   // the step-size controller runs on the host
    return 0;
}

}  // namespace

int launch_ode_copy(const P3& dst, long long off, const P3& src, long long n, cudaStream_t st) {
    for (int a = 0; a < 3; ++a)
        AUX_CUDA(cudaMemcpyAsync(dst.a[a] + off, src.a[a], sizeof(float) * (size_t)n, cudaMemcpyDeviceToDevice, st));
    return 0;
}

int launch_ode_axpy(int exact, const P3& y, float h, const P3& f, const P3& out, long long n, cudaStream_t st) {
    api_count_launch(1);
    if (exact) fk_ode_axpy_kernel<true><<<ode_grid(n), ODE_THREADS, 0, st>>>(y, h, f, out, n);
    else fk_ode_axpy_kernel<false><<<ode_grid(n), ODE_THREADS, 0, st>>>(y, h, f, out, n);
    AUX_CUDA(cudaGetLastError());
    return 0;
}

int launch_ode_init_norms(int exact, const P3& y, const P3& f, float rtol, float atol, long long n, const OdeScratch& S,
                          double* sums2_host, cudaStream_t st) {
    const dim3 g = ode_grid(n);
    api_count_launch(1);
    if (exact) fk_ode_init_norms_kernel<true><<<g, ODE_THREADS, 0, st>>>(y, f, rtol, atol, n, S.partial);
    else fk_ode_init_norms_kernel<false><<<g, ODE_THREADS, 0, st>>>(y, f, rtol, atol, n, S.partial);
    AUX_CUDA(cudaGetLastError());
    return fetch_sums(S, 3 * (int)g.x, 2, sums2_host, st);
}

int launch_ode_diff_norm(int exact, const P3& f1, const P3& f0, const P3& y, float rtol, float atol, long long n,
                         const OdeScratch& S, double* sum_host, cudaStream_t st) {
    const dim3 g = ode_grid(n);
    api_count_launch(1);
    if (exact) fk_ode_diff_norm_kernel<true><<<g, ODE_THREADS, 0, st>>>(f1, f0, y, rtol, atol, n, S.partial);
    else fk_ode_diff_norm_kernel<false><<<g, ODE_THREADS, 0, st>>>(f1, f0, y, rtol, atol, n, S.partial);
    AUX_CUDA(cudaGetLastError());
    return fetch_sums(S, 3 * (int)g.x, 1, sum_host, st);
}

int launch_ode_stage(int exact, const Dopri& T, int i, const P3& y, const P3* k, float dt, const P3& ys, long long n,
                     cudaStream_t st) {
    api_count_launch(1);
    if (exact) fk_ode_stage_kernel<true><<<ode_grid(n), ODE_THREADS, 0, st>>>(T, i, y, k7(k, 0), k7(k, 1), k7(k, 2), dt, ys, n);
    else fk_ode_stage_kernel<false><<<ode_grid(n), ODE_THREADS, 0, st>>>(T, i, y, k7(k, 0), k7(k, 1), k7(k, 2), dt, ys, n);
    AUX_CUDA(cudaGetLastError());
    return 0;
}

int launch_ode_finish(int exact, const Dopri& T, const P3& y, const P3* k, float dt, float rtol, float atol, const P3& yn,
                      const P3* c, long long n, const OdeScratch& S, double* sum_host, cudaStream_t st) {
    const dim3 g = ode_grid(n);
    api_count_launch(1);
    if (exact)
        fk_ode_finish_kernel<true><<<g, ODE_THREADS, 0, st>>>(T, y, k7(k, 0), k7(k, 1), k7(k, 2), dt, rtol, atol, yn, c5(c, 0),
                                                              c5(c, 1), c5(c, 2), n, S.partial);
    else
        fk_ode_finish_kernel<false><<<g, ODE_THREADS, 0, st>>>(T, y, k7(k, 0), k7(k, 1), k7(k, 2), dt, rtol, atol, yn, c5(c, 0),
                                                               c5(c, 1), c5(c, 2), n, S.partial);
    AUX_CUDA(cudaGetLastError());
    return fetch_sums(S, 3 * (int)g.x, 1, sum_host, st);
}

int launch_ode_interp(int exact, const P3* c, float r, const P3& out, long long off, long long n, cudaStream_t st) {
    api_count_launch(1);
    if (exact) fk_ode_interp_kernel<true><<<ode_grid(n), ODE_THREADS, 0, st>>>(c5(c, 0), c5(c, 1), c5(c, 2), r, out, off, n);
    else fk_ode_interp_kernel<false><<<ode_grid(n), ODE_THREADS, 0, st>>>(c5(c, 0), c5(c, 1), c5(c, 2), r, out, off, n);
    AUX_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace fk

// ---------------------------------------------------------------------------------------------------- C ABI
extern "C" {

static size_t aux_align(size_t x) { return (x + 255) / 256 * 256; }

size_t fk_resize_workspace_bytes(int H, int W, int Ho, int Wo, int n_planes) {
    if (H <= 0 || W <= 0 || Ho <= 0 || Wo <= 0 || n_planes <= 0) return 0;
    const fk::ResizeAxis ah = fk::make_resize_axis(H, Ho), aw = fk::make_resize_axis(W, Wo);
    return aux_align(4 * (size_t)Ho) + aux_align(4 * (size_t)Ho * ah.K) + aux_align(4 * (size_t)Wo) +
           aux_align(4 * (size_t)Wo * aw.K) + aux_align(sizeof(void*) * (size_t)n_planes);
}

int fk_resize_bilinear(const float* const* planes, int n_planes, int H, int W, float* out, int Ho, int Wo, void* workspace,
                       size_t workspace_bytes, int workspace_ready, void* stream) {
    using namespace fk;
    if (!planes || !out || !workspace) return api_fail(-1, "NULL pointer");
    if (H <= 0 || W <= 0 || Ho <= 0 || Wo <= 0 || n_planes <= 0 || n_planes > 65535) return api_fail(-1, "bad resize shape");
    if (workspace_bytes < fk_resize_workspace_bytes(H, W, Ho, Wo, n_planes)) return api_fail(-4, "workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    const ResizeAxis ah = make_resize_axis(H, Ho), aw = make_resize_axis(W, Wo);
    char* p = (char*)workspace;
    ResizeArgs A;
    memset(&A, 0, sizeof(A));
    A.n_planes = n_planes; A.H = H; A.W = W; A.Ho = Ho; A.Wo = Wo; A.Kh = ah.K; A.Kw = aw.K; A.out = out;
    // pageable sources: the runtime stages them before returning
    A.lo_h = (const int*)p;
    if (!workspace_ready) AUX_CUDA(cudaMemcpyAsync(p, ah.lo.data(), 4 * (size_t)Ho, cudaMemcpyHostToDevice, st));
    p += aux_align(4 * (size_t)Ho);
    A.wt_h = (const float*)p;
    if (!workspace_ready) AUX_CUDA(cudaMemcpyAsync(p, resize_tap_major(ah).data(), 4 * (size_t)Ho * ah.K, cudaMemcpyHostToDevice, st));
    p += aux_align(4 * (size_t)Ho * ah.K);
    A.lo_w = (const int*)p;
    if (!workspace_ready) AUX_CUDA(cudaMemcpyAsync(p, aw.lo.data(), 4 * (size_t)Wo, cudaMemcpyHostToDevice, st));
    p += aux_align(4 * (size_t)Wo);
    A.wt_w = (const float*)p;
    if (!workspace_ready) AUX_CUDA(cudaMemcpyAsync(p, resize_tap_major(aw).data(), 4 * (size_t)Wo * aw.K, cudaMemcpyHostToDevice, st));
    p += aux_align(4 * (size_t)Wo * aw.K);
    if (n_planes <= RESIZE_BY_VALUE) {
        for (int i = 0; i < n_planes; ++i) A.plane[i] = planes[i];
    } else {
        A.planes = (const float* const*)p;
        AUX_CUDA(cudaMemcpyAsync(p, planes, sizeof(void*) * (size_t)n_planes, cudaMemcpyHostToDevice, st));
    }
    api_count_launch(1);
    const dim3 grid((Wo + 31) / 32, (Ho + 7) / 8, n_planes);
    const size_t smem = sizeof(float) * ((size_t)A.Kw * 32 + (size_t)A.Kh * 8);
    if (smem <= 40 * 1024) fk_resize_kernel<true><<<grid, dim3(32, 8), smem, st>>>(A);
    else fk_resize_kernel<false><<<grid, dim3(32, 8), 0, st>>>(A);   // extreme reductions: weights straight from L1/L2
    AUX_CUDA(cudaGetLastError());
    return 0;
}

int fk_electrogram(const float* x, int frames, int H, int W, float p0, float p1, float* out, void* stream) {
    using namespace fk;
    if (!x || !out) return api_fail(-1, "NULL pointer");
    if (frames <= 0 || H <= 0 || W <= 0) return api_fail(-1, "bad electrogram shape");
    api_count_launch(1);
    fk_electrogram_kernel<<<frames, dim3(32, 32), 0, (cudaStream_t)stream>>>(x, H, W, p0, p1, out);
    AUX_CUDA(cudaGetLastError());
    return 0;
}

}  // extern "C"
