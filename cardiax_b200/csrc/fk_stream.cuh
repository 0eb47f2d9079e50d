// fk_stream.cuh -- CUDA kernel + launcher around the streaming body in fk_stream.h
#pragma once
#include <cuda_runtime.h>

#include "fk_stream.h"

namespace fk {

template <bool EXACT, int T>
__global__ void __launch_bounds__(256, 1)
fk_stream_kernel(const __grid_constant__ TileArgs A, const __grid_constant__ StreamGeom G) {
    extern __shared__ __align__(16) float fk_stream_smem[];
    StreamSmem<T> S;
    stream_carve<T>(fk_stream_smem, G, S);
    const int strip = blockIdx.x % G.nstrips, chunk = blockIdx.x / G.nstrips;
    StreamCta C;
    stream_cta_setup<T>(A, G, strip, chunk, blockIdx.y, C);
    StreamState<T> R;
    {
        float* f = reinterpret_cast<float*>(&R);
#pragma unroll
        for (int q = 0; q < (int)(sizeof(R) / sizeof(float)); ++q) f[q] = 0.0f;
    }
    const int tid = threadIdx.x;
#pragma unroll
    for (int j = 0; j < FK_PF; ++j) stream_prefetch<T>(A, G, C, S, j, tid, C.cs + 4 * tid < C.c_end);
    const int nfill = (8 * T < C.niter && stream_steady_ok<T>(C)) ? 8 * T : C.niter;
    int i = 0;
    for (; i < nfill; ++i) {   // pipeline fill (and launches with an active stimulus): fully conditional body
        stream_iter<EXACT, T, false>(A, G, C, S, R, i, tid);
        __syncthreads();
    }
    for (; i < C.niter; ++i) {  // steady state: every stage consumes and emits one row
        stream_iter<EXACT, T, true>(A, G, C, S, R, i, tid);
        __syncthreads();
    }
}

template <bool EXACT, int T>
inline int launch_stream_t(const StreamPlan& P, const TileArgs& A, int batch, cudaStream_t st) {
    static long long attr_set = 0;   // largest dynamic shared memory already allowed for this instantiation
    if (P.smem_bytes > attr_set && attr_set < 227 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(fk_stream_kernel<EXACT, T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)P.smem_bytes);
        if (e != cudaSuccess) return (int)e;
        attr_set = P.smem_bytes;
    }
    dim3 grid(P.G.nstrips * P.G.nchunks, batch);
    fk_stream_kernel<EXACT, T><<<grid, P.G.NT, P.smem_bytes, st>>>(A, P.G);
    return (int)cudaGetLastError();
}

template <bool EXACT, int T>
inline int stream_occupancy_t(int NT, long long smem) {
    // memoised: the planner asks for the same few configurations at every call
    static int memo_nt[64], memo_n[64], memo_cnt = 0;
    static long long memo_smem[64];
    for (int i = 0; i < memo_cnt; ++i)
        if (memo_nt[i] == NT && memo_smem[i] == smem) return memo_n[i];
    if (cudaFuncSetAttribute(fk_stream_kernel<EXACT, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) !=
        cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, fk_stream_kernel<EXACT, T>, NT, (size_t)smem) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    if (memo_cnt < 64) { memo_nt[memo_cnt] = NT; memo_smem[memo_cnt] = smem; memo_n[memo_cnt] = n; ++memo_cnt; }
    return n;
}

// resident CTAs per SM of the streaming kernel for (T, numerics, threads, shared memory)
inline int stream_occupancy(int T, int exact, int NT, long long smem) {
    switch (T) {
#define FK_CASE(TT) \
    case TT: return exact ? stream_occupancy_t<true, TT>(NT, smem) : stream_occupancy_t<false, TT>(NT, smem);
        FK_CASE(1) FK_CASE(2) FK_CASE(3) FK_CASE(4)
#undef FK_CASE
    }
    return 0;
}

// returns 0, a cudaError_t (> 0), or < 0 when T is unsupported
inline int launch_stream(const StreamPlan& P, const TileArgs& A, int exact, int batch, cudaStream_t st) {
    switch (P.T) {
#define FK_CASE(TT) \
    case TT: return exact ? launch_stream_t<true, TT>(P, A, batch, st) : launch_stream_t<false, TT>(P, A, batch, st);
        FK_CASE(1) FK_CASE(2) FK_CASE(3) FK_CASE(4)
#undef FK_CASE
    }
    return -5;
}

}  // namespace fk
