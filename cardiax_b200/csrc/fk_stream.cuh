// fk_stream.cuh -- CUDA kernel + launcher around the streaming body in fk_stream.h
#pragma once
#include <cuda_runtime.h>

#include "fk_stream.h"

namespace fk {

// cudaFuncSetAttribute and occupancy answers are PER DEVICE: every cache of them in this library is indexed by the
// current device ordinal, so a process that drives several GPUs (or switches device) opts in on each of them
enum { FK_MAX_DEVICES = 64 };
inline int cur_device() {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= FK_MAX_DEVICES) d = 0;
    return d;
}

// Which (strip, chunk) a block works on.  The CTAs of the tissue's first and last strip carry the one-sided column formulas
// (+20 % time, tools/probe_stream_timing.py) and a launch is one wave, so they come FIRST in block order: the hardware hands
// blocks out round-robin, which puts at most one of them on an SM as long as there are no more of them than SMs (4096^2: 98
// on 148), and a block that is resident first is also the one the warp schedulers favour.  In (strip, chunk) order every
// ninth SM held two of them and the launch ended 8 % after the average SM.  The other strips' last and first row chunks
// follow (their general-body iterations at the physical bottom / top edge made them the last blocks to finish).
__device__ __forceinline__ void stream_block_work(const StreamGeom& G, int b, int& strip, int& chunk) {
    if (G.nstrips < 3 || G.nchunks < 3) { strip = b % G.nstrips; chunk = b / G.nstrips; return; }
    const int ne = 2 * G.nchunks, ni = G.nstrips - 2;
    if (b < ne) { strip = (b & 1) ? G.nstrips - 1 : 0; chunk = b >> 1; return; }
    b -= ne;
    if (b < ni) { strip = 1 + b; chunk = G.nchunks - 1; return; }
    b -= ni;
    if (b < ni) { strip = 1 + b; chunk = 0; return; }
    b -= ni;
    strip = 1 + b % ni;
    chunk = 1 + b / ni;
}

// development (-DFK_STREAM_TIMING, tools/probe_stream_timing.py): every CTA records when it started and ended (globaltimer,
// ns) and on which SM it ran -- where a launch's time goes when its CTAs do not finish together
#ifdef FK_STREAM_TIMING
enum { FK_STREAM_TIMING_CTAS = 8192 };
__device__ unsigned long long fk_stream_timing_buf[3 * FK_STREAM_TIMING_CTAS];
__device__ __forceinline__ unsigned long long fk_globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#endif

template <bool EXACT, int T, bool UNI, int MODE = FK_STORE_PLAIN>
__global__ void __launch_bounds__(T == 2 ? 192 : 256, T <= 2 ? 2 : 1)
fk_stream_kernel(const __grid_constant__ TileArgs A, const __grid_constant__ StreamGeom G) {
    extern __shared__ __align__(16) float fk_stream_smem[];
#ifdef FK_STREAM_TIMING
    const unsigned long long t_start = fk_globaltimer();
#endif
    int strip, chunk;
    stream_block_work(G, (int)blockIdx.x, strip, chunk);
    StreamCta C;
    stream_cta_setup<T>(A, G, strip, chunk, blockIdx.y, C);
    const int tid = threadIdx.x;
    const StreamMem M = stream_mem<T>(fk_stream_smem, tid, G.NT);
    StreamState<T> R;
    stream_state_init<T>(A, C, tid, R);
    const int nfill = stream_nfill<T>(C);
    float* bar = stream_bar<T>(fk_stream_smem, G.NT);
    if (tid == 0) sb_init(bar, G.NT);
    int i = 0;
    if (!stream_use_warm<T>(C)) {
        // (a chunk at the physical top edge emits from iteration 4 on): plain prologue, general body from iteration 0
#pragma unroll
        for (int j = 0; j < FK_PF; ++j) stream_prefetch<T>(A, C, M, j, tid, C.cs + 4 * tid < C.c_end);
        __syncthreads();
    } else {
        stream_warm_load<T>(A, C, M, tid);
        async_wait<0>();
        __syncthreads();
        stream_warm_start<EXACT, T>(A, C, R, M, tid);   // stands for iterations 0 .. 7
        __syncthreads();
        i = FK_WARM;
    }
    for (; i < nfill; ++i) {   // pipeline fill (and launches with an active stimulus): fully conditional body
        stream_iter<EXACT, T, -1, UNI, true, 4, MODE>(A, C, R, M, i, tid, stream_ptrs_any<T>(M, i), nullptr);
        __syncthreads();
    }
    // steady state: every stage consumes and emits one row per iteration; unrolled U-fold so that every ring slot is a
    // per-body base register plus a compile-time offset.  Iterations are separated by the split barrier: phase k of it
    // completes when every thread has finished iteration k - 1 of this loop.
    constexpr int U = stream_unroll(T);
    const int i_end = nfill + U * stream_nbody<T>(C);   // (a chunk at the bottom edge ends its steady state early)
    if (i < i_end) {
        sb_arrive(bar);   // phase 0: the fill loop's last block barrier stands for "iteration -1"
        const bool edge = C.edgeL >= 0 || C.edgeR >= 0;
        StreamBody<T> Y = stream_body_at<T>(M, i);
#define FK_STEADY_LOOP(EDGE)                                                                                           \
    for (; i < i_end; i += U) {                                                                                        \
        stream_iter<EXACT, T, 0, UNI, EDGE, U, MODE>(A, C, R, M, i, tid, stream_ptrs_phase<T, U, 0>(Y), bar);               \
        stream_iter<EXACT, T, 1, UNI, EDGE, U, MODE>(A, C, R, M, i + 1, tid, stream_ptrs_phase<T, U, 1>(Y), bar);           \
        if (U == 4) {                                                                                                  \
            stream_iter<EXACT, T, U == 4 ? 2 : 0, UNI, EDGE, U, MODE>(A, C, R, M, i + 2, tid,                               \
                                                                stream_ptrs_phase<T, U, U == 4 ? 2 : 0>(Y), bar);      \
            stream_iter<EXACT, T, U == 4 ? 3 : 1, UNI, EDGE, U, MODE>(A, C, R, M, i + 3, tid,                               \
                                                                stream_ptrs_phase<T, U, U == 4 ? 3 : 1>(Y), bar);      \
        }                                                                                                              \
        stream_body_next<T, U>(Y);                                                                                     \
    }
        if (edge) { FK_STEADY_LOOP(true) } else { FK_STEADY_LOOP(false) }
#undef FK_STEADY_LOOP
        sb_wait(bar, 0);   // an even number of iterations later: everybody is through the last one
    }
    if (i < C.niter) {
        // tail: fewer than U rows left, or the stages of a chunk at the physical bottom edge running dry.  The CTA's
        // constants are rebuilt here rather than kept alive across the unrolled loop, where every uniform register
        // counts (with them live, ptxas moved the loop's constants from uniform to ordinary registers: -12 %)
        StreamCta Ct;
        stream_cta_setup<T>(A, G, strip, chunk, blockIdx.y, Ct);
        for (; i < Ct.niter; ++i) {
            stream_iter<EXACT, T, -1, UNI, true, 4, MODE>(A, Ct, R, M, i, tid, stream_ptrs_any<T>(M, i), nullptr);
            __syncthreads();
        }
    }
    // halo mirror: this thread's stores into the neighbouring GPUs' memory are performed, system wide, before the kernel
    // ends -- the flag the neighbour waits for is written by a later operation of the same stream
    if (MODE == FK_STORE_MIRROR) __threadfence_system();
#ifdef FK_STREAM_TIMING
    __syncthreads();
    if (threadIdx.x == 0 && gridDim.x * gridDim.y <= FK_STREAM_TIMING_CTAS) {
        unsigned smid;
        asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
        const int rec = (int)blockIdx.y * (int)gridDim.x + chunk * G.nstrips + strip;   // (records in (tissue, chunk, strip) order)
        fk_stream_timing_buf[3 * rec] = t_start;
        fk_stream_timing_buf[3 * rec + 1] = fk_globaltimer();
        fk_stream_timing_buf[3 * rec + 2] = smid;
    }
#endif
}

// (Programmatic dependent launch -- the next launch's blocks resident and waiting in griddepcontrol.wait as this one's
// retire -- was measured and dropped: 23.4 instead of 22.2 ms per 500-step segment of 4096^2, gpurun_out/r02zt.)
template <class K>
inline int launch_plain(K kernel, dim3 grid, const StreamPlan& P, const TileArgs& A, cudaStream_t st) {
    kernel<<<grid, P.G.NT, P.smem_bytes, st>>>(A, P.G);
    return (int)cudaGetLastError();
}

template <bool EXACT, int T, bool UNI>
inline int launch_stream_t(const StreamPlan& P, const TileArgs& A, int batch, cudaStream_t st) {
    static long long attr_set_dev[FK_MAX_DEVICES] = {0};   // largest dynamic shared memory already allowed, per device
    long long& attr_set = attr_set_dev[cur_device()];
    if (P.smem_bytes > attr_set && attr_set < 227 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(fk_stream_kernel<EXACT, T, UNI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)P.smem_bytes);
        if (e != cudaSuccess) return (int)e;
        attr_set = P.smem_bytes;
    }
    dim3 grid(P.G.nstrips * P.G.nchunks, batch);
    if (A.hy_u) {   // fast Heun: the instantiation whose last level stores y + (E - y) / 2 (fast numerics, T <= 2 only)
        if constexpr (!EXACT && T <= 2) {
            static long long attr_set_h_dev[FK_MAX_DEVICES] = {0};
            long long& attr_set_h = attr_set_h_dev[cur_device()];
            if (P.smem_bytes > attr_set_h && attr_set_h < 227 * 1024) {
                cudaError_t e = cudaFuncSetAttribute(fk_stream_kernel<EXACT, T, UNI, FK_STORE_HEUN>,
                                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P.smem_bytes);
                if (e != cudaSuccess) return (int)e;
                attr_set_h = P.smem_bytes;
            }
            return launch_plain(fk_stream_kernel<EXACT, T, UNI, FK_STORE_HEUN>, grid, P, A, st);
        } else {
            return -2;   // not built
        }
    }
    if (A.mir_u[0] || A.mir_u[1]) {   // row-slab decomposition: the instantiation that mirrors the band rows to the neighbours
        if constexpr (T <= 2) {
            static long long attr_set_m_dev[FK_MAX_DEVICES] = {0};
            long long& attr_set_m = attr_set_m_dev[cur_device()];
            if (P.smem_bytes > attr_set_m && attr_set_m < 227 * 1024) {
                cudaError_t e = cudaFuncSetAttribute(fk_stream_kernel<EXACT, T, UNI, FK_STORE_MIRROR>,
                                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P.smem_bytes);
                if (e != cudaSuccess) return (int)e;
                attr_set_m = P.smem_bytes;
            }
            return launch_plain(fk_stream_kernel<EXACT, T, UNI, FK_STORE_MIRROR>, grid, P, A, st);
        } else {
            return -2;   // not built: the caller copies the bands instead
        }
    }
    return launch_plain(fk_stream_kernel<EXACT, T, UNI>, grid, P, A, st);
}

template <bool EXACT, int T, bool UNI>
inline int stream_occupancy_t(int NT, long long smem) {
    // memoised: the planner asks for the same few configurations at every call
    static int memo_nt[64], memo_n[64], memo_dev[64], memo_cnt = 0;
    static long long memo_smem[64];
    const int dev = cur_device();
    for (int i = 0; i < memo_cnt; ++i)
        if (memo_nt[i] == NT && memo_smem[i] == smem && memo_dev[i] == dev) return memo_n[i];
    if (cudaFuncSetAttribute(fk_stream_kernel<EXACT, T, UNI>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) !=
        cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, fk_stream_kernel<EXACT, T, UNI>, NT, (size_t)smem) !=
        cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    if (memo_cnt < 64) { memo_nt[memo_cnt] = NT; memo_smem[memo_cnt] = smem; memo_n[memo_cnt] = n; memo_dev[memo_cnt] = dev; ++memo_cnt; }
    return n;
}

// One translation unit per (T, numerics) -- fk_stream_tu.cu compiled with -DFK_TU_T=<T> -DFK_TU_EXACT=<0|1> --
// instantiates the kernels of that depth and defines these two entry points, so that they compile in parallel.
#define FK_STREAM_TU_DECL(TT)                                                                                      \
    int launch_stream_T##TT##_E0(const StreamPlan& P, const TileArgs& A, int batch, cudaStream_t st);              \
    int launch_stream_T##TT##_E1(const StreamPlan& P, const TileArgs& A, int batch, cudaStream_t st);              \
    int stream_occupancy_T##TT##_E0(int uni, int NT, long long smem);                                              \
    int stream_occupancy_T##TT##_E1(int uni, int NT, long long smem);
FK_STREAM_TU_DECL(1) FK_STREAM_TU_DECL(2) FK_STREAM_TU_DECL(3) FK_STREAM_TU_DECL(4)
#undef FK_STREAM_TU_DECL
#ifndef FK_DEPTH_MASK        // development builds may link a subset of the depths (bit T set = depth T present)
#define FK_DEPTH_MASK 0x1e
#endif

// resident CTAs per SM of the streaming kernel for (T, numerics, diffusivity kind, threads, shared memory)
inline int stream_occupancy(int T, int exact, int uni, int NT, long long smem) {
    switch (T) {
#define FK_CASE(TT)                                                                                     \
    case TT:                                                                                            \
        if constexpr ((FK_DEPTH_MASK >> TT) & 1)                                                        \
            return exact ? stream_occupancy_T##TT##_E1(uni, NT, smem) : stream_occupancy_T##TT##_E0(uni, NT, smem); \
        break;
        FK_CASE(1) FK_CASE(2) FK_CASE(3) FK_CASE(4)
#undef FK_CASE
    }
    return 0;
}

// returns 0, a cudaError_t (> 0), or < 0 when T is unsupported
inline int launch_stream(const StreamPlan& P, const TileArgs& A, int exact, int batch, cudaStream_t st) {
    switch (P.T) {
#define FK_CASE(TT)                                                                                     \
    case TT:                                                                                            \
        if constexpr ((FK_DEPTH_MASK >> TT) & 1)                                                        \
            return exact ? launch_stream_T##TT##_E1(P, A, batch, st) : launch_stream_T##TT##_E0(P, A, batch, st); \
        break;
        FK_CASE(1) FK_CASE(2) FK_CASE(3) FK_CASE(4)
#undef FK_CASE
    }
    return -5;
}

}  // namespace fk
