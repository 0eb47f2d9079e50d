// fk_resident.cu -- the resident kernel (body: fk_resident.h) and its cooperative launcher.  sm_100a only.
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdlib.h>

#include "fk_resident.cuh"

namespace fk {

namespace {

// dynamic shared memory the kernel may ask for: the 227 KB opt-in limit minus its static shared memory (rounded up)
enum { FK_RES_SMEM_OPTIN = 227 * 1024 - 256 };

// development: cycle counters of CTA (0, 0), thread 0 (FK_RES_TIMING=1): ring, interior, halo, barrier; [6] = steps
__device__ u64 g_res_timing[8];
#define FK_TICK(k)                       \
    if (timed) {                         \
        const long long now_ = clock64();\
        acc##k += (u64)(now_ - t_);      \
        t_ = now_;                       \
    }

template <bool EXACT, int NC, bool MG, bool TP = false>
__global__ void __launch_bounds__(FK_RES_MAX_THREADS, 1)
fk_resident_kernel(const __grid_constant__ TileArgs A, const __grid_constant__ ResGeom G) {
    extern __shared__ __align__(16) float fk_res_smem[];
    // which stimuli are active at a step (the typed schedule of fk_core.h: doubles, fmod, 64-bit remainders) is worked out
    // for 32 steps at a time by the first warp, one step per lane -- not by one thread on every step's critical path
    __shared__ unsigned s_mask[32];
    ResCta X;
    res_setup(A, G, blockIdx.x, blockIdx.y, gridDim.y, fk_res_smem, X);
    const int tid = threadIdx.x, nthr = blockDim.x;
    ResThread T;
    res_thread_setup<NC>(A, G, X, tid, nthr, T);
    res_load(A, G, X, tid, nthr);
    __syncthreads();
    // development (FK_RES_TIMING=<cta>): thread 0 of that CTA times the phases; every warp of it also records when it
    // left each phase (max over warps, relative to the step's start): what the step really waits for
#ifndef FK_RES_WARP_TIMING
#define FK_RES_WARP_TIMING 0   // 1: development build whose timed CTA also records when its slowest warp leaves each phase
#endif
    const bool tcta = G.timing != nullptr && (int)blockIdx.x == (int)(G.spin_limit >> 25) && blockIdx.y == 0;
    const bool timed = tcta && tid == 0;
    __shared__ unsigned s_wmax[3];
    __shared__ long long s_t0;
    u64 acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0, acc4 = 0, acc5 = 0, acc7 = 0;
    long long t_ = timed ? clock64() : 0;
    for (int s = 0; s < G.nsteps; ++s) {
        if ((s & 31) == 0) {   // (every thread passed the barrier that ended step s - 1: nobody reads the old masks)
            if (tid < 32) s_mask[tid] = res_mask(A, X, s + tid);
            __syncthreads();
        }
        const unsigned mask = s_mask[s & 31];
        if (FK_RES_WARP_TIMING && tcta) {
            if (tid == 0) { s_wmax[0] = s_wmax[1] = s_wmax[2] = 0; s_t0 = clock64(); }
            __syncthreads();
        }
        if (TP) {   // two-pass step: first derivatives of the tile (and the cells at a physical edge), then the cells
            res_phase_a<EXACT, MG>(A, G, X, s, mask, tid, nthr);
            if (FK_RES_WARP_TIMING && tcta && (tid & 31) == 0) atomicMax(&s_wmax[2], (unsigned)(clock64() - s_t0));
            __syncthreads();
        }
#pragma unroll 1
        for (int phase = 0; phase < 2; ++phase) {
            // phase 0: the ring, published to the neighbours' mailboxes as it is computed; phase 1: the interior, while
            // those records travel.  ONE call site: the body exists once and stays inside the instruction cache.
            res_phase<EXACT, NC, MG, false, TP>(A, G, X, T, s, phase, mask, tid, nthr);
            if (FK_RES_WARP_TIMING && tcta && (tid & 31) == 0) atomicMax(&s_wmax[phase], (unsigned)(clock64() - s_t0));
            if (phase == 0) { FK_TICK(0) } else { FK_TICK(1) }
        }
        if (s == G.nsteps - 1) break;
        if (!res_halo(G, X, T, s, tid, nthr)) __trap();   // a lost neighbour must not hang the device
        FK_TICK(2)
        __syncthreads();
        FK_TICK(3)
        if (FK_RES_WARP_TIMING && timed) { acc4 += s_wmax[0]; acc5 += s_wmax[1]; acc7 += s_wmax[2]; }
    }
    if (timed) {
        G.timing[0] = acc0; G.timing[1] = acc1; G.timing[2] = acc2; G.timing[3] = acc3; G.timing[4] = acc4; G.timing[5] = acc5; G.timing[7] = acc7;
        G.timing[6] = (u64)G.nsteps;
    }
}

// The cluster transport (fk_resident.h): grid (tiles, tissues), the tiles of a tissue = one thread-block cluster.  Ring
// cells go straight into the neighbours' halos (distributed shared memory); one cluster barrier per Euler step orders
// them -- its release / acquire pair also stands for the step's block barrier.
template <bool EXACT, int NC>
__global__ void __launch_bounds__(FK_RES_MAX_THREADS, 1)
fk_cluster_kernel(const __grid_constant__ TileArgs A, const __grid_constant__ ResGeom G) {
    namespace cg = cooperative_groups;
    extern __shared__ __align__(16) float fk_res_smem[];
    __shared__ unsigned s_mask[32];
    cg::cluster_group cl = cg::this_cluster();
    ResCta X;
    res_setup(A, G, blockIdx.x, blockIdx.y, gridDim.y, fk_res_smem, X);
    // cluster rank == tile index (the cluster spans the grid's x dimension)
    if (X.has_n) X.peer[0] = cl.map_shared_rank(fk_res_smem, X.tile - G.ntc);
    if (X.has_s) X.peer[1] = cl.map_shared_rank(fk_res_smem, X.tile + G.ntc);
    if (X.has_w) X.peer[2] = cl.map_shared_rank(fk_res_smem, X.tile - 1);
    if (X.has_e) X.peer[3] = cl.map_shared_rank(fk_res_smem, X.tile + 1);
    const int tid = threadIdx.x, nthr = blockDim.x;
    ResThread T;
    res_thread_setup<NC>(A, G, X, tid, nthr, T);
    res_load(A, G, X, tid, nthr);
    cl.sync();   // every CTA of the cluster is running (its shared memory may be written) and has loaded its tile
    const bool timed = G.timing != nullptr && tid == 0 && blockIdx.x == 0 && blockIdx.y == 0;
    u64 acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0;
    long long t_ = timed ? clock64() : 0;
    for (int s = 0; s < G.nsteps; ++s) {
        if ((s & 31) == 0) {
            if (tid < 32) s_mask[tid] = res_mask(A, X, s + tid);
            __syncthreads();
        }
        const unsigned mask = s_mask[s & 31];
#pragma unroll 1
        for (int phase = 0; phase < 2; ++phase) {
            res_phase<EXACT, NC, false, true>(A, G, X, T, s, phase, mask, tid, nthr);
            if (phase == 0) { FK_TICK(0) } else { FK_TICK(1) }
        }
        if (s == G.nsteps - 1) break;
        // step s read buffer s & 1 (own cells + halo) and wrote the other one, here and in the neighbours' halos; after
        // the barrier everybody reads that one and overwrites buffer s & 1, which nobody reads any more
        if (G.spin_limit & 4u) __syncthreads();                       // (development: timing floor, wrong results)
        else if (G.spin_limit & 2u) {
            __syncthreads();
            asm volatile("barrier.cluster.arrive.relaxed.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
        } else cl.sync();
        FK_TICK(2)
    }
    if (timed) {
        G.timing[0] = acc0; G.timing[1] = acc1; G.timing[2] = acc2; G.timing[3] = acc3;
        G.timing[6] = (u64)G.nsteps;
    }
}

template <bool EXACT, int NC>
int launch_cluster_nc(const ResPlan& P, const TileArgs& A, int batch, cudaStream_t st) {
    static bool attr_set_dev[64] = {false};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
    cudaError_t e;
    if (!attr_set_dev[dev]) {
        e = cudaFuncSetAttribute(fk_cluster_kernel<EXACT, NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, FK_RES_SMEM_OPTIN);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(fk_cluster_kernel<EXACT, NC>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (e != cudaSuccess) return (int)e;
        attr_set_dev[dev] = true;
    }
    const int ntiles = P.G.ntr * P.G.ntc;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ntiles, batch);
    cfg.blockDim = dim3(P.threads);
    cfg.dynamicSmemBytes = (size_t)P.smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = ntiles; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    ResGeom G = P.G;
    static const bool timing = getenv("FK_RES_TIMING") != nullptr;
    if (timing) {
        void* sym = nullptr;
        if (cudaGetSymbolAddress(&sym, g_res_timing) == cudaSuccess) G.timing = (u64*)sym;
    }
    G.spin_limit = getenv("FK_CL_DEBUG") ? (unsigned)atoi(getenv("FK_CL_DEBUG")) : 0u;   // development switches
    return (int)cudaLaunchKernelEx(&cfg, fk_cluster_kernel<EXACT, NC>, A, G);
}

template <bool EXACT>
int launch_cluster_t(const ResPlan& P, const TileArgs& A, int batch, cudaStream_t st) {
    if (P.G.nc == 1) return launch_cluster_nc<EXACT, 1>(P, A, batch, st);
    if (P.G.nc == 2) return launch_cluster_nc<EXACT, 2>(P, A, batch, st);
    return launch_cluster_nc<EXACT, 4>(P, A, batch, st);
}

// clusters of this shape the device can hold at once (0: the shape cannot be launched)
template <bool EXACT, int NC>
int cluster_capacity_nc(int ntiles, int threads, size_t smem) {
    cudaError_t e = cudaFuncSetAttribute(fk_cluster_kernel<EXACT, NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, FK_RES_SMEM_OPTIN);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(fk_cluster_kernel<EXACT, NC>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ntiles, 1);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = ntiles; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    int n = 0;
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveClusters(&n, fk_cluster_kernel<EXACT, NC>, &cfg);
    if (e != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

template <bool EXACT, int NC, bool MG, bool TP = false>
int launch_nc(const ResPlan& P, const TileArgs& A, const ResGeom& G, int batch, cudaStream_t st) {
    static bool attr_set_dev[64] = {false};   // the opt-in is per device
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
    bool& attr_set = attr_set_dev[dev];
    cudaError_t e;
    if (!attr_set) {
        e = cudaFuncSetAttribute(fk_resident_kernel<EXACT, NC, MG, TP>, cudaFuncAttributeMaxDynamicSharedMemorySize, FK_RES_SMEM_OPTIN);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    void* args[2] = {(void*)&A, (void*)&G};
    return (int)cudaLaunchCooperativeKernel((const void*)fk_resident_kernel<EXACT, NC, MG, TP>, dim3(G.ntr * G.ntc, batch),
                                            dim3(P.threads), args, (size_t)P.smem_bytes, st);
}

template <bool EXACT>
int occupancy_nc(int nc, int mg, int threads, size_t smem) {
    int n = 0;
    cudaError_t e;
#define FK_OCC(NC, MG)                                                                                                       \
    e = cudaFuncSetAttribute(fk_resident_kernel<EXACT, NC, MG>, cudaFuncAttributeMaxDynamicSharedMemorySize, FK_RES_SMEM_OPTIN); \
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, fk_resident_kernel<EXACT, NC, MG>, threads, smem);
    if (mg) { if (nc == 1) { FK_OCC(1, true) } else if (nc == 2) { FK_OCC(2, true) } else { FK_OCC(4, true) } }
    else { if (nc == 1) { FK_OCC(1, false) } else if (nc == 2) { FK_OCC(2, false) } else { FK_OCC(4, false) } }
#undef FK_OCC
    if (e != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

template <bool EXACT>
int launch_t(const ResPlan& P, const TileArgs& A, int batch, cudaStream_t st) {
    // tags start at zero: a record is valid from the step that writes tag >= 1
    cudaError_t e = cudaMemsetAsync(P.G.xchg, 0, (size_t)P.xchg_bytes, st);
    if (e != cudaSuccess) return (int)e;
    ResGeom G = P.G;
    static const bool timing = getenv("FK_RES_TIMING") != nullptr;
    if (timing) {
        void* sym = nullptr;
        if (cudaGetSymbolAddress(&sym, g_res_timing) == cudaSuccess) G.timing = (u64*)sym;
        G.spin_limit = (G.spin_limit & ((1u << 25) - 1)) | ((unsigned)atoi(getenv("FK_RES_TIMING")) << 25);   // which CTA
    }
    if (G.tp) return G.mg ? launch_nc<EXACT, 4, true, true>(P, A, G, batch, st) : launch_nc<EXACT, 4, false, true>(P, A, G, batch, st);
    if (G.mg) {
        if (G.nc == 1) return launch_nc<EXACT, 1, true>(P, A, G, batch, st);
        if (G.nc == 2) return launch_nc<EXACT, 2, true>(P, A, G, batch, st);
        return launch_nc<EXACT, 4, true>(P, A, G, batch, st);
    }
    if (G.nc == 1) return launch_nc<EXACT, 1, false>(P, A, G, batch, st);
    if (G.nc == 2) return launch_nc<EXACT, 2, false>(P, A, G, batch, st);
    return launch_nc<EXACT, 4, false>(P, A, G, batch, st);
}

}  // namespace

int launch_resident(const ResPlan& P, const TileArgs& A, int exact, int batch, cudaStream_t st) {
    if (P.G.cluster) return exact ? launch_cluster_t<true>(P, A, batch, st) : launch_cluster_t<false>(P, A, batch, st);
    return exact ? launch_t<true>(P, A, batch, st) : launch_t<false>(P, A, batch, st);
}

int cluster_capacity(int exact, int nc, int ntiles, int threads, long long smem_bytes) {
    if (smem_bytes > FK_RES_SMEM_OPTIN || ntiles < 1 || ntiles > FK_CLUSTER_MAX) return 0;
    static int memo[2][3][FK_CLUSTER_MAX + 1][17][16];   // [exact][nc][tiles][threads / 32][device]: 0 unknown, else n + 1
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) dev = 0;
    const int ni = nc == 1 ? 0 : (nc == 2 ? 1 : 2), ti = threads / 32;
    // (the shared-memory size changes the answer only through "fits / does not fit one CTA per SM", tested above)
    int& m = memo[exact ? 1 : 0][ni][ntiles][ti > 16 ? 16 : ti][dev];
    if (m) return m - 1;
    int n;
    if (exact) n = nc == 1 ? cluster_capacity_nc<true, 1>(ntiles, threads, (size_t)smem_bytes)
                 : nc == 2 ? cluster_capacity_nc<true, 2>(ntiles, threads, (size_t)smem_bytes)
                           : cluster_capacity_nc<true, 4>(ntiles, threads, (size_t)smem_bytes);
    else n = nc == 1 ? cluster_capacity_nc<false, 1>(ntiles, threads, (size_t)smem_bytes)
             : nc == 2 ? cluster_capacity_nc<false, 2>(ntiles, threads, (size_t)smem_bytes)
                       : cluster_capacity_nc<false, 4>(ntiles, threads, (size_t)smem_bytes);
    m = n + 1;
    return n;
}

// development: the counters of the most recent timed launch (synchronises the device)
int resident_timing(unsigned long long* out8) {
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaMemcpyFromSymbol(out8, g_res_timing, sizeof(unsigned long long) * 8);
    return (int)e;
}

// CTAs of the resident kernel the device can hold at once (cooperative launch limit) for this CTA shape
int resident_capacity(int exact, int nc, int mg, int threads, long long smem_bytes, int num_sms) {
    if (smem_bytes > FK_RES_SMEM_OPTIN) return 0;
    const int n = exact ? occupancy_nc<true>(nc, mg, threads, (size_t)smem_bytes) : occupancy_nc<false>(nc, mg, threads, (size_t)smem_bytes);
    return n * num_sms;
}

}  // namespace fk
