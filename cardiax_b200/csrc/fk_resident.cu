// fk_resident.cu -- the resident kernel (body: fk_resident.h) and its cooperative launcher.  sm_100a only.
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

template <bool EXACT, int NC, bool MG>
__global__ void __launch_bounds__(FK_RES_MAX_THREADS, 1)
fk_resident_kernel(const __grid_constant__ TileArgs A, const __grid_constant__ ResGeom G) {
    extern __shared__ __align__(16) float fk_res_smem[];
    __shared__ unsigned s_mask[2];
    ResCta X;
    res_setup(A, G, blockIdx.x, blockIdx.y, gridDim.y, fk_res_smem, X);
    const int tid = threadIdx.x, nthr = blockDim.x;
    ResThread T;
    res_thread_setup<NC>(A, G, X, tid, nthr, T);
    res_load(A, G, X, tid, nthr);
    if (tid == 0) s_mask[0] = res_mask(A, X, 0);
    __syncthreads();
    const bool timed = G.timing != nullptr && tid == 0 && blockIdx.x == 0 && blockIdx.y == 0;
    u64 acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0;
    long long t_ = timed ? clock64() : 0;
    for (int s = 0; s < G.nsteps; ++s) {
        const unsigned mask = s_mask[s & 1];
#pragma unroll 1
        for (int phase = 0; phase < 2; ++phase) {
            // phase 0: the ring, published to the neighbours' mailboxes as it is computed; phase 1: the interior, while
            // those records travel.  ONE call site: the body exists once and stays inside the instruction cache.
            res_phase<EXACT, NC, MG>(A, G, X, T, s, phase, mask, tid, nthr);
            if (phase == 0) { FK_TICK(0) } else { FK_TICK(1) }
        }
        if (s == G.nsteps - 1) break;
        if (tid == 0) s_mask[(s + 1) & 1] = res_mask(A, X, s + 1);
        if (!res_halo(G, X, T, s, tid, nthr)) __trap();   // a lost neighbour must not hang the device
        FK_TICK(2)
        __syncthreads();
        FK_TICK(3)
    }
    if (timed) {
        G.timing[0] = acc0; G.timing[1] = acc1; G.timing[2] = acc2; G.timing[3] = acc3;
        G.timing[6] = (u64)G.nsteps;
    }
}

template <bool EXACT, int NC, bool MG>
int launch_nc(const ResPlan& P, const TileArgs& A, const ResGeom& G, int batch, cudaStream_t st) {
    static bool attr_set_dev[64] = {false};   // the opt-in is per device
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
    bool& attr_set = attr_set_dev[dev];
    cudaError_t e;
    if (!attr_set) {
        e = cudaFuncSetAttribute(fk_resident_kernel<EXACT, NC, MG>, cudaFuncAttributeMaxDynamicSharedMemorySize, FK_RES_SMEM_OPTIN);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    void* args[2] = {(void*)&A, (void*)&G};
    return (int)cudaLaunchCooperativeKernel((const void*)fk_resident_kernel<EXACT, NC, MG>, dim3(G.ntr * G.ntc, batch),
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
    }
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
    return exact ? launch_t<true>(P, A, batch, st) : launch_t<false>(P, A, batch, st);
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
