// fk_resident.cu -- the resident kernel (body: fk_resident.h) and its cooperative launcher.  sm_100a only.
#include <cuda_runtime.h>

#include "fk_resident.cuh"

namespace fk {

namespace {

__device__ __forceinline__ void flag_release(unsigned* p, unsigned v) {
    __threadfence();   // the CTA's ring stores (ordered before this thread by the block barrier) become visible first
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned flag_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

template <bool EXACT>
__global__ void __launch_bounds__(512, 1)
fk_resident_kernel(const __grid_constant__ TileArgs A, const __grid_constant__ ResGeom G) {
    extern __shared__ __align__(16) float fk_res_smem[];
    __shared__ unsigned s_mask[2];
    ResCta X;
    res_setup(A, G, blockIdx.x, blockIdx.y, fk_res_smem, X);
    const int tid = threadIdx.x, nthr = blockDim.x, warp = tid >> 5, lane = tid & 31, nwarps = nthr >> 5;
    res_load(A, G, X, tid, nthr);
    if (tid == 0) s_mask[0] = res_mask(A, X, 0);
    __syncthreads();
    unsigned* const flags = G.flags + (long long)blockIdx.y * (G.ntr * G.ntc);
    for (int s = 0; s < G.nsteps; ++s) {
        const unsigned mask = s_mask[s & 1];
        const bool last = s == G.nsteps - 1;
        res_phase<EXACT>(A, G, X, s, 0, mask, tid, nthr);          // ring: new u also to the exchange plane
        if (!last) {
            __syncthreads();
            if (tid == 0) {
                flag_release(flags + X.tile, (unsigned)(s + 1));
                s_mask[(s + 1) & 1] = res_mask(A, X, s + 1);
            }
        }
        res_phase<EXACT>(A, G, X, s, 1, mask, tid, nthr);          // interior, while the flag travels
        if (!last) {
            for (int j = warp; j < FK_RES_JOBS; j += nwarps) {
                const int nb = res_job_neighbour(G, X, j);
                if (nb < 0) continue;
                unsigned polls = 0;
                while (flag_acquire(flags + nb) < (unsigned)(s + 1))
                    if (++polls > G.spin_limit) __trap();           // a lost neighbour must not hang the device
                res_job_load(A, G, X, s, j, lane, 32);
            }
            __syncthreads();
        }
    }
}

template <bool EXACT>
int launch_t(const ResPlan& P, const TileArgs& A, int batch, cudaStream_t st) {
    static bool attr_set = false;
    cudaError_t e;
    if (!attr_set) {
        e = cudaFuncSetAttribute(fk_resident_kernel<EXACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    e = cudaMemsetAsync(P.G.flags, 0, sizeof(unsigned) * (size_t)P.G.ntr * P.G.ntc * batch, st);
    if (e != cudaSuccess) return (int)e;
    void* args[2] = {(void*)&A, (void*)&P.G};
    e = cudaLaunchCooperativeKernel((const void*)fk_resident_kernel<EXACT>, dim3(P.G.ntr * P.G.ntc, batch), dim3(P.threads),
                                    args, (size_t)P.smem_bytes, st);
    return (int)e;
}

}  // namespace

int launch_resident(const ResPlan& P, const TileArgs& A, int exact, int batch, cudaStream_t st) {
    return exact ? launch_t<true>(P, A, batch, st) : launch_t<false>(P, A, batch, st);
}

// CTAs of the resident kernel the device can hold at once (cooperative launch limit) for this CTA shape
int resident_capacity(int exact, int threads, long long smem_bytes, int num_sms) {
    if (smem_bytes > 227 * 1024) return 0;
    cudaError_t e = exact ? cudaFuncSetAttribute(fk_resident_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)
                          : cudaFuncSetAttribute(fk_resident_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    int n = 0;
    if (e == cudaSuccess)
        e = exact ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, fk_resident_kernel<true>, threads, (size_t)smem_bytes)
                  : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, fk_resident_kernel<false>, threads, (size_t)smem_bytes);
    if (e != cudaSuccess) { cudaGetLastError(); return 0; }
    return n * num_sms;
}

}  // namespace fk
