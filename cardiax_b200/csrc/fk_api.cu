// fk_api.cu -- kernels and the C ABI (include/fk.h) of libfk.so.  sm_100a only.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <mutex>
#include <vector>

#include "../../include/fk.h"
#include "fk_core.h"
#include "fk_tile.h"
#include "fk_stream.cuh"
#include "fk_resident.cuh"
#include "fk_driver.h"
#include "fk_aux.cuh"

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, const char* a = "") {
    snprintf(g_err, sizeof(g_err), fmt, a);
    return code;
}
int cuda_fail(cudaError_t e, const char* where) {
    snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString(e));
    return (int)e;
}
#define FK_CUDA(call)                                     \
    do {                                                  \
        cudaError_t e_ = (call);                          \
        if (e_ != cudaSuccess) return cuda_fail(e_, #call); \
    } while (0)

// diagnostics: the launch counter is shared (atomic); "last plan / last kernel" describe the calling thread's last call
std::atomic<long long> g_launches{0};
thread_local int g_last_plan[8] = {0, 0, 0, 0, 0, 0, 0, 0};
thread_local const char* g_last_kernel = "";
int num_sms() {
    static int sms[fk::FK_MAX_DEVICES] = {0};   // per device
    const int dev = fk::cur_device();
    if (!sms[dev]) {
        int n = 0;
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        sms[dev] = n > 0 ? n : 148;
    }
    return sms[dev];
}

}  // namespace

namespace fk {
int api_fail(int code, const char* msg) { return fail(code, "%s", msg); }
int api_cuda_fail(int e, const char* where) { return cuda_fail((cudaError_t)e, where); }
void api_count_launch(int n) { g_launches += n; }
}  // namespace fk

namespace {

// ------------------------------------------------------------------ kernels
template <bool EXACT>
__global__ void __launch_bounds__(256) fk_tile_kernel(const __grid_constant__ fk::TileArgs A) {
    extern __shared__ __align__(16) float fk_smem[];
    fk::TileCtx X;
    fk::tile_setup(A, blockIdx.x, blockIdx.y, fk_smem, X);
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5, ntx = 32, nty = blockDim.x >> 5;
    fk::tile_load(A, X, tx, ty, ntx, nty);
    __syncthreads();
    float* Uc = X.U0;
    float* Un = X.U1;
    for (int s = 1; s <= A.T; ++s) {
        fk::tile_grad<EXACT>(A, X, s, Uc, tx, ty, ntx, nty);
        __syncthreads();
        fk::tile_update<EXACT>(A, X, s, Uc, Un, tx, ty, ntx, nty);
        __syncthreads();
        float* t = Uc;
        Uc = Un;
        Un = t;
    }
}

// low-latency single-step kernel for small tissues (fk_wide.h): one thread per 4 cells of a row, no barrier
template <bool EXACT>
__global__ void __launch_bounds__(128) fk_wide_kernel(const __grid_constant__ fk::TileArgs A) {
    const int wq = A.W >> 2;
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= (long long)A.H * wq) return;
    const int row = (int)(t / wq), c = 4 * (int)(t - (long long)row * wq);
    fk::wide_thread<EXACT>(A, blockIdx.y, row, c, fk::wide_mask(A, blockIdx.y));
}

// solve.gradient on (outer, n, inner)
__global__ void fk_gradient_kernel(const float* __restrict__ a, float* __restrict__ out, long long outer, long long n,
                                   long long inner) {
    const long long total = outer * n * inner;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long k = idx % inner, i = (idx / inner) % n, o = idx / (inner * n);
        const float* base = a + o * n * inner + k;
        const int kind = i < 2 ? fk::FWD : (i >= n - 2 ? fk::BWD : fk::CEN);
        float k0, k1, k2, k3;
        int o0, o1, o2, o3;
        fk::kind_coeffs(kind, k0, k1, k2, k3, o0, o1, o2, o3);
        out[idx] = fk::tap4<true>(k0, k1, k2, k3, base[(i + o0) * inner], base[(i + o1) * inner], base[(i + o2) * inner],
                                  base[(i + o3) * inner]);
    }
}

// D_x, D_y of solve.py:53-54
__global__ void fk_dgrad_kernel(const float* __restrict__ D, float* __restrict__ DX, float* __restrict__ DY, int H, int W,
                                float dx, int phys_top, int phys_bot) {
    // a block walks whole rows (no 64-bit division per cell: the flat-index form of this kernel took 150 us on 4096^2)
    const long long plane = (long long)H * W;
    const float* Ds = D + blockIdx.y * plane;
    for (int row = (int)blockIdx.x; row < H; row += (int)gridDim.x) {
        const long long r0 = blockIdx.y * plane + (long long)row * W;
        for (int col = (int)threadIdx.x; col < W; col += (int)blockDim.x) {
            float gx, gy;
            fk::dgrad_cell(Ds, H, W, dx, phys_top, phys_bot, row, col, gx, gy);
            DX[r0 + col] = gx;
            DY[r0 + col] = gy;
        }
    }
}

// solve.stimulate on one array
__global__ void fk_stimulate_kernel(const float* __restrict__ x, float* __restrict__ out, long long n,
                                    const fk::StimDev* __restrict__ stims, int n_stim, double t, int t_is_int) {
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n;
         idx += (long long)gridDim.x * blockDim.x) {
        float st = 0.0f;
        for (int i = 0; i < n_stim; ++i) {
            const fk::StimDev s = stims[i];
            if (s.field && fk::stim_on(s, t, t_is_int)) {
                const float f = s.field[idx];
                if (f != 0.0f) st = f;
            }
        }
        out[idx] = st != 0.0f ? st : x[idx];
    }
}

// Heun stages (solve.py:73-85): out = y + k * h, or out = y + (k1 + k2) * h, on the three state arrays at once
template <bool EXACT>
__global__ void fk_heun_stage_kernel(const float* __restrict__ yv, const float* __restrict__ yw, const float* __restrict__ yu,
                                     const float* __restrict__ av, const float* __restrict__ aw, const float* __restrict__ au,
                                     const float* __restrict__ bv, const float* __restrict__ bw, const float* __restrict__ bu,
                                     float h, float* __restrict__ ov, float* __restrict__ ow, float* __restrict__ ou, long long n) {
    typedef fk::Num<EXACT> N;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float kv = av[i], kw = aw[i], ku = au[i];
        if (bv) { kv = N::add(kv, bv[i]); kw = N::add(kw, bw[i]); ku = N::add(ku, bu[i]); }
        ov[i] = fk::euler<EXACT>(yv[i], kv, h);
        ow[i] = fk::euler<EXACT>(yw[i], kw, h);
        ou[i] = fk::euler<EXACT>(yu[i], ku, h);
    }
}

// fast Heun: y + (E(E(y)) - y) / 2 on the three state arrays
__global__ void fk_heun_combine_kernel(const float* __restrict__ yv, const float* __restrict__ yw, const float* __restrict__ yu,
                                       const float* __restrict__ ev, const float* __restrict__ ew, const float* __restrict__ eu,
                                       float* __restrict__ ov, float* __restrict__ ow, float* __restrict__ ou, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float a = yv[i], b = yw[i], c = yu[i];
        ov[i] = fmaf(0.5f, __fsub_rn(ev[i], a), a);
        ow[i] = fmaf(0.5f, __fsub_rn(ew[i], b), b);
        ou[i] = fmaf(0.5f, __fsub_rn(eu[i], c), c);
    }
}

// exhaustive check of the exact-mode divisions against __fdiv_rn: every significand (both signs), every exponent at which
// a quotient can be a denormal number plus a few ordinary and extreme ones, every divisor of a run -- each through the
// sequence the cell update uses for it (fk_core.h: divd for dx and Cm, divd / divd_tie for tau_d and tau_0, divd above
// div_lo for the others)
__global__ void fk_divcheck_kernel(fk::Consts K, unsigned long long* bad) {
    const float divisors[10] = {K.dx, K.Cm, K.tau_d, K.tau_0, K.two_tau_si, K.tau_v_plus, K.tau_v1_minus, K.tau_v2_minus,
                                K.tau_w_plus, K.tau_w_minus};
    const double recips[10] = {K.yd_dx, K.yd_Cm, K.yd_tau_d, K.yd_tau_0, K.yd_two_tau_si, K.yd_tvp, K.yd_tvm1, K.yd_tvm2,
                               K.yd_twp, K.yd_twm};
    unsigned long long local = 0;
    for (unsigned m = blockIdx.x * blockDim.x + threadIdx.x; m < (1u << 23); m += gridDim.x * blockDim.x)
        for (int e = 0; e < 40; ++e) {
            // exponent fields 0 (denormals, +-0) .. 31, then 2^-40, 2^0, 2^40, 2^100 and the largest finite binade, inf / NaN
            const unsigned ex = e < 32 ? (unsigned)e : (e == 32 ? 87u : e == 33 ? 127u : e == 34 ? 167u : e == 35 ? 227u : e == 36 ? 254u : e == 37 ? 255u : 64u + e);
            const float a = __uint_as_float((ex << 23) | m);
            for (int i = 0; i < 10; ++i) {
                const float x = (i & 1) ? -a : a;
                float q;
                if (i < 2) q = fk::Num<true>::divd(x, divisors[i], recips[i]);
                else if (i < 4) q = K.tie_num ? fk::Num<true>::divd_tie(x, divisors[i], recips[i], i == 2 ? K.hb_tau_d : K.hb_tau_0)
                                              : fk::Num<true>::divd(x, divisors[i], recips[i]);
                else if (fabsf(x) > K.div_lo || x == 0.0f) q = fk::Num<true>::divd(x, divisors[i], recips[i]);
                else continue;   // below div_lo the cell is redone with IEEE divisions
                const float ref = __fdiv_rn(x, divisors[i]);
                if (__float_as_uint(q) != __float_as_uint(ref) && !(q != q && ref != ref)) ++local;
            }
        }
    // ... and the quotient inside tanh (Num<true>::div_nr) for every one of the 2^32 arguments
    for (unsigned long long x = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; x < (1ull << 32);
         x += (unsigned long long)gridDim.x * blockDim.x) {
        const float f = __uint_as_float((unsigned)x);
        if (__float_as_uint(fk::tanh_xla<true>(f, false)) != __float_as_uint(fk::tanh_xla<true>(f, true))) ++local;
    }
    if (local) atomicAdd(bad, local);
}

// ------------------------------------------------------------------ host helpers
fk::Consts make_consts(const FkParams& p, float dt, float dx) {
    static_assert(sizeof(FkParams) == 14 * sizeof(float), "FkParams layout");
    return fk::make_consts(&p.tau_v_plus, dt, dx);
}

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct Workspace {
    float *DX, *DY, *pv, *pw, *pu;
    fk::StimDev* stims;
    fk::u64* xchg;
    size_t xchg_bytes;
    size_t bytes;
};

Workspace carve(void* base, int H, int W, int batch, int n_stim, int d_batched) {
    Workspace ws;
    char* p = (char*)base;
    const size_t plane = (size_t)H * W * sizeof(float);
    const size_t dplanes = d_batched ? (size_t)batch : 1;
    size_t off = 0;
    ws.DX = (float*)(p + off); off = align_up(off + plane * dplanes, 256);
    ws.DY = (float*)(p + off); off = align_up(off + plane * dplanes, 256);
    ws.pv = (float*)(p + off); off = align_up(off + plane * batch, 256);
    ws.pw = (float*)(p + off); off = align_up(off + plane * batch, 256);
    ws.pu = (float*)(p + off); off = align_up(off + plane * batch, 256);
    ws.stims = (fk::StimDev*)(p + off); off = align_up(off + sizeof(fk::StimDev) * (size_t)std::max(1, batch * n_stim), 256);
    ws.xchg_bytes = (size_t)fk::res_xchg_bytes(H, W, batch);
    ws.xchg = ws.xchg_bytes ? (fk::u64*)(p + off) : nullptr; off = align_up(off + ws.xchg_bytes, 256);
    ws.bytes = off;
    return ws;
}

int upload_stims(const FkStimulus* stimuli, int count, fk::StimDev* dev, cudaStream_t st) {
    if (count <= 0) return 0;
    static_assert(sizeof(fk::StimDev) == sizeof(FkStimulus), "layout");
    // pageable source: the runtime stages it before returning, so the caller's array may die after the call
    FK_CUDA(cudaMemcpyAsync(dev, stimuli, sizeof(FkStimulus) * (size_t)count, cudaMemcpyHostToDevice, st));
    return 0;
}

int launch_dgrad(const float* D, float* DX, float* DY, int H, int W, int planes, float dx, int pt, int pb, cudaStream_t st) {
    const int blocks = std::min(H, 148 * 16);
    const int threads = W >= 256 ? 256 : (W >= 128 ? 128 : 64);
    ++g_launches;
    fk_dgrad_kernel<<<dim3(blocks, planes), threads, 0, st>>>(D, DX, DY, H, W, dx, pt, pb);
    FK_CUDA(cudaGetLastError());
    return 0;
}

// ---- optional per-launch timing (bench.py): CUDA events around every launch, on the launch stream
enum { FK_PROF_MAX_EVENTS = 1 << 20 };   // 2^19 launches between two fk_profile_collect calls; beyond: counted as dropped
struct Prof {
    std::mutex mu;                    // the event lists are shared by every thread that calls into the library
    bool on = false;
    long long dropped = 0;            // launches not timed because the event lists were full (reported, never silent)
    std::vector<cudaEvent_t> ev[2];   // [0] streaming kernel, [1] tile kernel : begin/end pairs
    std::vector<cudaEvent_t> pool;
    double stream_cs = 0;             // cell-steps the timed streaming launches produced
    cudaEvent_t get() {
        cudaEvent_t e;
        if (!pool.empty()) { e = pool.back(); pool.pop_back(); return e; }
        cudaEventCreate(&e);
        return e;
    }
} g_prof;
struct ProfScope {
    int kind; cudaStream_t st; bool active;
    ProfScope(int k, cudaStream_t s) : kind(k), st(s), active(false) {
        if (!g_prof.on) return;
        std::lock_guard<std::mutex> lock(g_prof.mu);
        active = g_prof.ev[k].size() < (size_t)FK_PROF_MAX_EVENTS;
        if (!active) { ++g_prof.dropped; return; }
        cudaEvent_t e = g_prof.get(); cudaEventRecord(e, st); g_prof.ev[kind].push_back(e);
    }
    ~ProfScope() {
        if (!active) return;
        std::lock_guard<std::mutex> lock(g_prof.mu);
        cudaEvent_t e = g_prof.get(); cudaEventRecord(e, st); g_prof.ev[kind].push_back(e);
    }
};

// ---- launch-queue throttle.  A long run enqueues a launch per T steps far faster than the GPU retires them (4096^2: 62 us
// of host time per 94 us launch once the queue is full), and a hardware queue that has filled up makes the driver wait for
// room on a slow path: 500-step segments of 4096^2 then took anything between 22.5 and 48 ms (gpurun_out/r02zs).  So the
// host never runs more than 2 x FK_THROTTLE_EVERY streaming launches ahead: it records an event every FK_THROTTLE_EVERY
// launches and first waits for the one recorded two periods earlier -- at most 768 launches (~70 ms of 4096^2 work) queued, never a full queue (~1000).
enum { FK_THROTTLE_EVERY = 384 };
struct Throttle {
    std::mutex mu;
    cudaEvent_t ev[2] = {nullptr, nullptr};
    bool used[2] = {false, false};
    long long n = 0;
    void tick(cudaStream_t st) {
        std::lock_guard<std::mutex> lock(mu);
        if (++n % FK_THROTTLE_EVERY) return;
        const int k = (int)((n / FK_THROTTLE_EVERY) & 1);
        if (!ev[k] && cudaEventCreateWithFlags(&ev[k], cudaEventDisableTiming) != cudaSuccess) { ev[k] = nullptr; cudaGetLastError(); return; }
        if (used[k]) cudaEventSynchronize(ev[k]);
        used[k] = cudaEventRecord(ev[k], st) == cudaSuccess;
    }
} g_throttle;

struct CudaBackend {
    cudaStream_t st;
    int num_sms() { return ::num_sms(); }
    int occupancy(int T, int exact, int uni, int NT, long long smem) { return fk::stream_occupancy(T, exact, uni, NT, smem); }
    int tiles(fk::TileArgs& A, int exact, int batch, bool) {
        if (A.heun) {   // development: FK_HEUN_TILE="rows,cols" overrides the fused Heun tile
            static int hth = -1, htw = 0;
            if (hth < 0) { const char* e = getenv("FK_HEUN_TILE"); hth = 0; if (e) sscanf(e, "%d,%d", &hth, &htw); }
            if (hth > 0 && htw > 0)
                for (int i = 0; i < A.nreg; ++i) {
                    A.reg[i].th = std::min(hth, A.reg[i].R1 - A.reg[i].R0);
                    A.reg[i].tw = std::min(htw, A.reg[i].C1 - A.reg[i].C0);
                }
        }
        long long floats = 0;
        const int total = fk::finish_regions(A, &floats);
        if (total == 0) return 0;
        const size_t smem = (size_t)floats * sizeof(float);
        if (smem > 227 * 1024) return fail(-3, "tile does not fit shared memory%s");
        ProfScope ps(1, st);
        g_last_kernel = "fk_tile_kernel";
        ++g_launches;
        static size_t attr_set_dev[fk::FK_MAX_DEVICES][2] = {{0, 0}};   // largest dynamic shared memory already allowed,
        size_t* attr_set = attr_set_dev[fk::cur_device()];              // per device and instantiation
        if (smem > attr_set[exact ? 1 : 0]) {
            if (exact) FK_CUDA(cudaFuncSetAttribute(fk_tile_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            else FK_CUDA(cudaFuncSetAttribute(fk_tile_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr_set[exact ? 1 : 0] = smem;
        }
        if (exact) fk_tile_kernel<true><<<dim3(total, batch), 256, smem, st>>>(A);
        else fk_tile_kernel<false><<<dim3(total, batch), 256, smem, st>>>(A);
        FK_CUDA(cudaGetLastError());
        return 0;
    }
    int wide(const fk::TileArgs& A, int exact, int batch) {
        const long long threads = (long long)A.H * (A.W >> 2);
        const unsigned blocks = (unsigned)((threads + 127) / 128);
        g_last_plan[0] = 1; g_last_plan[1] = 128; g_last_plan[2] = 0; g_last_plan[3] = A.W; g_last_plan[4] = 1;
        g_last_plan[5] = (int)blocks; g_last_plan[6] = 0; g_last_plan[7] = 0;
        ProfScope ps(0, st);
        g_last_kernel = "fk_wide_kernel";
        if (ps.active) g_prof.stream_cs += (double)A.H * A.W * batch;
        ++g_launches;
        if (exact) fk_wide_kernel<true><<<dim3(blocks, batch), 128, 0, st>>>(A);
        else fk_wide_kernel<false><<<dim3(blocks, batch), 128, 0, st>>>(A);
        FK_CUDA(cudaGetLastError());
        return 0;
    }
    long long resident_smem_limit() { return 227 * 1024 - 256; }
    bool cluster_ok(const fk::ResPlan& P, int exact) {
        return fk::cluster_capacity(exact, P.G.nc, P.G.ntr * P.G.ntc, P.threads, P.smem_bytes) >= 1;
    }
    int resident(const fk::ResPlan& P, const fk::TileArgs& A, int exact, int batch) {
        g_last_plan[0] = P.G.nsteps; g_last_plan[1] = P.threads; g_last_plan[2] = P.G.ntc; g_last_plan[3] = P.G.tw_max;
        g_last_plan[4] = P.G.th_max; g_last_plan[5] = P.G.ntr; g_last_plan[6] = P.G.nc + 8 * P.G.mg; g_last_plan[7] = (int)P.smem_bytes;
        if (!P.G.cluster) {
            const int cap = fk::resident_capacity(exact, P.G.nc, P.G.mg, P.threads, P.smem_bytes, num_sms());
            if ((long long)P.G.ntr * P.G.ntc * batch > cap) return fail(-3, "resident kernel: the tiles are not co-resident on this device%s");
        }
        ProfScope ps(0, st);
        g_last_kernel = P.G.cluster ? "fk_cluster_kernel" : "fk_resident_kernel";
        if (ps.active) g_prof.stream_cs += (double)A.H * A.W * batch * P.G.nsteps;
        ++g_launches;
        const int rc = fk::launch_resident(P, A, exact, batch, st);
        if (rc) return cuda_fail((cudaError_t)rc, "resident kernel launch");
        return 0;
    }
    int stream(const fk::StreamPlan& P, const fk::TileArgs& A, int exact, int batch) {
        g_last_plan[0] = P.T; g_last_plan[1] = P.G.NT; g_last_plan[2] = P.G.nstrips; g_last_plan[3] = P.G.cstride;
        g_last_plan[4] = P.G.RH; g_last_plan[5] = P.G.nchunks; g_last_plan[6] = P.occ; g_last_plan[7] = (int)P.smem_bytes;
        ProfScope ps(0, st);
        g_last_kernel = "fk_stream_kernel";
        if (ps.active) g_prof.stream_cs += (double)(P.G.row1 - P.G.row0) * A.W * P.T * batch;
        ++g_launches;
        g_throttle.tick(st);
        const int rc = fk::launch_stream(P, A, exact, batch, st);
        if (rc > 0) return cuda_fail((cudaError_t)rc, "streaming kernel launch");
        if (rc < 0) return fail(rc, "streaming kernel: unsupported steps_per_launch%s");
        return 0;
    }
};

int check_common(int H, int W, int batch, const FkParams* params, int n_stim, const FkStimulus* stimuli) {
    if (H < 3 || W < 3) return fail(-1, "tissue must be at least 3 x 3 (the reference's gradient needs n + 2 >= 5)%s");
    if (batch < 1) return fail(-1, "batch must be >= 1%s");
    if (!params) return fail(-1, "params is NULL%s");
    if (n_stim < 0 || n_stim > 32) return fail(-2, "n_stim must be in [0, 32]%s");
    if (n_stim > 0 && !stimuli) return fail(-1, "stimuli is NULL%s");
    return 0;
}

}  // namespace

// ====================================================================== C ABI
extern "C" {

int fk_abi_version(void) { return FK_ABI_VERSION; }

long long fk_launch_count(void) { return g_launches; }

void fk_last_plan(int* out8) {
    for (int i = 0; i < 8; ++i) out8[i] = g_last_plan[i];
}

const char* fk_last_kernel(void) { return g_last_kernel; }

int fk_resident_timing(unsigned long long* out8) {
    if (!out8) return fail(-1, "NULL pointer%s");
    const int rc = fk::resident_timing(out8);
    return rc ? cuda_fail((cudaError_t)rc, "fk_resident_timing") : 0;
}

void fk_profile_enable(int on) { g_prof.on = on != 0; }

int fk_profile_collect(double* stream_ms, long long* stream_launches, double* tile_ms, long long* tile_launches,
                       double* stream_cell_steps) {
    std::lock_guard<std::mutex> lock(g_prof.mu);
    if (stream_cell_steps) *stream_cell_steps = g_prof.stream_cs;
    g_prof.stream_cs = 0;
    double ms[2] = {0, 0};
    long long n[2] = {0, 0};
    for (int k = 0; k < 2; ++k) {
        std::vector<cudaEvent_t>& v = g_prof.ev[k];
        for (size_t i = 0; i + 1 < v.size(); i += 2) {
            FK_CUDA(cudaEventSynchronize(v[i + 1]));
            float t = 0;
            FK_CUDA(cudaEventElapsedTime(&t, v[i], v[i + 1]));
            ms[k] += t;
            ++n[k];
        }
        for (cudaEvent_t e : v) g_prof.pool.push_back(e);
        v.clear();
    }
    if (stream_ms) *stream_ms = ms[0];
    if (stream_launches) *stream_launches = n[0];
    if (tile_ms) *tile_ms = ms[1];
    if (tile_launches) *tile_launches = n[1];
    return 0;
}

long long fk_profile_dropped(void) {
    std::lock_guard<std::mutex> lock(g_prof.mu);
    const long long d = g_prof.dropped;
    g_prof.dropped = 0;
    return d;
}
const char* fk_last_error(void) { return g_err; }

void fk_default_options(FkOptions* opt) {
    if (!opt) return;
    memset(opt, 0, sizeof(*opt));
    opt->phys_top = 1;
    opt->phys_bottom = 1;
}

size_t fk_workspace_bytes(int H, int W, int batch, int n_stim, int d_batched) {
    if (H <= 0 || W <= 0 || batch <= 0 || n_stim < 0) return 0;
    return carve(nullptr, H, W, batch, n_stim, d_batched).bytes;
}

int fk_diffusivity_gradients(const float* D, float* DX, float* DY, int H, int W, int batch, float dx, int phys_top,
                             int phys_bottom, void* stream) {
    if (!D || !DX || !DY) return fail(-1, "NULL pointer%s");
    if (H < 3 || W < 3 || batch < 1) return fail(-1, "bad shape%s");
    return launch_dgrad(D, DX, DY, H, W, batch, dx, phys_top, phys_bottom, (cudaStream_t)stream);
}

int fk_gradient(const float* a, float* out, long long outer, long long n, long long inner, void* stream) {
    if (!a || !out) return fail(-1, "NULL pointer%s");
    if (n < 5) return fail(-1, "gradient needs at least 5 points along the axis%s");
    if (outer < 0 || inner < 0) return fail(-1, "bad shape%s");
    const long long total = outer * n * inner;
    if (total == 0) return 0;
    int blocks = (int)std::min<long long>((total + 255) / 256, 148LL * 32);
    fk_gradient_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(a, out, outer, n, inner);
    FK_CUDA(cudaGetLastError());
    return 0;
}

int fk_stimulate(double t, int t_is_int, const float* x, float* out, int H, int W, const FkStimulus* stimuli, int n_stim,
                 void* workspace, size_t workspace_bytes, void* stream) {
    if (!x || !out) return fail(-1, "NULL pointer%s");
    if (H < 1 || W < 1 || n_stim < 0) return fail(-1, "bad shape%s");
    if (n_stim > 0 && (!workspace || workspace_bytes < sizeof(fk::StimDev) * (size_t)n_stim))
        return fail(-4, "workspace too small%s");
    cudaStream_t st = (cudaStream_t)stream;
    int rc = upload_stims(stimuli, n_stim, (fk::StimDev*)workspace, st);
    if (rc) return rc;
    const long long n = (long long)H * W;
    int blocks = (int)std::min<long long>((n + 255) / 256, 148LL * 32);
    fk_stimulate_kernel<<<blocks, 256, 0, st>>>(x, out, n, (const fk::StimDev*)workspace, n_stim, t, t_is_int);
    FK_CUDA(cudaGetLastError());
    return 0;
}

static int run_euler(const float* v_in, const float* w_in, const float* u_in, float* v_out, float* w_out, float* u_out,
                     const float* D, int d_batched, int H, int W, int batch, const FkParams* params,
                     const FkStimulus* stimuli, int n_stim, double t0, long long nsteps, float dt, float dx,
                     const FkOptions* opt_in, int rhs_mode, void* workspace, size_t workspace_bytes, void* stream,
                     const float* DXext = nullptr, const float* DYext = nullptr, int row0 = 0, int row1 = 0,
                     const fk::SlabMirror* mirror = nullptr, bool* mirrored = nullptr) {
    int rc = check_common(H, W, batch, params, n_stim, stimuli);
    if (rc) return rc;
    if (!v_in || !w_in || !u_in || !v_out || !w_out || !u_out || !D) return fail(-1, "NULL pointer%s");
    FkOptions opt;
    if (opt_in) opt = *opt_in; else fk_default_options(&opt);
    if (opt.steps_per_launch < 0 || opt.steps_per_launch > 8) return fail(-1, "steps_per_launch must be in [0, 8]%s");
    const bool rows_mode = row1 > 0;  // single launch into a row window: no ping-pong scratch, maps from the caller
    const size_t need = rows_mode ? sizeof(fk::StimDev) * (size_t)std::max(1, batch * n_stim)
                                  : fk_workspace_bytes(H, W, batch, n_stim, d_batched);
    if (need && (!workspace || workspace_bytes < need)) return fail(-4, "workspace too small%s");
    if (rows_mode && (!DXext || !DYext)) return fail(-1, "row-window calls need the D_x, D_y maps%s");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t plane_bytes = (size_t)H * W * sizeof(float) * batch;
    if (nsteps <= 0 && !rhs_mode) {
        FK_CUDA(cudaMemcpyAsync(v_out, v_in, plane_bytes, cudaMemcpyDeviceToDevice, st));
        FK_CUDA(cudaMemcpyAsync(w_out, w_in, plane_bytes, cudaMemcpyDeviceToDevice, st));
        FK_CUDA(cudaMemcpyAsync(u_out, u_in, plane_bytes, cudaMemcpyDeviceToDevice, st));
        return 0;
    }
    Workspace ws;
    if (rows_mode) {
        memset(&ws, 0, sizeof(ws));
        ws.stims = (fk::StimDev*)workspace;
    } else {
        ws = carve(workspace, H, W, batch, n_stim, d_batched);
    }
    // a launch during which no stimulus can fire runs without the stimulus machinery (fk_core.h: stims_quiet)
    static_assert(sizeof(fk::StimDev) == sizeof(FkStimulus), "layout");
    if (n_stim > 0 && !rhs_mode && fk::stims_quiet((const fk::StimDev*)stimuli, batch * n_stim, t0, nsteps)) n_stim = 0;
    rc = upload_stims(stimuli, batch * n_stim, ws.stims, st);
    if (rc) return rc;
    if (!DXext || !DYext) {
        rc = launch_dgrad(D, ws.DX, ws.DY, H, W, d_batched ? batch : 1, dx, opt.phys_top, opt.phys_bottom, st);
        if (rc) return rc;
    }
    fk::DriveBuffers B = fk::DriveBuffers();
    B.v_in = v_in; B.w_in = w_in; B.u_in = u_in; B.v_out = v_out; B.w_out = w_out; B.u_out = u_out;
    B.pv = ws.pv; B.pw = ws.pw; B.pu = ws.pu; B.D = D; B.stims = ws.stims;
    B.xchg = rows_mode ? nullptr : ws.xchg;
    B.xchg_bytes = rows_mode ? 0 : (long long)ws.xchg_bytes;
    B.DX = DXext ? DXext : ws.DX;
    B.DY = DYext ? DYext : ws.DY;
    B.mirror = mirror; B.mirrored = mirrored;
    fk::DriveOptions o;
    o.exact = opt.exact; o.steps_per_launch = opt.steps_per_launch; o.kernel = opt.kernel;
    o.phys_top = opt.phys_top; o.phys_bottom = opt.phys_bottom; o.cta_threads = opt.cta_threads;
    o.rows_per_cta = opt.rows_per_cta; o.uniform_diffusivity = opt.uniform_diffusivity;
    o.row0 = row0; o.row1 = row1;
    o.tiles_r = opt.tiles_r; o.tiles_c = opt.tiles_c; o.cells_per_thread = opt.cells_per_thread;
    o.edge_rows = opt.edge_rows; o.edge_colgroups = opt.edge_colgroups; o.maps_global = opt.maps_global;
    o.t_is_int = opt.counter_is_int;
    CudaBackend be;
    be.st = st;
    const char* why = "";
    fk::Consts K = make_consts(*params, dt, dx);
    if (opt.safe_division) fk::set_safe_division(K);   // every exact-mode division the IEEE way
    rc = fk::drive_euler(be, B, d_batched, H, W, batch, K, n_stim, t0, nsteps, o, rhs_mode, &why);
    if (rc && why[0]) return fail(rc, "%s", why);
    return rc;
}

int fk_forward_euler(const float* v_in, const float* w_in, const float* u_in, float* v_out, float* w_out, float* u_out,
                     const float* D, int d_batched, int H, int W, int batch, const FkParams* params,
                     const FkStimulus* stimuli, int n_stim, double t0, double t1, float dt, float dx, const FkOptions* opt,
                     void* workspace, size_t workspace_bytes, void* stream) {
    const long long nsteps = fk::count_steps(t0, t1);
    return run_euler(v_in, w_in, u_in, v_out, w_out, u_out, D, d_batched, H, W, batch, params, stimuli, n_stim, t0, nsteps,
                     dt, dx, opt, 0, workspace, workspace_bytes, stream);
}

size_t fk_heun_workspace_bytes(int H, int W, int batch, int n_stim, int d_batched) {
    const size_t base = fk_workspace_bytes(H, W, batch, n_stim, d_batched);
    if (!base) return 0;
    return base + 9 * align_up((size_t)H * W * sizeof(float) * batch, 256);   // k1, the predictor state, k2
}

int fk_forward_heun(const float* v_in, const float* w_in, const float* u_in, float* v_out, float* w_out, float* u_out,
                    const float* D, int d_batched, int H, int W, int batch, const FkParams* params,
                    const FkStimulus* stimuli, int n_stim, double t0, double t1, float dt, float dx, const FkOptions* opt_in,
                    void* workspace, size_t workspace_bytes, void* stream) {
    int rc = check_common(H, W, batch, params, n_stim, stimuli);
    if (rc) return rc;
    if (!v_in || !w_in || !u_in || !v_out || !w_out || !u_out || !D) return fail(-1, "NULL pointer%s");
    FkOptions opt;
    if (opt_in) opt = *opt_in; else fk_default_options(&opt);
    if (!opt.phys_top || !opt.phys_bottom) return fail(-1, "fk_forward_heun works on whole tissues%s");
    const size_t need = fk_heun_workspace_bytes(H, W, batch, n_stim, d_batched);
    if (!workspace || workspace_bytes < need) return fail(-4, "workspace too small%s");
    cudaStream_t st = (cudaStream_t)stream;
    const long long nsteps = fk::count_steps(t0, t1);
    const size_t plane_bytes = (size_t)H * W * sizeof(float) * batch;
    if (nsteps <= 0) {
        FK_CUDA(cudaMemcpyAsync(v_out, v_in, plane_bytes, cudaMemcpyDeviceToDevice, st));
        FK_CUDA(cudaMemcpyAsync(w_out, w_in, plane_bytes, cudaMemcpyDeviceToDevice, st));
        FK_CUDA(cudaMemcpyAsync(u_out, u_in, plane_bytes, cudaMemcpyDeviceToDevice, st));
        return 0;
    }
    Workspace ws = carve(workspace, H, W, batch, n_stim, d_batched);
    float* extra[9];
    {
        char* p = (char*)workspace + ws.bytes;
        const size_t stride = align_up(plane_bytes, 256);
        for (int i = 0; i < 9; ++i) extra[i] = (float*)(p + i * stride);
    }
    float *k1[3] = {extra[0], extra[1], extra[2]}, *y1[3] = {extra[3], extra[4], extra[5]}, *k2[3] = {extra[6], extra[7], extra[8]};
    rc = upload_stims(stimuli, batch * n_stim, ws.stims, st);
    if (rc) return rc;
    rc = launch_dgrad(D, ws.DX, ws.DY, H, W, d_batched ? batch : 1, dx, 1, 1, st);
    if (rc) return rc;
    fk::DriveOptions o;
    memset(&o, 0, sizeof(o));
    o.exact = opt.exact; o.phys_top = 1; o.phys_bottom = 1; o.kernel = 1; o.t_is_int = opt.counter_is_int;
    CudaBackend be;
    be.st = st;
    fk::Consts K = make_consts(*params, dt, dx);
    if (opt.safe_division) fk::set_safe_division(K);
    const long long n = (long long)H * W * batch;
    const int blocks = (int)std::min<long long>((n + 255) / 256, 148LL * 16);
    const float h_half = (float)((double)dt * 0.5);   // `dt * 0.5` of solve.py:83: exact halving of fl32(dt)
    const float *yv = v_in, *yw = w_in, *yu = u_in;
    const char* why = "";
    auto rhs = [&](const float* sv, const float* sw, const float* su, float** k, double t) {
        fk::DriveBuffers B = fk::DriveBuffers();
        memset(&B, 0, sizeof(B));
        B.v_in = sv; B.w_in = sw; B.u_in = su; B.v_out = k[0]; B.w_out = k[1]; B.u_out = k[2];
        B.D = D; B.DX = ws.DX; B.DY = ws.DY; B.stims = ws.stims;
        return fk::drive_euler(be, B, d_batched, H, W, batch, K, n_stim, t, 1, o, 1, &why);
    };
    if (!opt.exact && opt.steps_per_launch == 0) {
        // fast numerics: a Heun step through the FAST Euler kernels.  With E(y) = y + dt f(y) (one Euler step at counter
        // t): y1 = E(y), E(y1) = y1 + dt k2, hence y + (k1 + k2) dt / 2 = y + (E(E(y)) - y) / 2 -- two single-step
        // launches at the SAME counter (streaming kernel, or the wide kernel on small tissues) and one combine pass,
        // instead of the general tile kernel: 4x the throughput on large tissues at one extra rounding per step.
        // When no stimulus is active at t or t + 1, E at counter t + 1 equals E at counter t, so E(E(y)) is ONE
        // temporally blocked two-step call (streaming kernel T = 2, or the resident kernel on small tissues).
        // The closing pass is folded into the store of the launch that produces E(E(y)) (fk::drive_heun_fast).
        fk::HeunFastBuffers HB;
        HB.v_in = v_in; HB.w_in = w_in; HB.u_in = u_in; HB.v_out = v_out; HB.w_out = w_out; HB.u_out = u_out;
        HB.pv = ws.pv; HB.pw = ws.pw; HB.pu = ws.pu;
        for (int a = 0; a < 3; ++a) { HB.s1[a] = k1[a]; HB.s2[a] = y1[a]; HB.s3[a] = k2[a]; }
        HB.D = D; HB.DX = ws.DX; HB.DY = ws.DY; HB.stims = ws.stims; HB.xchg = ws.xchg; HB.xchg_bytes = (long long)ws.xchg_bytes;
        struct HeunBackend : CudaBackend {
            int blocks;
            int combine(const float* yv, const float* yw, const float* yu, const float* ev, const float* ew, const float* eu,
                        float* ov, float* ow, float* ou, long long n) {
                ++g_launches;
                fk_heun_combine_kernel<<<blocks, 256, 0, st>>>(yv, yw, yu, ev, ew, eu, ov, ow, ou, n);
                FK_CUDA(cudaGetLastError());
                return 0;
            }
            int copy(float* dst, const float* src, long long n) {
                FK_CUDA(cudaMemcpyAsync(dst, src, sizeof(float) * (size_t)n, cudaMemcpyDeviceToDevice, st));
                return 0;
            }
        } hb;
        hb.st = st; hb.blocks = blocks;
        rc = fk::drive_heun_fast(hb, HB, d_batched, H, W, batch, K, (const fk::StimDev*)stimuli, n_stim, t0, nsteps,
                                 opt.uniform_diffusivity, opt.cta_threads, opt.rows_per_cta, /*fold=*/opt.kernel != 1,
                                 /*try_resident=*/opt.kernel == 4,   // opt-in: 24-32 us per step vs 16.5 with two wide launches
                                 &why, opt.counter_is_int);
        return rc ? (why[0] ? fail(rc, "%s", why) : rc) : 0;
    }
    if (opt.steps_per_launch == 2 || (opt.steps_per_launch == 0 && (long long)H * W * batch < (1LL << 20))) {
        // exact numerics, tissues that do not fill the machine (measured: 1.4 - 1.7x the unfused sequence below 2^20 cells,
        // 0.7 - 0.8x above -- the tile's 8-cell apron is recomputed): ONE launch per Heun step (fk::drive_heun: predictor and corrector are the tile kernel's two levels)
        fk::DriveBuffers B = fk::DriveBuffers();
        memset(&B, 0, sizeof(B));
        B.v_in = v_in; B.w_in = w_in; B.u_in = u_in; B.v_out = v_out; B.w_out = w_out; B.u_out = u_out;
        B.pv = ws.pv; B.pw = ws.pw; B.pu = ws.pu;
        B.D = D; B.DX = ws.DX; B.DY = ws.DY; B.stims = ws.stims;
        rc = fk::drive_heun(be, B, d_batched, H, W, batch, K, n_stim, t0, nsteps, opt.exact, h_half, opt.counter_is_int);
        return rc;
    }
    // the unfused sequence (two right-hand-side launches + two stage kernels per step); steps_per_launch = 1 forces it
    for (long long l = 0; l < nsteps; ++l) {
        const double t = t0 + (double)l;   // both stages at the same counter (solve.py:78, 80)
        const bool to_out = ((nsteps - 1 - l) % 2 == 0);
        float *nv = to_out ? v_out : ws.pv, *nw = to_out ? w_out : ws.pw, *nu = to_out ? u_out : ws.pu;
        rc = rhs(yv, yw, yu, k1, t);
        if (rc) return why[0] ? fail(rc, "%s", why) : rc;
        g_launches += 2;
        if (opt.exact) fk_heun_stage_kernel<true><<<blocks, 256, 0, st>>>(yv, yw, yu, k1[0], k1[1], k1[2], nullptr, nullptr, nullptr, dt, y1[0], y1[1], y1[2], n);
        else fk_heun_stage_kernel<false><<<blocks, 256, 0, st>>>(yv, yw, yu, k1[0], k1[1], k1[2], nullptr, nullptr, nullptr, dt, y1[0], y1[1], y1[2], n);
        rc = rhs(y1[0], y1[1], y1[2], k2, t);
        if (rc) return why[0] ? fail(rc, "%s", why) : rc;
        if (opt.exact) fk_heun_stage_kernel<true><<<blocks, 256, 0, st>>>(yv, yw, yu, k1[0], k1[1], k1[2], k2[0], k2[1], k2[2], h_half, nv, nw, nu, n);
        else fk_heun_stage_kernel<false><<<blocks, 256, 0, st>>>(yv, yw, yu, k1[0], k1[1], k1[2], k2[0], k2[1], k2[2], h_half, nv, nw, nu, n);
        FK_CUDA(cudaGetLastError());
        yv = nv; yw = nw; yu = nu;
    }
    return 0;
}

// ---- solve._forward_dormandprince: the CUDA backend of fk::drive_dopri5 (fk_ode.h)
size_t fk_dopri5_workspace_bytes(int H, int W, int batch, int n_stim, int d_batched) {
    const size_t base = fk_workspace_bytes(H, W, batch, n_stim, d_batched);
    if (!base) return 0;
    // y, stage, candidate, k[7], two sets of five interpolation coefficients: 20 States
    return base + 60 * align_up((size_t)H * W * sizeof(float) * batch, 256) + align_up(fk::ode_scratch_bytes(), 256);
}

int fk_odeint_dopri5(const float* v0, const float* w0, const float* u0, float* v_out, float* w_out, float* u_out,
                     const float* D, int d_batched, int H, int W, int batch, const FkParams* params,
                     const FkStimulus* stimuli, int n_stim, const float* ts, int n_ts, float dx, float rtol, float atol,
                     double mxstep, const FkOptions* opt_in, void* workspace, size_t workspace_bytes, void* stream,
                     long long* stats3) {
    int rc = check_common(H, W, batch, params, n_stim, stimuli);
    if (rc) return rc;
    if (!v0 || !w0 || !u0 || !v_out || !w_out || !u_out || !D || !ts) return fail(-1, "NULL pointer%s");
    if (n_ts < 1) return fail(-1, "fk_odeint_dopri5 needs at least one output time%s");
    FkOptions opt;
    if (opt_in) opt = *opt_in; else fk_default_options(&opt);
    const size_t need = fk_dopri5_workspace_bytes(H, W, batch, n_stim, d_batched);
    if (!workspace || workspace_bytes < need) return fail(-4, "workspace too small%s");
    cudaStream_t st = (cudaStream_t)stream;
    Workspace ws = carve(workspace, H, W, batch, n_stim, d_batched);
    const size_t plane_bytes = (size_t)H * W * sizeof(float) * batch;
    const size_t stride = align_up(plane_bytes, 256);
    char* p = (char*)workspace + ws.bytes;
    auto next3 = [&]() { fk::P3 r; for (int a = 0; a < 3; ++a) { r.a[a] = (float*)p; p += stride; } return r; };
    fk::OdeBuffers B;
    B.n = (long long)H * W * batch;
    B.y = next3(); B.ys = next3(); B.yn = next3();
    for (int s = 0; s < 7; ++s) B.k[s] = next3();
    for (int c = 0; c < 2; ++c) for (int j = 0; j < 5; ++j) B.c[c][j] = next3();
    B.out.a[0] = v_out; B.out.a[1] = w_out; B.out.a[2] = u_out;
    fk::OdeScratch scratch;
    scratch.partial = (double*)p;
    FK_CUDA(cudaMemcpyAsync(B.y.a[0], v0, plane_bytes, cudaMemcpyDeviceToDevice, st));
    FK_CUDA(cudaMemcpyAsync(B.y.a[1], w0, plane_bytes, cudaMemcpyDeviceToDevice, st));
    FK_CUDA(cudaMemcpyAsync(B.y.a[2], u0, plane_bytes, cudaMemcpyDeviceToDevice, st));
    rc = upload_stims(stimuli, batch * n_stim, ws.stims, st);
    if (rc) return rc;
    rc = launch_dgrad(D, ws.DX, ws.DY, H, W, d_batched ? batch : 1, dx, 1, 1, st);
    if (rc) return rc;

    struct Backend {
        CudaBackend be;
        fk::DriveOptions o;
        fk::Consts K;
        fk::Dopri T;
        fk::OdeScratch S;
        const float *D, *DX, *DY;
        const fk::StimDev* stims;
        int d_batched, H, W, batch, n_stim, exact;
        long long n;
        const char* why;
        int rhs(const fk::P3& y, const fk::P3& k, float t) {
            fk::DriveBuffers Bf = fk::DriveBuffers();
            memset(&Bf, 0, sizeof(Bf));
            Bf.v_in = y.a[0]; Bf.w_in = y.a[1]; Bf.u_in = y.a[2]; Bf.v_out = k.a[0]; Bf.w_out = k.a[1]; Bf.u_out = k.a[2];
            Bf.D = D; Bf.DX = DX; Bf.DY = DY; Bf.stims = stims;
            const int rc = fk::drive_euler(be, Bf, d_batched, H, W, batch, K, n_stim, (double)t, 1, o, 1, &why);
            return rc ? (why[0] ? fail(rc, "%s", why) : rc) : 0;
        }
        int copy(const fk::P3& dst, long long off, const fk::P3& src) { return fk::launch_ode_copy(dst, off, src, n, be.st); }
        int init_norms(const fk::P3& y, const fk::P3& f, float rtol, float atol, double* s2) {
            return fk::launch_ode_init_norms(exact, y, f, rtol, atol, n, S, s2, be.st);
        }
        int axpy(const fk::P3& y, float h, const fk::P3& f, const fk::P3& out) { return fk::launch_ode_axpy(exact, y, h, f, out, n, be.st); }
        int diff_norm(const fk::P3& f1, const fk::P3& f0, const fk::P3& y, float rtol, float atol, double* s) {
            return fk::launch_ode_diff_norm(exact, f1, f0, y, rtol, atol, n, S, s, be.st);
        }
        int stage(int i, const fk::P3& y, const fk::P3* k, float dt, const fk::P3& ys) {
            return fk::launch_ode_stage(exact, T, i, y, k, dt, ys, n, be.st);
        }
        int finish(const fk::P3& y, const fk::P3* k, float dt, float rtol, float atol, const fk::P3& yn, const fk::P3* c, double* s) {
            return fk::launch_ode_finish(exact, T, y, k, dt, rtol, atol, yn, c, n, S, s, be.st);
        }
        int interp(const fk::P3* c, float r, const fk::P3& out, long long off) { return fk::launch_ode_interp(exact, c, r, out, off, n, be.st); }
    } be;
    be.be.st = st;
    memset(&be.o, 0, sizeof(be.o));
    be.o.exact = opt.exact; be.o.phys_top = 1; be.o.phys_bottom = 1; be.o.kernel = 1;
    be.K = make_consts(*params, 0.0f, dx);
    if (opt.safe_division) fk::set_safe_division(be.K);
    be.T = fk::make_dopri();
    be.S = scratch;
    be.D = D; be.DX = ws.DX; be.DY = ws.DY; be.stims = ws.stims;
    be.d_batched = d_batched; be.H = H; be.W = W; be.batch = batch; be.n_stim = n_stim; be.exact = opt.exact;
    be.n = B.n; be.why = "";
    fk::OdeStats S = {0, 0, 0};
    rc = fk::drive_dopri5(be, B, n_ts, ts, rtol, atol, mxstep, &S);
    if (rc) return rc;
    if (stats3) { stats3[0] = S.attempts; stats3[1] = S.accepted; stats3[2] = S.rhs_evals; }
    return 0;
}

int fk_check_exact_division(const FkParams* params, float dx, long long* mismatches, void* stream) {
    if (!params || !mismatches) return fail(-1, "NULL pointer%s");
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long* dev = nullptr;
    FK_CUDA(cudaMalloc(&dev, sizeof(unsigned long long)));
    FK_CUDA(cudaMemsetAsync(dev, 0, sizeof(unsigned long long), st));
    fk_divcheck_kernel<<<148 * 8, 256, 0, st>>>(make_consts(*params, 0.01f, dx), dev);
    unsigned long long host = 0;
    cudaError_t e = cudaMemcpyAsync(&host, dev, sizeof(host), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(dev);
    if (e != cudaSuccess) return cuda_fail(e, "fk_check_exact_division");
    *mismatches = (long long)host;
    return 0;
}

int fk_euler_rows(const float* v_in, const float* w_in, const float* u_in, float* v_out, float* w_out, float* u_out,
                  const float* D, const float* DX, const float* DY, int H, int W, const FkParams* params,
                  const FkStimulus* stimuli, int n_stim, double t0, int nsteps, float dt, float dx, const FkOptions* opt,
                  int row0, int row1, void* workspace, size_t workspace_bytes, void* stream) {
    if (nsteps < 1 || nsteps > 4) return fail(-1, "fk_euler_rows advances 1..4 steps per call%s");
    if (row1 <= row0 || row0 < 0 || row1 > H) return fail(-1, "bad row window%s");
    FkOptions o;
    if (opt) o = *opt; else fk_default_options(&o);
    o.steps_per_launch = nsteps;
    return run_euler(v_in, w_in, u_in, v_out, w_out, u_out, D, 0, H, W, 1, params, stimuli, n_stim, t0, nsteps, dt, dx, &o,
                     0, workspace, workspace_bytes, stream, DX, DY, row0, row1);
}

// ---- peer memory for the slab decomposition
namespace {
__global__ void fk_flag_signal_kernel(unsigned int* flag, unsigned int value) {
    __threadfence_system();   // everything this stream wrote before (kernel boundary) is ordered before the flag
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(value) : "memory");
}
// fallback when the driver offers no stream memory operations: one thread polls (and gives up after ~20 s rather than hang)
__global__ void fk_flag_wait_kernel(const unsigned int* flag, unsigned int value) {
    const long long t0 = clock64();
    for (;;) {
        unsigned int v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        if ((int)(v - value) >= 0) return;
        if (clock64() - t0 > 40000000000LL) __trap();
        __nanosleep(256);
    }
}
typedef int (*StreamWaitValue32Fn)(cudaStream_t, unsigned long long, unsigned int, unsigned int);
StreamWaitValue32Fn stream_wait_value32() {
    static StreamWaitValue32Fn fn = nullptr;
    static bool looked = false;
    if (!looked) {
        looked = true;
        if (!getenv("FK_PEER_WAIT_KERNEL")) {
            void* p = nullptr;
            cudaDriverEntryPointQueryResult q;
            if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &p, cudaEnableDefault, &q) == cudaSuccess &&
                q == cudaDriverEntryPointSuccess)
                fn = (StreamWaitValue32Fn)p;
            else
                cudaGetLastError();
        }
    }
    return fn;
}
}  // namespace

int fk_peer_alloc(size_t bytes, void** dev_ptr_out, unsigned char* handle64_out) {
    if (!dev_ptr_out || !handle64_out || bytes == 0) return fail(-1, "fk_peer_alloc: bad arguments%s");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    void* p = nullptr;
    FK_CUDA(cudaMalloc(&p, bytes));
    cudaError_t e = cudaMemset(p, 0, bytes);
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { cudaFree(p); return cuda_fail(e, "fk_peer_alloc"); }
    memcpy(handle64_out, &h, 64);
    *dev_ptr_out = p;
    return 0;
}

int fk_peer_open(const unsigned char* handle64, void** dev_ptr_out) {
    if (!handle64 || !dev_ptr_out) return fail(-1, "fk_peer_open: bad arguments%s");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    FK_CUDA(cudaIpcOpenMemHandle(dev_ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}

int fk_peer_close(void* mapped_dev_ptr) {
    if (mapped_dev_ptr) FK_CUDA(cudaIpcCloseMemHandle(mapped_dev_ptr));
    return 0;
}

int fk_peer_free(void* dev_ptr) {
    if (dev_ptr) FK_CUDA(cudaFree(dev_ptr));
    return 0;
}

int fk_peer_signal(unsigned int* flag_peer_dev, unsigned int value, void* stream) {
    if (!flag_peer_dev) return fail(-1, "fk_peer_signal: NULL flag%s");
    ++g_launches;
    fk_flag_signal_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(flag_peer_dev, value);
    FK_CUDA(cudaGetLastError());
    return 0;
}

int fk_peer_wait(const unsigned int* flag_local_dev, unsigned int value, void* stream) {
    if (!flag_local_dev) return fail(-1, "fk_peer_wait: NULL flag%s");
    if (StreamWaitValue32Fn fn = stream_wait_value32()) {
        const int rc = fn((cudaStream_t)stream, (unsigned long long)(uintptr_t)flag_local_dev, value, 0x0 /* CU_STREAM_WAIT_VALUE_GEQ */);
        if (rc == 0) return 0;
        // (a driver that refuses the operation: fall through to the polling kernel)
    }
    ++g_launches;
    fk_flag_wait_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(flag_local_dev, value);
    FK_CUDA(cudaGetLastError());
    return 0;
}

int fk_peer_copy(void* dst_dev, const void* src_dev, size_t bytes, void* stream) {
    if (!dst_dev || !src_dev) return fail(-1, "fk_peer_copy: NULL pointer%s");
    if (bytes) FK_CUDA(cudaMemcpyAsync(dst_dev, src_dev, bytes, cudaMemcpyDefault, (cudaStream_t)stream));
    return 0;
}

int fk_euler_rows_peer(const float* v_in, const float* w_in, const float* u_in, float* v_out, float* w_out, float* u_out,
                       const float* D, const float* DX, const float* DY, int H, int W, const FkParams* params,
                       const FkStimulus* stimuli, int n_stim, double t0, int nsteps, float dt, float dx, const FkOptions* opt,
                       int row0, int row1, void* workspace, size_t workspace_bytes, void* stream, const FkPeerMirror* mirror,
                       int* fused_out) {
    if (fused_out) *fused_out = 0;
    if (nsteps < 1 || nsteps > 4) return fail(-1, "fk_euler_rows advances 1..4 steps per call%s");
    if (row1 <= row0 || row0 < 0 || row1 > H) return fail(-1, "bad row window%s");
    fk::SlabMirror M;
    memset(&M, 0, sizeof(M));
    bool any = false;
    if (mirror)
        for (int nb = 0; nb < 2; ++nb) {
            if (!mirror->u[nb]) continue;
            if (!mirror->v[nb] || !mirror->w[nb]) return fail(-1, "mirror: NULL array%s");
            if (mirror->row0[nb] < row0 || mirror->row1[nb] > row1 || mirror->row1[nb] < mirror->row0[nb] || mirror->dst_row0[nb] < 0)
                return fail(-1, "mirror rows must lie inside the rows this call writes%s");
            M.u[nb] = mirror->u[nb]; M.v[nb] = mirror->v[nb]; M.w[nb] = mirror->w[nb];
            M.row0[nb] = mirror->row0[nb]; M.row1[nb] = mirror->row1[nb]; M.dst_row0[nb] = mirror->dst_row0[nb];
            any = any || M.row1[nb] > M.row0[nb];
        }
    FkOptions o;
    if (opt) o = *opt; else fk_default_options(&o);
    o.steps_per_launch = nsteps;
    bool mirrored = false;
    int rc = run_euler(v_in, w_in, u_in, v_out, w_out, u_out, D, 0, H, W, 1, params, stimuli, n_stim, t0, nsteps, dt, dx, &o,
                       0, workspace, workspace_bytes, stream, DX, DY, row0, row1, any ? &M : nullptr, &mirrored);
    if (rc) return rc;
    if (any && !mirrored) {   // not a streaming launch: the band rows follow as copies on the same stream
        cudaStream_t st = (cudaStream_t)stream;
        for (int nb = 0; nb < 2; ++nb) {
            if (!M.u[nb] || M.row1[nb] <= M.row0[nb]) continue;
            const size_t bytes = (size_t)(M.row1[nb] - M.row0[nb]) * W * sizeof(float);
            const size_t so = (size_t)M.row0[nb] * W, dof = (size_t)M.dst_row0[nb] * W;
            FK_CUDA(cudaMemcpyAsync(M.v[nb] + dof, v_out + so, bytes, cudaMemcpyDefault, st));
            FK_CUDA(cudaMemcpyAsync(M.w[nb] + dof, w_out + so, bytes, cudaMemcpyDefault, st));
            FK_CUDA(cudaMemcpyAsync(M.u[nb] + dof, u_out + so, bytes, cudaMemcpyDefault, st));
        }
    }
    if (fused_out) *fused_out = mirrored ? 1 : 0;
    return 0;
}

int fk_rhs(const float* v, const float* w, const float* u, float* dv, float* dw, float* du, const float* D, int d_batched,
           int H, int W, int batch, const FkParams* params, const FkStimulus* stimuli, int n_stim, double t, float dx,
           const FkOptions* opt, void* workspace, size_t workspace_bytes, void* stream) {
    return run_euler(v, w, u, dv, dw, du, D, d_batched, H, W, batch, params, stimuli, n_stim, t, 1, 0.0f, dx, opt, 1,
                     workspace, workspace_bytes, stream);
}

}  // extern "C"
