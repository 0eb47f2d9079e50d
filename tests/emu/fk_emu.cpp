// TEST INFRASTRUCTURE ONLY -- CPU emulation of the kernel bodies in cardiax_b200/csrc.
// The same FK_HD phase functions the CUDA kernels call are run here by ONE emulated thread per
// tile (phases in order, barriers implicit) or, for the streaming kernel, by all threads of a
// block one after the other inside each row iteration; the launch sequence is the product's own
// fk::drive_euler.  g++ -ffp-contract=off keeps EXACT mode bit-identical to nvcc's
// __fmul_rn/__fadd_rn/__fdiv_rn build.  Never loaded by the product.
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../cardiax_b200/csrc/fk_core.h"
#include "../../cardiax_b200/csrc/fk_tile.h"
#include "../../cardiax_b200/csrc/fk_stream.h"
#include "../../cardiax_b200/csrc/fk_driver.h"
#include "../../cardiax_b200/csrc/fk_ode.h"

namespace {

struct EmuBackend {
    int reverse = 0;       // run the threads of a streaming block in reverse order (hazard check)
    int launches_tile = 0, launches_stream = 0;
    int num_sms() { return 148; }
    int occupancy(int, int, int, int NT, long long) { return NT <= 128 ? 2 : 1; }
    int tiles(fk::TileArgs& A, int exact, int batch, bool) {
        long long floats = 0;
        const int total = fk::finish_regions(A, &floats);
        if (total == 0) return 0;
        ++launches_tile;
        std::vector<float> smem((size_t)floats);
        for (int sim = 0; sim < batch; ++sim)
            for (int tile = 0; tile < total; ++tile) {
                // poison shared memory so a read of something never written shows up as NaN
                for (auto& x : smem) x = __builtin_nanf("");
                fk::TileCtx X;
                fk::tile_setup(A, tile, sim, smem.data(), X);
                fk::tile_load(A, X, 0, 0, 1, 1);
                float *Uc = X.U0, *Un = X.U1;
                for (int s = 1; s <= A.T; ++s) {
                    if (exact) {
                        fk::tile_grad<true>(A, X, s, Uc, 0, 0, 1, 1);
                        fk::tile_update<true>(A, X, s, Uc, Un, 0, 0, 1, 1);
                    } else {
                        fk::tile_grad<false>(A, X, s, Uc, 0, 0, 1, 1);
                        fk::tile_update<false>(A, X, s, Uc, Un, 0, 0, 1, 1);
                    }
                    float* t = Uc; Uc = Un; Un = t;
                }
            }
        return 0;
    }
    int launches_wide = 0;
    int wide(const fk::TileArgs& A, int exact, int batch) {
        ++launches_wide;
        for (int sim = 0; sim < batch; ++sim) {
            const unsigned mask = fk::wide_mask(A, sim);
            for (int row = 0; row < A.H; ++row)
                for (int c = 0; c < A.W; c += 4) {
                    if (exact) fk::wide_thread<true>(A, sim, row, c, mask);
                    else fk::wide_thread<false>(A, sim, row, c, mask);
                }
        }
        return 0;
    }
    int launches_res = 0;
    long long resident_smem_limit() { return 227 * 1024 - 256; }
    bool cluster_ok(const fk::ResPlan&, int) { return true; }
    int resident(const fk::ResPlan& P, const fk::TileArgs& A, int exact, int batch) {
        ++launches_res;
        last_res = P;
        return fk::emu_resident_launch(P, A, batch, exact);
    }
    fk::ResPlan last_res;
    int stream(const fk::StreamPlan& P, const fk::TileArgs& A, int exact, int batch) {
        ++launches_stream;
        return fk::emu_stream_launch(P, A, batch, exact, reverse);
    }
};

}  // namespace

extern "C" {

struct EmuStim {
    const float* field;
    double start, duration, period;
    int int_mask, reserved;
};

// options: {exact, steps_per_launch, kernel, phys_top, phys_bottom, cta_threads, rows_per_cta, uniform_diffusivity, reverse,
//           row0, row1, tiles_r, tiles_c, cells_per_thread, edge_rows, edge_colgroups, maps_global, counter_is_int}
// info (optional, 2 ints): tile launches, stream launches + 1000 * wide launches + 1000000 * resident launches
// slab halo mirror of the NEXT fk_emu_euler call (fk::SlabMirror; consumed by that call): the emulated streaming kernel
// stores rows [row0[n], row1[n]) of its result into these host arrays as well, like the device kernel does into peer memory
static fk::SlabMirror g_mirror;
static bool g_mirror_set = false, g_mirror_done = false;
int fk_emu_set_mirror(float* const* vwu_up, float* const* vwu_down, const int* row0, const int* row1, const int* dst_row0) {
    memset(&g_mirror, 0, sizeof(g_mirror));
    float* const* nb[2] = {vwu_up, vwu_down};
    for (int n = 0; n < 2; ++n) {
        if (nb[n]) { g_mirror.v[n] = nb[n][0]; g_mirror.w[n] = nb[n][1]; g_mirror.u[n] = nb[n][2]; }
        g_mirror.row0[n] = row0[n]; g_mirror.row1[n] = row1[n]; g_mirror.dst_row0[n] = dst_row0[n];
    }
    g_mirror_set = true;
    return 0;
}
int fk_emu_mirror_was_fused(void) { return g_mirror_done ? 1 : 0; }

int fk_emu_euler(const float* v_in, const float* w_in, const float* u_in, float* v_out, float* w_out, float* u_out,
                 const float* D, int d_batched, int H, int W, int batch, const float* params14, const EmuStim* stims,
                 int n_stim, double t0, double t1, float dt, float dx, const int* options, int rhs_mode, int* info) {
    static_assert(sizeof(EmuStim) == sizeof(fk::StimDev), "layout");
    const size_t plane = (size_t)H * W;
    const size_t dplanes = d_batched ? batch : 1;
    std::vector<float> DX(plane * dplanes), DY(plane * dplanes), pv(plane * batch), pw(plane * batch), pu(plane * batch);
    for (size_t b = 0; b < dplanes; ++b)
        for (int r = 0; r < H; ++r)
            for (int c = 0; c < W; ++c)
                fk::dgrad_cell(D + b * plane, H, W, dx, options[3], options[4], r, c, DX[b * plane + (size_t)r * W + c],
                               DY[b * plane + (size_t)r * W + c]);
    fk::DriveBuffers B = fk::DriveBuffers();
    B.v_in = v_in; B.w_in = w_in; B.u_in = u_in; B.v_out = v_out; B.w_out = w_out; B.u_out = u_out;
    B.pv = pv.data(); B.pw = pw.data(); B.pu = pu.data(); B.D = D; B.DX = DX.data(); B.DY = DY.data();
    B.stims = (const fk::StimDev*)stims;
    g_mirror_done = false;
    if (g_mirror_set) { B.mirror = &g_mirror; B.mirrored = &g_mirror_done; g_mirror_set = false; }
    fk::u64 dummy[1] = {0};
    B.xchg = dummy;          // the emulation allocates its own mailboxes
    B.xchg_bytes = fk::res_xchg_bytes(H, W, batch);
    fk::DriveOptions o;
    o.exact = options[0]; o.steps_per_launch = options[1]; o.kernel = options[2]; o.phys_top = options[3];
    o.phys_bottom = options[4]; o.cta_threads = options[5]; o.rows_per_cta = options[6]; o.uniform_diffusivity = options[7];
    o.row0 = options[9]; o.row1 = options[10]; o.tiles_r = options[11]; o.tiles_c = options[12]; o.cells_per_thread = options[13]; o.edge_rows = options[14]; o.edge_colgroups = options[15]; o.maps_global = options[16];
    o.t_is_int = options[17];
    EmuBackend be;
    be.reverse = options[8];
    const long long nsteps = rhs_mode ? 1 : fk::count_steps(t0, t1);
    if (nsteps <= 0 && !rhs_mode) {
        memcpy(v_out, v_in, plane * batch * 4); memcpy(w_out, w_in, plane * batch * 4); memcpy(u_out, u_in, plane * batch * 4);
        return 0;
    }
    const char* why = "";
    const int rc = fk::drive_euler(be, B, d_batched, H, W, batch, fk::make_consts(params14, dt, dx), n_stim, t0, nsteps, o,
                                   rhs_mode, &why);
    if (info) { info[0] = be.launches_tile; info[1] = be.launches_stream + 1000 * be.launches_wide + 1000000 * be.launches_res; }
    return rc;
}

int fk_emu_stim_active(float t, float start, float duration, float period) {
    return fk::stim_active(t, start, duration, period) ? 1 : 0;
}
int fk_emu_stim_active_typed(double t, int t_is_int, double start, double duration, double period, int kinds) {
    return fk::stim_active_typed(t, t_is_int, start, duration, period, kinds) ? 1 : 0;
}
// fk_core.h: stims_quiet for one stimulus (tests compare it with the schedule evaluated step by step)
int fk_emu_stims_quiet(double t0, long long nsteps, double start, double duration, double period, int kinds) {
    static float dummy = 1.0f;
    fk::StimDev sd;
    sd.field = &dummy; sd.start = start; sd.duration = duration; sd.period = period; sd.kinds = kinds; sd.reserved = 0;
    return fk::stims_quiet(&sd, 1, t0, nsteps) ? 1 : 0;
}

}  // extern "C"

// planner probe (tests): occupancy modelled as min(65536 / (regs * NT), 227 KB / smem)
extern "C" int fk_emu_plan(int H, int W, int batch, int T, int cta_threads, int rows_per_cta, int regs, int* out) {
    fk::StreamPlan P;
    const bool ok = fk::plan_stream(4 * T, H - 4 * T, W, batch, T, cta_threads, rows_per_cta, 148, 0, 256,
                                    [&](int NT, long long smem) {
                                        const int a = 65536 / (((regs + 7) / 8 * 8) * NT);
                                        const int b = (int)((228 * 1024) / (smem + 1024));
                                        return a < b ? a : b;
                                    }, P);
    if (!ok) return 0;
    out[0] = P.G.NT; out[1] = P.G.nstrips; out[2] = P.G.cstride; out[3] = P.G.RH; out[4] = P.G.nchunks;
    out[5] = (int)P.smem_bytes;
    return 1;
}

// the whole-tissue plan with balanced row chunks (the chunks at the physical top / bottom edge top_off / bot_off rows shorter):
// out = {nchunks, first row of every chunk ..., H}
extern "C" int fk_emu_plan_rows(int H, int W, int batch, int T, int regs, int top_off, int bot_off, int* out, int cap) {
    fk::StreamPlan P;
    const bool ok = fk::plan_stream(0, H, W, batch, T, 0, 0, 148, 0, 256,
                                    [&](int NT, long long smem) {
                                        const int a = 65536 / (((regs + 7) / 8 * 8) * NT);
                                        const int b = (int)((228 * 1024) / (smem + 1024));
                                        return a < b ? a : b;
                                    }, P, top_off, bot_off);
    if (!ok || P.G.nchunks + 2 > cap) return 0;
    out[0] = P.G.nchunks;
    for (int c = 0; c <= P.G.nchunks; ++c) out[1 + c] = fk::stream_chunk_row(P.G, c);
    return 1;
}

// resident planner probe (tests): {ntr, ntc, th_max, tw_max, threads, smem bytes, cells per thread, mailbox bytes}
extern "C" int fk_emu_plan_resident(int H, int W, int batch, int* out, const int* force6) {   // force6[6] = maps_global
    fk::ResPlan P;
    if (!fk::plan_resident(H, W, batch, 148, 227 * 1024 - 256, fk::res_xchg_bytes(H, W, batch), force6[0], force6[1], force6[2],
                           force6[3], force6[4], force6[5], force6[6], P)) return 0;
    out[0] = P.G.ntr; out[1] = P.G.ntc; out[2] = P.G.th_max; out[3] = P.G.tw_max; out[4] = P.threads; out[5] = (int)P.smem_bytes;
    out[6] = P.G.nc; out[7] = (int)P.xchg_bytes; out[8] = P.G.eh; out[9] = P.G.ewq; out[10] = P.G.single; out[11] = P.G.mg; out[12] = P.G.tp;
    return 1;
}

extern "C" int fk_emu_dgrad(const float* D, float* DX, float* DY, int H, int W, float dx, int phys_top, int phys_bot) {
    for (int r = 0; r < H; ++r)
        for (int c = 0; c < W; ++c)
            fk::dgrad_cell(D, H, W, dx, phys_top, phys_bot, r, c, DX[(long long)r * W + c], DY[(long long)r * W + c]);
    return 0;
}

// ---- Dormand-Prince: the product's fk::drive_dopri5 on a CPU backend (element bodies of fk_aux.h, sums in fp64)
namespace {
struct EmuOde {
    EmuBackend be;
    fk::DriveOptions o;
    fk::Consts K;
    fk::Dopri T;
    const float *D, *DX, *DY;
    const fk::StimDev* stims;
    int d_batched, H, W, batch, n_stim, exact;
    long long n;
    int rhs(const fk::P3& y, const fk::P3& k, float t) {
        fk::DriveBuffers B = fk::DriveBuffers();
        memset(&B, 0, sizeof(B));
        B.v_in = y.a[0]; B.w_in = y.a[1]; B.u_in = y.a[2]; B.v_out = k.a[0]; B.w_out = k.a[1]; B.u_out = k.a[2];
        B.D = D; B.DX = DX; B.DY = DY; B.stims = stims;
        const char* why = "";
        return fk::drive_euler(be, B, d_batched, H, W, batch, K, n_stim, (double)t, 1, o, 1, &why);
    }
    int copy(const fk::P3& dst, long long off, const fk::P3& src) {
        for (int a = 0; a < 3; ++a) memcpy(dst.a[a] + off, src.a[a], sizeof(float) * (size_t)n);
        return 0;
    }
    template <bool E> void init_norms_t(const fk::P3& y, const fk::P3& f, float rtol, float atol, double* s2) {
        s2[0] = s2[1] = 0.0;
        for (int a = 0; a < 3; ++a)
            for (long long e = 0; e < n; ++e) {
                float qy, qf;
                fk::ode_scaled<E>(y.a[a][e], f.a[a][e], rtol, atol, qy, qf);
                s2[0] += (double)(qy * qy); s2[1] += (double)(qf * qf);
            }
    }
    int init_norms(const fk::P3& y, const fk::P3& f, float rtol, float atol, double* s2) {
        if (exact) init_norms_t<true>(y, f, rtol, atol, s2); else init_norms_t<false>(y, f, rtol, atol, s2);
        return 0;
    }
    int axpy(const fk::P3& y, float h, const fk::P3& f, const fk::P3& out) {
        for (int a = 0; a < 3; ++a)
            for (long long e = 0; e < n; ++e) out.a[a][e] = y.a[a][e] + h * f.a[a][e];
        return 0;
    }
    int diff_norm(const fk::P3& f1, const fk::P3& f0, const fk::P3& y, float rtol, float atol, double* s) {
        *s = 0.0;
        for (int a = 0; a < 3; ++a)
            for (long long e = 0; e < n; ++e) {
                const float scale = atol + fabsf(y.a[a][e]) * rtol;
                const float q = (f1.a[a][e] - f0.a[a][e]) / scale;
                *s += (double)(q * q);
            }
        return 0;
    }
    int stage(int i, const fk::P3& y, const fk::P3* k, float dt, const fk::P3& ys) {
        for (int a = 0; a < 3; ++a)
            for (long long e = 0; e < n; ++e) {
                float kk[6];
                for (int s = 0; s < 6; ++s) kk[s] = s < i ? k[s].a[a][e] : 0.0f;
                ys.a[a][e] = exact ? fk::ode_stage<true>(y.a[a][e], T.beta[i - 1], kk, i, dt)
                                   : fk::ode_stage<false>(y.a[a][e], T.beta[i - 1], kk, i, dt);
            }
        return 0;
    }
    int finish(const fk::P3& y, const fk::P3* k, float dt, float rtol, float atol, const fk::P3& yn, const fk::P3* c, double* s) {
        *s = 0.0;
        for (int a = 0; a < 3; ++a)
            for (long long e = 0; e < n; ++e) {
                float kk[7], y1, r2, coef[5];
                for (int j = 0; j < 7; ++j) kk[j] = k[j].a[a][e];
                if (exact) fk::ode_finish<true>(T, y.a[a][e], kk, dt, rtol, atol, y1, r2, coef);
                else fk::ode_finish<false>(T, y.a[a][e], kk, dt, rtol, atol, y1, r2, coef);
                yn.a[a][e] = y1;
                for (int j = 0; j < 5; ++j) c[j].a[a][e] = coef[j];
                *s += (double)r2;
            }
        return 0;
    }
    int interp(const fk::P3* c, float r, const fk::P3& out, long long off) {
        for (int a = 0; a < 3; ++a)
            for (long long e = 0; e < n; ++e) {
                float coef[5];
                for (int j = 0; j < 5; ++j) coef[j] = c[j].a[a][e];
                out.a[a][off + e] = exact ? fk::ode_interp<true>(coef, r) : fk::ode_interp<false>(coef, r);
            }
        return 0;
    }
};
}  // namespace

// stats3: {attempts, accepted, rhs evaluations}
extern "C" int fk_emu_dopri5(const float* v0, const float* w0, const float* u0, float* v_out, float* w_out, float* u_out,
                             const float* D, int H, int W, const float* params14, const EmuStim* stims, int n_stim,
                             const float* ts, int n_ts, float dx, float rtol, float atol, double mxstep, int exact,
                             long long* stats3) {
    const size_t n = (size_t)H * W;
    std::vector<float> DX(n), DY(n), store(60 * n);
    for (int r = 0; r < H; ++r)
        for (int c = 0; c < W; ++c) fk::dgrad_cell(D, H, W, dx, 1, 1, r, c, DX[(size_t)r * W + c], DY[(size_t)r * W + c]);
    float* p = store.data();
    auto next3 = [&]() { fk::P3 r; for (int a = 0; a < 3; ++a) { r.a[a] = p; p += n; } return r; };
    fk::OdeBuffers B;
    B.n = (long long)n;
    B.y = next3(); B.ys = next3(); B.yn = next3();
    for (int s = 0; s < 7; ++s) B.k[s] = next3();
    for (int c = 0; c < 2; ++c) for (int j = 0; j < 5; ++j) B.c[c][j] = next3();
    B.out.a[0] = v_out; B.out.a[1] = w_out; B.out.a[2] = u_out;
    memcpy(B.y.a[0], v0, n * 4); memcpy(B.y.a[1], w0, n * 4); memcpy(B.y.a[2], u0, n * 4);
    EmuOde be;
    memset(&be.o, 0, sizeof(be.o));
    be.o.exact = exact; be.o.phys_top = 1; be.o.phys_bottom = 1; be.o.kernel = 1;
    be.K = fk::make_consts(params14, 0.0f, dx);
    be.T = fk::make_dopri();
    be.D = D; be.DX = DX.data(); be.DY = DY.data(); be.stims = (const fk::StimDev*)stims;
    be.d_batched = 0; be.H = H; be.W = W; be.batch = 1; be.n_stim = n_stim; be.exact = exact; be.n = (long long)n;
    fk::OdeStats S = {0, 0, 0};
    const int rc = fk::drive_dopri5(be, B, n_ts, ts, rtol, atol, mxstep, &S);
    if (stats3) { stats3[0] = S.attempts; stats3[1] = S.accepted; stats3[2] = S.rhs_evals; }
    return rc;
}

extern "C" int fk_emu_resize(const float* in, int planes, int H, int W, float* out, int Ho, int Wo) {
    const fk::ResizeAxis ah = fk::make_resize_axis(H, Ho), aw = fk::make_resize_axis(W, Wo);
    for (int p = 0; p < planes; ++p)
        for (int i = 0; i < Ho; ++i)
            for (int j = 0; j < Wo; ++j)
                out[((size_t)p * Ho + i) * Wo + j] = fk::resize_pixel(in + (size_t)p * H * W, W, ah.lo[i], ah.wt.data() + (size_t)i * ah.K, 1,
                                                                      ah.K, aw.lo[j], aw.wt.data() + (size_t)j * aw.K, 1, aw.K);
    return 0;
}

extern "C" int fk_emu_electrogram(const float* x, int frames, int H, int W, float p0, float p1, float* out) {
    for (int f = 0; f < frames; ++f) {
        double acc = 0.0;
        for (int i = 0; i < H; ++i)
            for (int j = 0; j < W; ++j) acc += (double)(x[((size_t)f * H + i) * W + j] * fk::egm_weight(i, j, p0, p1));
        out[f] = (float)acc;
    }
    return 0;
}

// ---- fused Heun: the product's fk::drive_heun (one tile-kernel launch per step) on the CPU backend
extern "C" int fk_emu_heun(const float* v_in, const float* w_in, const float* u_in, float* v_out, float* w_out, float* u_out,
                           const float* D, int H, int W, int batch, const float* params14, const EmuStim* stims, int n_stim,
                           double t0, double t1, float dt, float dx, int exact, int* launches) {
    const size_t plane = (size_t)H * W;
    std::vector<float> DX(plane), DY(plane), pv(plane * batch), pw(plane * batch), pu(plane * batch);
    for (int r = 0; r < H; ++r)
        for (int c = 0; c < W; ++c) fk::dgrad_cell(D, H, W, dx, 1, 1, r, c, DX[(size_t)r * W + c], DY[(size_t)r * W + c]);
    fk::DriveBuffers B = fk::DriveBuffers();
    memset(&B, 0, sizeof(B));
    B.v_in = v_in; B.w_in = w_in; B.u_in = u_in; B.v_out = v_out; B.w_out = w_out; B.u_out = u_out;
    B.pv = pv.data(); B.pw = pw.data(); B.pu = pu.data(); B.D = D; B.DX = DX.data(); B.DY = DY.data();
    B.stims = (const fk::StimDev*)stims;
    EmuBackend be;
    const long long nsteps = fk::count_steps(t0, t1);
    if (nsteps <= 0) {
        memcpy(v_out, v_in, plane * batch * 4); memcpy(w_out, w_in, plane * batch * 4); memcpy(u_out, u_in, plane * batch * 4);
        return 0;
    }
    const int rc = fk::drive_heun(be, B, 0, H, W, batch, fk::make_consts(params14, dt, dx), n_stim, t0, nsteps, exact,
                                  (float)((double)dt * 0.5));
    if (launches) *launches = be.launches_tile;
    return rc;
}

// ---- fast Heun: the product's fk::drive_heun_fast (Euler kernels + folded closing pass) on the CPU backend
namespace {
struct EmuHeunBackend : EmuBackend {
    int combines = 0;
    int combine(const float* yv, const float* yw, const float* yu, const float* ev, const float* ew, const float* eu,
                float* ov, float* ow, float* ou, long long n) {
        ++combines;
        for (long long i = 0; i < n; ++i) {
            ov[i] = fmaf(0.5f, ev[i] - yv[i], yv[i]); ow[i] = fmaf(0.5f, ew[i] - yw[i], yw[i]); ou[i] = fmaf(0.5f, eu[i] - yu[i], yu[i]);
        }
        return 0;
    }
    int copy(float* dst, const float* src, long long n) { memcpy(dst, src, sizeof(float) * (size_t)n); return 0; }
};
}  // namespace

// info (4 ints): tile launches, stream launches, wide launches, combine passes
extern "C" int fk_emu_heun_fast(const float* v_in, const float* w_in, const float* u_in, float* v_out, float* w_out, float* u_out,
                                const float* D, int H, int W, int batch, const float* params14, const EmuStim* stims, int n_stim,
                                double t0, double t1, float dt, float dx, int fold, int force_stream, int* info) {
    const size_t plane = (size_t)H * W, all = plane * batch;
    std::vector<float> DX(plane), DY(plane), store(12 * all);
    for (int r = 0; r < H; ++r)
        for (int c = 0; c < W; ++c) fk::dgrad_cell(D, H, W, dx, 1, 1, r, c, DX[(size_t)r * W + c], DY[(size_t)r * W + c]);
    fk::HeunFastBuffers HB;
    memset(&HB, 0, sizeof(HB));
    HB.v_in = v_in; HB.w_in = w_in; HB.u_in = u_in; HB.v_out = v_out; HB.w_out = w_out; HB.u_out = u_out;
    float* p = store.data();
    HB.pv = p; HB.pw = p + all; HB.pu = p + 2 * all; p += 3 * all;
    for (int a = 0; a < 3; ++a) { HB.s1[a] = p; p += all; }
    for (int a = 0; a < 3; ++a) { HB.s2[a] = p; p += all; }
    for (int a = 0; a < 3; ++a) { HB.s3[a] = p; p += all; }
    HB.D = D; HB.DX = DX.data(); HB.DY = DY.data(); HB.stims = (const fk::StimDev*)stims;
    EmuHeunBackend be;
    const long long nsteps = fk::count_steps(t0, t1);
    const char* why = "";
    // force_stream: pretend the tissue is large (cta_threads / rows_per_cta chosen by the test make it streamable)
    const int rc = fk::drive_heun_fast(be, HB, 0, H, W, batch, fk::make_consts(params14, dt, dx), (const fk::StimDev*)stims, n_stim,
                                       t0, nsteps, 0, force_stream ? 32 : 0, force_stream ? 24 : 0, fold != 0, false, &why);
    if (info) { info[0] = be.launches_tile; info[1] = be.launches_stream; info[2] = be.launches_wide; info[3] = be.combines; }
    return rc;
}
