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
    float start, duration, period;
};

// options: {exact, steps_per_launch, kernel, phys_top, phys_bottom, cta_threads, rows_per_cta, uniform_diffusivity, reverse,
//           row0, row1, tiles_r, tiles_c, cells_per_thread, edge_rows, edge_colgroups, maps_global}
// info (optional, 2 ints): tile launches, stream launches + 1000 * wide launches + 1000000 * resident launches
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
    fk::DriveBuffers B;
    B.v_in = v_in; B.w_in = w_in; B.u_in = u_in; B.v_out = v_out; B.w_out = w_out; B.u_out = u_out;
    B.pv = pv.data(); B.pw = pw.data(); B.pu = pu.data(); B.D = D; B.DX = DX.data(); B.DY = DY.data();
    B.stims = (const fk::StimDev*)stims;
    fk::u64 dummy[1] = {0};
    B.xchg = dummy;          // the emulation allocates its own mailboxes
    B.xchg_bytes = fk::res_xchg_bytes(H, W, batch);
    fk::DriveOptions o;
    o.exact = options[0]; o.steps_per_launch = options[1]; o.kernel = options[2]; o.phys_top = options[3];
    o.phys_bottom = options[4]; o.cta_threads = options[5]; o.rows_per_cta = options[6]; o.uniform_diffusivity = options[7];
    o.row0 = options[9]; o.row1 = options[10]; o.tiles_r = options[11]; o.tiles_c = options[12]; o.cells_per_thread = options[13]; o.edge_rows = options[14]; o.edge_colgroups = options[15]; o.maps_global = options[16];
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

// resident planner probe (tests): {ntr, ntc, th_max, tw_max, threads, smem bytes, cells per thread, mailbox bytes}
extern "C" int fk_emu_plan_resident(int H, int W, int batch, int* out, const int* force6) {   // force6[6] = maps_global
    fk::ResPlan P;
    if (!fk::plan_resident(H, W, batch, 148, 227 * 1024 - 256, fk::res_xchg_bytes(H, W, batch), force6[0], force6[1], force6[2],
                           force6[3], force6[4], force6[5], force6[6], P)) return 0;
    out[0] = P.G.ntr; out[1] = P.G.ntc; out[2] = P.G.th_max; out[3] = P.G.tw_max; out[4] = P.threads; out[5] = (int)P.smem_bytes;
    out[6] = P.G.nc; out[7] = (int)P.xchg_bytes; out[8] = P.G.eh; out[9] = P.G.ewq; out[10] = P.G.single; out[11] = P.G.mg;
    return 1;
}

extern "C" int fk_emu_dgrad(const float* D, float* DX, float* DY, int H, int W, float dx, int phys_top, int phys_bot) {
    for (int r = 0; r < H; ++r)
        for (int c = 0; c < W; ++c)
            fk::dgrad_cell(D, H, W, dx, phys_top, phys_bot, r, c, DX[(long long)r * W + c], DY[(long long)r * W + c]);
    return 0;
}
