// fk_driver.h -- the launch sequence of solve._forward_euler / solve.step, independent of where
// the kernels run.  fk_api.cu instantiates it with the CUDA backend; tests/emu with the CPU
// emulation, so the ping-pong / frame / tail logic is unit-tested without a GPU.
#pragma once
#include <cstdlib>
#include "fk_stream.h"
#include "fk_tile.h"
#include "fk_wide.h"
#include "fk_resident.h"

namespace fk {

struct DriveOptions {
    int exact;
    int steps_per_launch;   // 0 = default
    int kernel;             // 0 auto, 1 tiles only, 2 streaming required, 3 low-latency wide kernel required,
                            // 4 resident kernel required
    int phys_top, phys_bottom;
    int cta_threads, rows_per_cta;
    int uniform_diffusivity;
    int row0, row1;         // output rows of a single-launch call (slab building block); row1 <= 0: all the rows owned
    int tiles_r, tiles_c;   // resident kernel: tile grid (0 = planner's choice)
    int cells_per_thread;   // resident kernel: 1, 2 or 4 adjacent cells per thread (0 = planner's choice)
    int edge_rows, edge_colgroups;   // resident kernel: size of the tiles at the tissue's edges (0 auto, < 0 even split)
    int maps_global;        // resident kernel: 1 = diffusivity maps read from global memory (L2), 0 = planner's choice
    int t_is_int;           // the loop counter is an int32 (stimulus schedule typing, fk_core.h: stim_active_typed)
};

// row-slab decomposition: rows [row0[n], row1[n]) of a call's result are also wanted in the memory of the neighbouring
// GPU n (0 = the slab above, 1 = the slab below) -- peer-mapped (Hb', W) arrays -- starting at its row dst_row0[n]
struct SlabMirror {
    float *v[2], *w[2], *u[2];   // null: no neighbour on that side
    int row0[2], row1[2], dst_row0[2];
};

struct DriveBuffers {
    const float *v_in, *w_in, *u_in;
    float *v_out, *w_out, *u_out;
    float *pv, *pw, *pu;            // ping-pong scratch, (batch, H, W) each
    const float *D, *DX, *DY;
    const StimDev* stims;           // device/emulated copy of the stimulus table, or null
    u64* xchg;                      // resident kernel: mailboxes (zeroed by the backend), xchg_bytes long, or null
    long long xchg_bytes;
    const float *hy_v, *hy_w, *hy_u;   // fast Heun: y of the step whose E(E(y)) this call computes, or null; applied by
    bool* hy_folded;                   // the call's last launch if it is a streaming / wide one (*hy_folded says so)
    const SlabMirror* mirror;          // slab halo mirror of the call's last launch, or null; done by the launch itself when
    bool* mirrored;                    // it is a streaming one (*mirrored says so), else left to the caller (plain copies)
};

enum { FK_DEFAULT_T = 2, FK_RES_MAX_CTAS = 1024 };
// tissues up to this many cells go to the cluster transport of the resident kernel by default (fk_resident.h)
#ifndef FK_CLUSTER_MAX_CELLS
#define FK_CLUSTER_MAX_CELLS (80 * 80)   // measured (profiles/probe_cluster_r02.md): ahead of the mailbox form only up to ~64^2
#endif
// mailbox bytes fk_workspace_bytes provisions for the resident kernel: 32 per cell of problems it may take, else none
inline long long res_xchg_bytes(int H, int W, int batch) {
    const long long cells = (long long)H * W * batch;
    return cells <= (1LL << 21) ? 32 * cells + 4096 : 0;
}
// Tissues (x batch) up to this many cells that fit the SMs' shared memory run whole calls in ONE resident launch.
#ifndef FK_RES_MAX_CELLS
#define FK_RES_MAX_CELLS (1LL << 21)
#endif

// finalises the tile counts of A.reg[0..nreg)
inline int finish_regions(TileArgs& A, long long* smem_floats) {
    int total = 0;
    long long maxfloats = 0;
    for (int i = 0; i < A.nreg; ++i) {
        TileRegion& R = A.reg[i];
        // The tissue's edge row/column reaches FIVE cells inwards (forward/backward formulas applied twice,
        // solve.py:232-235 on top of itself), one more than a level's 4-cell apron: a tile must never be a
        // single row or column, so a remainder of 1 is folded into the last tile (tile_setup).
        R.ntr = tile_count(R.R1 - R.R0, R.th);
        R.ntc = tile_count(R.C1 - R.C0, R.tw);
        R.first = total;
        total += R.ntr * R.ntc;
        const long long f = tile_smem_floats(R.th + 1, R.tw + 1, A.T, A.heun);
        if (f > maxfloats) maxfloats = f;
    }
    if (smem_floats) *smem_floats = maxfloats;
    return total;
}

inline void add_region(TileArgs& A, int R0, int R1, int C0, int C1, int th, int tw) {
    if (R1 <= R0 || C1 <= C0) return;
    TileRegion& R = A.reg[A.nreg++];
    R.R0 = R0; R.R1 = R1; R.C0 = C0; R.C1 = C1;
    R.th = th < R1 - R0 ? th : R1 - R0;
    R.tw = tw < C1 - C0 ? tw : C1 - C0;
}

inline void pick_tile(int rows, int W, int T, int batch, int sms, int& th, int& tw) {
    // whole-tissue coverage by general tiles: 32 x 64 output cells (+ 4T apron) keeps two CTAs per SM at T <= 2
    th = 32; tw = 64;
    while (tile_smem_floats(th + 1, tw + 1, T) * 4 > 110 * 1024 && th > 8) th -= 8;
    while (tile_smem_floats(th + 1, tw + 1, T) * 4 > 220 * 1024 && tw > 16) tw -= 16;
    if (th > rows) th = rows;
    if (tw > W) tw = W;
    // a tissue that 32 x 64 tiles cannot spread over the machine (256^2: 32 tiles for 148 SMs) gets smaller ones: a tile
    // launch there is latency bound (solve.step, the Dormand-Prince stages, exact Heun), not apron bound
    auto count = [&]() { return (long long)tile_count(rows, th) * tile_count(W, tw) * batch; };
    while (count() < sms && (th > 8 || tw > 32)) {
        if (th >= tw / 2 && th > 8) th = (th + 1) / 2 < 8 ? 8 : (th + 1) / 2;
        else if (tw > 32) tw = tw / 2 < 32 ? 32 : tw / 2;
        else th = (th + 1) / 2 < 8 ? 8 : (th + 1) / 2;
    }
}

// Backend: int tiles(TileArgs&, int exact, int batch, bool side);
//          int stream(const StreamPlan&, const TileArgs&, int exact, int batch);
//          int wide(const TileArgs&, int exact, int batch);
//          int resident(const ResPlan&, const TileArgs&, int exact, int batch);  long long resident_smem_limit();
//          bool cluster_ok(const ResPlan&, int exact);   // can this cluster shape be launched?
//          int num_sms(); int occupancy(int T, int exact, int uniform, int NT, long long smem_bytes);
// Returns 0 or the backend's error code; *why gets a static message on argument errors.
template <class Backend>
int drive_euler(Backend& be, const DriveBuffers& B, int d_batched, int H, int W, int batch, const Consts& K, int n_stim,
                double t0, long long nsteps, const DriveOptions& opt, int rhs_mode, const char** why) {
    *why = "";
    // exact numerics are latency bound (unfused dependent chains, 215 instructions per cell-step): one step per launch
    // runs at 128 registers with twice the resident warps (68 vs 57 Gcell-steps/s on 4096^2, profiles/probe_exact_r02.md)
    int Tmax = rhs_mode ? 1 : (opt.steps_per_launch ? opt.steps_per_launch : (opt.exact ? 1 : FK_DEFAULT_T));
    const bool slab = !opt.phys_top || !opt.phys_bottom;
    if (slab) {
        // slab decomposition: halo rows are inputs only and must be re-exchanged after every launch
        if (nsteps > Tmax) { *why = "a slab call advances at most steps_per_launch steps (halos must be exchanged)"; return -1; }
        Tmax = (int)nsteps;
        if (H < 8 * Tmax + 1) { *why = "slab too thin for its halo"; return -1; }
    }
    // output rows [R0, R1) of a launch of T steps.  The streaming kernel takes all of them, physical top / bottom edge
    // included (its first / last row chunk uses the reference's one-sided formulas there) -- unless a row window starts
    // or ends INSIDE the 4T rows next to a physical edge, where neither a halo nor the edge itself is available to it.
    auto rows = [&](int T, int& R0, int& R1, bool& streamable) {
        if (opt.row1 > 0) { R0 = opt.row0; R1 = opt.row1; }
        else { R0 = opt.phys_top ? 0 : 4 * T; R1 = opt.phys_bottom ? H : H - 4 * T; }
        streamable = !(opt.phys_top && R0 > 0 && R0 < 4 * T) && !(opt.phys_bottom && R1 < H && R1 > H - 4 * T);
    };
    if (opt.row1 > 0) {
        if (nsteps > Tmax) { *why = "a row-window call is a single launch: nsteps <= steps_per_launch"; return -1; }
        Tmax = (int)nsteps;
        const int lo = opt.row0 - 4 * Tmax, hi = opt.row1 + 4 * Tmax;
        if (opt.row0 < 0 || opt.row1 > H || opt.row0 >= opt.row1 || (lo < 0 && !(opt.phys_top && opt.row0 == 0)) ||
            (hi > H && !(opt.phys_bottom && opt.row1 == H))) {
            *why = "row window needs 4*nsteps rows of input on each side (or a physical edge)";
            return -1;
        }
    }
    // tissues that fit the machine's shared memory: the whole call in one resident launch (fk_resident.h)
    ResPlan rplan;
    bool use_res = false;
    // ... a small tissue as ONE thread-block cluster (halos through distributed shared memory); kernel = 5 forces it
    if (!rhs_mode && !slab && opt.row1 <= 0 && nsteps < (1LL << 30) &&
        (opt.kernel == 5 || (opt.kernel == 0 && opt.steps_per_launch == 0 && nsteps >= 4 && batch == 1 &&
                             (long long)H * W <= FK_CLUSTER_MAX_CELLS))) {
        use_res = plan_cluster(H, W, be.resident_smem_limit(), opt.tiles_r, opt.tiles_c, opt.cta_threads, opt.cells_per_thread, rplan) &&
                  be.cluster_ok(rplan, opt.exact);
    }
    if (opt.kernel == 5 && !use_res) { *why = "cluster kernel not applicable (needs W % 4 == 0, a tissue of at most 16 tiles that fit shared memory)"; return -5; }
    if (!use_res && !rhs_mode && !slab && opt.row1 <= 0 && B.xchg && nsteps < (1LL << 30) &&
        (opt.kernel == 4 || (opt.kernel == 0 && opt.steps_per_launch == 0 && nsteps >= 4 &&
                             (long long)H * W * batch <= FK_RES_MAX_CELLS))) {
        const int cap = be.num_sms() < FK_RES_MAX_CTAS ? be.num_sms() : FK_RES_MAX_CTAS;
        use_res = plan_resident(H, W, batch, cap, be.resident_smem_limit(), B.xchg_bytes, opt.tiles_r, opt.tiles_c,
                                opt.cta_threads, opt.cells_per_thread, opt.edge_rows, opt.edge_colgroups, opt.maps_global, rplan);
    }
    if (opt.kernel == 4 && !use_res) { *why = "resident kernel not applicable (needs W % 4 == 0, a whole tissue that fits shared memory)"; return -5; }
    // tissues too small to fill the machine: one launch of the barrier-free wide kernel per step
    const bool wide_ok = !rhs_mode && !slab && opt.row1 <= 0 && W % 4 == 0 && H >= 3 && W >= 4;
    const bool use_wide = !use_res && wide_ok && (opt.kernel == 3 || (opt.kernel == 0 && opt.steps_per_launch == 0 &&
                                                          (long long)H * W * batch < (1LL << 20)));
    if (opt.kernel == 3 && !use_wide) { *why = "wide kernel not applicable (needs W % 4 == 0, whole tissue)"; return -5; }
    StreamPlan plan;
    bool use_stream = false;
    if (!rhs_mode && opt.kernel != 1 && !use_wide && !use_res) {
        auto try_plan = [&](int T, StreamPlan& P) {
            int R0, R1;
            bool ok;
            rows(T, R0, R1, ok);
            return ok && R1 > R0 && plan_stream(R0, R1, W, batch, T, opt.cta_threads, opt.rows_per_cta, be.num_sms(),
                                          opt.uniform_diffusivity, stream_max_threads(T),
                                          [&](int NT, long long smem) { return be.occupancy(T, opt.exact, opt.uniform_diffusivity, NT, smem); }, P,
                                          (opt.phys_top && R0 == 0) ? stream_top_off(T) : 0,
                                          (opt.phys_bottom && R1 == H) ? stream_bot_off(T) : 0);
        };
        use_stream = try_plan(Tmax, plan);
        if (opt.steps_per_launch == 0 && !slab && opt.row1 <= 0 && nsteps > 1 && (long long)H * W * batch < (1LL << 21)) {
            // no depth requested: a tissue too small to fill the machine is latency-bound, and there one step per
            // launch (8 rows of pipeline fill instead of 16) beats two.  ~12 units of launch overhead per launch.
            StreamPlan p1;
            if (try_plan(1, p1) && (!use_stream || p1.cost + 12.0 < (plan.cost + 12.0) / Tmax)) {
                plan = p1;
                Tmax = 1;
                use_stream = true;
            }
        }
        if (!use_stream && opt.kernel == 2) { *why = "streaming kernel not applicable to this shape"; return -5; }
    }

    TileArgs A;
    A = TileArgs();
    A.D = B.D; A.DX = B.DX; A.DY = B.DY;
    A.plane = (long long)H * W;
    A.plane_D = d_batched ? A.plane : 0;
    A.H = H; A.W = W;
    A.phys_top = opt.phys_top; A.phys_bot = opt.phys_bottom; A.phys_left = 1; A.phys_right = 1;
    A.rhs_mode = rhs_mode;
    A.K = K;
    A.stims = n_stim ? B.stims : nullptr;
    A.n_stim = n_stim;
    A.t_is_int = opt.t_is_int;

    if (use_res) {
        A.u_in = B.u_in; A.v_in = B.v_in; A.w_in = B.w_in;
        A.u_out = B.u_out; A.v_out = B.v_out; A.w_out = B.w_out;
        A.T = 1; A.t0 = t0;
        rplan.G.nsteps = (int)nsteps;
        rplan.G.xchg = B.xchg; rplan.G.timing = nullptr;
        rplan.G.spin_limit = 1u << 24;
        return be.resident(rplan, A, opt.exact, batch);
    }
    if (use_wide) Tmax = 1;
    const long long nl = rhs_mode ? 1 : (nsteps + Tmax - 1) / Tmax;
    const float *sv = B.v_in, *sw = B.w_in, *su = B.u_in;
    long long remaining = rhs_mode ? 1 : nsteps;
    double t = t0;
    for (long long l = 0; l < nl; ++l) {
        const int T = (int)(remaining < Tmax ? remaining : Tmax);
        // the last launch lands in the caller's output; the ones before alternate with the scratch
        const bool to_out = ((nl - 1 - l) % 2 == 0);
        A.u_in = su; A.v_in = sv; A.w_in = sw;
        A.u_out = to_out ? B.u_out : B.pu;
        A.v_out = to_out ? B.v_out : B.pv;
        A.w_out = to_out ? B.w_out : B.pw;
        A.T = T; A.t0 = t;
        A.hy_u = A.hy_v = A.hy_w = nullptr;
        if (B.hy_u && l == nl - 1 && !rhs_mode && (use_wide || (use_stream && T == plan.T))) {
            A.hy_u = B.hy_u; A.hy_v = B.hy_v; A.hy_w = B.hy_w;
            if (B.hy_folded) *B.hy_folded = true;
        }
        for (int nb = 0; nb < 2; ++nb) { A.mir_u[nb] = A.mir_v[nb] = A.mir_w[nb] = nullptr; A.mir_r0[nb] = A.mir_r1[nb] = 0; A.mir_off[nb] = 0; }
        if (B.mirror && l == nl - 1 && !rhs_mode && batch == 1 && use_stream && T == plan.T && T <= 2) {
            for (int nb = 0; nb < 2; ++nb)
                if (B.mirror->u[nb] && B.mirror->row1[nb] > B.mirror->row0[nb]) {
                    A.mir_u[nb] = B.mirror->u[nb]; A.mir_v[nb] = B.mirror->v[nb]; A.mir_w[nb] = B.mirror->w[nb];
                    A.mir_r0[nb] = B.mirror->row0[nb]; A.mir_r1[nb] = B.mirror->row1[nb];
                    A.mir_off[nb] = (long long)(B.mirror->dst_row0[nb] - B.mirror->row0[nb]) * W;
                }
            if (B.mirrored) *B.mirrored = true;
        }
        int rc;
        if (use_wide) {
            rc = be.wide(A, opt.exact, batch);
            if (rc) return rc;
        } else if (use_stream && T == plan.T) {
            rc = be.stream(plan, A, opt.exact, batch);
            if (rc) return rc;
        } else {
            int R0, R1;
            bool ok;
            rows(T, R0, R1, ok);
            int th, tw;
            pick_tile(R1 - R0, W, T, batch, be.num_sms(), th, tw);
            A.nreg = 0;
            add_region(A, R0, R1, 0, W, th, tw);
            rc = be.tiles(A, opt.exact, batch, false);
            if (rc) return rc;
        }
        su = A.u_out; sv = A.v_out; sw = A.w_out;
        t += T;
        remaining -= T;
    }
    return 0;
}

// solve._forward_heun (solve.py:73-85, 103-111): ONE launch of the general tile kernel per Heun step -- its two levels
// are the predictor and the corrector, both at the same counter; y and k1 of a tile's own cells wait in shared memory.
template <class Backend>
int drive_heun(Backend& be, const DriveBuffers& B, int d_batched, int H, int W, int batch, const Consts& K, int n_stim,
               double t0, long long nsteps, int exact, float h_half, int t_is_int = 0) {
    TileArgs A;
    A = TileArgs();
    A.D = B.D; A.DX = B.DX; A.DY = B.DY;
    A.plane = (long long)H * W;
    A.plane_D = d_batched ? A.plane : 0;
    A.H = H; A.W = W;
    A.phys_top = A.phys_bot = A.phys_left = A.phys_right = 1;
    A.K = K;
    A.stims = n_stim ? B.stims : nullptr;
    A.n_stim = n_stim;
    A.t_is_int = t_is_int;
    A.T = 2; A.heun = 1; A.h_half = h_half;
    int th = 16, tw = 64;   // 88 KB: two CTAs per SM; measured best of {32x64, 16x64, 24x96, 16x128} on 256^2 ... 4096^2
    const float *sv = B.v_in, *sw = B.w_in, *su = B.u_in;
    for (long long l = 0; l < nsteps; ++l) {
        const bool to_out = ((nsteps - 1 - l) % 2 == 0);
        A.u_in = su; A.v_in = sv; A.w_in = sw;
        A.u_out = to_out ? B.u_out : B.pu;
        A.v_out = to_out ? B.v_out : B.pv;
        A.w_out = to_out ? B.w_out : B.pw;
        A.t0 = t0 + (double)l;
        A.nreg = 0;
        add_region(A, 0, H, 0, W, th, tw);
        const int rc = be.tiles(A, exact, batch, false);
        if (rc) return rc;
        su = A.u_out; sv = A.v_out; sw = A.w_out;
    }
    return 0;
}

// solve._forward_heun in FAST numerics: a Heun step through the fast Euler kernels.  With E = one Euler step at counter
// t, y + (k1 + k2) dt / 2 = y + (E(E(y)) - y) / 2.  E(E(y)) is one temporally blocked two-step call whenever no stimulus is
// active at t or t + 1 (E at t + 1 then equals E at t), two single-step calls at counter t otherwise; the closing pass is
// folded into the store of the launch that produces E(E(y)) when that is a streaming / wide launch (TileArgs::hy_*),
// else a combine pass follows.  Backend: drive_euler's + combine(y.., e.., out.., n) + copy(dst, src, n).
struct HeunFastBuffers {
    const float *v_in, *w_in, *u_in;
    float *v_out, *w_out, *u_out;
    float *pv, *pw, *pu;            // ping-pong of the step loop
    float *s1[3], *s2[3], *s3[3];   // three scratch States
    const float *D, *DX, *DY;
    const StimDev* stims;           // device table
    u64* xchg;
    long long xchg_bytes;
};

template <class Backend>
int drive_heun_fast(Backend& be, const HeunFastBuffers& HB, int d_batched, int H, int W, int batch, const Consts& K,
                    const StimDev* host_stims, int n_stim, double t0, long long nsteps, int uniform, int cta_threads,
                    int rows_per_cta, bool fold, bool try_resident, const char** why, int t_is_int = 0) {
    *why = "";
    DriveOptions oe = DriveOptions();
    oe.phys_top = 1; oe.phys_bottom = 1; oe.t_is_int = t_is_int;
    oe.uniform_diffusivity = uniform; oe.cta_threads = cta_threads; oe.rows_per_cta = rows_per_cta;
    DriveOptions o1 = oe, o2 = oe;
    const bool small = cta_threads == 0 && (long long)H * W * batch < (1LL << 20) && W % 4 == 0 && H >= 3;
    if (small) o1.kernel = 3;            // one wide launch per stage
    else o1.steps_per_launch = 1;        // streaming kernel, T = 1
    if (cta_threads) { o1.kernel = 2; o2.kernel = 2; o2.steps_per_launch = 2; }   // a caller who sizes the CTAs wants streaming
    const long long n = (long long)H * W * batch;
    auto quiet = [&](double t) {
        for (int i = 0; i < batch * n_stim; ++i)
            if (host_stims[i].field && (stim_on(host_stims[i], t, t_is_int) || stim_on(host_stims[i], t + 1.0, t_is_int)))
                return false;
        return true;
    };
    const float *yv = HB.v_in, *yw = HB.w_in, *yu = HB.u_in;
    int rc;
    for (long long l = 0; l < nsteps; ++l) {
        const double t = t0 + (double)l;
        const bool to_out = ((nsteps - 1 - l) % 2 == 0);
        float *nv = to_out ? HB.v_out : HB.pv, *nw = to_out ? HB.w_out : HB.pw, *nu = to_out ? HB.u_out : HB.pu;
        DriveBuffers B = DriveBuffers();
        B.D = HB.D; B.DX = HB.DX; B.DY = HB.DY; B.stims = HB.stims;
        B.pv = HB.s3[0]; B.pw = HB.s3[1]; B.pu = HB.s3[2];   // ping-pong scratch of a two-launch call
        B.xchg = HB.xchg; B.xchg_bytes = HB.xchg_bytes;
        bool folded = false;
        float* e2[3] = {HB.s2[0], HB.s2[1], HB.s2[2]};
        if (fold) { e2[0] = nv; e2[1] = nw; e2[2] = nu; }
        auto arm_fold = [&]() {   // for the call that produces E(E(y)) only
            if (fold) { B.hy_v = yv; B.hy_w = yw; B.hy_u = yu; B.hy_folded = &folded; }
        };
        if (quiet(t)) {
            arm_fold();
            B.v_in = yv; B.w_in = yw; B.u_in = yu; B.v_out = e2[0]; B.w_out = e2[1]; B.u_out = e2[2];
            rc = -5;
            if (try_resident) {
                o2.kernel = 4;
                rc = drive_euler(be, B, d_batched, H, W, batch, K, n_stim, t, 2, o2, 0, why);
                if (rc == -5 || rc == -3) { try_resident = false; o2.kernel = 0; folded = false; }
            }
            if (rc == -5 || rc == -3) rc = drive_euler(be, B, d_batched, H, W, batch, K, n_stim, t, 2, o2, 0, why);
            if (rc) return rc;
        } else {
            B.v_in = yv; B.w_in = yw; B.u_in = yu; B.v_out = HB.s1[0]; B.w_out = HB.s1[1]; B.u_out = HB.s1[2];
            rc = drive_euler(be, B, d_batched, H, W, batch, K, n_stim, t, 1, o1, 0, why);
            if (rc) return rc;
            arm_fold();
            B.v_in = HB.s1[0]; B.w_in = HB.s1[1]; B.u_in = HB.s1[2]; B.v_out = e2[0]; B.w_out = e2[1]; B.u_out = e2[2];
            rc = drive_euler(be, B, d_batched, H, W, batch, K, n_stim, t, 1, o1, 0, why);
            if (rc) return rc;
        }
        if (!folded) {
            if (fold) {   // a tile-kernel / resident launch: E(E(y)) sits where the new state belongs
                if ((rc = be.copy(HB.s2[0], nv, n)) || (rc = be.copy(HB.s2[1], nw, n)) || (rc = be.copy(HB.s2[2], nu, n))) return rc;
            }
            if ((rc = be.combine(yv, yw, yu, HB.s2[0], HB.s2[1], HB.s2[2], nv, nw, nu, n))) return rc;
        }
        yv = nv; yw = nw; yu = nu;
    }
    return 0;
}

// lax.fori_loop(t0, t1, ...) with a float counter: iterations while t0 + k < t1
inline long long count_steps(double t0, double t1) {
    long long n = 0;
    if (t1 > t0) {
        n = (long long)ceil(t1 - t0);
        while (n > 0 && t0 + (double)(n - 1) >= t1) --n;
        while (t0 + (double)n < t1) ++n;
    }
    return n;
}

}  // namespace fk
