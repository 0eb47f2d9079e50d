// fk_resident.h -- the RESIDENT kernel body: a whole call (hundreds of Euler steps) of a tissue that fits the
// machine's shared memory in ONE launch.
//
// cardiax's own workloads are small tissues stepped many times (128^2 x 1e3 steps, 512^2 x 1e5 steps in 500-step
// segments: deepx data generation, BASELINE configs 1 and 2).  Launch by launch such a tissue is latency bound (one
// launch + two L2 round trips per step, ~8 us).  Here the tissue is cut into one tile per CTA, at most one CTA per SM,
// all co-resident (cooperative launch), and the state never leaves the SM between steps:
//
//   u        two shared-memory buffers (ping-pong) of (th + 8) x (tw + 8): the tile plus a 4-cell halo -- the reach of
//            the reference's two-pass derivative (solve.py:49-52) -- on the sides that have a neighbour
//   v, w     shared memory, touched by the owning thread only
//   D, D_x, D_y  shared memory, loaded once
//
// Per step a CTA (1) computes its RING groups -- the cells within 4 of a tile edge -- and stores their new u to a
// global exchange plane (L2), (2) publishes "step s done" with a release store to its flag, (3) computes its INTERIOR
// groups while that store is in flight, (4) one warp per halo side/row spins on the neighbour's flag (acquire) and
// copies the neighbour's ring cells from L2 into the halo of the next u buffer.  Only the four edge neighbours are
// involved: the stencil is plus-shaped (no cross terms), so corners are never read.  Exchange planes alternate by step
// parity; a CTA cannot overwrite a plane before its neighbours have read it because it needs THEIR next flag first.
//
// A thread computes groups of 4 adjacent cells of one row exactly like the low-latency wide kernel (fk_wide.h): the
// vertical and horizontal two-pass derivatives are rebuilt in registers from the shared-memory window, physical edges
// use the reference's one-sided formulas on an index-clamped window (= the reference's edge padding, solve.py:29-31).
// Same FK_HD source for the CPU emulation (tests/emu), which runs the CTAs phase by phase.
#pragma once
#include "fk_core.h"
#include "fk_stream.h"
#include "fk_tile.h"
#include "fk_wide.h"

namespace fk {

struct ResGeom {
    int ntr, ntc;        // tiles along rows / columns: balanced split, columns in groups of 4 cells
    int th_max, tw_max;  // largest tile
    int pitch;           // floats per row of a u buffer = tw_max + 8
    int nsteps;          // Euler steps of the launch
    float* xb[2];        // exchange planes (batch, H, W): plane (s + 1) & 1 receives the ring cells of step s
    unsigned* flags;     // (batch, ntr * ntc), zero at launch: flag = number of steps whose ring is published
    unsigned spin_limit; // polls of one flag after which the kernel traps instead of hanging the device
};

struct ResCta {
    int ti, tj, tile, sim;
    int r0, r1, c0, c1, th, tw, q;   // q = tw / 4 groups per row
    int nring, ninner;               // groups within 4 cells of a tile edge / the others
    float *U[2], *V, *Wd, *Dm, *DXm, *DYm;
    long long boff, boffD;
    const StimDev* stims;
};

FK_HD long long res_smem_floats(int th_max, int tw_max) {
    return 2LL * (th_max + 8) * (tw_max + 8) + 5LL * th_max * tw_max;
}

FK_HD void res_setup(const TileArgs& A, const ResGeom& G, int tile, int sim, float* smem, ResCta& X) {
    X.tile = tile; X.sim = sim;
    X.ti = tile / G.ntc; X.tj = tile - X.ti * G.ntc;
    const int Q = A.W >> 2;
    X.r0 = (int)((long long)A.H * X.ti / G.ntr);
    X.r1 = (int)((long long)A.H * (X.ti + 1) / G.ntr);
    X.c0 = 4 * (int)((long long)Q * X.tj / G.ntc);
    X.c1 = 4 * (int)((long long)Q * (X.tj + 1) / G.ntc);
    X.th = X.r1 - X.r0; X.tw = X.c1 - X.c0; X.q = X.tw >> 2;
    if (X.th <= 8 || X.q <= 2) { X.nring = X.th * X.q; X.ninner = 0; }
    else { X.nring = 8 * X.q + 2 * (X.th - 8); X.ninner = (X.th - 8) * (X.q - 2); }
    const long long nu = (long long)(G.th_max + 8) * G.pitch, nv = (long long)G.th_max * G.tw_max;
    X.U[0] = smem; X.U[1] = smem + nu;
    X.V = smem + 2 * nu; X.Wd = X.V + nv; X.Dm = X.Wd + nv; X.DXm = X.Dm + nv; X.DYm = X.DXm + nv;
    X.boff = (long long)sim * A.plane;
    X.boffD = (long long)sim * A.plane_D;
    X.stims = A.stims ? A.stims + (long long)sim * A.n_stim : nullptr;
}

// group i of the ring (phase 0) or of the interior (phase 1) -> local row and column of its first cell
FK_HD void res_locate(const ResCta& X, int phase, int i, int& lr, int& lc) {
    const int q = X.q;
    if (phase) { const int r = i / (q - 2); lr = 4 + r; lc = 4 * (1 + i - r * (q - 2)); return; }
    if (X.ninner == 0 || i < 4 * q) { const int r = i / q; lr = r; lc = 4 * (i - r * q); return; }
    if (i < 8 * q) { const int j = i - 4 * q, r = j / q; lr = X.th - 4 + r; lc = 4 * (j - r * q); return; }
    const int j = i - 8 * q;
    lr = 4 + (j >> 1);
    lc = (j & 1) ? 4 * (q - 1) : 0;
}

// coherent 16-byte load of data other SMs wrote during this launch (L2, never a stale L1 line)
FK_HD F4 ldcg4(const float* p) {
#if defined(__CUDA_ARCH__)
    const float4 t = __ldcg(reinterpret_cast<const float4*>(p));
    F4 r; r.x = t.x; r.y = t.y; r.z = t.z; r.w = t.w;
    return r;
#else
    return *reinterpret_cast<const F4*>(p);
#endif
}

// level-0 state -> shared memory: u with its halo straight from the input (complete before the launch), v, w, maps
FK_HD void res_load(const TileArgs& A, const ResGeom& G, const ResCta& X, int tid, int nthr) {
    const int ra = X.r0 - 4 < 0 ? 0 : X.r0 - 4, rb = X.r1 + 4 > A.H ? A.H : X.r1 + 4;
    const int ca = X.c0 - 4 < 0 ? 0 : X.c0 - 4, cb = X.c1 + 4 > A.W ? A.W : X.c1 + 4;
    const int ng = (cb - ca) >> 2, total = (rb - ra) * ng;
    for (int i = tid; i < total; i += nthr) {
        const int r = i / ng, row = ra + r, c = ca + 4 * (i - r * ng);
        float t[4];
        unpack4(ldg4(A.u_in + X.boff + (long long)row * A.W + c), t);
        st4(X.U[0] + (row - X.r0 + 4) * G.pitch + (c - X.c0 + 4), t);
    }
    const int own = X.th * X.q;
    for (int i = tid; i < own; i += nthr) {
        const int lr = i / X.q, lc = 4 * (i - lr * X.q);
        const long long g = (long long)(X.r0 + lr) * A.W + X.c0 + lc;
        const int o = lr * G.tw_max + lc;
        float t[4];
        unpack4(ldg4(A.v_in + X.boff + g), t); st4(X.V + o, t);
        unpack4(ldg4(A.w_in + X.boff + g), t); st4(X.Wd + o, t);
        unpack4(ldg4(A.D + X.boffD + g), t); st4(X.Dm + o, t);
        unpack4(ldg4(A.DX + X.boffD + g), t); st4(X.DXm + o, t);
        unpack4(ldg4(A.DY + X.boffD + g), t); st4(X.DYm + o, t);
    }
}

// stimuli of this CTA's tissue that are active at step s of the launch (solve.py:262-267, fp32 counter)
FK_HD unsigned res_mask(const TileArgs& A, const ResCta& X, int s) {
    unsigned m = 0;
    const float t = (float)(A.t0 + (double)s);
    for (int i = 0; i < A.n_stim; ++i) {
        const StimDev sd = X.stims[i];
        if (sd.field && stim_active(t, sd.start, sd.duration, sd.period)) m |= 1u << i;
    }
    return m;
}

// one Euler step of cells (row, c .. c+3), local (lr, lc): reads `cur` (+ halo), writes `nxt`, v, w in place; the new u
// also goes to the exchange plane (ring groups, xb != null) or, with v and w, to the caller's output (last step)
template <bool EXACT>
FK_HD void res_group(const TileArgs& A, const ResGeom& G, const ResCta& X, const float* cur, float* nxt, int lr, int lc,
                     unsigned mask, bool last, float* xb) {
    const int H = A.H, W = A.W, P = G.pitch, row = X.r0 + lr, c = X.c0 + lc;
    const float* uc0 = cur + (lr + 4) * P + (lc + 4);
    const bool row_in = row >= 4 && row + 5 <= H;   // every vertical formula central, no clamped row
    const bool col_in = c >= 4 && c + 8 <= W;       // same for the columns of all four cells
    float u_x[4], u_y[4], u_xx[4], u_yy[4], uc[4];
    unpack4(ld4(uc0), uc);
    const int o = lr * G.tw_max + lc;
    float v[4], w[4], Dv[4], DXv[4], DYv[4], stim[4] = {0.f, 0.f, 0.f, 0.f};
    unpack4(ld4(X.V + o), v);
    unpack4(ld4(X.Wd + o), w);
    unpack4(ld4(X.Dm + o), Dv);
    unpack4(ld4(X.DXm + o), DXv);
    unpack4(ld4(X.DYm + o), DYv);
    const long long g = (long long)row * W + c;
    if (mask) {  // solve.py:260-269: later stimuli override earlier ones, zero cells never stimulate
        for (int i = 0; i < A.n_stim; ++i)
            if (mask >> i & 1u) {
                float f[4];
                unpack4(ldg4(X.stims[i].field + g), f);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (f[k] != 0.0f) stim[k] = f[k];
            }
    }
    // ---- vertical (axis 0)
    if (row_in) {
        float ur[9][4];
#pragma unroll
        for (int j = 0; j < 9; ++j) {
            if (j == 4) {
#pragma unroll
                for (int k = 0; k < 4; ++k) ur[4][k] = uc[k];
            } else {
                unpack4(ld4(uc0 + (j - 4) * P), ur[j]);
            }
        }
        float gxv[5][4];
#pragma unroll
        for (int j = 0; j < 5; ++j)
#pragma unroll
            for (int k = 0; k < 4; ++k) gxv[j][k] = dcen<EXACT>(A.K, ur[j][k], ur[j + 1][k], ur[j + 3][k], ur[j + 4][k]);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            u_x[k] = gxv[2][k];
            u_xx[k] = dcen<EXACT>(A.K, gxv[0][k], gxv[1][k], gxv[3][k], gxv[4][k]);
        }
    } else {
        // padded rows P-6 .. P+6 = tissue rows row-6 .. row+6 clamped (solve.py:31); entries the applicable formulas
        // never use may fall outside the tile's buffer rows and are clamped into it
        float ur[13][4];
        const int nrows = G.th_max + 8;
#pragma unroll
        for (int j = 0; j < 13; ++j) {
            const int sr = clampi(clampi(row + j - 6, 0, H - 1) - X.r0 + 4, 0, nrows - 1);
            unpack4(ld4(cur + sr * P + (lc + 4)), ur[j]);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float col[13];
#pragma unroll
            for (int j = 0; j < 13; ++j) col[j] = ur[j][k];
            wide_axis_general<EXACT>(A.K, col, row + 1, H, u_x[k], u_xx[k]);
        }
    }
    // ---- horizontal (axis 1)
    if (col_in) {
        float e[12];   // columns c-4 .. c+7
        unpack4(ld4(uc0 - 4), e);
        unpack4(ld4(uc0 + 4), e + 8);
#pragma unroll
        for (int k = 0; k < 4; ++k) e[4 + k] = uc[k];
        float gyv[8];
#pragma unroll
        for (int m = 0; m < 8; ++m) gyv[m] = dcen<EXACT>(A.K, e[m], e[m + 1], e[m + 3], e[m + 4]);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            u_y[k] = gyv[k + 2];
            u_yy[k] = dcen<EXACT>(A.K, gyv[k], gyv[k + 1], gyv[k + 3], gyv[k + 4]);
        }
    } else {
        float e[16];   // padded columns Q0-6 .. Q0+9, Q0 = c + 1: tissue columns c-6 .. c+9 clamped
        const float* urow = cur + (lr + 4) * P;
#pragma unroll
        for (int m = 0; m < 16; ++m)
            e[m] = (m >= 6 && m < 10) ? uc[m - 6] : urow[clampi(clampi(c + m - 6, 0, W - 1) - X.c0 + 4, 0, P - 1)];
        float gyv[10];   // u_y at padded columns Q0-3 .. Q0+6, shared by the four cells
#pragma unroll
        for (int m = 0; m < 10; ++m) gyv[m] = wide_deriv<EXACT>(A.K, kind_of(c + m - 2, W, 1, 1), e + m);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            u_y[k] = gyv[k + 3];
            u_yy[k] = wide_deriv<EXACT>(A.K, kind_of(c + k + 1, W, 1, 1), gyv + k);
        }
    }
    float un[4], vn[4], wn[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float del_u = diffusion<EXACT>(Dv[k], DXv[k], DYv[k], u_x[k], u_y[k], u_xx[k], u_yy[k]);
        float d_v, d_w, d_u;
        cell_rhs<EXACT>(A.K, uc[k], v[k], w[k], del_u, stim[k], d_v, d_w, d_u);
        vn[k] = euler<EXACT>(v[k], d_v, A.K.dt);
        wn[k] = euler<EXACT>(w[k], d_w, A.K.dt);
        un[k] = euler<EXACT>(uc[k], d_u, A.K.dt);
    }
    if (last) {
        st4(A.u_out + X.boff + g, un);
        st4(A.v_out + X.boff + g, vn);
        st4(A.w_out + X.boff + g, wn);
    } else {
        st4(nxt + (lr + 4) * P + (lc + 4), un);
        st4(X.V + o, vn);
        st4(X.Wd + o, wn);
        if (xb) st4(xb + g, un);
    }
}

// the ring (phase 0) or interior (phase 1) groups of step s, strided over the CTA's threads
template <bool EXACT>
FK_HD void res_phase(const TileArgs& A, const ResGeom& G, const ResCta& X, int s, int phase, unsigned mask, int tid,
                     int nthr) {
    const float* cur = X.U[s & 1];
    float* nxt = X.U[(s + 1) & 1];
    const bool last = s == G.nsteps - 1;
    float* xb = (last || phase) ? nullptr : G.xb[(s + 1) & 1] + X.boff;
    const int n = phase ? X.ninner : X.nring;
    for (int i = tid; i < n; i += nthr) {
        int lr, lc;
        res_locate(X, phase, i, lr, lc);
        res_group<EXACT>(A, G, X, cur, nxt, lr, lc, mask, last, xb);
    }
}

// Halo jobs of a step, one warp each: 0..3 = rows of the north halo, 4..7 = rows of the south halo, 8 = west, 9 = east.
enum { FK_RES_JOBS = 10 };

// the tile a job reads from, or -1 at a physical edge
FK_HD int res_job_neighbour(const ResGeom& G, const ResCta& X, int j) {
    if (j < 4) return X.ti > 0 ? X.tile - G.ntc : -1;
    if (j < 8) return X.ti < G.ntr - 1 ? X.tile + G.ntc : -1;
    if (j == 8) return X.tj > 0 ? X.tile - 1 : -1;
    return X.tj < G.ntc - 1 ? X.tile + 1 : -1;
}

// copy the neighbour's ring cells of step s from the exchange plane into the halo of the next u buffer
FK_HD void res_job_load(const TileArgs& A, const ResGeom& G, const ResCta& X, int s, int j, int lane, int nlanes) {
    const float* src = G.xb[(s + 1) & 1] + X.boff;
    float* nxt = X.U[(s + 1) & 1];
    float t[4];
    if (j < 8) {
        const int row = j < 4 ? X.r0 - 4 + j : X.r1 + (j - 4);
        float* dst = nxt + (row - X.r0 + 4) * G.pitch + 4;
        for (int k = lane; k < X.q; k += nlanes) {
            unpack4(ldcg4(src + (long long)row * A.W + X.c0 + 4 * k), t);
            st4(dst + 4 * k, t);
        }
    } else {
        const int c = j == 8 ? X.c0 - 4 : X.c1;
        for (int r = lane; r < X.th; r += nlanes) {
            unpack4(ldcg4(src + (long long)(X.r0 + r) * A.W + c), t);
            st4(nxt + (r + 4) * G.pitch + (c - X.c0 + 4), t);
        }
    }
}

// ------------------------------------------------------------------ launch geometry (host)
struct ResPlan {
    ResGeom G;
    int threads;
    long long smem_bytes;
};

// Tiles for a (batch, H, W) problem on `capacity` co-resident CTAs with `smem_limit` bytes each.  A tile is at least
// 8 x 8 (a neighbour's halo comes from ONE tile, and the edge formulas reach 7 cells into the tissue).  Cost: the serial
// rounds of groups a CTA runs per step (ring and interior separately -- the flag is published in between), then the
// size of the largest tile; ties go to fewer, wider tiles (row halos are contiguous in the exchange plane).
inline bool plan_resident(int H, int W, int batch, int capacity, long long smem_limit, int force_ntr, int force_ntc,
                          int force_threads, ResPlan& P) {
    if (W % 4 != 0 || H < 3 || W < 4 || batch < 1) return false;
    const int Q = W >> 2;
    double best = 1e300;
    bool found = false;
    for (int ntr = 1; ntr <= H; ++ntr) {
        if (ntr > 1 && H / ntr < 8) break;
        if (force_ntr > 0 && ntr != force_ntr) continue;
        for (int ntc = 1; ntc <= Q; ++ntc) {
            if (ntc > 1 && Q / ntc < 2) break;
            if ((long long)ntr * ntc * batch > capacity) break;
            if (force_ntc > 0 && ntc != force_ntc) continue;
            const int th = (H + ntr - 1) / ntr, q = (Q + ntc - 1) / ntc, tw = 4 * q;
            const long long smem = res_smem_floats(th, tw) * 4;
            if (smem > smem_limit) continue;
            int nring, ninner;
            if (th <= 8 || q <= 2) { nring = th * q; ninner = 0; }
            else { nring = 8 * q + 2 * (th - 8); ninner = (th - 8) * (q - 2); }
            int threads = force_threads;
            if (threads <= 0) {
                const int m = nring > ninner ? nring : ninner;
                threads = (m + 31) / 32 * 32;
                if (threads < 128) threads = 128;
                if (threads > 512) threads = 512;
            }
            const int rounds = (nring + threads - 1) / threads + (ninner + threads - 1) / threads;
            // one round ~ 1 unit; exchange latency ~ 2 units whatever the geometry; a CTA's warps share 4 schedulers
            const double cost = rounds * (threads > 128 ? threads / 128.0 : 1.0) + 1e-4 * th * tw + 1e-3 * ntr * ntc +
                                1e-3 * th;
            if (cost < best) {
                best = cost;
                found = true;
                P.G.ntr = ntr; P.G.ntc = ntc; P.G.th_max = th; P.G.tw_max = tw; P.G.pitch = tw + 8;
                P.threads = threads;
                P.smem_bytes = smem;
            }
        }
    }
    return found;
}

// ------------------------------------------------------------------ CPU emulation of one launch (tests/emu)
#if !defined(__CUDACC__)
}  // namespace fk
#include <vector>
namespace fk {
// CTAs advance in lock step, phase by phase, which is one legal interleaving of the flag protocol; shared memory is
// poisoned with NaN so that a read of a halo nobody filled shows up in the result.
inline int emu_resident_launch(const ResPlan& P, const TileArgs& A, int batch, int exact) {
    const ResGeom& G = P.G;
    const int ntiles = G.ntr * G.ntc;
    const long long floats = res_smem_floats(G.th_max, G.tw_max);
    std::vector<std::vector<float>> smem((size_t)ntiles * batch, std::vector<float>((size_t)floats, __builtin_nanf("")));
    std::vector<ResCta> X((size_t)ntiles * batch);
    std::vector<unsigned> flags((size_t)ntiles * batch, 0u);
    for (int sim = 0; sim < batch; ++sim)
        for (int t = 0; t < ntiles; ++t) {
            ResCta& x = X[(size_t)sim * ntiles + t];
            res_setup(A, G, t, sim, smem[(size_t)sim * ntiles + t].data(), x);
            res_load(A, G, x, 0, 1);
        }
    for (int s = 0; s < G.nsteps; ++s) {
        const bool last = s == G.nsteps - 1;
        for (size_t i = 0; i < X.size(); ++i) {
            const unsigned mask = res_mask(A, X[i], s);
            for (int phase = 0; phase < 2; ++phase) {
                if (exact) res_phase<true>(A, G, X[i], s, phase, mask, 0, 1);
                else res_phase<false>(A, G, X[i], s, phase, mask, 0, 1);
            }
            flags[i] = (unsigned)(s + 1);
        }
        if (last) break;
        for (size_t i = 0; i < X.size(); ++i)
            for (int j = 0; j < FK_RES_JOBS; ++j) {
                const int nb = res_job_neighbour(G, X[i], j);
                if (nb < 0) continue;
                if (flags[(size_t)X[i].sim * ntiles + nb] < (unsigned)(s + 1)) return -7;
                res_job_load(A, G, X[i], s, j, 0, 1);
            }
    }
    return 0;
}
#endif

}  // namespace fk
