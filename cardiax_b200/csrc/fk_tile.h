// fk_tile.h -- the GENERAL multi-step tile: exact boundary semantics of the reference
// (edge-pad by one, two passes of solve.gradient on the PADDED array, crop -- solve.py:29-32,
// 49-55, 61-65) for any rectangle of the tissue, T Euler steps per launch with all
// intermediate levels in shared memory.
//
// It is used (a) for the frame of 4T cells around the tissue that the streaming kernel
// (fk_stream.h) leaves out, (b) for whole tissues that are too small or oddly shaped for the
// streaming kernel, and (c) for solve.step (rhs_mode: T = 1, returns d_v, d_w, d_u).
//
// The body is written as PHASES separated by block barriers; each phase is a loop over items
// strided by the thread grid (tx, ty, ntx, nty) and only talks to other items through shared
// memory across a barrier.  The CUDA kernel calls the phases with __syncthreads() between them;
// tests/emu calls the same phases with one emulated thread.
#pragma once
#include "fk_core.h"

namespace fk {

struct TileRegion {  // a rectangle of output cells cut into th x tw tiles
    int R0, R1, C0, C1;
    int th, tw;
    int ntr, ntc;    // tile counts
    int first;       // index of this region's first tile in the launch
};

struct TileArgs {
    const float *u_in, *v_in, *w_in;   // (batch, H, W)
    float *u_out, *v_out, *w_out;
    const float *D, *DX, *DY;          // diffusivity and its two derivative maps (D_x, D_y of solve.py:53-54)
    long long plane;                   // H*W   (stride between simulations)
    long long plane_D;                 // H*W, or 0 when all simulations share one diffusivity map
    int H, W;
    int phys_top, phys_bot, phys_left, phys_right;  // which buffer edges are physical tissue edges
    int T;                             // Euler steps in this launch (levels)
    int rhs_mode;                      // 1: write (d_v, d_w, d_u) of a single step instead of the new state
    int heun;                          // 1: the launch's two levels are the two stages of ONE Heun step (solve.py:73-85):
                                       //    level 1 = predictor y1 = y + k1 dt, level 2 evaluates k2 = f(y1) at the SAME
                                       //    counter and writes y + (k1 + k2) * h_half
    float h_half;                      // dt * 0.5 (solve.py:83)
    const float *hy_v, *hy_w, *hy_u;   // fast Heun (streaming / wide kernels, last level only): when set, the launch stores
                                       // y + (E - y) / 2 instead of its Euler result E, y = these arrays (fk_forward_heun)
    // row-slab decomposition (streaming kernel, last level only): rows [mir_r0[n], mir_r1[n]) of the result are ALSO
    // stored into the neighbour's halo, a peer-mapped array of GPU n (0 = the slab above, 1 = the slab below), at element
    // index (this launch's index) + mir_off[n] -- the halo exchange travels over NVLink as the rows are produced
    float *mir_u[2], *mir_v[2], *mir_w[2];
    long long mir_off[2];
    int mir_r0[2], mir_r1[2];
    Consts K;
    const StimDev* stims;              // (batch, n_stim)
    int n_stim;
    double t0;                         // counter value of the first level
    int t_is_int;                      // the counter is an int32 (deepx.generate.sequence) rather than a float32 (solve.forward)
    TileRegion reg[4];
    int nreg;
};

struct TileCtx {  // per-tile geometry, identical for every thread of the block
    int r0, r1, c0, c1;      // output rectangle
    int ra, rb, ca, cb;      // level-0 (input) rectangle
    int nr, nc, SG;
    float *U0, *U1, *V, *Wd, *GX, *GY;  // shared memory
    float* HS;               // Heun: y.v, y.w, k1.v, k1.w, k1.u of the output cells (5 planes of how cells)
    int how;                 // cells of the output rectangle
    unsigned mask[8];        // active stimuli per level (bit i = stimulus i)
    long long boff, boffD;   // batch offsets
    const StimDev* stims;
};

// tiles along an axis of `len` cells: ceil(len / t), except that a last tile of ONE cell is folded into its neighbour
FK_HD int tile_count(int len, int t) {
    const int n = (len + t - 1) / t;
    return (n > 1 && len % t == 1) ? n - 1 : n;
}

FK_HD long long tile_smem_floats(int th, int tw, int T, int heun = 0) {
    long long nr = th + 8LL * T, nc = tw + 8LL * T;
    return 4 * nr * nc + (nr + 3) * nc + nr * (nc + 3) + (heun ? 5LL * th * tw : 0);
}

FK_HD void level_rect(const TileArgs& A, const TileCtx& X, int s, int& a, int& b, int& c, int& d) {
    const int e = 4 * (A.T - s);
    a = X.r0 - e < 0 ? 0 : X.r0 - e;
    b = X.r1 + e > A.H ? A.H : X.r1 + e;
    c = X.c0 - e < 0 ? 0 : X.c0 - e;
    d = X.c1 + e > A.W ? A.W : X.c1 + e;
}

// locate tile `tile` of the launch and fill the context (smem carve-up included)
FK_HD void tile_setup(const TileArgs& A, int tile, int sim, float* smem, TileCtx& X) {
    int g = 0;
    for (int i = 1; i < A.nreg; ++i)
        if (tile >= A.reg[i].first) g = i;
    const TileRegion& R = A.reg[g];
    const int lt = tile - R.first;
    const int tr = lt / R.ntc, tc = lt - tr * R.ntc;
    X.r0 = R.R0 + tr * R.th;
    X.r1 = tr == R.ntr - 1 ? R.R1 : X.r0 + R.th;   // the last tile takes a remainder of one cell as well
    X.c0 = R.C0 + tc * R.tw;
    X.c1 = tc == R.ntc - 1 ? R.C1 : X.c0 + R.tw;
    level_rect(A, X, 0, X.ra, X.rb, X.ca, X.cb);
    X.nr = X.rb - X.ra;
    X.nc = X.cb - X.ca;
    X.SG = X.nc + 3;
    const long long n = (long long)X.nr * X.nc;
    X.U0 = smem;
    X.U1 = X.U0 + n;
    X.V = X.U1 + n;
    X.Wd = X.V + n;
    X.GX = X.Wd + n;
    X.GY = X.GX + (long long)(X.nr + 3) * X.nc;
    X.HS = X.GY + (long long)X.nr * X.SG;
    X.how = (X.r1 - X.r0) * (X.c1 - X.c0);
    X.boff = (long long)sim * A.plane;
    X.boffD = (long long)sim * A.plane_D;
    X.stims = A.stims ? A.stims + (long long)sim * A.n_stim : nullptr;
    for (int s = 0; s < 8; ++s) X.mask[s] = 0;
    for (int s = 0; s < A.T; ++s) {
        const double t = A.t0 + (double)s;
        unsigned m = 0;
        for (int i = 0; i < A.n_stim; ++i) {
            const StimDev sd = X.stims[i];
            if (sd.field && stim_on(sd, t, A.t_is_int)) m |= 1u << i;
        }
        X.mask[s] = m;
    }
    if (A.heun) X.mask[1] = X.mask[0];   // both stages see the same counter (solve.py:78, 80)
}

// phase 0: level-0 state -> shared memory
FK_HD void tile_load(const TileArgs& A, const TileCtx& X, int tx, int ty, int ntx, int nty) {
    for (int r = X.ra + ty; r < X.rb; r += nty)
        for (int c0 = X.ca + tx; c0 < X.cb; c0 += 4 * ntx) {   // batches of 4 cells: 12 loads in flight per thread
            float uu[4], vv[4], ww[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int c = c0 + q * ntx;
                if (c < X.cb) {
                    const long long g = X.boff + (long long)r * A.W + c;
                    uu[q] = ldg1(A.u_in + g); vv[q] = ldg1(A.v_in + g); ww[q] = ldg1(A.w_in + g);
                }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int c = c0 + q * ntx;
                if (c < X.cb) {
                    const int i = (r - X.ra) * X.nc + (c - X.ca);
                    X.U0[i] = uu[q]; X.V[i] = vv[q]; X.Wd[i] = ww[q];
                }
            }
        }
}

// phase A of level s: u_x and u_y on the padded array (solve.py:49-50), from level s-1 in Uc
template <bool EXACT>
FK_HD void tile_grad(const TileArgs& A, const TileCtx& X, int s, const float* Uc, int tx, int ty, int ntx, int nty) {
    int a, b, c, d;
    level_rect(A, X, s, a, b, c, d);
    // u_x at padded rows PA..PB, tissue columns c..d-1
    const int PA = A.phys_top ? (a - 1 < 0 ? 0 : a - 1) : a - 1;
    const int PB = A.phys_bot ? (b + 2 > A.H + 1 ? A.H + 1 : b + 2) : b + 2;
    for (int P = PA + ty; P <= PB; P += nty) {
        float k0, k1, k2, k3;
        int o0, o1, o2, o3;
        const int kind = kind_of(P, A.H, A.phys_top, A.phys_bot);
        kind_coeffs(kind, k0, k1, k2, k3, o0, o1, o2, o3);
        const int i0 = (clampi(P + o0 - 1, 0, A.H - 1) - X.ra) * X.nc, i1 = (clampi(P + o1 - 1, 0, A.H - 1) - X.ra) * X.nc,
                  i2 = (clampi(P + o2 - 1, 0, A.H - 1) - X.ra) * X.nc, i3 = (clampi(P + o3 - 1, 0, A.H - 1) - X.ra) * X.nc;
        for (int col = c + tx; col < d; col += ntx) {
            const int j = col - X.ca;
            X.GX[(P - X.ra + 1) * X.nc + j] =
                deriv<EXACT>(A.K, kind, k0, k1, k2, k3, Uc[i0 + j], Uc[i1 + j], Uc[i2 + j], Uc[i3 + j]);
        }
    }
    // u_y at tissue rows a..b-1, padded columns QA..QB
    const int QA = A.phys_left ? (c - 1 < 0 ? 0 : c - 1) : c - 1;
    const int QB = A.phys_right ? (d + 2 > A.W + 1 ? A.W + 1 : d + 2) : d + 2;
    for (int row = a + ty; row < b; row += nty) {
        const float* Ur = Uc + (row - X.ra) * X.nc - X.ca;
        for (int Q = QA + tx; Q <= QB; Q += ntx) {
            float k0, k1, k2, k3;
            int o0, o1, o2, o3;
            const int kind = kind_of(Q, A.W, A.phys_left, A.phys_right);
            kind_coeffs(kind, k0, k1, k2, k3, o0, o1, o2, o3);
            X.GY[(row - X.ra) * X.SG + (Q - X.ca + 1)] =
                deriv<EXACT>(A.K, kind, k0, k1, k2, k3, Ur[clampi(Q + o0 - 1, 0, A.W - 1)], Ur[clampi(Q + o1 - 1, 0, A.W - 1)],
                             Ur[clampi(Q + o2 - 1, 0, A.W - 1)], Ur[clampi(Q + o3 - 1, 0, A.W - 1)]);
        }
    }
}

// phase B of level s: second derivatives, reaction, stimulus, Euler update (solve.py:35-59, 70)
template <bool EXACT>
FK_HD void tile_update(const TileArgs& A, const TileCtx& X, int s, const float* Uc, float* Un, int tx, int ty, int ntx,
                       int nty) {
    int a, b, c, d;
    level_rect(A, X, s, a, b, c, d);
    const bool last = (s == A.T);
    const unsigned mask = X.mask[s - 1];
    for (int row = a + ty; row < b; row += nty) {
        float k0, k1, k2, k3;
        int o0, o1, o2, o3;
        const int P = row + 1;
        const int kindr = kind_of(P, A.H, A.phys_top, A.phys_bot);
        kind_coeffs(kindr, k0, k1, k2, k3, o0, o1, o2, o3);
        const float* G0 = X.GX + (P + o0 - X.ra + 1) * X.nc - X.ca;
        const float* G1 = X.GX + (P + o1 - X.ra + 1) * X.nc - X.ca;
        const float* G2 = X.GX + (P + o2 - X.ra + 1) * X.nc - X.ca;
        const float* G3 = X.GX + (P + o3 - X.ra + 1) * X.nc - X.ca;
        const float* GC = X.GX + (P - X.ra + 1) * X.nc - X.ca;
        const float* GYr = X.GY + (row - X.ra) * X.SG - X.ca + 1;
        for (int col0 = c + tx; col0 < d; col0 += 4 * ntx) {
            // batches of 4 cells: the diffusivity maps (and stimulus fields) of all four are requested before any is used
            float Dq[4], DXq[4], DYq[4], stq[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int col = col0 + q * ntx;
                Dq[q] = DXq[q] = DYq[q] = stq[q] = 0.0f;
                if (col < d) {
                    const long long gd = X.boffD + (long long)row * A.W + col;
                    Dq[q] = ldg1(A.D + gd); DXq[q] = ldg1(A.DX + gd); DYq[q] = ldg1(A.DY + gd);
                    if (mask) {  // solve.py:260-269: later stimuli override earlier ones, zero cells never stimulate
                        const long long g = (long long)row * A.W + col;
                        for (int i = 0; i < A.n_stim; ++i)
                            if (mask >> i & 1u) {
                                const float f = ldg1(X.stims[i].field + g);
                                if (f != 0.0f) stq[q] = f;
                            }
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int col = col0 + q * ntx;
                if (col >= d) continue;
                const int Q = col + 1;
                const float u_xx = deriv<EXACT>(A.K, kindr, k0, k1, k2, k3, G0[col], G1[col], G2[col], G3[col]);
                float q0, q1, q2, q3;
                int p0, p1, p2, p3;
                const int kindc = kind_of(Q, A.W, A.phys_left, A.phys_right);
                kind_coeffs(kindc, q0, q1, q2, q3, p0, p1, p2, p3);
                const float u_yy =
                    deriv<EXACT>(A.K, kindc, q0, q1, q2, q3, GYr[Q + p0], GYr[Q + p1], GYr[Q + p2], GYr[Q + p3]);
                const float u_x = GC[col], u_y = GYr[Q];
                const long long g = (long long)row * A.W + col;
                const int i = (row - X.ra) * X.nc + (col - X.ca);
                const float u = Uc[i], v = X.V[i], w = X.Wd[i];
                float d_v = 0.0f, d_w = 0.0f, d_u = 0.0f, vn, wn, un;
                if (A.rhs_mode || A.heun) {   // the derivatives themselves are wanted (solve.step, the Heun stages)
                    const float del_u = diffusion<EXACT>(Dq[q], DXq[q], DYq[q], u_x, u_y, u_xx, u_yy);
                    cell_rhs<EXACT>(A.K, u, v, w, del_u, stq[q], d_v, d_w, d_u);
                    vn = euler<EXACT>(v, d_v, A.K.dt); wn = euler<EXACT>(w, d_w, A.K.dt); un = euler<EXACT>(u, d_u, A.K.dt);
                } else {                      // the ordinary Euler update: the same cell function as every other kernel
                    cell_step<EXACT>(A.K, u, v, w, Dq[q], DXq[q], DYq[q], u_x, u_y, u_xx, u_yy, stq[q], un, vn, wn);
                }
                if (A.rhs_mode) {
                    A.v_out[X.boff + g] = d_v;
                    A.w_out[X.boff + g] = d_w;
                    A.u_out[X.boff + g] = d_u;
                    continue;
                }
                if (A.heun) {
                    const bool mine = row >= X.r0 && row < X.r1 && col >= X.c0 && col < X.c1;
                    const int o = (row - X.r0) * (X.c1 - X.c0) + (col - X.c0);
                    if (!last) {   // predictor: keep y and k1 of the cells this tile will write
                        if (mine) {
                            X.HS[o] = v; X.HS[X.how + o] = w;
                            X.HS[2 * X.how + o] = d_v; X.HS[3 * X.how + o] = d_w; X.HS[4 * X.how + o] = d_u;
                        }
                    } else {       // corrector: y + (k1 + k2) * (dt * 0.5)
                        if (mine) {
                            typedef Num<EXACT> N;
                            A.v_out[X.boff + g] = euler<EXACT>(X.HS[o], N::add(X.HS[2 * X.how + o], d_v), A.h_half);
                            A.w_out[X.boff + g] = euler<EXACT>(X.HS[X.how + o], N::add(X.HS[3 * X.how + o], d_w), A.h_half);
                            A.u_out[X.boff + g] = euler<EXACT>(X.U0[i], N::add(X.HS[4 * X.how + o], d_u), A.h_half);
                        }
                        continue;
                    }
                }
                if (last) {
                    if (row >= X.r0 && row < X.r1 && col >= X.c0 && col < X.c1) {
                        A.v_out[X.boff + g] = vn;
                        A.w_out[X.boff + g] = wn;
                        A.u_out[X.boff + g] = un;
                    }
                } else {
                    Un[i] = un;
                    X.V[i] = vn;
                    X.Wd[i] = wn;
                }
            }
        }
    }
}

}  // namespace fk
