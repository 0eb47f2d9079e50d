// fk_wide.h -- the LOW-LATENCY single-step kernel body for tissues too small to fill the machine.
//
// The streaming kernel marches down the rows, so one Euler step of a 128^2 ... 512^2 tissue costs a serial chain of
// >= 16 row iterations (~10 us) whatever the size.  Here every thread owns 4 adjacent cells of ONE row and nothing
// is sequential: it reads the 9 x 4 u values above/below and the 12 values left/right of its cells straight from
// L1/L2 (such tissues are cache resident), rebuilds the reference's two-pass derivative in registers (u_x at 5 rows,
// u_y at 8 columns, then the second pass -- same operations, same order as solve.py:49-52), and finishes the cell.
// No shared memory, no barrier, the whole tissue in one launch including all four physical edges: rows/columns
// within 5 cells of an edge take a general path that evaluates the padded-array formulas cell by cell (warp-uniform
// for rows, two warps per row for columns).  One launch = one step (T = 1).
#pragma once
#include "fk_core.h"
#include "fk_stream.h"
#include "fk_tile.h"

namespace fk {

// u of the edge-padded array (solve.py:31) at padded indices (P, Q)
FK_HD float wide_upad(const float* u, int H, int W, int P, int Q) {
    return u[(long long)clampi(P - 1, 0, H - 1) * W + clampi(Q - 1, 0, W - 1)];
}

// u_x (axis 0) / u_y (axis 1) of the padded array at padded (P, Q): solve.py:49-50
template <bool EXACT>
FK_HD float wide_g1(const Consts& K, const float* u, int H, int W, int axis, int P, int Q) {
    float k0, k1, k2, k3;
    int o0, o1, o2, o3;
    if (axis == 0) {
        const int kind = kind_of(P, H, 1, 1);
        kind_coeffs(kind, k0, k1, k2, k3, o0, o1, o2, o3);
        return deriv<EXACT>(K, kind, k0, k1, k2, k3, wide_upad(u, H, W, P + o0, Q), wide_upad(u, H, W, P + o1, Q),
                            wide_upad(u, H, W, P + o2, Q), wide_upad(u, H, W, P + o3, Q));
    }
    const int kind = kind_of(Q, W, 1, 1);
    kind_coeffs(kind, k0, k1, k2, k3, o0, o1, o2, o3);
    return deriv<EXACT>(K, kind, k0, k1, k2, k3, wide_upad(u, H, W, P, Q + o0), wide_upad(u, H, W, P, Q + o1),
                        wide_upad(u, H, W, P, Q + o2), wide_upad(u, H, W, P, Q + o3));
}

// general path: first and second derivatives of one cell from the padded-array formulas (any position)
template <bool EXACT>
FK_HD void wide_cell_general(const Consts& K, const float* u, int H, int W, int row, int col, float& u_x, float& u_y,
                             float& u_xx, float& u_yy) {
    const int P = row + 1, Q = col + 1;
    float k0, k1, k2, k3;
    int o0, o1, o2, o3;
    u_x = wide_g1<EXACT>(K, u, H, W, 0, P, Q);
    u_y = wide_g1<EXACT>(K, u, H, W, 1, P, Q);
    int kind = kind_of(P, H, 1, 1);
    kind_coeffs(kind, k0, k1, k2, k3, o0, o1, o2, o3);
    u_xx = deriv<EXACT>(K, kind, k0, k1, k2, k3, wide_g1<EXACT>(K, u, H, W, 0, P + o0, Q),
                        wide_g1<EXACT>(K, u, H, W, 0, P + o1, Q), wide_g1<EXACT>(K, u, H, W, 0, P + o2, Q),
                        wide_g1<EXACT>(K, u, H, W, 0, P + o3, Q));
    kind = kind_of(Q, W, 1, 1);
    kind_coeffs(kind, k0, k1, k2, k3, o0, o1, o2, o3);
    u_yy = deriv<EXACT>(K, kind, k0, k1, k2, k3, wide_g1<EXACT>(K, u, H, W, 1, P, Q + o0),
                        wide_g1<EXACT>(K, u, H, W, 1, P, Q + o1), wide_g1<EXACT>(K, u, H, W, 1, P, Q + o2),
                        wide_g1<EXACT>(K, u, H, W, 1, P, Q + o3));
}

// one thread: cells (row, c .. c+3) of tissue `sim`; mask = stimuli active at this step
template <bool EXACT>
FK_HD void wide_thread(const TileArgs& A, int sim, int row, int c, unsigned mask) {
    const int H = A.H, W = A.W;
    const long long boff = (long long)sim * A.plane, boffD = (long long)sim * A.plane_D;
    const float* u = A.u_in + boff;
    const long long g = (long long)row * W + c;
    float u_x[4], u_y[4], u_xx[4], u_yy[4], uc[4];
    const bool row_in = row >= 4 && row + 5 <= H;   // every vertical formula central, no clamped row
    const bool col_in = c >= 4 && c + 8 <= W;       // same for the columns of all four cells
    if (row_in && col_in) {
        float ur[9][4];
#pragma unroll
        for (int j = 0; j < 9; ++j) unpack4(ld4(u + g + (long long)(j - 4) * W), ur[j]);
        float gxv[5][4];
#pragma unroll
        for (int j = 0; j < 5; ++j)
#pragma unroll
            for (int k = 0; k < 4; ++k) gxv[j][k] = dcen<EXACT>(A.K, ur[j][k], ur[j + 1][k], ur[j + 3][k], ur[j + 4][k]);
        float e[12];
        unpack4(ld4(u + g - 4), e);
        unpack4(ld4(u + g + 4), e + 8);
#pragma unroll
        for (int k = 0; k < 4; ++k) e[4 + k] = ur[4][k];
        float gyv[8];
#pragma unroll
        for (int m = 0; m < 8; ++m) gyv[m] = dcen<EXACT>(A.K, e[m], e[m + 1], e[m + 3], e[m + 4]);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            uc[k] = ur[4][k];
            u_x[k] = gxv[2][k];
            u_xx[k] = dcen<EXACT>(A.K, gxv[0][k], gxv[1][k], gxv[3][k], gxv[4][k]);
            u_y[k] = gyv[k + 2];
            u_yy[k] = dcen<EXACT>(A.K, gyv[k], gyv[k + 1], gyv[k + 3], gyv[k + 4]);
        }
    } else {
        unpack4(ld4(u + g), uc);
        for (int k = 0; k < 4; ++k) wide_cell_general<EXACT>(A.K, u, H, W, row, c + k, u_x[k], u_y[k], u_xx[k], u_yy[k]);
    }
    float v[4], w[4], Dv[4], DXv[4], DYv[4], stim[4] = {0.f, 0.f, 0.f, 0.f};
    unpack4(ld4(A.v_in + boff + g), v);
    unpack4(ld4(A.w_in + boff + g), w);
    unpack4(ld4(A.D + boffD + g), Dv);
    unpack4(ld4(A.DX + boffD + g), DXv);
    unpack4(ld4(A.DY + boffD + g), DYv);
    if (mask) {  // solve.py:260-269
        const StimDev* st = A.stims + (long long)sim * A.n_stim;
        for (int q = 0; q < A.n_stim; ++q)
            if (mask >> q & 1u) {
                float f[4];
                unpack4(ld4(st[q].field + g), f);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (f[k] != 0.0f) stim[k] = f[k];
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
    st4(A.u_out + boff + g, un);
    st4(A.v_out + boff + g, vn);
    st4(A.w_out + boff + g, wn);
}

// stimuli of tissue `sim` active at counter A.t0
FK_HD unsigned wide_mask(const TileArgs& A, int sim) {
    unsigned m = 0;
    if (A.n_stim) {
        const StimDev* st = A.stims + (long long)sim * A.n_stim;
        const float t = (float)A.t0;
        for (int i = 0; i < A.n_stim; ++i) {
            const StimDev sd = st[i];
            if (sd.field && stim_active(t, sd.start, sd.duration, sd.period)) m |= 1u << i;
        }
    }
    return m;
}

}  // namespace fk
