// fk_wide.h -- the LOW-LATENCY single-step kernel body for tissues too small to fill the machine.
//
// The streaming kernel marches down the rows, so one Euler step of a 128^2 ... 512^2 tissue costs a serial chain of
// >= 16 row iterations (~10 us) whatever the size.  Here every thread owns 4 adjacent cells of ONE row and nothing
// is sequential: it reads the 9 x 4 u values above/below and the 12 values left/right of its cells straight from
// L1/L2 (such tissues are cache resident), rebuilds the reference's two-pass derivative in registers (u_x at 5 rows,
// u_y at 8 columns, then the second pass -- same operations, same order as solve.py:49-52), and finishes the cell.
// No shared memory, no barrier, the whole tissue in one launch including all four physical edges: per axis, a thread
// within 5 cells of an edge evaluates the padded-array formulas of whichever kind applies on a slightly wider,
// index-clamped window that still lives in registers (wide_axis_general).  One launch = one step (T = 1).
#pragma once
#include "fk_core.h"
#include "fk_stream.h"
#include "fk_tile.h"

namespace fk {

// first derivative / dx (solve.py:225-254) of whichever kind, from a window a[0..6] centred on a[3]: the central
// formula reads a[1], a[2], a[4], a[5]; the forward one a[3..6]; the backward one a[0..3].  Every index is a
// compile-time constant, so the windows stay in registers.
template <bool EXACT>
FK_HD float wide_deriv(const Consts& K, int kind, const float* a) {
    if (kind == CEN) return dcen<EXACT>(K, a[1], a[2], a[4], a[5]);
    const bool f = kind == FWD;
    float k0, k1, k2, k3;
    int o0, o1, o2, o3;
    kind_coeffs(kind, k0, k1, k2, k3, o0, o1, o2, o3);
    return deriv<EXACT>(K, kind, k0, k1, k2, k3, f ? a[3] : a[0], f ? a[4] : a[1], f ? a[5] : a[2], f ? a[6] : a[3]);
}

// First and second derivative along one axis of the edge-padded array (solve.py:29-31, 49-52, crop :61-65) at padded
// index P (tissue index P - 1) of an axis of n cells, anywhere on the axis.  up[0..12] are the padded values
// P-6 .. P+6 (pad = repeat of the edge cell; entries beyond the pad are never used by the formulas that apply).
template <bool EXACT>
FK_HD void wide_axis_general(const Consts& K, const float* up, int P, int n, float& d1, float& d2) {
    float g[7];   // first derivative at padded P-3 .. P+3
#pragma unroll
    for (int m = 0; m < 7; ++m) g[m] = wide_deriv<EXACT>(K, kind_of(P - 3 + m, n, 1, 1), up + m);
    d1 = g[3];
    d2 = wide_deriv<EXACT>(K, kind_of(P, n, 1, 1), g);
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
    // ---- every load of the thread is issued up front, whichever formulas apply: the kernel is latency bound (such
    // tissues are cache resident), and one round trip to L2 instead of four is most of its run time
    float ur[13][4];   // padded rows P-6 .. P+6, P = row + 1: tissue rows row-6 .. row+6 clamped (solve.py:31)
#pragma unroll
    for (int j = 0; j < 13; ++j) unpack4(ld4(u + (long long)clampi(row + j - 6, 0, H - 1) * W + c), ur[j]);
    float e[16];       // padded columns Q0-6 .. Q0+9, Q0 = c + 1: tissue columns c-6 .. c+9 clamped
    {
        const float* urow = u + (long long)row * W;
#pragma unroll
        for (int m = 0; m < 16; ++m) e[m] = (m >= 6 && m < 10) ? 0.0f : urow[clampi(c + m - 6, 0, W - 1)];
    }
    float v[4], w[4], Dv[4], DXv[4], DYv[4], stim[4] = {0.f, 0.f, 0.f, 0.f};
    unpack4(ld4(A.v_in + boff + g), v);
    unpack4(ld4(A.w_in + boff + g), w);
    unpack4(ldg4(A.D + boffD + g), Dv);
    unpack4(ldg4(A.DX + boffD + g), DXv);
    unpack4(ldg4(A.DY + boffD + g), DYv);
    if (mask) {  // solve.py:260-269
        const StimDev* st = A.stims + (long long)sim * A.n_stim;
        for (int q = 0; q < A.n_stim; ++q)
            if (mask >> q & 1u) {
                float f[4];
                unpack4(ldg4(st[q].field + g), f);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (f[k] != 0.0f) stim[k] = f[k];
            }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) { uc[k] = ur[6][k]; e[6 + k] = ur[6][k]; }
    // ---- vertical (axis 0), per column: central formulas on rows row-4 .. row+4 in the interior, the general ones on
    // the 13 clamped rows near the top / bottom edge
    if (row_in) {
        float gxv[5][4];
#pragma unroll
        for (int j = 0; j < 5; ++j)
#pragma unroll
            for (int k = 0; k < 4; ++k)
                gxv[j][k] = dcen<EXACT>(A.K, ur[j + 2][k], ur[j + 3][k], ur[j + 5][k], ur[j + 6][k]);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            u_x[k] = gxv[2][k];
            u_xx[k] = dcen<EXACT>(A.K, gxv[0][k], gxv[1][k], gxv[3][k], gxv[4][k]);
        }
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float col[13];
#pragma unroll
            for (int j = 0; j < 13; ++j) col[j] = ur[j][k];
            wide_axis_general<EXACT>(A.K, col, row + 1, H, u_x[k], u_xx[k]);
        }
    }
    // ---- horizontal (axis 1): central formulas on columns c-4 .. c+7 in the interior, the general ones on the 16
    // clamped columns for the threads at the left / right edge
    if (col_in) {
        float gyv[8];
#pragma unroll
        for (int m = 0; m < 8; ++m) gyv[m] = dcen<EXACT>(A.K, e[m + 2], e[m + 3], e[m + 5], e[m + 6]);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            u_y[k] = gyv[k + 2];
            u_yy[k] = dcen<EXACT>(A.K, gyv[k], gyv[k + 1], gyv[k + 3], gyv[k + 4]);
        }
    } else {
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
        cell_step<EXACT>(A.K, uc[k], v[k], w[k], Dv[k], DXv[k], DYv[k], u_x[k], u_y[k], u_xx[k], u_yy[k], stim[k], un[k], vn[k],
                         wn[k]);
    }
    if (A.hy_u) {   // fast Heun (see fk_stream.h)
        float y4[4];
        unpack4(ldg4(A.hy_u + boff + g), y4); heun_fold4(y4, un);
        unpack4(ldg4(A.hy_v + boff + g), y4); heun_fold4(y4, vn);
        unpack4(ldg4(A.hy_w + boff + g), y4); heun_fold4(y4, wn);
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
        for (int i = 0; i < A.n_stim; ++i) {
            const StimDev sd = st[i];
            if (sd.field && stim_on(sd, A.t0, A.t_is_int)) m |= 1u << i;
        }
    }
    return m;
}

}  // namespace fk
