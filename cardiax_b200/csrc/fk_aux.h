// fk_aux.h -- element bodies of the kernels AROUND the Euler hot path: the snapshot resize of
// cardiax/io.py:118-124, the adaptive Dormand-Prince integrator behind solve._forward_dormandprince
// (cardiax/solve.py:88-89, 114-124 -> jax.experimental.ode.odeint) and metrics.electrogram (cardiax/metrics.py:13-22).
//
// Like fk_core.h everything is FK_HD and free of CUDA builtins: tests/emu compiles the same bodies with g++.
#pragma once
#include <math.h>

#include <vector>

#include "fk_core.h"

namespace fk {

// ------------------------------------------------------------------------------------------------ resize
// jax.image.resize(a, shape, "bilinear") (antialias=True): a separable triangle filter of radius
// max(n_in / n_out, 1) around the half-pixel-centred sample position, weights renormalised to sum 1 and zeroed when the
// sample falls outside the input (jax/_src/image/scale.py: compute_weight_mat).  Per output index the non-zero taps are
// one contiguous run; the tables hold its first input index and K weights (zero padded).
struct ResizeAxis {
    int n_in, n_out, K;
    std::vector<int> lo;      // n_out
    std::vector<float> wt;    // n_out * K
};

inline ResizeAxis make_resize_axis(int n_in, int n_out) {
    ResizeAxis A;
    A.n_in = n_in; A.n_out = n_out;
    const double inv_scale = (double)n_in / (double)n_out;
    const double ks = inv_scale > 1.0 ? inv_scale : 1.0;
    A.K = (int)floor(2.0 * ks) + 2;
    if (A.K > n_in) A.K = n_in;
    A.lo.assign(n_out, 0);
    A.wt.assign((size_t)n_out * A.K, 0.0f);
    std::vector<double> w(A.K);
    for (int o = 0; o < n_out; ++o) {
        const double sf = ((double)o + 0.5) * inv_scale - 0.5;
        int lo = (int)ceil(sf - ks);
        if (lo > n_in - A.K) lo = n_in - A.K;
        if (lo < 0) lo = 0;
        double total = 0.0;
        for (int k = 0; k < A.K; ++k) {
            const double x = fabs(sf - (double)(lo + k)) / ks;
            w[k] = x < 1.0 ? 1.0 - x : 0.0;
            total += w[k];
        }
        const bool inside = sf >= -0.5 && sf <= (double)n_in - 0.5;
        const bool ok = fabs(total) > 1000.0 * 1.1920928955078125e-07;
        A.lo[o] = lo;
        for (int k = 0; k < A.K; ++k) A.wt[(size_t)o * A.K + k] = (inside && ok) ? (float)(w[k] / total) : 0.0f;
    }
    return A;
}

enum { RESIZE_BY_VALUE = 8 };
struct ResizeArgs {
    const float* plane[RESIZE_BY_VALUE];   // n_planes <= 8: the pointers themselves (a snapshot has three)
    const float* const* planes;            // otherwise: device table of n_planes pointers to (H, W) arrays
    float* out;                            // (n_planes, Ho, Wo)
    int n_planes, H, W, Ho, Wo, Kh, Kw;
    const int* lo_h; const float* wt_h;    // Ho ints; Kh x Ho weights, TAP-major (coalesced across output indices)
    const int* lo_w; const float* wt_w;    // Wo ints; Kw x Wo weights
};

// tap-major copy of an axis' weights: t[k * n_out + o]
inline std::vector<float> resize_tap_major(const ResizeAxis& A) {
    std::vector<float> t((size_t)A.K * A.n_out);
    for (int o = 0; o < A.n_out; ++o)
        for (int k = 0; k < A.K; ++k) t[(size_t)k * A.n_out + o] = A.wt[(size_t)o * A.K + k];
    return t;
}

// one output pixel: rows outer, columns inner, fused multiply-adds in tap order; wh[r * sh], ww[q * sw]
FK_HD float resize_pixel(const float* __restrict__ in, int W, int r0, const float* __restrict__ wh, int sh, int Kh, int c0,
                         const float* __restrict__ ww, int sw, int Kw) {
    float acc = 0.0f;
    for (int r = 0; r < Kh; ++r) {
        const float* row = in + (size_t)(r0 + r) * W + c0;
        float racc = 0.0f;
#pragma unroll 4
        for (int q = 0; q < Kw; ++q) racc = fmaf(row[q], ww[q * sw], racc);
        acc = fmaf(wh[r * sh], racc, acc);
    }
    return acc;
}

// ------------------------------------------------------------------------------------------------ Dopri5
// Butcher tableau, error and mid-point weights of jax.experimental.ode (runge_kutta_step, interp_fit_dopri): Python
// doubles rounded to fp32 when they meet the fp32 state.
struct Dopri {
    float alpha[7];
    float beta[6][7];
    float c_sol[7], c_err[7], c_mid[7];
};

inline Dopri make_dopri() {
    Dopri T;
    const double alpha[7] = {1. / 5, 3. / 10, 4. / 5, 8. / 9, 1., 1., 0};
    const double beta[6][7] = {{1. / 5, 0, 0, 0, 0, 0, 0},
                               {3. / 40, 9. / 40, 0, 0, 0, 0, 0},
                               {44. / 45, -56. / 15, 32. / 9, 0, 0, 0, 0},
                               {19372. / 6561, -25360. / 2187, 64448. / 6561, -212. / 729, 0, 0, 0},
                               {9017. / 3168, -355. / 33, 46732. / 5247, 49. / 176, -5103. / 18656, 0, 0},
                               {35. / 384, 0, 500. / 1113, 125. / 192, -2187. / 6784, 11. / 84, 0}};
    const double c_sol[7] = {35. / 384, 0, 500. / 1113, 125. / 192, -2187. / 6784, 11. / 84, 0};
    const double c_err[7] = {35. / 384 - 1951. / 21600, 0, 500. / 1113 - 22642. / 50085, 125. / 192 - 451. / 720,
                             -2187. / 6784 - -12231. / 42400, 11. / 84 - 649. / 6300, -1. / 60.};
    const double c_mid[7] = {6025192743. / 30085553152. / 2, 0, 51252292925. / 65400821598. / 2,
                             -2691868925. / 45128329728. / 2, 187940372067. / 1594534317056. / 2,
                             -1776094331. / 19743644256. / 2, 11237099. / 235043384. / 2};
    for (int i = 0; i < 7; ++i) {
        T.alpha[i] = (float)alpha[i]; T.c_sol[i] = (float)c_sol[i]; T.c_err[i] = (float)c_err[i]; T.c_mid[i] = (float)c_mid[i];
        for (int r = 0; r < 6; ++r) T.beta[r][i] = (float)beta[r][i];
    }
    return T;
}

template <bool EXACT>
FK_HD float ode_div(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fdiv_rn(a, b);
#else
    return a / b;
#endif
}

// dot(c, k) over the first n entries: products accumulated left to right, zero weights skipped (0 * k adds nothing)
template <bool EXACT>
FK_HD float ode_dot(const float* c, const float* k, int n) {
    typedef Num<EXACT> N;
    float acc = N::mul(c[0], k[0]);
    for (int j = 1; j < n; ++j)
        if (c[j] != 0.0f) acc = N::mad(c[j], k[j], acc);
    return acc;
}

// yi = y0 + dt * dot(beta[i-1], k)   (runge_kutta_step.body_fun)
template <bool EXACT>
FK_HD float ode_stage(float y0, const float* beta_row, const float* k, int i, float dt) {
    typedef Num<EXACT> N;
    return N::add(y0, N::mul(dt, ode_dot<EXACT>(beta_row, k, i)));
}

// the end of one attempt: y1, (err / tol)^2 and the five coefficients of the 4th-order interpolant
// (runge_kutta_step tail, error_ratio, interp_fit_dopri + fit_4th_order_polynomial)
template <bool EXACT>
FK_HD void ode_finish(const Dopri& T, float y0, const float* k, float dt, float rtol, float atol, float& y1, float& ratio2,
                      float* coef) {
    typedef Num<EXACT> N;
    y1 = N::add(N::mul(dt, ode_dot<EXACT>(T.c_sol, k, 7)), y0);
    const float err = N::mul(dt, ode_dot<EXACT>(T.c_err, k, 7));
    const float tol = N::add(atol, N::mul(rtol, fmaxf(fabsf(y0), fabsf(y1))));
    const float q = ode_div<EXACT>(err, tol);
    ratio2 = N::mul(q, q);
    const float ym = N::add(y0, N::mul(dt, ode_dot<EXACT>(T.c_mid, k, 7)));
    const float dy0 = k[0], dy1 = k[6];
    // a = -2.*dt*dy0 + 2.*dt*dy1 -  8.*y0 -  8.*y1 + 16.*y_mid        (left to right, scalar * dt first)
    float a = N::mul(N::mul(-2.0f, dt), dy0);
    a = N::add(a, N::mul(N::mul(2.0f, dt), dy1));
    a = N::sub(a, N::mul(8.0f, y0));
    a = N::sub(a, N::mul(8.0f, y1));
    a = N::add(a, N::mul(16.0f, ym));
    // b =  5.*dt*dy0 - 3.*dt*dy1 + 18.*y0 + 14.*y1 - 32.*y_mid
    float b = N::mul(N::mul(5.0f, dt), dy0);
    b = N::sub(b, N::mul(N::mul(3.0f, dt), dy1));
    b = N::add(b, N::mul(18.0f, y0));
    b = N::add(b, N::mul(14.0f, y1));
    b = N::sub(b, N::mul(32.0f, ym));
    // c = -4.*dt*dy0 + dt*dy1 - 11.*y0 - 5.*y1 + 16.*y_mid
    float c = N::mul(N::mul(-4.0f, dt), dy0);
    c = N::add(c, N::mul(dt, dy1));
    c = N::sub(c, N::mul(11.0f, y0));
    c = N::sub(c, N::mul(5.0f, y1));
    c = N::add(c, N::mul(16.0f, ym));
    coef[0] = a; coef[1] = b; coef[2] = c;
    coef[3] = N::mul(dt, dy0);
    coef[4] = y0;
}

// jnp.polyval(coef, r): Horner from the leading coefficient, multiply then add
template <bool EXACT>
FK_HD float ode_interp(const float* coef, float r) {
    typedef Num<EXACT> N;
    float y = coef[0];
    for (int j = 1; j < 5; ++j) y = N::add(N::mul(y, r), coef[j]);
    return y;
}

// initial_step_size: y / scale, f / scale with scale = atol + |y| * rtol
template <bool EXACT>
FK_HD void ode_scaled(float y, float f, float rtol, float atol, float& qy, float& qf) {
    typedef Num<EXACT> N;
    const float scale = N::add(atol, N::mul(fabsf(y), rtol));
    qy = ode_div<EXACT>(y, scale);
    qf = ode_div<EXACT>(f, scale);
}

// ------------------------------------------------------------------------------------------------ electrogram
// cardiax/metrics.py:13-22: sum over the frame of x[i][j] * sqrt((j - p0)^2 + (i - p1)^2) (`ogrid[:W, :H]` there, so
// the reference only broadcasts for square frames)
FK_HD float egm_weight(int i, int j, float p0, float p1) {
    typedef Num<true> N;
    const float dx = N::sub((float)j, p0), dy = N::sub((float)i, p1);
    return sqrtf(N::add(N::mul(dx, dx), N::mul(dy, dy)));
}

}  // namespace fk
