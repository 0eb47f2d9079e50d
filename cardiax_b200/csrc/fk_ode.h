// fk_ode.h -- the adaptive Dormand-Prince loop behind solve._forward_dormandprince (cardiax/solve.py:114-124), i.e.
// jax.experimental.ode.odeint(step, state, ts, params, diffusivity, stimuli, dx) with its defaults rtol = atol = 1.4e-8,
// mxstep = inf.  jax is an un-vendored dependency of the reference (install_jax.sh:2 pins jaxlib 0.1.64, jax itself
// unpinned); this restates the published algorithm of that vintage (jax/experimental/ode.py: initial_step_size,
// runge_kutta_step, error_ratio, optimal_step_size, interp_fit_dopri, _odeint.scan_fun).
//
// The state arrays never leave the device; the step-size controller is a handful of fp32 scalars on the host (each
// operation rounded to fp32 like the traced fp32 scalars of the reference), fed by ONE reduced number per attempt.
// Backend-independent like fk_driver.h: fk_api.cu runs it on CUDA kernels, tests/emu on the CPU emulation.
#pragma once
#include <math.h>

#include "fk_aux.h"

namespace fk {

struct P3 {   // the three arrays of a State (v, w, u)
    float* a[3];
};

struct OdeBuffers {
    P3 y, ys, yn;        // current state, stage state, candidate state
    P3 k[7];             // stage derivatives
    P3 c[2][5];          // interpolation coefficients: [which][a, b, c, d, e]
    P3 out;              // (n_ts, n) per array
    long long n;         // elements per array (batch * H * W)
};

struct OdeStats {
    long long attempts, accepted, rhs_evals;
};

namespace ode {
inline float f32(double x) { return (float)x; }
inline float add(float a, float b) { return f32((double)a + (double)b); }
inline float sub(float a, float b) { return f32((double)a - (double)b); }
inline float mul(float a, float b) { return f32((double)a * (double)b); }
inline float div(float a, float b) { return f32((double)a / (double)b); }
inline float sqrt_(float a) { return f32(sqrt((double)a)); }
inline float pow_(float a, float b) { return f32(pow((double)a, (double)b)); }   // double pow, rounded once
inline float norm(double sumsq) { return sqrt_(f32(sumsq)); }

// optimal_step_size(last_step, mean_error_ratio, safety=0.9, ifactor=10.0, dfactor=0.2, order=5.0)
inline float optimal_step_size(float last_step, float mean_error_ratio) {
    const float safety = f32(0.9), ifactor = f32(10.0);
    const float dfactor = mean_error_ratio < 1.0f ? 1.0f : f32(0.2);
    const float err_ratio = sqrt_(mean_error_ratio);
    float factor = div(pow_(err_ratio, f32(1.0 / 5.0)), safety);
    const float hi = div(1.0f, dfactor);
    if (hi < factor) factor = hi;                       // jnp.minimum(.., 1 / dfactor)
    if (f32(1.0 / 10.0) > factor) factor = f32(1.0 / 10.0);   // jnp.maximum(1 / ifactor, ..)
    return mean_error_ratio == 0.0f ? mul(last_step, ifactor) : div(last_step, factor);
}
}  // namespace ode

// Backend BE:
//   int rhs(const P3& y, const P3& k, float t)                      k = step(y, t)
//   int copy(const P3& dst, long long dst_off, const P3& src)       dst[dst_off ..] = src
//   int init_norms(const P3& y, const P3& f, float rtol, float atol, double* sumsq2)   sum (y/scale)^2, sum (f/scale)^2
//   int axpy(const P3& y, float h, const P3& f, const P3& out)      out = y + h f
//   int diff_norm(const P3& f1, const P3& f0, const P3& y, float rtol, float atol, double* sumsq)
//   int stage(int i, const P3& y, const P3* k, float dt, const P3& ys)
//   int finish(const P3& y, const P3* k, float dt, float rtol, float atol, const P3& yn, const P3* c, double* sum)
//   int interp(const P3* c, float r, const P3& out, long long off)
template <class BE>
int drive_dopri5(BE& be, OdeBuffers& B, int n_ts, const float* ts, float rtol, float atol, double mxstep, OdeStats* stats) {
    using namespace ode;
    const Dopri T = make_dopri();
    OdeStats S = {0, 0, 0};
    int rc;
    if (n_ts <= 0) return 0;
    const float nf = f32((double)(3 * B.n));
    // f0 = func(y0, ts[0]); dt = initial_step_size(func, ts[0], y0, 4, rtol, atol, f0)
    float t = ts[0];
    if ((rc = be.rhs(B.y, B.k[0], t))) return rc;
    ++S.rhs_evals;
    double s2[2], s1;
    if ((rc = be.init_norms(B.y, B.k[0], rtol, atol, s2))) return rc;
    const float d0 = norm(s2[0]), d1 = norm(s2[1]);
    const float h0 = (d0 < f32(1e-5) || d1 < f32(1e-5)) ? f32(1e-6) : div(mul(f32(0.01), d0), d1);
    if ((rc = be.axpy(B.y, h0, B.k[0], B.ys))) return rc;
    if ((rc = be.rhs(B.ys, B.k[1], add(t, h0)))) return rc;
    ++S.rhs_evals;
    if ((rc = be.diff_norm(B.k[1], B.k[0], B.y, rtol, atol, &s1))) return rc;
    const float d2 = div(norm(s1), h0);
    float h1;
    if (d1 <= f32(1e-15) && d2 <= f32(1e-15)) {
        const float a = f32(1e-6), b = mul(h0, f32(1e-3));
        h1 = a > b ? a : b;
    } else {
        h1 = pow_(div(f32(0.01), add(d1, d2)), f32(1.0 / (4 + 1.0)));
    }
    float dt = mul(f32(100.0), h0) < h1 ? mul(f32(100.0), h0) : h1;
    // interp_coeff = [y0] * 5; carry = [y0, f0, ts[0], dt, ts[0], interp_coeff]
    int cur = 0;
    for (int j = 0; j < 5; ++j)
        if ((rc = be.copy(B.c[cur][j], 0, B.y))) return rc;
    if ((rc = be.copy(B.out, 0, B.y))) return rc;   // jnp.concatenate((y0[None], ys))
    float last_t = t;
    for (int it = 1; it < n_ts; ++it) {
        const float target = ts[it];
        double i = 0;
        while (t < target && i < mxstep && dt > 0.0f) {
            // runge_kutta_step
            for (int s = 1; s < 7; ++s) {
                const float ti = add(t, mul(dt, T.alpha[s - 1]));
                if ((rc = be.stage(s, B.y, B.k, dt, B.ys))) return rc;
                if ((rc = be.rhs(B.ys, B.k[s], ti))) return rc;
                ++S.rhs_evals;
            }
            double sum;
            if ((rc = be.finish(B.y, B.k, dt, rtol, atol, B.yn, B.c[cur ^ 1], &sum))) return rc;
            const float ratio = div(f32(sum), nf);          // jnp.mean(err_ratio ** 2)
            const float next_t = add(t, dt);
            const float new_dt = optimal_step_size(dt, ratio);
            ++S.attempts;
            if (ratio <= 1.0f) {                             // accept: y, f, t, last_t, interp_coeff move on
                ++S.accepted;
                P3 tmp = B.y; B.y = B.yn; B.yn = tmp;
                tmp = B.k[0]; B.k[0] = B.k[6]; B.k[6] = tmp;
                cur ^= 1;
                last_t = t;
                t = next_t;
            }
            dt = new_dt;
            i += 1;
        }
        // relative_output_time = (target_t - last_t) / (t - last_t); y_target = polyval(interp_coeff, ..)
        const float r = div(sub(target, last_t), sub(t, last_t));
        if ((rc = be.interp(B.c[cur], r, B.out, (long long)it * B.n))) return rc;
    }
    if (stats) *stats = S;
    return 0;
}

}  // namespace fk
