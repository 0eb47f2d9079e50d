/* TEST INFRASTRUCTURE ONLY -- C restatement of the reference's FK Euler loop.
 *
 * Included twice by fk_oracle.c with REAL = float / double and SUF = f32 / f64.
 * Follows /root/reference/cardiax/solve.py line by line, operation order kept:
 *   step            solve.py:26-65      gradient   solve.py:225-254
 *   step_euler      solve.py:68-70      stimulate  solve.py:257-271
 *   _forward_euler  solve.py:92-100
 * Compiled with -ffp-contract=off so no product/add is fused: the f32 build is
 * bit-identical to oracle/fk_oracle.py (checked by tests/test_oracle.py).
 * PARITY UNPINNED for u/v/w values -- see oracle/__init__.py.
 */

#define CAT_(a, b) a##_##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUF)

/* solve.py:225-254 -- one output of gradient() along a line of length n (stride s), times dx. */
static inline REAL FN(grad1)(const REAL *a, long n, long s, long i) {
    if (i < 2)
        return (((REAL)(-11.0 / 6.0) * a[i * s] + (REAL)3 * a[(i + 1) * s]) - (REAL)(3.0 / 2.0) * a[(i + 2) * s]) +
               (REAL)(1.0 / 3.0) * a[(i + 3) * s];
    if (i >= n - 2)
        return (((REAL)(-1.0 / 3.0) * a[(i - 3) * s] + (REAL)(3.0 / 2.0) * a[(i - 2) * s]) - (REAL)3 * a[(i - 1) * s]) +
               (REAL)(11.0 / 6.0) * a[i * s];
    return (((REAL)(1.0 / 12.0) * a[(i - 2) * s] - (REAL)(2.0 / 3.0) * a[(i - 1) * s]) + (REAL)(2.0 / 3.0) * a[(i + 1) * s]) -
           (REAL)(1.0 / 12.0) * a[(i + 2) * s];
}

static REAL FN(tanh_impl)(REAL x, int tanh_mode) {
#ifdef REAL_IS_FLOAT
    if (tanh_mode == 0) { /* XLA EmitFastTanh (jaxlib 0.1.64), un-contracted */
        static const float nc[7] = {-2.76076847742355e-16f, 2.00018790482477e-13f, -8.60467152213735e-11f,
                                    5.12229709037114e-08f,  1.48572235717979e-05f, 6.37261928875436e-04f,
                                    4.89352455891786e-03f};
        static const float dc[4] = {1.19825839466702e-06f, 1.18534705686654e-04f, 2.26843463243900e-03f,
                                    4.89352518554385e-03f};
        float xc = x < -9.0f ? -9.0f : x;
        xc = xc > 9.0f ? 9.0f : xc;
        float x2 = xc * xc;
        float num = nc[0];
        for (int i = 1; i < 7; ++i) num = x2 * num + nc[i];
        num = xc * num;
        float den = dc[0];
        for (int i = 1; i < 4; ++i) den = x2 * den + dc[i];
        return fabsf(x) < 0.0004f ? x : num / den;
    }
    return tanhf(x);
#else
    (void)tanh_mode;
    return tanh(x);
#endif
}

/* edge-pad a (H,W) array into (H+2,W+2): solve.py:29-32 */
static void FN(pad_edge)(const REAL *a, REAL *p, long H, long W) {
    long Wp = W + 2;
#pragma omp parallel for schedule(static)
    for (long r = 0; r < H + 2; ++r) {
        long rr = r - 1 < 0 ? 0 : (r - 1 > H - 1 ? H - 1 : r - 1);
        const REAL *src = a + rr * W;
        REAL *dst = p + r * Wp;
        dst[0] = src[0];
        for (long c = 0; c < W; ++c) dst[c + 1] = src[c];
        dst[Wp - 1] = src[W - 1];
    }
}

/* solve.py:262-267 in the counter's floating type (float for the `forward` path). */
int FN(fk_oracle_stim_active)(REAL t, REAL start, REAL duration, REAL period) {
    if (!(t >= start)) return 0;
    REAL x = start - t + (REAL)1;
#ifdef REAL_IS_FLOAT
    REAL m = fmodf(x, period);
#else
    REAL m = fmod(x, period);
#endif
    if (m != 0 && ((m < 0) != (period < 0))) m += period; /* jnp.mod: sign of the divisor */
    return m < duration;
}

/* Advance (v,w,u) in place by nsteps Euler steps starting at counter t0.
 * params: 14 values in cardiax/params.py:4-18 order.  stim_fields: n_stim pointers to (H,W) arrays;
 * stim_proto: n_stim x 3 (start, duration, period).  Returns 0, or -1 on allocation failure. */
int FN(fk_oracle_forward_euler)(REAL *v, REAL *w, REAL *u, const REAL *D, long H, long W, const REAL *params,
                                const REAL *const *stim_fields, const REAL *stim_proto, int n_stim, double t0,
                                long nsteps, REAL dt, REAL dx, int tanh_mode) {
    const REAL tau_v_plus = params[0], tau_v1_minus = params[1], tau_v2_minus = params[2], tau_w_plus = params[3],
               tau_w_minus = params[4], tau_d = params[5], tau_0 = params[6], tau_r = params[7], tau_si = params[8],
               k = params[9], V_csi = params[10], V_c = params[11], V_v = params[12], Cm = params[13];
    const long Hp = H + 2, Wp = W + 2, NP = Hp * Wp;
    const REAL one = (REAL)1;
    REAL *up = malloc(sizeof(REAL) * NP), *Dp = malloc(sizeof(REAL) * NP), *Dx = malloc(sizeof(REAL) * NP),
         *Dy = malloc(sizeof(REAL) * NP), *ux = malloc(sizeof(REAL) * NP), *uy = malloc(sizeof(REAL) * NP),
         *un = malloc(sizeof(REAL) * H * W);
    if (!up || !Dp || !Dx || !Dy || !ux || !uy || !un) return -1;

    /* :32, :53-54 -- D is static: D_x, D_y once */
    FN(pad_edge)(D, Dp, H, W);
#pragma omp parallel for schedule(static)
    for (long r = 0; r < Hp; ++r)
        for (long c = 0; c < Wp; ++c) {
            Dx[r * Wp + c] = FN(grad1)(Dp + c, Hp, Wp, r) / dx;
            Dy[r * Wp + c] = FN(grad1)(Dp + r * Wp, Wp, 1, c) / dx;
        }

    for (long s = 0; s < nsteps; ++s) {
        REAL t = (REAL)(t0 + (double)s);
        int active[64];
        int any = 0;
        for (int i = 0; i < n_stim && i < 64; ++i) {
            active[i] = FN(fk_oracle_stim_active)(t, stim_proto[3 * i], stim_proto[3 * i + 1], stim_proto[3 * i + 2]);
            any |= active[i];
        }
        FN(pad_edge)(u, up, H, W); /* :31 */
        /* :49-50 u_x, u_y on the padded array */
#pragma omp parallel for schedule(static)
        for (long r = 0; r < Hp; ++r)
            for (long c = 0; c < Wp; ++c) {
                ux[r * Wp + c] = FN(grad1)(up + c, Hp, Wp, r) / dx;
                uy[r * Wp + c] = FN(grad1)(up + r * Wp, Wp, 1, c) / dx;
            }
#pragma omp parallel for schedule(static)
        for (long r = 1; r <= H; ++r)
            for (long c = 1; c <= W; ++c) {
                const long ip = r * Wp + c, ic = (r - 1) * W + (c - 1);
                const REAL uu = up[ip], vv = v[ic], ww = w[ic];
                /* :35-37 */
                const REAL p = uu >= V_c ? one : (REAL)0, q = uu >= V_v ? one : (REAL)0;
                const REAL tau_v_minus = (one - q) * tau_v1_minus + q * tau_v2_minus;
                /* :39-42 */
                const REAL j_fi = -vv * p * (uu - V_c) * (one - uu) / tau_d;
                const REAL j_so = (uu * (one - p) / tau_0) + (p / tau_r);
                const REAL j_si = -(ww * (one + FN(tanh_impl)(k * (uu - V_csi), tanh_mode))) / ((REAL)2 * tau_si);
                REAL j_ion = -(j_fi + j_so + j_si) / Cm;
                /* :45-46, :257-271 */
                if (any) {
                    REAL st = 0;
                    for (int i = 0; i < n_stim; ++i) {
                        REAL f = stim_fields[i][ic];
                        if (active[i] && f != 0) st = f;
                    }
                    if (st != 0) j_ion = st;
                }
                /* :51-55 */
                const REAL u_xx = FN(grad1)(ux + c, Hp, Wp, r) / dx;
                const REAL u_yy = FN(grad1)(uy + r * Wp, Wp, 1, c) / dx;
                const REAL del_u = (Dp[ip] * (u_xx + u_yy) + (Dx[ip] * ux[ip])) + (Dy[ip] * uy[ip]);
                /* :57-59 */
                const REAL d_v = ((one - p) * (one - vv) / tau_v_minus) - ((p * vv) / tau_v_plus);
                const REAL d_w = ((one - p) * (one - ww) / tau_w_minus) - ((p * ww) / tau_w_plus);
                const REAL d_u = del_u + j_ion;
                /* :70 */
                v[ic] = vv + d_v * dt;
                w[ic] = ww + d_w * dt;
                un[ic] = uu + d_u * dt;
            }
        memcpy(u, un, sizeof(REAL) * H * W);
    }
    free(up); free(Dp); free(Dx); free(Dy); free(ux); free(uy); free(un);
    return 0;
}

#undef FN
#undef CAT
#undef CAT_
