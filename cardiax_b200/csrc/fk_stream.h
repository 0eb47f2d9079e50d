// fk_stream.h -- the STREAMING multi-step kernel body (interior of the tissue).
//
// A CTA owns a strip of CW = 4 * NT columns and a chunk of rows and marches down the rows once.
// T Euler steps ("stages") are pipelined behind each other, stage s+1 lagging stage s by 4 rows
// (the stencil radius), so u, v, w are read from HBM once and written once per T steps:
//
//   row n of level 0  --stage 0-->  row n-4 of level 1  --stage 1-->  row n-8 of level 2 ...
//
// Per thread (4 adjacent columns, float4) and per stage the vertical part of the reference's
// two-pass derivative lives in REGISTERS: a 5-row window of u and a 4-row window of u_x
// (solve.py:49,51: u_x is rounded to fp32, then differentiated again).  The horizontal part
// needs the neighbours' u and u_y (solve.py:50,52): each stage keeps its last 4 input rows and
// its last 2 u_y rows in shared-memory rings; v and w (pointwise) wait 4 rows in a private ring.
// One block barrier per row iteration orders all ring traffic.
//
// Only cells whose whole dependency cone is the 4th-order central formula are produced here:
// rows/cols [4T, N-4T).  The frame of 4T cells is done by the general tile kernel (fk_tile.h).
// Strip edges: columns within 4s of a strip edge are garbage at level s and are never stored.
//
// Written as FK_HD code over an explicit per-thread state so tests/emu can run it on the CPU
// (threads of a block executed one after the other inside each iteration).
#pragma once
#include "fk_core.h"
#include "fk_tile.h"

namespace fk {

struct alignas(8) F2 { float x, y; };
struct alignas(16) F4 { float x, y, z, w; };
FK_HD F2 ld2(const float* p) { return *reinterpret_cast<const F2*>(p); }
FK_HD F4 ld4(const float* p) { return *reinterpret_cast<const F4*>(p); }
FK_HD void st4(float* p, const float* v) {
    F4 t; t.x = v[0]; t.y = v[1]; t.z = v[2]; t.w = v[3];
    *reinterpret_cast<F4*>(p) = t;
}
FK_HD void unpack4(const F4& t, float* v) { v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }

// 16-byte asynchronous global -> shared copy (cp.async / LDGSTS on the device; the CPU emulation
// copies immediately, which is the earliest legal completion).
FK_HD void async_copy16(float* sdst, const float* gsrc) {
#if defined(__CUDA_ARCH__)
    const unsigned sa = (unsigned)__cvta_generic_to_shared(sdst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gsrc) : "memory");
#else
    *reinterpret_cast<F4*>(sdst) = *reinterpret_cast<const F4*>(gsrc);
#endif
}
FK_HD void async_commit() {
#if defined(__CUDA_ARCH__)
    asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
template <int N>
FK_HD void async_wait() {
#if defined(__CUDA_ARCH__)
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
#endif
}

enum { FK_PF = 2 };      // level-0 rows are fetched this many iterations ahead
enum { FK_U0DEP = 8 };   // stage-0 u ring: 5 window rows + FK_PF in flight, rounded to a power of two
enum { FK_VWDEP = 4 };   // level-0 v, w staging ring (>= FK_PF + 1)

struct StreamGeom {  // per launch
    int NT;          // threads per CTA
    int CW;          // 4 * NT, columns a CTA reads
    int RS;          // ring row stride in floats (CW + 8: 4 pad floats each side)
    int RH;          // output rows per CTA
    int nstrips, nchunks;
    int cstride;     // output columns per strip (<= CW - 8T)
    int uniformD;    // diffusivity (and so D_x, D_y) is one constant over the interior
    int row0, row1;  // output rows of this launch (every one needs 4T rows of input above and below)
};

template <int T>
struct StreamSmem {
    float* uring[T];  // stage 0: [FK_U0DEP][RS] level-0 rows (cp.async target); stage s >= 1: [5][RS] level-s rows
    float* gyx[T];    // [2][RS] u_y rows of stage s, double buffered
    float* vring[T];  // stage 0: [FK_VWDEP][CW] level-0 v (cp.async target); s >= 1: [5][CW] level-s v waiting 4 rows
    float* wring[T];
};

FK_HD long long stream_smem_floats(int T, int NT) {
    const long long CW = 4LL * NT, RS = CW + 8;
    return (long long)(FK_U0DEP + 5 * (T - 1)) * RS + (long long)T * 2 * RS + 2LL * FK_VWDEP * CW +
           (long long)(T - 1) * 2 * 5 * CW;
}

template <int T>
FK_HD void stream_carve(float* smem, const StreamGeom& G, StreamSmem<T>& S) {
    float* p = smem;
    for (int s = 0; s < T; ++s) { S.uring[s] = p; p += (s == 0 ? (int)FK_U0DEP : 5) * G.RS; }
    for (int s = 0; s < T; ++s) { S.gyx[s] = p; p += 2 * G.RS; }
    for (int s = 0; s < T; ++s) {
        const int dep = s == 0 ? (int)FK_VWDEP : 5;
        S.vring[s] = p; p += dep * G.CW;
        S.wring[s] = p; p += dep * G.CW;
    }
}

template <int T>
struct StreamState {     // registers of one thread
    float GX[T][4][4];   // u_x rows rho-2 .. rho+1 of each stage
    float gy[T][4];      // u_y of row rho (made one iteration ahead)
    float gypad[T];      // edge threads only: u_y of the PAD column next to the tissue edge (padded index 0 / W+1)
};

struct StreamCta {       // uniform per CTA
    int cs;              // first column of the strip
    int r0, r1;          // output rows
    int rin0, rin_end;   // level-0 rows read
    int out_c0, out_c1;  // output columns
    int c_end;           // end of the columns this CTA needs
    int edgeL, edgeR;    // thread that holds the tissue's first / last column (-1: not in this strip)
    long long boff, boffD;
    float Dc, DXc, DYc;  // uniform-diffusivity constants
    float DYcL, DYcR;    // D_y of a constant map in the tissue's first / last column (one-sided formula: not DYc)
    unsigned mask[8];
    const StimDev* stims;
    int niter;
};

template <int T>
FK_HD void stream_cta_setup(const TileArgs& A, const StreamGeom& G, int strip, int chunk, int sim, StreamCta& C) {
    // strip j owns output columns [j stride, (j+1) stride); it reads 4T more on each side unless that side is the
    // tissue's physical left/right edge, which the edge thread handles with the reference's one-sided formulas
    C.cs = strip * G.cstride - 4 * T < 0 ? 0 : strip * G.cstride - 4 * T;
    C.r0 = G.row0 + chunk * G.RH;
    C.r1 = C.r0 + G.RH < G.row1 ? C.r0 + G.RH : G.row1;
    C.rin0 = C.r0 - 4 * T;
    C.rin_end = C.r1 + 4 * T;
    C.out_c0 = strip * G.cstride;
    C.out_c1 = C.out_c0 + G.cstride < A.W ? C.out_c0 + G.cstride : A.W;
    C.c_end = C.out_c1 + 4 * T < A.W ? C.out_c1 + 4 * T : A.W;  // columns past this are nobody's input
    C.edgeL = strip == 0 ? 0 : -1;
    C.edgeR = C.out_c1 == A.W ? (A.W - 4 - C.cs) / 4 : -1;
    C.boff = (long long)sim * A.plane;
    C.boffD = (long long)sim * A.plane_D;
    C.stims = A.stims ? A.stims + (long long)sim * A.n_stim : nullptr;
    C.niter = (C.r1 - C.r0) + 8 * T;
    for (int s = 0; s < 8; ++s) C.mask[s] = 0;
    for (int s = 0; s < T; ++s) {
        const float t = (float)(A.t0 + (double)s);
        unsigned m = 0;
        for (int i = 0; i < A.n_stim; ++i) {
            const StimDev sd = C.stims[i];
            if (sd.field && stim_active(t, sd.start, sd.duration, sd.period)) m |= 1u << i;
        }
        C.mask[s] = m;
    }
    C.Dc = C.DXc = C.DYc = C.DYcL = C.DYcR = 0.0f;
    if (G.uniformD) {
        const long long g = C.boffD + (long long)G.row0 * A.W + 4 * T;  // any cell 2+ away from every edge
        C.Dc = A.D[g]; C.DXc = A.DX[g]; C.DYc = A.DY[g];
        C.DYcL = A.DY[g - 4 * T];
        C.DYcR = A.DY[g - 4 * T + A.W - 1];
    }
}

// start the asynchronous fetch of the level-0 rows iteration `i` will consume; always commits a group
template <int T>
FK_HD void stream_prefetch(const TileArgs& A, const StreamGeom& G, const StreamCta& C, const StreamSmem<T>& S, int i,
                           int tid, bool act) {
    const int n = C.rin0 + i;  // u row pushed at iteration i
    const int c = C.cs + 4 * tid;
    if (act && n < C.rin_end)
        async_copy16(S.uring[0] + (n & (FK_U0DEP - 1)) * G.RS + 4 + 4 * tid, A.u_in + C.boff + (long long)n * A.W + c);
    const int rho = n - 4;     // v, w row stage 0 emits at iteration i
    if (act && rho >= C.r0 - 4 * (T - 1) && rho < C.r1 + 4 * (T - 1)) {
        const long long g = C.boff + (long long)rho * A.W + c;
        async_copy16(S.vring[0] + (rho & (FK_VWDEP - 1)) * G.CW + 4 * tid, A.v_in + g);
        async_copy16(S.wring[0] + (rho & (FK_VWDEP - 1)) * G.CW + 4 * tid, A.w_in + g);
    }
    async_commit();
}

FK_HD int mod5(int x) { return x >= 5 ? x - 5 : x; }  // for 0 <= x < 10

// one-sided first derivative / dx (solve.py:232-235, 246-249) on four consecutive values
template <bool EXACT>
FK_HD float edge_deriv(const Consts& K, int kind, float a0, float a1, float a2, float a3) {
    float k0, k1, k2, k3;
    int o0, o1, o2, o3;
    kind_coeffs(kind, k0, k1, k2, k3, o0, o1, o2, o3);
    return deriv<EXACT>(K, kind, k0, k1, k2, k3, a0, a1, a2, a3);
}

// may iterations i >= 8T use the condition-free body?  (no stimulus active in any level of this launch)
template <int T>
FK_HD bool stream_steady_ok(const StreamCta& C) {
    unsigned m = 0;
    for (int s = 0; s < T; ++s) m |= C.mask[s];
    return m == 0;
}

// second derivatives, reaction, stimulus and Euler update of one row of one stage (4 cells)
template <bool EXACT, bool HAS_STIM>
FK_HD void stream_emit(const Consts& K, const float* u0, const float* v, const float* w, const float* gxm2,
                       const float* gxm1, const float* gx0, const float* gxp1, const float* gxp2, const float* g,
                       const float* gy0, const float* Dv, const float* DXv, const float* DYv, const float* stim,
                       bool edgeL, bool edgeR, float* un, float* vn, float* wn) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float u_xx = dcen<EXACT>(K, gxm2[k], gxm1[k], gxp1[k], gxp2[k]);    // solve.py:51
        float u_yy = dcen<EXACT>(K, g[k], g[k + 1], g[k + 3], g[k + 4]);          // solve.py:52
        // the tissue's first / last column: forward / backward formula on u_y of padded columns 1..4 / W-3..W
        if (k == 0 && edgeL) u_yy = edge_deriv<EXACT>(K, FWD, g[2], g[3], g[4], g[5]);
        if (k == 3 && edgeR) u_yy = edge_deriv<EXACT>(K, BWD, g[2], g[3], g[4], g[5]);
        const float del_u = diffusion<EXACT>(Dv[k], DXv[k], DYv[k], gx0[k], gy0[k], u_xx, u_yy);
        float d_v, d_w, d_u;
        cell_rhs<EXACT, HAS_STIM>(K, u0[k], v[k], w[k], del_u, HAS_STIM ? stim[k] : 0.0f, d_v, d_w, d_u);
        vn[k] = euler<EXACT>(v[k], d_v, K.dt);
        wn[k] = euler<EXACT>(w[k], d_w, K.dt);
        un[k] = euler<EXACT>(u0[k], d_u, K.dt);
    }
}

// one row iteration of one thread.  `tid` in [0, NT), its 4 columns start at C.cs + 4 tid.
// STEADY: the caller guarantees that every stage has input and emits in this iteration (i >= 8T) and that no
// stimulus is active in this launch, so all the per-stage conditions fold away.
template <bool EXACT, int T, bool STEADY>
FK_HD void stream_iter(const TileArgs& A, const StreamGeom& G, const StreamCta& C, const StreamSmem<T>& S,
                       StreamState<T>& R, int i, int tid) {
    const int c = C.cs + 4 * tid;
    const bool act = c < C.c_end;
    // rows fetched FK_PF iterations ago have landed (this thread's own columns; the neighbours'
    // columns of a row are only read three barriers later)
    stream_prefetch<T>(A, G, C, S, i + FK_PF, tid, act);
    async_wait<FK_PF>();
    if (!act) return;
    const int own = 4 + 4 * tid;  // offset of the thread's columns inside a padded ring row
    const bool edgeL = tid == C.edgeL, edgeR = tid == C.edgeR;
    const int n0 = C.rin0 + i;
    const int m8 = n0 & (FK_U0DEP - 1);
    const int m5 = n0 % 5;
    float in_u[4] = {0.f, 0.f, 0.f, 0.f};
    bool have_in = STEADY || n0 < C.rin_end;
#pragma unroll
    for (int s = 0; s < T; ++s) {
        const int rho = n0 - 4 * (s + 1);  // row this stage emits (level s+1); its newest input row is rho+4
        const int lo = C.r0 - 4 * (T - 1 - s), hi = C.r1 + 4 * (T - 1 - s);
        // ring slots of input rows rho+4 (new), rho+3, rho+1, rho
        int sl_new, sl_r3, sl_r1, sl_r0;
        if (s == 0) {
            sl_new = m8; sl_r3 = (m8 + 7) & 7; sl_r1 = (m8 + 5) & 7; sl_r0 = (m8 + 4) & 7;
        } else {
            sl_new = mod5(m5 + (s % 5)); sl_r3 = mod5(sl_new + 4); sl_r1 = mod5(sl_new + 2); sl_r0 = mod5(sl_new + 1);
        }
        float* ur = S.uring[s] + own;
        if (s == 0) {
            if (have_in) unpack4(ld4(ur + sl_new * G.RS), in_u);
        } else if (have_in) {
            st4(ur + sl_new * G.RS, in_u);
        }
        float u0[4], u1[4], u3[4];
        unpack4(ld4(ur + sl_r0 * G.RS), u0);
        unpack4(ld4(ur + sl_r1 * G.RS), u1);
        unpack4(ld4(ur + sl_r3 * G.RS), u3);
        // u_x of row rho+2 (solve.py:49)
        float ngx[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) ngx[k] = dcen<EXACT>(A.K, u0[k], u1[k], u3[k], in_u[k]);
        const bool emit = STEADY || (rho >= lo && rho < hi);
        if (emit) {
            const float* gr = S.gyx[s] + ((i + 1) & 1) * G.RS + own;
            const F2 L = ld2(gr - 2), Rr = ld2(gr + 4);
            // at a tissue edge the neighbour is the pad column, whose u_y this thread made itself
            const float g[8] = {L.x, edgeL ? R.gypad[s] : L.y, R.gy[s][0], R.gy[s][1], R.gy[s][2], R.gy[s][3],
                                edgeR ? R.gypad[s] : Rr.x, Rr.y};
            float v[4], w[4];
            const int vslot = s == 0 ? (rho & (FK_VWDEP - 1)) : sl_r0;
            unpack4(ld4(S.vring[s] + vslot * G.CW + 4 * tid), v);
            unpack4(ld4(S.wring[s] + vslot * G.CW + 4 * tid), w);
            float Dv[4], DXv[4], DYv[4];
            const long long grow = (long long)rho * A.W + c;
            if (G.uniformD) {
#pragma unroll
                for (int k = 0; k < 4; ++k) { Dv[k] = C.Dc; DXv[k] = C.DXc; DYv[k] = C.DYc; }
                if (edgeL) DYv[0] = C.DYcL;
                if (edgeR) DYv[3] = C.DYcR;
            } else {
                unpack4(ld4(A.D + C.boffD + grow), Dv);
                unpack4(ld4(A.DX + C.boffD + grow), DXv);
                unpack4(ld4(A.DY + C.boffD + grow), DYv);
            }
            float un[4], vn[4], wn[4];
            const unsigned mask = STEADY ? 0u : C.mask[s];
            if (mask) {  // solve.py:260-269: later stimuli win, zero cells never stimulate
                float stim[4] = {0.f, 0.f, 0.f, 0.f};
                for (int q = 0; q < A.n_stim; ++q)
                    if (mask >> q & 1u) {
                        float f[4];
                        unpack4(ld4(C.stims[q].field + grow), f);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            if (f[k] != 0.0f) stim[k] = f[k];
                    }
                stream_emit<EXACT, true>(A.K, u0, v, w, R.GX[s][0], R.GX[s][1], R.GX[s][2], R.GX[s][3], ngx, g, R.gy[s], Dv,
                                         DXv, DYv, stim, edgeL, edgeR, un, vn, wn);
            } else {
                stream_emit<EXACT, false>(A.K, u0, v, w, R.GX[s][0], R.GX[s][1], R.GX[s][2], R.GX[s][3], ngx, g, R.gy[s],
                                          Dv, DXv, DYv, nullptr, edgeL, edgeR, un, vn, wn);
            }
            if (s == T - 1) {
                if (c >= C.out_c0 && c < C.out_c1) {
                    st4(A.u_out + C.boff + grow, un);
                    st4(A.v_out + C.boff + grow, vn);
                    st4(A.w_out + C.boff + grow, wn);
                }
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) in_u[k] = un[k];
                // level s+1 row rho == newest input row of stage s+1: same mod-5 phase as its u ring
                const int ns = mod5(m5 + ((s + 1) % 5));
                st4(S.vring[s + 1] + ns * G.CW + 4 * tid, vn);
                st4(S.wring[s + 1] + ns * G.CW + 4 * tid, wn);
            }
        }
        // u_y of row rho+1 (solve.py:50) for the next iteration, published for the neighbours
        if (STEADY || (rho + 1 >= lo && rho + 1 < hi)) {
            const F2 L = ld2(ur + sl_r1 * G.RS - 2), Rr = ld2(ur + sl_r1 * G.RS + 4);
            // edge-pad (solve.py:31): the column outside the tissue repeats the edge column
            const float e[8] = {L.x, edgeL ? u1[0] : L.y, u1[0], u1[1], u1[2], u1[3], edgeR ? u1[3] : Rr.x, Rr.y};
#pragma unroll
            for (int k = 0; k < 4; ++k) R.gy[s][k] = dcen<EXACT>(A.K, e[k], e[k + 1], e[k + 3], e[k + 4]);
            if (edgeL) {  // padded columns 0 (the pad) and 1 (tissue column 0) use the forward formula
                R.gypad[s] = edge_deriv<EXACT>(A.K, FWD, e[1], e[2], e[3], e[4]);
                R.gy[s][0] = edge_deriv<EXACT>(A.K, FWD, e[2], e[3], e[4], e[5]);
            }
            if (edgeR) {  // padded columns W (tissue column W-1) and W+1 (the pad) use the backward formula
                R.gy[s][3] = edge_deriv<EXACT>(A.K, BWD, e[2], e[3], e[4], e[5]);
                R.gypad[s] = edge_deriv<EXACT>(A.K, BWD, e[3], e[4], e[5], e[6]);
            }
            st4(S.gyx[s] + (i & 1) * G.RS + own, R.gy[s]);
        }
        // slide the u_x window
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            R.GX[s][0][k] = R.GX[s][1][k];
            R.GX[s][1][k] = R.GX[s][2][k];
            R.GX[s][2][k] = R.GX[s][3][k];
            R.GX[s][3][k] = ngx[k];
        }
        have_in = emit;
    }
}

// ---------------------------------------------------------------- host-side planning (no CUDA calls)
struct StreamPlan {
    int T;
    StreamGeom G;
    long long smem_bytes;
    int occ;      // resident CTAs per SM the plan assumed
    double cost;  // modelled time of one launch (arbitrary units; compare plans of the same problem only)
};

// Applicability + geometry.  cta_threads / rows_per_cta: 0 = choose.
// occ(NT, smem_bytes) -> resident CTAs per SM for that configuration (0 = cannot launch).
//
// Strips are balanced (equal widths, multiple of 4 columns) and the row chunks are sized so that the
// grid is a whole number of waves of num_sms * occ CTAs; candidates are ranked by a simple model:
// rounds * iterations-per-CTA * resident warps / issue-efficiency(resident warps).
template <class OccFn>
inline bool plan_stream(int row0, int row1, int W, int batch, int T, int cta_threads, int rows_per_cta, int num_sms,
                        int uniformD, int max_threads, OccFn occ, StreamPlan& P) {
    if (T < 1 || T > 4) return false;
    if (W % 4 != 0) return false;                       // float4 rows
    if (row1 - row0 < 8 || W < 8 * T + 32) return false;  // too small: the general tile kernel does it all
    const int Wint = W, Hint = row1 - row0;
    double best = -1.0;
    const int ns_min = (Wint + (4 * max_threads - 8 * T) - 1) / (4 * max_threads - 8 * T);
    for (int ns = ns_min; ns < ns_min + 24; ++ns) {
        int stride = (Wint + ns - 1) / ns;
        stride = (stride + 3) / 4 * 4;
        const int need = stride + 8 * T;
        int NT = (need + 127) / 128 * 32;
        if (cta_threads > 0) {
            if (NT > cta_threads) continue;
            NT = cta_threads;
        }
        if (NT > max_threads || NT < 32) continue;
        const int nstrips = (Wint + stride - 1) / stride;
        const long long smem = stream_smem_floats(T, NT) * 4;
        if (smem > 227 * 1024) continue;
        const int o = occ(NT, smem);
        if (o < 1) continue;
        const long long slots = (long long)num_sms * o, units = (long long)nstrips * batch;
        for (int waves = 1; waves <= 3; ++waves) {
            int RH = rows_per_cta;
            if (RH <= 0) {
                long long nch = waves * slots / units;
                if (nch < 1) nch = 1;
                RH = (int)((Hint + nch - 1) / nch);
                if (RH < 8) RH = 8;
            }
            if (RH > Hint) RH = Hint;
            const int nchunks = (Hint + RH - 1) / RH;
            const long long ncta = units * nchunks;
            const double rounds = (double)((ncta + slots - 1) / slots);
            // CTAs actually resident on an SM (a small tissue does not fill the machine), their active warps, and
            // the time of one row iteration: T stages, each at least a lone warp's dependent chain (~2 units) and
            // growing with the warps that share the four schedulers
            long long res = (ncta + num_sms - 1) / num_sms;
            if (res > o) res = o;
            const double warps = (double)((need + 127) / 128) * (double)res;
            const double cost = rounds * (RH + 8.0 * T) * T * (warps + 8.0) / 4.0;
            if (best < 0 || cost < best) {
                best = cost;
                P.G.NT = NT; P.G.CW = 4 * NT; P.G.RS = 4 * NT + 8; P.G.RH = RH;
                P.G.nstrips = nstrips; P.G.nchunks = nchunks; P.G.cstride = stride; P.G.uniformD = uniformD;
                P.G.row0 = row0; P.G.row1 = row1;
                P.T = T;
                P.smem_bytes = smem;
                P.occ = o;
                P.cost = cost;
            }
            if (rows_per_cta > 0) break;
        }
        if (cta_threads > 0) break;
    }
    return best >= 0;
}

// ---------------------------------------------------------------- CPU emulation of one launch (tests only)
#if !defined(__CUDACC__)
}  // namespace fk
#include <vector>
namespace fk {
template <bool EXACT, int T>
inline void emu_stream_cta(const TileArgs& A, const StreamGeom& G, int strip, int chunk, int sim, bool reverse) {
    std::vector<float> smem((size_t)stream_smem_floats(T, G.NT), __builtin_nanf(""));
    StreamSmem<T> S;
    stream_carve<T>(smem.data(), G, S);
    StreamCta C;
    stream_cta_setup<T>(A, G, strip, chunk, sim, C);
    std::vector<StreamState<T>> R((size_t)G.NT);
    for (auto& r : R) {
        float* f = reinterpret_cast<float*>(&r);
        for (size_t q = 0; q < sizeof(r) / sizeof(float); ++q) f[q] = __builtin_nanf("");
    }
    for (int tid = 0; tid < G.NT; ++tid)
        for (int j = 0; j < FK_PF; ++j) stream_prefetch<T>(A, G, C, S, j, tid, C.cs + 4 * tid < C.c_end);
    const bool steady_ok = stream_steady_ok<T>(C);
    for (int i = 0; i < C.niter; ++i) {
        const bool steady = steady_ok && i >= 8 * T;
        for (int q = 0; q < G.NT; ++q) {
            const int tid = reverse ? G.NT - 1 - q : q;
            if (steady) stream_iter<EXACT, T, true>(A, G, C, S, R[tid], i, tid);
            else stream_iter<EXACT, T, false>(A, G, C, S, R[tid], i, tid);
        }
    }
}

inline int emu_stream_launch(const StreamPlan& P, const TileArgs& A, int batch, int exact, int reverse) {
    for (int sim = 0; sim < batch; ++sim)
        for (int chunk = 0; chunk < P.G.nchunks; ++chunk)
            for (int strip = 0; strip < P.G.nstrips; ++strip) {
#define FK_EMU_CASE(TT)                                                                      \
    case TT:                                                                                 \
        if (exact) emu_stream_cta<true, TT>(A, P.G, strip, chunk, sim, reverse != 0);        \
        else emu_stream_cta<false, TT>(A, P.G, strip, chunk, sim, reverse != 0);             \
        break;
                switch (P.T) {
                    FK_EMU_CASE(1) FK_EMU_CASE(2) FK_EMU_CASE(3) FK_EMU_CASE(4)
                    default: return -5;
                }
#undef FK_EMU_CASE
            }
    return 0;
}
#endif

}  // namespace fk
