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

struct StreamGeom {  // per launch
    int NT;          // threads per CTA
    int CW;          // 4 * NT, columns a CTA reads
    int RS;          // ring row stride in floats (CW + 8: 4 pad floats each side)
    int RH;          // output rows per CTA
    int nstrips, nchunks;
    int cstride;     // CW - 8T: output columns per strip
    int uniformD;    // diffusivity (and so D_x, D_y) is one constant over the interior
};

template <int T>
struct StreamSmem {
    float* uring[T];  // [4][RS] last four input rows of stage s (level-s values)
    float* gyx[T];    // [2][RS] u_y rows of stage s, double buffered
    float* vring[T];  // [5][CW] level-s v waiting for stage s (s >= 1)
    float* wring[T];
};

FK_HD long long stream_smem_floats(int T, int NT) {
    const long long CW = 4LL * NT, RS = CW + 8;
    return (long long)T * 4 * RS + (long long)T * 2 * RS + (long long)(T - 1) * 2 * 5 * CW;
}

template <int T>
FK_HD void stream_carve(float* smem, const StreamGeom& G, StreamSmem<T>& S) {
    float* p = smem;
    for (int s = 0; s < T; ++s) { S.uring[s] = p; p += 4 * G.RS; }
    for (int s = 0; s < T; ++s) { S.gyx[s] = p; p += 2 * G.RS; }
    S.vring[0] = S.wring[0] = nullptr;
    for (int s = 1; s < T; ++s) { S.vring[s] = p; p += 5 * G.CW; S.wring[s] = p; p += 5 * G.CW; }
}

template <int T>
struct StreamState {     // registers of one thread
    float U[T][5][4];    // rows rho .. rho+4 of the stage's input level
    float GX[T][4][4];   // u_x rows rho-2 .. rho+1
    float gy[T][4];      // u_y of row rho (made one iteration ahead)
    float nu[4], nv[4], nw[4];  // level-0 rows prefetched for the next iteration
};

struct StreamCta {       // uniform per CTA
    int cs;              // first column of the strip
    int r0, r1;          // output rows
    int rin0, rin_end;   // level-0 rows read
    int out_c0, out_c1;  // output columns
    int c_end;           // end of the columns this CTA needs
    long long boff, boffD;
    float Dc, DXc, DYc;  // uniform-diffusivity constants
    unsigned mask[8];
    const StimDev* stims;
    int niter;
};

template <int T>
FK_HD void stream_cta_setup(const TileArgs& A, const StreamGeom& G, int strip, int chunk, int sim, StreamCta& C) {
    C.cs = strip * G.cstride;
    C.r0 = 4 * T + chunk * G.RH;
    C.r1 = C.r0 + G.RH < A.H - 4 * T ? C.r0 + G.RH : A.H - 4 * T;
    C.rin0 = C.r0 - 4 * T;
    C.rin_end = C.r1 + 4 * T;
    C.out_c0 = C.cs + 4 * T;
    C.out_c1 = C.out_c0 + G.cstride < A.W - 4 * T ? C.out_c0 + G.cstride : A.W - 4 * T;
    C.c_end = C.out_c1 + 4 * T;  // columns past this are nobody's input
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
    C.Dc = C.DXc = C.DYc = 0.0f;
    if (G.uniformD) {
        const long long g = C.boffD + (long long)(4 * T) * A.W + 4 * T;
        C.Dc = A.D[g]; C.DXc = A.DX[g]; C.DYc = A.DY[g];
    }
}

// prefetch the level-0 rows iteration `i` will consume
template <int T>
FK_HD void stream_prefetch(const TileArgs& A, const StreamCta& C, StreamState<T>& R, int i, int c, bool act) {
    const int n = C.rin0 + i;  // u row pushed at iteration i
    if (act && n < C.rin_end) unpack4(ld4(A.u_in + C.boff + (long long)n * A.W + c), R.nu);
    const int rho = n - 4;     // v, w row stage 0 emits at iteration i
    if (act && rho >= C.r0 - 4 * (T - 1) && rho < C.r1 + 4 * (T - 1)) {
        unpack4(ld4(A.v_in + C.boff + (long long)rho * A.W + c), R.nv);
        unpack4(ld4(A.w_in + C.boff + (long long)rho * A.W + c), R.nw);
    }
}

// one row iteration of one thread.  `tid` in [0, NT), c = first of its 4 columns.
template <bool EXACT, int T>
FK_HD void stream_iter(const TileArgs& A, const StreamGeom& G, const StreamCta& C, const StreamSmem<T>& S,
                       StreamState<T>& R, int i, int tid) {
    const int c = C.cs + 4 * tid;
    const bool act = c < C.c_end;
    const int own = 4 + 4 * tid;  // offset of the thread's columns inside a ring row
    float in_u[4], in_v[4], in_w[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { in_u[k] = R.nu[k]; in_v[k] = R.nv[k]; in_w[k] = R.nw[k]; }
    stream_prefetch<T>(A, C, R, i + 1, c, act);
    if (!act) return;
    bool have_in = (C.rin0 + i) < C.rin_end;
#pragma unroll
    for (int s = 0; s < T; ++s) {
        const int rho = C.rin0 + i - 4 * (s + 1);  // row this stage emits (level s+1)
        const int n = rho + 4;                     // newest input row (level s)
        const int lo = C.r0 - 4 * (T - 1 - s), hi = C.r1 + 4 * (T - 1 - s);
        if (have_in) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int k = 0; k < 4; ++k) R.U[s][j][k] = R.U[s][j + 1][k];
#pragma unroll
            for (int k = 0; k < 4; ++k) R.U[s][4][k] = in_u[k];
            st4(S.uring[s] + (n & 3) * G.RS + own, in_u);
        }
        // u_x of row rho+2 (solve.py:49)
        float ngx[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) ngx[k] = dcen<EXACT>(A.K, R.U[s][0][k], R.U[s][1][k], R.U[s][3][k], R.U[s][4][k]);
        // u_y of row rho+1 (solve.py:50), published for the neighbours
        float ngy[4] = {0.f, 0.f, 0.f, 0.f};
        if (rho + 1 >= lo && rho + 1 < hi) {
            const float* ur = S.uring[s] + ((rho + 1) & 3) * G.RS + own;
            const F2 L = ld2(ur - 2), Rr = ld2(ur + 4);
            const float e[8] = {L.x, L.y, R.U[s][1][0], R.U[s][1][1], R.U[s][1][2], R.U[s][1][3], Rr.x, Rr.y};
#pragma unroll
            for (int k = 0; k < 4; ++k) ngy[k] = dcen<EXACT>(A.K, e[k], e[k + 1], e[k + 3], e[k + 4]);
            st4(S.gyx[s] + (i & 1) * G.RS + own, ngy);
        }
        const bool emit = rho >= lo && rho < hi;
        if (emit) {
            const float* gr = S.gyx[s] + ((i + 1) & 1) * G.RS + own;
            const F2 L = ld2(gr - 2), Rr = ld2(gr + 4);
            const float g[8] = {L.x, L.y, R.gy[s][0], R.gy[s][1], R.gy[s][2], R.gy[s][3], Rr.x, Rr.y};
            float v[4], w[4];
            if (s == 0) {
#pragma unroll
                for (int k = 0; k < 4; ++k) { v[k] = in_v[k]; w[k] = in_w[k]; }
            } else {
                unpack4(ld4(S.vring[s] + (rho % 5) * G.CW + 4 * tid), v);
                unpack4(ld4(S.wring[s] + (rho % 5) * G.CW + 4 * tid), w);
            }
            float Dv[4], DXv[4], DYv[4];
            const long long grow = (long long)rho * A.W + c;
            if (G.uniformD) {
#pragma unroll
                for (int k = 0; k < 4; ++k) { Dv[k] = C.Dc; DXv[k] = C.DXc; DYv[k] = C.DYc; }
            } else {
                unpack4(ld4(A.D + C.boffD + grow), Dv);
                unpack4(ld4(A.DX + C.boffD + grow), DXv);
                unpack4(ld4(A.DY + C.boffD + grow), DYv);
            }
            float stim[4] = {0.f, 0.f, 0.f, 0.f};
            const unsigned mask = C.mask[s];
            if (mask) {  // solve.py:260-269
                for (int q = 0; q < A.n_stim; ++q)
                    if (mask >> q & 1u) {
                        float f[4];
                        unpack4(ld4(C.stims[q].field + grow), f);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            if (f[k] != 0.0f) stim[k] = f[k];
                    }
            }
            float un[4], vn[4], wn[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float u_xx = dcen<EXACT>(A.K, R.GX[s][0][k], R.GX[s][1][k], R.GX[s][3][k], ngx[k]);  // :51
                const float u_yy = dcen<EXACT>(A.K, g[k], g[k + 1], g[k + 3], g[k + 4]);                   // :52
                const float del_u = diffusion<EXACT>(Dv[k], DXv[k], DYv[k], R.GX[s][2][k], R.gy[s][k], u_xx, u_yy);
                float d_v, d_w, d_u;
                cell_rhs<EXACT>(A.K, R.U[s][0][k], v[k], w[k], del_u, stim[k], d_v, d_w, d_u);
                vn[k] = euler<EXACT>(v[k], d_v, A.K.dt);
                wn[k] = euler<EXACT>(w[k], d_w, A.K.dt);
                un[k] = euler<EXACT>(R.U[s][0][k], d_u, A.K.dt);
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
                st4(S.vring[s + 1] + (rho % 5) * G.CW + 4 * tid, vn);
                st4(S.wring[s + 1] + (rho % 5) * G.CW + 4 * tid, wn);
            }
        }
        // slide the u_x window, keep u_y of the next row
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            R.GX[s][0][k] = R.GX[s][1][k];
            R.GX[s][1][k] = R.GX[s][2][k];
            R.GX[s][2][k] = R.GX[s][3][k];
            R.GX[s][3][k] = ngx[k];
            R.gy[s][k] = ngy[k];
        }
        have_in = emit;
    }
}

// ---------------------------------------------------------------- host-side planning (no CUDA calls)
struct StreamPlan {
    int T;
    StreamGeom G;
    long long smem_bytes;
};

// Applicability + geometry.  cta_threads / rows_per_cta: 0 = choose.
inline bool plan_stream(int H, int W, int batch, int T, int cta_threads, int rows_per_cta, int num_sms, int uniformD,
                        StreamPlan& P) {
    if (T < 1 || T > 4) return false;
    if (W % 4 != 0) return false;               // float4 rows
    if (H < 8 * T + 8 || W < 8 * T + 32) return false;  // too small: the general tile kernel does it all
    const int Wint = W - 8 * T, Hint = H - 8 * T;
    int NT = cta_threads;
    if (NT <= 0) {
        // widest CTA that the interior can fill, 64..256 threads
        NT = 256;
        while (NT > 64 && (NT / 2) * 4 - 8 * T >= Wint) NT /= 2;
    }
    if (NT % 32 != 0 || NT < 32 || NT > 1024) return false;
    StreamGeom& G = P.G;
    G.NT = NT;
    G.CW = 4 * NT;
    G.RS = G.CW + 8;
    if (G.CW - 8 * T < 4) return false;
    // balanced strips: the fewest strips that cover the interior, equal widths (multiple of 4)
    const int maxstride = G.CW - 8 * T;
    G.nstrips = (Wint + maxstride - 1) / maxstride;
    int stride = (Wint + G.nstrips - 1) / G.nstrips;
    stride = (stride + 3) / 4 * 4;
    G.cstride = stride;
    int RH = rows_per_cta;
    if (RH <= 0) {
        // about one CTA per SM-slot; at least 32 rows so the 8T-row pipeline fill stays small
        const long long slots = 2LL * num_sms;
        long long per = (slots + (long long)G.nstrips * batch - 1) / ((long long)G.nstrips * batch);
        if (per < 1) per = 1;
        RH = (int)((Hint + per - 1) / per);
        if (RH < 32) RH = 32;
    }
    if (RH > Hint) RH = Hint;
    G.RH = RH;
    G.nchunks = (Hint + RH - 1) / RH;
    G.uniformD = uniformD;
    P.T = T;
    P.smem_bytes = stream_smem_floats(T, NT) * 4;
    return P.smem_bytes <= 227 * 1024;
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
    for (int tid = 0; tid < G.NT; ++tid) stream_prefetch<T>(A, C, R[tid], 0, C.cs + 4 * tid, C.cs + 4 * tid < C.c_end);
    for (int i = 0; i < C.niter; ++i) {
        if (!reverse)
            for (int tid = 0; tid < G.NT; ++tid) stream_iter<EXACT, T>(A, G, C, S, R[tid], i, tid);
        else
            for (int tid = G.NT - 1; tid >= 0; --tid) stream_iter<EXACT, T>(A, G, C, S, R[tid], i, tid);
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
