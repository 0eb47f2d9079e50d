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
// Per step a CTA (1) computes its RING groups -- the cells within 4 of a tile edge -- and publishes their new u to the
// neighbours, (2) computes its INTERIOR groups while those stores travel, (3) copies the neighbours' ring cells into
// the halo of the next u buffer, (4) one block barrier.  Only the four edge neighbours are involved: the stencil is
// plus-shaped (no cross terms), so corners are never read.
//
// The exchange needs no fence and no flag: a published cell is ONE 8-byte record {fp32 value, step tag} written with a
// single relaxed gpu-scope store (8-byte accesses are single-copy atomic), into a per-tile mailbox in global memory
// (L2 resident); the consumer re-reads a record until its tag is the step it waits for (the "LL" protocol of NCCL).
// Mailboxes alternate by step parity: a CTA can only overwrite the records of step s-1 with those of step s+1 after
// it received its neighbours' step-s records, which they computed from its step s-1 ones.
//
// A thread computes groups of NC = 1, 2 or 4 adjacent cells of one row (the planner takes the smallest NC that keeps
// the groups of a tile within one round of 512 threads: a small tile is latency bound and wants every lane busy).
// The vertical and horizontal two-pass derivatives are rebuilt in registers from the shared-memory window like the
// low-latency wide kernel does (fk_wide.h); a group next to a physical edge evaluates the reference's one-sided
// formulas on an index-clamped window (= the reference's edge padding, solve.py:29-31).
// Same FK_HD source for the CPU emulation (tests/emu), which runs the CTAs phase by phase.
#pragma once
#include "fk_core.h"
#include "fk_stream.h"
#include "fk_tile.h"
#include "fk_wide.h"

namespace fk {

typedef unsigned long long u64;

struct ResGeom {
    int ntr, ntc;        // tiles along rows / columns: balanced split, columns in groups of 4 cells
    int th_max, tw_max;  // largest tile
    int pitch;           // floats per row of a u buffer = tw_max + 8
    int nsteps;          // Euler steps of the launch
    int nc;              // adjacent cells per thread group: 1, 2 or 4
    int slots;           // mailbox records per tile and parity: 8 * tw_max (4 top + 4 bottom rows) + 8 * th_max (columns)
    u64* xchg;           // mailboxes [2 parities][batch * ntr * ntc tiles][slots], tags zero at launch
    u64* timing;         // null, or cycle counters CTA (0, 0) fills (development: FK_RES_TIMING=1)
    unsigned spin_limit; // reads of one record after which the kernel traps instead of hanging the device
};

struct ResCta {
    int ti, tj, tile, sim;
    int r0, r1, c0, c1, th, tw, q;   // q = tw / nc groups per row
    int qe;                          // groups per row that lie within 4 cells of a tile edge, per side = 4 / nc
    int nring, ninner;               // groups within 4 cells of a tile edge / the others
    int has_n, has_s, has_w, has_e;  // neighbours (0 at a physical edge)
    int e_nt, e_nb, e_nl, e_nr, nedge;   // cells within 4 of a PHYSICAL edge: top/bottom rows, left/right columns, total
    int nhalo[4];                    // 16-byte units (2 records) of the north, south, west, east halo
    float *U0, *V, *Wd, *Dm, *DXm, *DYm;   // u buffer of parity p: U0 + p * nu
    int nu;                                // floats per u buffer
    long long boff, boffD;
    long long mbox;                  // this tile's mailbox: xchg + mbox (+ parity * pstride)
    long long pstride;
    const StimDev* stims;
};

FK_HD long long res_smem_floats(int th_max, int tw_max) {
    return 2LL * (th_max + 8) * (tw_max + 8) + 5LL * th_max * tw_max;
}

// the cells of tile rows [r0, r1) x columns [c0, c1) that lie within 4 cells of a physical edge of the H x W tissue
// (the ones whose formulas are not all central): nt top and nb bottom rows of the tile, nl / nr columns of the rest
FK_HD int res_edge_counts(int H, int W, int r0, int r1, int c0, int c1, int& nt, int& nb, int& nl, int& nr) {
    const int th = r1 - r0, tw = c1 - c0;
    nt = (r1 < 4 ? r1 : 4) - r0; if (nt < 0) nt = 0;
    nb = r1 - (r0 > H - 4 ? r0 : H - 4); if (nb < 0) nb = 0;
    if (nt + nb > th) { nt = th; nb = 0; }
    nl = (c1 < 4 ? c1 : 4) - c0; if (nl < 0) nl = 0;
    nr = c1 - (c0 > W - 4 ? c0 : W - 4); if (nr < 0) nr = 0;
    if (nl + nr > tw) { nl = tw; nr = 0; }
    return (nt + nb) * tw + (th - nt - nb) * (nl + nr);
}

FK_HD void res_setup(const TileArgs& A, const ResGeom& G, int tile, int sim, int batch, float* smem, ResCta& X) {
    X.tile = tile; X.sim = sim;
    X.ti = tile / G.ntc; X.tj = tile - X.ti * G.ntc;
    const int Q = A.W >> 2;
    X.r0 = (int)((long long)A.H * X.ti / G.ntr);
    X.r1 = (int)((long long)A.H * (X.ti + 1) / G.ntr);
    X.c0 = 4 * (int)((long long)Q * X.tj / G.ntc);
    X.c1 = 4 * (int)((long long)Q * (X.tj + 1) / G.ntc);
    X.th = X.r1 - X.r0; X.tw = X.c1 - X.c0; X.q = X.tw / G.nc; X.qe = 4 / G.nc;
    if (X.th <= 8 || X.tw <= 8) { X.nring = X.th * X.q; X.ninner = 0; }
    else { X.nring = 8 * X.q + 2 * X.qe * (X.th - 8); X.ninner = (X.th - 8) * (X.q - 2 * X.qe); }
    X.nedge = res_edge_counts(A.H, A.W, X.r0, X.r1, X.c0, X.c1, X.e_nt, X.e_nb, X.e_nl, X.e_nr);
    X.has_n = X.ti > 0; X.has_s = X.ti < G.ntr - 1; X.has_w = X.tj > 0; X.has_e = X.tj < G.ntc - 1;
    X.nhalo[0] = X.has_n ? 2 * X.tw : 0; X.nhalo[1] = X.has_s ? 2 * X.tw : 0;
    X.nhalo[2] = X.has_w ? 2 * X.th : 0; X.nhalo[3] = X.has_e ? 2 * X.th : 0;
    const long long nu = (long long)(G.th_max + 8) * G.pitch, nv = (long long)G.th_max * G.tw_max;
    X.U0 = smem; X.nu = (int)nu;
    X.V = smem + 2 * nu; X.Wd = X.V + nv; X.Dm = X.Wd + nv; X.DXm = X.Dm + nv; X.DYm = X.DXm + nv;
    X.boff = (long long)sim * A.plane;
    X.boffD = (long long)sim * A.plane_D;
    const long long ntiles = (long long)G.ntr * G.ntc;
    X.mbox = ((long long)sim * ntiles + tile) * G.slots;
    X.pstride = (long long)batch * ntiles * G.slots;
    X.stims = A.stims ? A.stims + (long long)sim * A.n_stim : nullptr;
}

// group i of the ring (phase 0) or of the interior (phase 1) -> local row and column of its first cell
FK_HD void res_locate(const ResGeom& G, const ResCta& X, int phase, int i, int& lr, int& lc) {
    const int q = X.q, qe = X.qe, nc = G.nc;
    if (phase) { const int qi = q - 2 * qe, r = i / qi; lr = 4 + r; lc = nc * (qe + i - r * qi); return; }
    if (X.ninner == 0 || i < 4 * q) { const int r = i / q; lr = r; lc = nc * (i - r * q); return; }
    if (i < 8 * q) { const int j = i - 4 * q, r = j / q; lr = X.th - 4 + r; lc = nc * (j - r * q); return; }
    const int j = i - 8 * q, r = j / (2 * qe), k = j - r * 2 * qe;   // middle rows: qe groups at each end
    lr = 4 + r;
    lc = nc * (k < qe ? k : q - 2 * qe + k);
}

// edge cell e (see res_edge_counts) -> local row and column
FK_HD void res_locate_edge(const ResCta& X, int e, int& lr, int& lc) {
    const int band = (X.e_nt + X.e_nb) * X.tw;
    if (e < band) {
        const int r = e / X.tw;
        lc = e - r * X.tw;
        lr = r < X.e_nt ? r : X.th - X.e_nb + (r - X.e_nt);
        return;
    }
    e -= band;
    const int m = X.e_nl + X.e_nr, r = e / m, k = e - r * m;
    lr = X.e_nt + r;
    lc = k < X.e_nl ? k : X.tw - X.e_nr + (k - X.e_nl);
}

// ------------------------------------------------------------------ mailbox records
FK_HD u64 ll_pack(float v, unsigned tag) {
#if defined(__CUDA_ARCH__)
    return ((u64)tag << 32) | (u64)__float_as_uint(v);
#else
    unsigned b;
    __builtin_memcpy(&b, &v, 4);
    return ((u64)tag << 32) | (u64)b;
#endif
}
FK_HD float ll_value(u64 r) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float((unsigned)r);
#else
    const unsigned b = (unsigned)r;
    float v;
    __builtin_memcpy(&v, &b, 4);
    return v;
#endif
}
FK_HD unsigned ll_tag(u64 r) { return (unsigned)(r >> 32); }
FK_HD void ll_store1(u64* p, u64 a) {
#if defined(__CUDA_ARCH__)
    asm volatile("st.relaxed.gpu.global.b64 [%0], %1;" ::"l"(p), "l"(a) : "memory");
#else
    *p = a;
#endif
}
FK_HD void ll_store2(u64* p, u64 a, u64 b) {   // 16-byte aligned pair; each record is atomic on its own
#if defined(__CUDA_ARCH__)
    asm volatile("st.relaxed.gpu.global.v2.b64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
#else
    p[0] = a; p[1] = b;
#endif
}
FK_HD void ll_load2(const u64* p, u64& a, u64& b) {
#if defined(__CUDA_ARCH__)
    asm volatile("ld.relaxed.gpu.global.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
#else
    a = p[0]; b = p[1];
#endif
}

// publish NC adjacent cells (local row lr, first local column lc) of the new u towards every neighbour whose halo they
// are part of.  Mailbox layout: record (rr, lc) at rr * tw_max + lc for rr = 0..3 (my top rows, read by the NORTH
// neighbour) and 4..7 (bottom rows, SOUTH neighbour); record (lr, cc) at 8 * tw_max + 8 * lr + cc for cc = 0..3 (my
// first columns, WEST neighbour) and 4..7 (last columns, EAST neighbour).
template <int NC>
FK_HD void res_publish(const ResGeom& G, const ResCta& X, u64* box, int lr, int lc, const float* un, unsigned tag) {
    u64 r[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) r[k] = ll_pack(un[k], tag);
    auto put = [&](u64* p) {
        if (NC == 1) ll_store1(p, r[0]);
        else {
#pragma unroll
            for (int k = 0; k + 1 < NC; k += 2) ll_store2(p + k, r[k], r[k + 1]);
        }
    };
    if (X.has_n && lr < 4) put(box + lr * G.tw_max + lc);
    if (X.has_s && lr >= X.th - 4) put(box + (4 + lr - (X.th - 4)) * G.tw_max + lc);
    if (X.has_w && lc < 4) put(box + 8 * G.tw_max + 8 * lr + lc);
    if (X.has_e && lc >= X.tw - 4) put(box + 8 * G.tw_max + 8 * lr + 4 + lc - (X.tw - 4));
}

// halo unit i (2 records) of step s: where it comes from (the neighbour's mailbox) and where it goes (next u buffer)
FK_HD void res_halo_unit(const ResGeom& G, const ResCta& X, float* nxt, const u64* xp, int i, const u64*& src, float*& dst) {
    const int P = G.pitch, hw = X.tw >> 1;
    if (i < X.nhalo[0]) {          // north halo rows -4 .. -1  <-  bottom rows (rr = 4 .. 7) of the tile above
        const int j = i / hw, k = 2 * (i - j * hw);
        src = xp + (X.mbox - (long long)G.ntc * G.slots) + (4 + j) * G.tw_max + k;
        dst = nxt + j * P + 4 + k;
        return;
    }
    i -= X.nhalo[0];
    if (i < X.nhalo[1]) {          // south halo rows th .. th+3  <-  top rows (rr = 0 .. 3) of the tile below
        const int j = i / hw, k = 2 * (i - j * hw);
        src = xp + (X.mbox + (long long)G.ntc * G.slots) + j * G.tw_max + k;
        dst = nxt + (X.th + 4 + j) * P + 4 + k;
        return;
    }
    i -= X.nhalo[1];
    if (i < X.nhalo[2]) {          // west halo columns -4 .. -1  <-  last columns (cc = 4 .. 7) of the tile to the left
        const int r = i >> 1, k = 2 * (i & 1);
        src = xp + (X.mbox - G.slots) + 8 * G.tw_max + 8 * r + 4 + k;
        dst = nxt + (r + 4) * P + k;
        return;
    }
    i -= X.nhalo[2];               // east halo columns tw .. tw+3  <-  first columns (cc = 0 .. 3) of the tile to the right
    const int r = i >> 1, k = 2 * (i & 1);
    src = xp + (X.mbox + G.slots) + 8 * G.tw_max + 8 * r + k;
    dst = nxt + (r + 4) * P + X.tw + 4 + k;
}

// receive the halo of step s (tag s + 1) into the next u buffer; returns false if a record never arrived
FK_HD bool res_halo(const ResGeom& G, const ResCta& X, int s, int tid, int nthr) {
    float* nxt = X.U0 + ((s + 1) & 1) * X.nu;
    const u64* xp = G.xchg + ((s + 1) & 1) * X.pstride;
    const unsigned tag = (unsigned)(s + 1);
    const int U = X.nhalo[0] + X.nhalo[1] + X.nhalo[2] + X.nhalo[3];
    for (int i = tid; i < U; i += 2 * nthr) {   // two units in flight per thread
        const u64 *s0, *s1 = nullptr;
        float *d0, *d1 = nullptr;
        res_halo_unit(G, X, nxt, xp, i, s0, d0);
        const bool two = i + nthr < U;
        if (two) res_halo_unit(G, X, nxt, xp, i + nthr, s1, d1);
        u64 a0, b0, a1 = 0, b1 = 0;
        ll_load2(s0, a0, b0);
        if (two) ll_load2(s1, a1, b1);
        unsigned spins = 0;
        while (ll_tag(a0) != tag || ll_tag(b0) != tag) {
#if defined(__CUDA_ARCH__)
            if (++spins > G.spin_limit) return false;
            ll_load2(s0, a0, b0);
#else
            return false;   // the emulation runs the CTAs in lock step: the record must be there
#endif
        }
        d0[0] = ll_value(a0); d0[1] = ll_value(b0);
        if (two) {
            while (ll_tag(a1) != tag || ll_tag(b1) != tag) {
#if defined(__CUDA_ARCH__)
                if (++spins > G.spin_limit) return false;
                ll_load2(s1, a1, b1);
#else
                return false;
#endif
            }
            d1[0] = ll_value(a1); d1[1] = ll_value(b1);
        }
    }
    return true;
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
        st4(X.U0 + (row - X.r0 + 4) * G.pitch + (c - X.c0 + 4), t);
    }
    const int q4 = X.tw >> 2, own = X.th * q4;
    for (int i = tid; i < own; i += nthr) {
        const int lr = i / q4, lc = 4 * (i - lr * q4);
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

// stimulus value of one cell: the last active stimulus whose field is non-zero there (solve.py:260-269), else 0
FK_HD float res_stim(const TileArgs& A, const ResCta& X, unsigned mask, long long g) {
    float st = 0.0f;
    for (int i = 0; i < A.n_stim; ++i)
        if (mask >> i & 1u) {
            const float f = ldg1(X.stims[i].field + g);
            if (f != 0.0f) st = f;
        }
    return st;
}

// NC-wide shared-memory accesses (NC = 1, 2, 4 floats, naturally aligned)
template <int NC>
FK_HD void ldn(const float* p, float* v) {
    if (NC == 4) unpack4(ld4(p), v);
    else if (NC == 2) { const F2 t = ld2(p); v[0] = t.x; v[1] = t.y; }
    else v[0] = p[0];
}
template <int NC>
FK_HD void stn(float* p, const float* v) {
    if (NC == 4) st4(p, v);
    else if (NC == 2) { F2 t; t.x = v[0]; t.y = v[1]; *reinterpret_cast<F2*>(p) = t; }
    else p[0] = v[0];
}

// first derivative / dx of whichever kind (solve.py:232-249) from a window a[0..6] centred on a[3]
template <bool EXACT>
FK_HD float res_deriv(const Consts& K, int kind, const float* a) {
    if (kind == CEN) return dcen<EXACT>(K, a[1], a[2], a[4], a[5]);
    if (kind == FWD)
        return deriv<EXACT>(K, FWD, (float)(-11.0 / 6.0), 3.0f, -(float)(3.0 / 2.0), (float)(1.0 / 3.0), a[3], a[4], a[5], a[6]);
    return deriv<EXACT>(K, BWD, (float)(-1.0 / 3.0), (float)(3.0 / 2.0), -3.0f, (float)(11.0 / 6.0), a[0], a[1], a[2], a[3]);
}

// First and second derivative along one axis of the edge-padded array (solve.py:29-31, 49-52, crop :61-65) at padded
// index P of an axis of n cells, anywhere on the axis: up[0..12] are the padded values P-6 .. P+6.  Only the first
// derivatives the second pass reads are evaluated (same values as fk_wide.h's wide_axis_general).
template <bool EXACT>
FK_HD void res_axis_general(const Consts& K, const float* up, int P, int n, float& d1, float& d2) {
    const int kind = kind_of(P, n, 1, 1);
    // the second pass reads g[1], g[2], g[4], g[5] (central; d1 = g[3]), g[3..6] (forward) or g[0..3] (backward)
    const int lo = kind == BWD ? 0 : (kind == CEN ? 1 : 3), hi = kind == FWD ? 6 : (kind == CEN ? 5 : 3);
    float g[7];
#pragma unroll
    for (int m = 0; m < 7; ++m) {   // ONE inlined copy of the derivative per slot keeps the kernel small
        g[m] = 0.0f;
        if (m >= lo && m <= hi) g[m] = res_deriv<EXACT>(K, kind_of(P - 3 + m, n, 1, 1), up + m);
    }
    d1 = g[3];
    d2 = res_deriv<EXACT>(K, kind, g);
}

// one Euler step of NC cells (row, c .. c+NC-1), local (lr, lc): reads `cur` (+ halo), writes `nxt`, v, w in place; ring
// groups (box != null) also publish the new u; the last step writes the caller's output instead
// GENERAL = false: the caller guarantees that every formula of the group is central (no cell within 4 of a physical
// edge) and only that path is compiled.
template <bool EXACT, int NC, bool GENERAL>
FK_HD void res_group(const TileArgs& A, const ResGeom& G, const ResCta& X, const float* cur, float* nxt, int lr, int lc,
                     unsigned mask, bool last, u64* box, unsigned tag) {
    const int H = A.H, W = A.W, P = G.pitch, row = X.r0 + lr, c = X.c0 + lc;
    const float* uc0 = cur + (lr + 4) * P + (lc + 4);
    float u_x[NC], u_y[NC], u_xx[NC], u_yy[NC], uc[NC];
    ldn<NC>(uc0, uc);
    const int o = lr * G.tw_max + lc;
    float v[NC], w[NC], Dv[NC], DXv[NC], DYv[NC], stim[NC];
    ldn<NC>(X.V + o, v);
    ldn<NC>(X.Wd + o, w);
    ldn<NC>(X.Dm + o, Dv);
    ldn<NC>(X.DXm + o, DXv);
    ldn<NC>(X.DYm + o, DYv);
    const long long g = (long long)row * W + c;
#pragma unroll
    for (int k = 0; k < NC; ++k) stim[k] = 0.0f;
    if (mask) {
#pragma unroll
        for (int k = 0; k < NC; ++k) stim[k] = res_stim(A, X, mask, g + k);
    }
    // ---- vertical (axis 0)
    if (!GENERAL || (row >= 4 && row + 5 <= H)) {   // every formula central, no clamped row: u_x at rows row-2 .. row+2, then u_xx
        float ur[9][NC];
#pragma unroll
        for (int j = 0; j < 9; ++j) {
            if (j == 4) {
#pragma unroll
                for (int k = 0; k < NC; ++k) ur[4][k] = uc[k];
            } else {
                ldn<NC>(uc0 + (j - 4) * P, ur[j]);
            }
        }
#pragma unroll
        for (int k = 0; k < NC; ++k) {
            float gx[5];
#pragma unroll
            for (int j = 0; j < 5; ++j) gx[j] = dcen<EXACT>(A.K, ur[j][k], ur[j + 1][k], ur[j + 3][k], ur[j + 4][k]);
            u_x[k] = gx[2];
            u_xx[k] = dcen<EXACT>(A.K, gx[0], gx[1], gx[3], gx[4]);
        }
    } else {
        // padded rows P-6 .. P+6 = tissue rows row-6 .. row+6 clamped (solve.py:31); entries the applicable formulas
        // never use may fall outside the tile's buffer rows and are clamped into it
        float ur[13][NC];
        const int nrows = G.th_max + 8;
#pragma unroll
        for (int j = 0; j < 13; ++j)
            ldn<NC>(cur + clampi(clampi(row + j - 6, 0, H - 1) - X.r0 + 4, 0, nrows - 1) * P + (lc + 4), ur[j]);
#pragma unroll
        for (int k = 0; k < NC; ++k) {
            float col[13];
#pragma unroll
            for (int j = 0; j < 13; ++j) col[j] = ur[j][k];
            res_axis_general<EXACT>(A.K, col, row + 1, H, u_x[k], u_xx[k]);
        }
    }
    // ---- horizontal (axis 1)
    if (!GENERAL || (c >= 4 && c + NC + 4 <= W)) {   // central for all NC cells: u_y at columns c-2 .. c+NC+1, then u_yy
        float e[NC + 8];               // columns c-4 .. c+NC+3
        const float* urow = uc0 - 4;
        if (NC == 4) { unpack4(ld4(urow), e); unpack4(ld4(urow + 8), e + 8); }
        else if (NC == 2) {
            const F2 a = ld2(urow), b = ld2(urow + 2), cc = ld2(urow + 6), d = ld2(urow + 8);
            e[0] = a.x; e[1] = a.y; e[2] = b.x; e[3] = b.y; e[6] = cc.x; e[7] = cc.y; e[8] = d.x; e[9] = d.y;
        } else {
#pragma unroll
            for (int m = 0; m < 9; ++m)
                if (m != 4) e[m] = urow[m];
        }
#pragma unroll
        for (int k = 0; k < NC; ++k) e[4 + k] = uc[k];
        float gy[NC + 4];
#pragma unroll
        for (int m = 0; m < NC + 4; ++m) gy[m] = dcen<EXACT>(A.K, e[m], e[m + 1], e[m + 3], e[m + 4]);
#pragma unroll
        for (int k = 0; k < NC; ++k) {
            u_y[k] = gy[k + 2];
            u_yy[k] = dcen<EXACT>(A.K, gy[k], gy[k + 1], gy[k + 3], gy[k + 4]);
        }
    } else {
        const float* urow = cur + (lr + 4) * P;
#pragma unroll
        for (int k = 0; k < NC; ++k) {
            float win[13];   // padded columns Q-6 .. Q+6 of cell c + k, clamped like the rows above
#pragma unroll
            for (int j = 0; j < 13; ++j) win[j] = urow[clampi(clampi(c + k + j - 6, 0, W - 1) - X.c0 + 4, 0, P - 1)];
            res_axis_general<EXACT>(A.K, win, c + k + 1, W, u_y[k], u_yy[k]);
        }
    }
    float un[NC], vn[NC], wn[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) {
        const float del_u = diffusion<EXACT>(Dv[k], DXv[k], DYv[k], u_x[k], u_y[k], u_xx[k], u_yy[k]);
        float d_v, d_w, d_u;
        cell_rhs<EXACT>(A.K, uc[k], v[k], w[k], del_u, stim[k], d_v, d_w, d_u);
        vn[k] = euler<EXACT>(v[k], d_v, A.K.dt);
        wn[k] = euler<EXACT>(w[k], d_w, A.K.dt);
        un[k] = euler<EXACT>(uc[k], d_u, A.K.dt);
    }
    if (last) {
        stn<NC>(A.u_out + X.boff + g, un);
        stn<NC>(A.v_out + X.boff + g, vn);
        stn<NC>(A.w_out + X.boff + g, wn);
    } else {
        if (box) res_publish<NC>(G, X, box, lr, lc, un, tag);
        stn<NC>(nxt + (lr + 4) * P + (lc + 4), un);
        stn<NC>(X.V + o, vn);
        stn<NC>(X.Wd + o, wn);
    }
}

// The ring (phase 0) or interior (phase 1) groups of step s, strided over the CTA's threads.  With NC > 1 the groups
// that touch a physical edge are skipped and their cells done ONE PER THREAD as extra items of phase 0: the one-sided
// formulas cost several times the central ones, and a tile at the tissue's edge must not be slower than the others
// (every CTA waits for its neighbours each step).
template <bool EXACT, int NC>
FK_HD void res_phase(const TileArgs& A, const ResGeom& G, const ResCta& X, int s, int phase, unsigned mask, int tid,
                     int nthr) {   // (the kernel calls this from ONE site, in a phase loop: a single copy of the body)
    const float* cur = X.U0 + (s & 1) * X.nu;
    float* nxt = X.U0 + ((s + 1) & 1) * X.nu;
    const bool last = s == G.nsteps - 1;
    u64* box = (last || phase) ? nullptr : G.xchg + ((s + 1) & 1) * X.pstride + X.mbox;
    const unsigned tag = (unsigned)(s + 1);
    const int n = phase ? X.ninner : X.nring;
    const int ne = (phase || NC == 1) ? 0 : X.nedge;
    for (int i = tid; i < n + ne; i += nthr) {
        int lr, lc;
        if (NC == 1) {
            res_locate(G, X, phase, i, lr, lc);
            res_group<EXACT, 1, true>(A, G, X, cur, nxt, lr, lc, mask, last, box, tag);
        } else if (i < n) {
            res_locate(G, X, phase, i, lr, lc);
            const int row = X.r0 + lr, c = X.c0 + lc;
            if (row >= 4 && row + 5 <= A.H && c >= 4 && c + NC + 4 <= A.W)
                res_group<EXACT, NC, false>(A, G, X, cur, nxt, lr, lc, mask, last, box, tag);
        } else {
            res_locate_edge(X, i - n, lr, lc);
            res_group<EXACT, 1, true>(A, G, X, cur, nxt, lr, lc, mask, last, box, tag);
        }
    }
}

// ------------------------------------------------------------------ launch geometry (host)
struct ResPlan {
    ResGeom G;
    int threads;
    long long smem_bytes;
    long long xchg_bytes;   // mailboxes of the launch (both parities)
};

enum { FK_RES_MAX_THREADS = 512 };

FK_HD void res_group_counts(int th, int tw, int nc, int& nring, int& ninner) {
    const int q = tw / nc, qe = 4 / nc;
    if (th <= 8 || tw <= 8) { nring = th * q; ninner = 0; }
    else { nring = 8 * q + 2 * qe * (th - 8); ninner = (th - 8) * (q - 2 * qe); }
}

// Tiles for a (batch, H, W) problem on `capacity` co-resident CTAs with `smem_limit` bytes each and `xchg_limit` bytes
// of mailboxes.  A tile next to another tile is at least 8 cells wide/tall (a halo comes from ONE tile, and the edge
// formulas reach 7 cells into the tissue).  Cells per thread group: the smallest NC whose groups fit one round of 512
// threads -- a small tile is latency bound and wants every lane busy; a large one wants the fewest instructions.
// Cost of a plan: the serial rounds a CTA runs per step, ring and interior separately (the halo travels in between),
// each at least a lone warp's dependent chain, then the tile size; ties go to fewer tiles.
inline bool plan_resident(int H, int W, int batch, int capacity, long long smem_limit, long long xchg_limit, int force_ntr,
                          int force_ntc, int force_threads, int force_nc, ResPlan& P) {
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
            const int th = (H + ntr - 1) / ntr, tw = 4 * ((Q + ntc - 1) / ntc);
            const long long smem = res_smem_floats(th, tw) * 4;
            if (smem > smem_limit) continue;
            const int slots = 8 * tw + 8 * th;
            const long long xbytes = 2LL * batch * ntr * ntc * slots * (long long)sizeof(u64);
            if (xbytes > xchg_limit) continue;
            int nc = force_nc;
            if (nc != 1 && nc != 2 && nc != 4) nc = th * tw <= FK_RES_MAX_THREADS ? 1 : (th * tw <= 2 * FK_RES_MAX_THREADS ? 2 : 4);
            int nring, ninner;
            res_group_counts(th, tw, nc, nring, ninner);
            if (nc > 1) {   // the corner tile's cells next to a physical edge are extra one-cell items of the ring phase
                int nt, nb, nl, nr;
                nring += res_edge_counts(H, W, 0, th, 0, tw, nt, nb, nl, nr);
            }
            int threads = force_threads;
            if (threads <= 0) {
                const int m = nring > ninner ? nring : ninner;
                threads = (m + 31) / 32 * 32;
                if (threads < 64) threads = 64;
                if (threads > FK_RES_MAX_THREADS) threads = FK_RES_MAX_THREADS;
            }
            // a round of one group per thread: a dependent chain ~ (4 + 3 nc) units long, or the issue time of its warps
            auto phase_cost = [&](int n) {
                if (n == 0) return 0.0;
                const int rounds = (n + threads - 1) / threads;
                const double chain = 4.0 + 3.0 * nc, issue = (double)((n + 31) / 32) * (2.0 + 2.5 * nc) / 4.0;
                return rounds * chain > issue ? rounds * chain : issue;
            };
            const double cost = phase_cost(nring) + phase_cost(ninner) + 1e-3 * ntr * ntc + 1e-4 * (th + tw);
            if (cost < best) {
                best = cost;
                found = true;
                P.G.ntr = ntr; P.G.ntc = ntc; P.G.th_max = th; P.G.tw_max = tw; P.G.pitch = tw + 8;
                P.G.nc = nc; P.G.slots = slots;
                P.threads = threads;
                P.smem_bytes = smem;
                P.xchg_bytes = xbytes;
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
template <bool EXACT>
inline void emu_res_phase(const TileArgs& A, const ResGeom& G, const ResCta& X, int s, int phase, unsigned mask) {
    if (G.nc == 1) res_phase<EXACT, 1>(A, G, X, s, phase, mask, 0, 1);
    else if (G.nc == 2) res_phase<EXACT, 2>(A, G, X, s, phase, mask, 0, 1);
    else res_phase<EXACT, 4>(A, G, X, s, phase, mask, 0, 1);
}
// CTAs advance in lock step, phase by phase, which is one legal interleaving of the mailbox protocol; shared memory is
// poisoned with NaN so that a read of a halo nobody filled shows up in the result, and a record that has not arrived
// when its consumer looks for it is an error (-7).
inline int emu_resident_launch(const ResPlan& P, const TileArgs& A, int batch, int exact) {
    ResGeom G = P.G;
    const int ntiles = G.ntr * G.ntc;
    const long long floats = res_smem_floats(G.th_max, G.tw_max);
    std::vector<std::vector<float>> smem((size_t)ntiles * batch, std::vector<float>((size_t)floats, __builtin_nanf("")));
    std::vector<u64> xchg((size_t)(P.xchg_bytes / sizeof(u64)), 0ull);
    G.xchg = xchg.data();
    std::vector<ResCta> X((size_t)ntiles * batch);
    for (int sim = 0; sim < batch; ++sim)
        for (int t = 0; t < ntiles; ++t) {
            ResCta& x = X[(size_t)sim * ntiles + t];
            res_setup(A, G, t, sim, batch, smem[(size_t)sim * ntiles + t].data(), x);
            res_load(A, G, x, 0, 1);
        }
    for (int s = 0; s < G.nsteps; ++s) {
        for (size_t i = 0; i < X.size(); ++i) {
            const unsigned mask = res_mask(A, X[i], s);
            for (int phase = 0; phase < 2; ++phase) {
                if (exact) emu_res_phase<true>(A, G, X[i], s, phase, mask);
                else emu_res_phase<false>(A, G, X[i], s, phase, mask);
            }
        }
        if (s == G.nsteps - 1) break;
        for (size_t i = 0; i < X.size(); ++i)
            if (!res_halo(G, X[i], s, 0, 1)) return -7;
    }
    return 0;
}
#endif

}  // namespace fk
