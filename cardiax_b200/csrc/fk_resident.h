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
// CLUSTER transport (ResGeom::cluster, fk_cluster_kernel): a tissue of at most 16 tiles is ONE thread-block cluster.  A
// ring cell's new u is then stored straight into the halo of the neighbouring CTA's next u buffer through distributed
// shared memory, and one cluster barrier per step (arrive.release / wait.acquire) replaces mailboxes, tags, polling and
// the halo copy: ~0.2 us per exchange instead of the ~1.1 us a store -> L2 -> polled load round trip costs on B200.
// Clusters are independent, so a batch of small tissues needs no co-residency (any batch size, scheduled in waves).
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

// Interior items as blocks of 2 rows x 4 cells (res_block2: 22 % fewer derivative instructions per cell).  Measured
// on B200 against plain 4-cell groups: 512^2 3.42 -> 4.84 us per step (half the items, so fewer warps to hide latency
// behind, and the block spills at the 128-register cap of a 512-thread CTA), 1024^2 7.11 -> 7.61, 1200^2 8.86 -> 8.89.
// Off; kept for a CTA shape with more registers per thread.
#ifndef FK_RES_R2
#define FK_RES_R2 0
#endif
// development switches for A/B builds
#ifndef FK_RES_TWO_PASS
#define FK_RES_TWO_PASS 0       // 1: the planner may choose the two-pass step (ResGeom::tp).  EXPERIMENT, off: bit-identical
#endif                          // (CPU emulation and GPU suite green with it on) but slower on B200 -- 512^2 4.03 vs 3.30 us per
                                // step, 1024^2 8.7 vs 7.6: a phase of this kernel is bound by the latency of ONE item's
                                // dependent chain at ~3 warps per scheduler (IPC ~0.25), not by its instruction count, so
                                // three short phases + a barrier lose to two long ones (profiles/probe_cluster_r02.md)
#ifndef FK_RES_EDGE_SPECIAL
#define FK_RES_EDGE_SPECIAL 1   // physical-edge cells through res_axis_edge (specialised by position) instead of res_axis_general
#endif

namespace fk {

typedef unsigned long long u64;

struct ResGeom {
    int ntr, ntc;        // tiles along rows / columns (columns are cut in groups of 4 cells)
    int eh, ewq;         // rows / column groups of the tiles at the tissue's edges (their one-sided formulas cost more,
                         // so they get fewer cells); the rest is split evenly.  0 = split everything evenly
    int th_max, tw_max;  // largest tile
    int pitch;           // floats per row of a u buffer = tw_max + 8
    int nsteps;          // Euler steps of the launch
    int nc;              // adjacent cells per thread group: 1, 2 or 4
    int single;          // 1: every tile's items fit one round of the CTA's threads -- no interior phase (all "ring")
    int mg;              // 1: the diffusivity maps stay in global memory (L2) instead of shared memory (larger tissues)
    int r2;              // 1 (nc = 4 only): interior items are blocks of 2 rows x 4 cells
    int slots;           // mailbox records per tile and parity: 8 * tw_max (4 top + 4 bottom rows) + 8 * th_max (columns)
    int cluster;         // 1: the tiles of a tissue form one thread-block cluster and exchange through distributed shared memory
    int tp;              // 1 (nc = 4): TWO-PASS step -- first derivatives of the whole tile into shared memory, then the cells
    u64* xchg;           // mailboxes [2 parities][batch * ntr * ntc tiles][slots], tags zero at launch
    u64* timing;         // null, or cycle counters CTA (0, 0) fills (development: FK_RES_TIMING=1)
    unsigned spin_limit; // reads of one record after which the kernel traps instead of hanging the device
};

struct ResCta {
    int ti, tj, tile, sim;
    int r0, r1, c0, c1, th, tw, q;   // q = tw / nc groups per row
    int qe;                          // groups per row that lie within 4 cells of a tile edge, per side = 4 / nc
    int has_n, has_s, has_w, has_e;  // neighbours (0 at a physical edge)
    int ir0, ir1, ig0, ig1;          // INTERIOR = rows [ir0, ir1) x groups [ig0, ig1): 4 cells away from every side
    int nring, ninner;               // RING groups (the rest) / interior items (groups, or 2-row blocks + a last odd row)
    int npair;                       // r2: items [0, npair) of the interior phase are 2-row blocks
    int e_nt, e_nb, e_nl, e_nr;      // cells within 4 of a PHYSICAL edge: top/bottom rows, left/right columns of the rest
    int nedge;                       // ... their number
    int nhalo[4];                    // 16-byte units (2 records) of the north, south, west, east halo
    int th_n, tw_w;                  // rows of the tile above, columns of the tile to the left (cluster transport)
    float* peer[4];                  // cluster transport: the north, south, west, east neighbour's U0 (its shared memory)
    float *U0, *V, *Wd, *Dm, *DXm, *DYm;   // u buffer of parity p: U0 + p * nu
    float *GX, *GY;                        // two-pass form: u_x of rows -2 .. th+1 (pitch tw_max), u_y of columns -4 .. tw+3 (pitch)
    int nu;                                // floats per u buffer
    long long boff, boffD;
    long long mbox;                  // this tile's mailbox: xchg + mbox (+ parity * pstride)
    long long pstride;
    const StimDev* stims;
};

FK_HD long long res_smem_floats(int th_max, int tw_max, int mg, int tp = 0) {
    return 2LL * (th_max + 8) * (tw_max + 8) + (mg ? 2LL : 5LL) * th_max * tw_max +
           (tp ? (long long)(th_max + 4) * tw_max + (long long)th_max * (tw_max + 8) : 0LL);
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

// split of n units into nt parts with e units in the first and last part (0: all parts even): start of part t
FK_HD int res_split(int n, int nt, int e, int t) {
    if (e <= 0 || nt < 3) return (int)((long long)n * t / nt);
    if (t == 0) return 0;
    if (t == nt) return n;
    return e + (int)((long long)(n - 2 * e) * (t - 1) / (nt - 2));
}

// geometry of one tile (no pointers): shared by the kernel and the planner
FK_HD void res_tile_geom(int H, int W, const ResGeom& G, int tile, ResCta& X) {
    X.tile = tile;
    X.ti = tile / G.ntc; X.tj = tile - X.ti * G.ntc;
    X.r0 = res_split(H, G.ntr, G.eh, X.ti);
    X.r1 = res_split(H, G.ntr, G.eh, X.ti + 1);
    X.c0 = 4 * res_split(W >> 2, G.ntc, G.ewq, X.tj);
    X.c1 = 4 * res_split(W >> 2, G.ntc, G.ewq, X.tj + 1);
    X.th = X.r1 - X.r0; X.tw = X.c1 - X.c0; X.q = X.tw / G.nc; X.qe = 4 / G.nc;
    X.has_n = X.ti > 0; X.has_s = X.ti < G.ntr - 1; X.has_w = X.tj > 0; X.has_e = X.tj < G.ntc - 1;
    // The ring is geometric -- 4 cells along EVERY side, physical edges included -- and the cells at a physical edge
    // all belong to phase 0.  (Tried: ring = only what a neighbour needs, the rest in phase 1.  Slower at every size:
    // the expensive one-sided cells then queue up behind the interior as a second, nearly empty round.)
    X.ir0 = 4; X.ir1 = X.th - 4; X.ig0 = X.qe; X.ig1 = X.q - X.qe;
    if (G.single || X.ir1 <= X.ir0 || X.ig1 <= X.ig0) { X.ir0 = X.ir1 = 0; X.ig0 = X.ig1 = 0; }   // no interior: all ring
    X.ninner = (X.ir1 - X.ir0) * (X.ig1 - X.ig0);
    X.nring = X.th * X.q - X.ninner;
    X.npair = 0;
    if (FK_RES_R2 && G.r2 && G.nc == 4) {   // interior rows in pairs; an odd last row stays a row of plain groups
        const int rows = X.ir1 - X.ir0, qi = X.ig1 - X.ig0;
        X.npair = (rows / 2) * qi;
        X.ninner = X.npair + (rows & 1) * qi;
    }
    X.nedge = res_edge_counts(H, W, X.r0, X.r1, X.c0, X.c1, X.e_nt, X.e_nb, X.e_nl, X.e_nr);
    X.th_n = X.has_n ? X.r0 - res_split(H, G.ntr, G.eh, X.ti - 1) : 0;
    X.tw_w = X.has_w ? X.c0 - 4 * res_split(W >> 2, G.ntc, G.ewq, X.tj - 1) : 0;
    X.peer[0] = X.peer[1] = X.peer[2] = X.peer[3] = nullptr;
    X.nhalo[0] = X.has_n ? 2 * X.tw : 0; X.nhalo[1] = X.has_s ? 2 * X.tw : 0;
    X.nhalo[2] = X.has_w ? 2 * X.th : 0; X.nhalo[3] = X.has_e ? 2 * X.th : 0;
}

FK_HD void res_setup(const TileArgs& A, const ResGeom& G, int tile, int sim, int batch, float* smem, ResCta& X) {
    res_tile_geom(A.H, A.W, G, tile, X);
    X.sim = sim;
    const long long nu = (long long)(G.th_max + 8) * G.pitch, nv = (long long)G.th_max * G.tw_max;
    X.U0 = smem; X.nu = (int)nu;
    X.V = smem + 2 * nu; X.Wd = X.V + nv; X.Dm = X.Wd + nv; X.DXm = X.Dm + nv; X.DYm = X.DXm + nv;   // (maps: only if !G.mg)
    X.GX = G.mg ? X.Dm : X.DYm + nv;   // (two-pass form only)
    X.GY = X.GX + (long long)(G.th_max + 4) * G.tw_max;
    X.boff = (long long)sim * A.plane;
    X.boffD = (long long)sim * A.plane_D;
    const long long ntiles = (long long)G.ntr * G.ntc;
    X.mbox = ((long long)sim * ntiles + tile) * G.slots;
    X.pstride = (long long)batch * ntiles * G.slots;
    X.stims = A.stims ? A.stims + (long long)sim * A.n_stim : nullptr;
}

// does a neighbour need cell (lr, lc)'s new u (is it within 4 cells of a side that has a neighbour)?
FK_HD bool res_publishes(const ResCta& X, int lr, int lc) {
    return (X.has_n && lr < 4) || (X.has_s && lr >= X.th - 4) || (X.has_w && lc < 4) || (X.has_e && lc >= X.tw - 4);
}

// group i of the ring (phase 0) or of the interior (phase 1) -> local row and column of its first cell
FK_HD void res_locate(const ResGeom& G, const ResCta& X, int phase, int i, int& lr, int& lc) {
    const int q = X.q, nc = G.nc;
    if (phase) {
        const int qi = X.ig1 - X.ig0;
        if (FK_RES_R2 && X.npair) {   // 2-row blocks first, then the odd last row
            if (i >= X.npair) { lr = X.ir1 - 1; lc = nc * (X.ig0 + i - X.npair); return; }
            const int r = i / qi;
            lr = X.ir0 + 2 * r; lc = nc * (X.ig0 + i - r * qi);
            return;
        }
        const int r = i / qi;
        lr = X.ir0 + r; lc = nc * (X.ig0 + i - r * qi);
        return;
    }
    if (X.ninner == 0) { const int r = i / q; lr = r; lc = nc * (i - r * q); return; }
    const int ntop = X.ir0 * q;
    if (i < ntop) { const int r = i / q; lr = r; lc = nc * (i - r * q); return; }
    i -= ntop;
    const int nbot = (X.th - X.ir1) * q;
    if (i < nbot) { const int r = i / q; lr = X.ir1 + r; lc = nc * (i - r * q); return; }
    i -= nbot;
    const int m = X.ig0 + q - X.ig1, r = i / m, k = i - r * m;   // middle rows: the groups left and right of the interior
    lr = X.ir0 + r;
    lc = nc * (k < X.ig0 ? k : X.ig1 + (k - X.ig0));
}

// edge cell e (of the set res_edge_counts describes) -> local row and column
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

// halo unit i (2 records): where it comes from (offset into a parity's mailboxes) and where it goes (offset into a u
// buffer).  The same for every step, so each thread works its first two units out once, before the step loop.
FK_HD void res_halo_unit(const ResGeom& G, const ResCta& X, int i, long long& src, int& dst) {
    const int P = G.pitch, hw = X.tw >> 1;
    if (i < X.nhalo[0]) {          // north halo rows -4 .. -1  <-  bottom rows (rr = 4 .. 7) of the tile above
        const int j = i / hw, k = 2 * (i - j * hw);
        src = (X.mbox - (long long)G.ntc * G.slots) + (4 + j) * G.tw_max + k;
        dst = j * P + 4 + k;
        return;
    }
    i -= X.nhalo[0];
    if (i < X.nhalo[1]) {          // south halo rows th .. th+3  <-  top rows (rr = 0 .. 3) of the tile below
        const int j = i / hw, k = 2 * (i - j * hw);
        src = (X.mbox + (long long)G.ntc * G.slots) + j * G.tw_max + k;
        dst = (X.th + 4 + j) * P + 4 + k;
        return;
    }
    i -= X.nhalo[1];
    if (i < X.nhalo[2]) {          // west halo columns -4 .. -1  <-  last columns (cc = 4 .. 7) of the tile to the left
        const int r = i >> 1, k = 2 * (i & 1);
        src = (X.mbox - G.slots) + 8 * G.tw_max + 8 * r + 4 + k;
        dst = (r + 4) * P + k;
        return;
    }
    i -= X.nhalo[2];               // east halo columns tw .. tw+3  <-  first columns (cc = 0 .. 3) of the tile to the right
    const int r = i >> 1, k = 2 * (i & 1);
    src = (X.mbox + G.slots) + 8 * G.tw_max + 8 * r + k;
    dst = (r + 4) * P + X.tw + 4 + k;
}

// what a thread does first in each phase and in the halo copy: the same every step, worked out once (the integer
// divisions of the item -> cell maps would otherwise sit on the critical path of every step)
struct ResThread {
    int ty0, lr0, lc0;   // first item of the ring phase: type (0 nothing, 1 central group of NC cells, 2 one general cell)
    int ty1, lr1, lc1;   // first item of the interior phase
    long long hs0, hs1;  // first two halo units: source offsets
    int hd0, hd1;        //                       destination offsets
};

// receive the halo of step s (tag s + 1) into the next u buffer; returns false if a record never arrived
FK_HD bool res_halo(const ResGeom& G, const ResCta& X, const ResThread& T, int s, int tid, int nthr) {
    float* nxt = X.U0 + ((s + 1) & 1) * X.nu;
    const u64* xp = G.xchg + ((s + 1) & 1) * X.pstride;
    const unsigned tag = (unsigned)(s + 1);
    const int U = X.nhalo[0] + X.nhalo[1] + X.nhalo[2] + X.nhalo[3];
    for (int i = tid; i < U; i += 2 * nthr) {   // two units in flight per thread
        long long s0 = T.hs0, s1 = T.hs1;
        int d0 = T.hd0, d1 = T.hd1;
        const bool two = i + nthr < U;
        if (i != tid) {
            res_halo_unit(G, X, i, s0, d0);
            if (two) res_halo_unit(G, X, i + nthr, s1, d1);
        }
        u64 a0, b0, a1 = 0, b1 = 0;
        ll_load2(xp + s0, a0, b0);
        if (two) ll_load2(xp + s1, a1, b1);
        unsigned spins = 0;
        while (ll_tag(a0) != tag || ll_tag(b0) != tag) {
#if defined(__CUDA_ARCH__)
            if (++spins > G.spin_limit) return false;
            ll_load2(xp + s0, a0, b0);
#else
            return false;   // the emulation runs the CTAs in lock step: the record must be there
#endif
        }
        nxt[d0] = ll_value(a0); nxt[d0 + 1] = ll_value(b0);
        if (two) {
            while (ll_tag(a1) != tag || ll_tag(b1) != tag) {
#if defined(__CUDA_ARCH__)
                if (++spins > G.spin_limit) return false;
                ll_load2(xp + s1, a1, b1);
#else
                return false;
#endif
            }
            nxt[d1] = ll_value(a1); nxt[d1 + 1] = ll_value(b1);
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
        if (!G.mg) {
            unpack4(ldg4(A.D + X.boffD + g), t); st4(X.Dm + o, t);
            unpack4(ldg4(A.DX + X.boffD + g), t); st4(X.DXm + o, t);
            unpack4(ldg4(A.DY + X.boffD + g), t); st4(X.DYm + o, t);
        }
    }
}

// stimuli of this CTA's tissue that are active at step s of the launch (solve.py:262-267, the caller's typing)
FK_HD unsigned res_mask(const TileArgs& A, const ResCta& X, int s) {
    unsigned m = 0;
    const double t = A.t0 + (double)s;
    for (int i = 0; i < A.n_stim; ++i) {
        const StimDev sd = X.stims[i];
        if (sd.field && stim_on(sd, t, A.t_is_int)) m |= 1u << i;
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
// NC-wide read-only global load (diffusivity maps kept in L2)
template <int NC>
FK_HD void ldgn(const float* p, float* v) {
    if (NC == 4) unpack4(ldg4(p), v);
    else {
#pragma unroll
        for (int k = 0; k < NC; ++k) v[k] = ldg1(p + k);
    }
}
template <int NC>
FK_HD void stn(float* p, const float* v) {
    if (NC == 4) st4(p, v);
    else if (NC == 2) { F2 t; t.x = v[0]; t.y = v[1]; *reinterpret_cast<F2*>(p) = t; }
    else p[0] = v[0];
}

// cluster transport: the same cells stored straight into the neighbours' halos (of the u buffer at float offset `off`
// from U0: the one the step writes), through distributed shared memory.  Layouts are identical in every CTA of the
// cluster (same pitch, same buffer offsets); only the neighbour's own height / width enters the address.
template <int NC>
FK_HD void res_publish_cl(const ResGeom& G, const ResCta& X, long long off, int lr, int lc, const float* un) {
    const int P = G.pitch;
    if (X.has_n && lr < 4) stn<NC>(X.peer[0] + off + (X.th_n + 4 + lr) * P + 4 + lc, un);
    if (X.has_s && lr >= X.th - 4) stn<NC>(X.peer[1] + off + (lr - (X.th - 4)) * P + 4 + lc, un);
    if (X.has_w && lc < 4) stn<NC>(X.peer[2] + off + (lr + 4) * P + X.tw_w + 4 + lc, un);
    if (X.has_e && lc >= X.tw - 4) stn<NC>(X.peer[3] + off + (lr + 4) * P + (lc - (X.tw - 4)), un);
}

// first derivative / dx of whichever kind (solve.py:232-249) from a window a[0..6] centred on a[3].  Branch free: the
// central value and ONE one-sided value (coefficients and operands selected) are both computed and one is kept -- a
// cell at a physical edge is a lone dependent chain on the step's critical path, and straight-line code lets its
// eight derivatives overlap.  Same operations on the same operands as fk_core.h's deriv(): identical bits.
template <bool EXACT>
FK_HD float res_deriv(const Consts& K, int kind, const float* a) {
    const float cen = dcen<EXACT>(K, a[1], a[2], a[4], a[5]);
    const bool f = kind == FWD;
    const float k0 = f ? (float)(-11.0 / 6.0) : (float)(-1.0 / 3.0), k1 = f ? 3.0f : (float)(3.0 / 2.0);
    const float k2 = f ? -(float)(3.0 / 2.0) : -3.0f, k3 = f ? (float)(1.0 / 3.0) : (float)(11.0 / 6.0);
    const float t = tap4<EXACT>(k0, k1, k2, k3, f ? a[3] : a[0], f ? a[4] : a[1], f ? a[5] : a[2], f ? a[6] : a[3]);
    float one;
    if (EXACT) one = div_dx(K, t);
    else one = Num<false>::mul(t, K.r_dx);
    return kind == CEN ? cen : one;
}

// First and second derivative along one axis of the edge-padded array (solve.py:29-31, 49-52, crop :61-65) at padded
// index P of an axis of n cells, anywhere on the axis: up[0..12] are the padded values P-6 .. P+6 (entries beyond the
// pad are never used by the formulas that apply).  Same values as fk_wide.h's wide_axis_general.
template <bool EXACT>
FK_HD void res_axis_general(const Consts& K, const float* up, int P, int n, float& d1, float& d2) {
    float g[7];   // first derivative at padded P-3 .. P+3
#pragma unroll
    for (int m = 0; m < 7; ++m) g[m] = res_deriv<EXACT>(K, kind_of(P - 3 + m, n, 1, 1), up + m);
    d1 = g[3];
    d2 = res_deriv<EXACT>(K, kind_of(P, n, 1, 1), g);
}

// The same two derivatives for a cell within 4 of a physical edge of a long axis (n >= 16), SPECIALISED by where the cell
// sits: side 0 = D cells from the low edge (padded P = D + 1), side 1 = D cells from the high edge (P = n - D), D = 0..3.
// The kinds of the first-pass derivatives the second pass reads are then compile-time facts, so only those 4-5 of the 7
// are computed, each by its own formula, without selects -- 5 four-tap sums instead of the 16 of res_axis_general (both
// variants of all 7 plus 2 for the second pass).  Same operations on the same operands: identical bits.  A thread's edge
// cell is the same every step, so the switch is perfectly predicted; it is warp-coherent where the tiles are at least a
// warp wide (the cells of a row sit in consecutive threads), and only there is it used: measured (B200, profiles/
// probe_cluster_r02.md) 512^2 3.46 -> 3.30 us per step, but 128^2 (tiles 16 wide: 4-8 cases per warp) 1.53 -> 1.84.
template <bool EXACT>
FK_HD float res_d_cen(const Consts& K, const float* a) { return dcen<EXACT>(K, a[1], a[2], a[4], a[5]); }
template <bool EXACT>
FK_HD float res_d_one(const Consts& K, bool f, const float* a) {   // f: forward (a[3..6]), else backward (a[0..3])
    const float t = f ? tap4<EXACT>((float)(-11.0 / 6.0), 3.0f, -(float)(3.0 / 2.0), (float)(1.0 / 3.0), a[3], a[4], a[5], a[6])
                      : tap4<EXACT>((float)(-1.0 / 3.0), (float)(3.0 / 2.0), -3.0f, (float)(11.0 / 6.0), a[0], a[1], a[2], a[3]);
    if (EXACT) return div_dx(K, t);
    return Num<false>::mul(t, K.r_dx);
}
template <bool EXACT, int SIDE, int D>
FK_HD void res_axis_edge_t(const Consts& K, const float* up, float& d1, float& d2) {
    float g[7];   // first derivative at padded P-3 .. P+3: only the entries the second pass reads
    if (D == 3) {                       // every formula central (the windows reach the pad: clamped loads)
        g[1] = res_d_cen<EXACT>(K, up + 1); g[2] = res_d_cen<EXACT>(K, up + 2); g[3] = res_d_cen<EXACT>(K, up + 3);
        g[4] = res_d_cen<EXACT>(K, up + 4); g[5] = res_d_cen<EXACT>(K, up + 5);
        d2 = res_d_cen<EXACT>(K, g);
    } else if (SIDE == 0) {
        if (D == 0) {                   // P = 1: first pass forward at padded 1, second pass forward
            g[3] = res_d_one<EXACT>(K, true, up + 3); g[4] = res_d_cen<EXACT>(K, up + 4); g[5] = res_d_cen<EXACT>(K, up + 5);
            g[6] = res_d_cen<EXACT>(K, up + 6);
            d2 = res_d_one<EXACT>(K, true, g);
        } else {                        // P = 2, 3: padded 0 and 1 forward
            g[1] = res_d_one<EXACT>(K, true, up + 1);
            g[2] = D == 1 ? res_d_one<EXACT>(K, true, up + 2) : res_d_cen<EXACT>(K, up + 2);
            g[3] = res_d_cen<EXACT>(K, up + 3); g[4] = res_d_cen<EXACT>(K, up + 4); g[5] = res_d_cen<EXACT>(K, up + 5);
            d2 = res_d_cen<EXACT>(K, g);
        }
    } else {
        if (D == 0) {                   // P = n: first pass backward at padded n, second pass backward
            g[0] = res_d_cen<EXACT>(K, up + 0); g[1] = res_d_cen<EXACT>(K, up + 1); g[2] = res_d_cen<EXACT>(K, up + 2);
            g[3] = res_d_one<EXACT>(K, false, up + 3);
            d2 = res_d_one<EXACT>(K, false, g);
        } else {                        // P = n - 1, n - 2: padded n and n + 1 backward
            g[1] = res_d_cen<EXACT>(K, up + 1); g[2] = res_d_cen<EXACT>(K, up + 2); g[3] = res_d_cen<EXACT>(K, up + 3);
            g[4] = D == 1 ? res_d_one<EXACT>(K, false, up + 4) : res_d_cen<EXACT>(K, up + 4);
            g[5] = res_d_one<EXACT>(K, false, up + 5);
            d2 = res_d_cen<EXACT>(K, g);
        }
    }
    d1 = g[3];
}
// code = D (low side) or 4 + D (high side)
template <bool EXACT>
FK_HD void res_axis_edge(const Consts& K, const float* up, int code, float& d1, float& d2) {
    switch (code) {
        case 0: res_axis_edge_t<EXACT, 0, 0>(K, up, d1, d2); break;
        case 1: res_axis_edge_t<EXACT, 0, 1>(K, up, d1, d2); break;
        case 2: res_axis_edge_t<EXACT, 0, 2>(K, up, d1, d2); break;
        case 4: res_axis_edge_t<EXACT, 1, 0>(K, up, d1, d2); break;
        case 5: res_axis_edge_t<EXACT, 1, 1>(K, up, d1, d2); break;
        case 6: res_axis_edge_t<EXACT, 1, 2>(K, up, d1, d2); break;
        default: res_axis_edge_t<EXACT, 0, 3>(K, up, d1, d2); break;
    }
}
// cell `x` of an axis of n >= 16 cells, within 4 of one of its ends -> code
FK_HD int res_edge_code(int x, int n) { return x < 4 ? x : 4 + (n - 1 - x); }

// one Euler step of an INTERIOR block of 2 rows x 4 cells (rows lr, lr + 1; every formula central, nothing published):
// the six u_x rows the two cells of a column need are computed once (6 + 2 first-pass/second-pass derivatives per
// column instead of 2 x (5 + 1)) and eight independent cells per thread hide more latency.
template <bool EXACT, bool MG>
FK_HD void res_block2(const TileArgs& A, const ResGeom& G, const ResCta& X, const float* cur, float* nxt, int lr, int lc,
                      unsigned mask, bool last) {
    const int W = A.W, P = G.pitch, row = X.r0 + lr, c = X.c0 + lc;
    const float* uc0 = cur + (lr + 4) * P + (lc + 4);
    float u_x[2][4], u_xx[2][4], uc[2][4];
    {   // ---- vertical: u_x at rows row-2 .. row+3 from rows row-4 .. row+5
        float ur[10][4];
#pragma unroll
        for (int j = 0; j < 10; ++j) unpack4(ld4(uc0 + (j - 4) * P), ur[j]);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float gx[6];
#pragma unroll
            for (int j = 0; j < 6; ++j) gx[j] = dcen<EXACT>(A.K, ur[j][k], ur[j + 1][k], ur[j + 3][k], ur[j + 4][k]);
            u_x[0][k] = gx[2]; u_x[1][k] = gx[3];
            u_xx[0][k] = dcen<EXACT>(A.K, gx[0], gx[1], gx[3], gx[4]);
            u_xx[1][k] = dcen<EXACT>(A.K, gx[1], gx[2], gx[4], gx[5]);
            uc[0][k] = ur[4][k]; uc[1][k] = ur[5][k];
        }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const float* ucr = uc0 + r * P;
        const int o = (lr + r) * G.tw_max + lc;
        const long long g = (long long)(row + r) * W + c;
        float Dv[4], DXv[4], DYv[4], v[4], w[4], stim[4] = {0.f, 0.f, 0.f, 0.f};
        if (MG) {
            unpack4(ldg4(A.D + X.boffD + g), Dv); unpack4(ldg4(A.DX + X.boffD + g), DXv); unpack4(ldg4(A.DY + X.boffD + g), DYv);
        } else {
            unpack4(ld4(X.Dm + o), Dv); unpack4(ld4(X.DXm + o), DXv); unpack4(ld4(X.DYm + o), DYv);
        }
        unpack4(ld4(X.V + o), v);
        unpack4(ld4(X.Wd + o), w);
        if (mask) {
#pragma unroll
            for (int k = 0; k < 4; ++k) stim[k] = res_stim(A, X, mask, g + k);
        }
        // ---- horizontal: u_y at columns c-2 .. c+5 from columns c-4 .. c+7, then u_yy
        float e[12];
        unpack4(ld4(ucr - 4), e);
        unpack4(ld4(ucr + 4), e + 8);
#pragma unroll
        for (int k = 0; k < 4; ++k) e[4 + k] = uc[r][k];
        float gy[8];
#pragma unroll
        for (int m = 0; m < 8; ++m) gy[m] = dcen<EXACT>(A.K, e[m], e[m + 1], e[m + 3], e[m + 4]);
        float un[4], vn[4], wn[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float u_yy = dcen<EXACT>(A.K, gy[k], gy[k + 1], gy[k + 3], gy[k + 4]);
            cell_step<EXACT>(A.K, uc[r][k], v[k], w[k], Dv[k], DXv[k], DYv[k], u_x[r][k], gy[k + 2], u_xx[r][k], u_yy, stim[k],
                             un[k], vn[k], wn[k]);
        }
        if (last) {
            st4(A.u_out + X.boff + g, un);
            st4(A.v_out + X.boff + g, vn);
            st4(A.w_out + X.boff + g, wn);
        } else {
            st4(nxt + (lr + r + 4) * P + (lc + 4), un);
            st4(X.V + o, vn);
            st4(X.Wd + o, wn);
        }
    }
}

// one Euler step of NC cells (row, c .. c+NC-1), local (lr, lc): reads `cur` (+ halo), writes `nxt`, v, w in place; ring
// groups (box != null) also publish the new u; the last step writes the caller's output instead
// GENERAL = false: the caller guarantees that every formula of the group is central (no cell within 4 of a physical
// edge) and only that path is compiled.
// MG: the diffusivity maps are read from global memory (L2 resident) instead of shared memory.
template <bool EXACT, int NC, bool GENERAL, bool MG, bool CL = false>
FK_HD void res_group(const TileArgs& A, const ResGeom& G, const ResCta& X, const float* cur, float* nxt, int lr, int lc,
                     unsigned mask, bool last, u64* box, unsigned tag) {
    const int H = A.H, W = A.W, P = G.pitch, row = X.r0 + lr, c = X.c0 + lc;
    const float* uc0 = cur + (lr + 4) * P + (lc + 4);
    float u_x[NC], u_y[NC], u_xx[NC], u_yy[NC], uc[NC];
    ldn<NC>(uc0, uc);
    const int o = lr * G.tw_max + lc;
    float v[NC], w[NC], Dv[NC], DXv[NC], DYv[NC], stim[NC];
    const long long g = (long long)row * W + c;
    if (MG) {   // requested first: the longest latency of the group
        ldgn<NC>(A.D + X.boffD + g, Dv);
        ldgn<NC>(A.DX + X.boffD + g, DXv);
        ldgn<NC>(A.DY + X.boffD + g, DYv);
    } else {
        ldn<NC>(X.Dm + o, Dv);
        ldn<NC>(X.DXm + o, DXv);
        ldn<NC>(X.DYm + o, DYv);
    }
    ldn<NC>(X.V + o, v);
    ldn<NC>(X.Wd + o, w);
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
        // padded rows P-6 .. P+6 = tissue rows row-6 .. row+6 clamped to the tissue (solve.py:31).  They all lie in the
        // tile's buffer: a tile with a neighbour above/below is >= 8 rows, so a cell within 4 rows of a physical edge
        // sits in the tile AT that edge, and 6 rows the other way is inside the tile or its 4-row halo.
        float ur[13][NC];
        const int lo = -row, hi = H - 1 - row;
#pragma unroll
        for (int j = 0; j < 13; ++j) ldn<NC>(uc0 + clampi(j - 6, lo, hi) * P, ur[j]);
#pragma unroll
        for (int k = 0; k < NC; ++k) {
            float col[13];
#pragma unroll
            for (int j = 0; j < 13; ++j) col[j] = ur[j][k];
            if (FK_RES_EDGE_SPECIAL && H >= 16 && G.tw_max >= 32) res_axis_edge<EXACT>(A.K, col, res_edge_code(row, H), u_x[k], u_xx[k]);
            else res_axis_general<EXACT>(A.K, col, row + 1, H, u_x[k], u_xx[k]);
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
#pragma unroll
        for (int k = 0; k < NC; ++k) {
            float win[13];   // padded columns Q-6 .. Q+6 of cell c + k, clamped like the rows above
            const int lo = -(c + k), hi = W - 1 - (c + k);
#pragma unroll
            for (int j = 0; j < 13; ++j) win[j] = uc0[k + clampi(j - 6, lo, hi)];
            if (FK_RES_EDGE_SPECIAL && W >= 16 && G.tw_max >= 32) res_axis_edge<EXACT>(A.K, win, res_edge_code(c + k, W), u_y[k], u_yy[k]);
            else res_axis_general<EXACT>(A.K, win, c + k + 1, W, u_y[k], u_yy[k]);
        }
    }
    float un[NC], vn[NC], wn[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) {
        cell_step<EXACT>(A.K, uc[k], v[k], w[k], Dv[k], DXv[k], DYv[k], u_x[k], u_y[k], u_xx[k], u_yy[k], stim[k], un[k], vn[k],
                         wn[k]);
    }
    if (last) {
        stn<NC>(A.u_out + X.boff + g, un);
        stn<NC>(A.v_out + X.boff + g, vn);
        stn<NC>(A.w_out + X.boff + g, wn);
    } else {
        if (box) {
            if (CL) { if (!(G.spin_limit & 1u)) res_publish_cl<NC>(G, X, (long long)(nxt - X.U0), lr, lc, un); }
            else res_publish<NC>(G, X, box, lr, lc, un, tag);
        }
        stn<NC>(nxt + (lr + 4) * P + (lc + 4), un);
        stn<NC>(X.V + o, vn);
        stn<NC>(X.Wd + o, wn);
    }
}

// ------------------------------------------------------------------ the TWO-PASS step (ResGeom::tp, nc = 4)
// res_group rebuilds every first derivative a cell needs from the u window: 9 four-tap sums per cell where the reference's
// two passes need 4, and 16 shared-memory loads per group.  Here the step is done the way the reference does it -- gradient,
// then gradient of the gradient (solve.py:49-52) -- with the first pass kept in shared memory:
//   pass A   u_x of tile rows -2 .. th+1 and u_y of tile columns -2 .. tw+1 (what the second pass of the tile's cells
//            reads; the apron comes from the u halo) -> GX, GY; the cells within 4 of a PHYSICAL edge are done in this
//            phase too, one per thread through the general formulas (they read u only);
//   barrier
//   pass B   per group of 4 cells: u_xx from 5 GX rows, u_yy from GY, then the cell update -- stream_emit, the streaming
//            kernel's own (two cells per instruction in fast numerics); ring groups first (published), then the interior.
// Same operations on the same operands as res_group: identical bits.  Per cell ~5 four-tap sums and ~6 16-byte loads.
FK_HD int res_grad_items(const ResCta& X) { return (X.th + 4) * (X.tw >> 2) + 2 * X.th; }

template <bool EXACT>
FK_HD void res_grad_item(const TileArgs& A, const ResGeom& G, const ResCta& X, const float* cur, int i) {
    const int q4 = X.tw >> 2, nmain = (X.th + 4) * q4, P = G.pitch;
    if (i < nmain) {
        const int rr = i / q4, lr = rr - 2, lc = 4 * (i - rr * q4);
        const int R = X.r0 + lr;
        if (R < 0 || R >= A.H) return;
        const float* uc = cur + (lr + 4) * P + (lc + 4);
        if (R >= 2 && R <= A.H - 3) {   // central along the rows: solve.py:49 at a row whose window is inside the tissue
            float a0[4], a1[4], a3[4], a4[4], gx[4];
            unpack4(ld4(uc - 2 * P), a0); unpack4(ld4(uc - P), a1); unpack4(ld4(uc + P), a3); unpack4(ld4(uc + 2 * P), a4);
            dcen_rows4<EXACT>(A.K, a0, a1, a3, a4, gx);
            st4(X.GX + (lr + 2) * G.tw_max + lc, gx);
        }
        if (lr >= 0 && lr < X.th) {     // solve.py:50 (groups at a physical left / right edge: their two outer values are
            float e[12], gy[4];         //  never read -- the cells that would read them go through the general formulas)
            unpack4(ld4(uc - 4), e); unpack4(ld4(uc), e + 4); unpack4(ld4(uc + 4), e + 8);
            dcen_span4<EXACT>(A.K, e + 2, gy);
            st4(X.GY + lr * P + (lc + 4), gy);
        }
        return;
    }
    // the two u_y values beyond each side of the tile (only where a neighbour's halo is there)
    const int j = i - nmain, lr = j >> 1, right = j & 1;
    if (right ? !X.has_e : !X.has_w) return;
    const float* uc = cur + (lr + 4) * P + 4 + (right ? X.tw - 2 : -4);   // e[0..5]: columns -4 .. 1, or tw-2 .. tw+3
    const F2 a = ld2(uc), b = ld2(uc + 2), c = ld2(uc + 4);
    F2 g;
    g.x = dcen<EXACT>(A.K, a.x, a.y, b.y, c.x);
    g.y = dcen<EXACT>(A.K, a.y, b.x, c.x, c.y);
    *reinterpret_cast<F2*>(X.GY + lr * P + 4 + (right ? X.tw : -2)) = g;
}

template <bool EXACT, bool MG>
FK_HD void res_group2p(const TileArgs& A, const ResGeom& G, const ResCta& X, const float* cur, float* nxt, int lr, int lc,
                       unsigned mask, bool last, u64* box, unsigned tag) {
    const int P = G.pitch, TW = G.tw_max, W = A.W, row = X.r0 + lr, c = X.c0 + lc;
    const int o = lr * TW + lc;
    const long long g = (long long)row * W + c;
    float Dv[4], DXv[4], DYv[4], v[4], w[4], stim[4] = {0.f, 0.f, 0.f, 0.f};
    if (MG) {   // requested first: the longest latency of the group
        unpack4(ldg4(A.D + X.boffD + g), Dv); unpack4(ldg4(A.DX + X.boffD + g), DXv); unpack4(ldg4(A.DY + X.boffD + g), DYv);
    } else {
        unpack4(ld4(X.Dm + o), Dv); unpack4(ld4(X.DXm + o), DXv); unpack4(ld4(X.DYm + o), DYv);
    }
    const float* gxp = X.GX + (lr + 2) * TW + lc;
    float gm2[4], gm1[4], g0[4], gp1[4], gp2[4], e[12], u0[4];
    unpack4(ld4(gxp - 2 * TW), gm2); unpack4(ld4(gxp - TW), gm1); unpack4(ld4(gxp), g0);
    unpack4(ld4(gxp + TW), gp1); unpack4(ld4(gxp + 2 * TW), gp2);
    const float* gyp = X.GY + lr * P + (lc + 4);
    unpack4(ld4(gyp - 4), e); unpack4(ld4(gyp), e + 4); unpack4(ld4(gyp + 4), e + 8);
    unpack4(ld4(cur + (lr + 4) * P + (lc + 4)), u0);
    unpack4(ld4(X.V + o), v);
    unpack4(ld4(X.Wd + o), w);
    float un[4], vn[4], wn[4];
    if (mask) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int k = 0; k < 4; ++k) stim[k] = res_stim(A, X, mask, g + k);
        stream_emit<EXACT, true, false>(A.K, u0, v, w, gm2, gm1, g0, gp1, gp2, e + 2, e + 4, Dv, DXv, DYv, stim, false, false, un, vn, wn);
    } else {
        stream_emit<EXACT, false, false>(A.K, u0, v, w, gm2, gm1, g0, gp1, gp2, e + 2, e + 4, Dv, DXv, DYv, stim, false, false, un, vn, wn);
    }
    if (last) {
        st4(A.u_out + X.boff + g, un);
        st4(A.v_out + X.boff + g, vn);
        st4(A.w_out + X.boff + g, wn);
    } else {
        if (box) res_publish<4>(G, X, box, lr, lc, un, tag);
        st4(nxt + (lr + 4) * P + (lc + 4), un);
        st4(X.V + o, vn);
        st4(X.Wd + o, wn);
    }
}

// pass A of step s: the gradient items, then the cells at a physical edge (general formulas, published like ring cells)
template <bool EXACT, bool MG>
FK_HD void res_phase_a(const TileArgs& A, const ResGeom& G, const ResCta& X, int s, unsigned mask, int tid, int nthr) {
    const float* cur = X.U0 + (s & 1) * X.nu;
    float* nxt = X.U0 + ((s + 1) & 1) * X.nu;
    const bool last = s == G.nsteps - 1;
    u64* box = last ? nullptr : G.xchg + ((s + 1) & 1) * X.pstride + X.mbox;
    const unsigned tag = (unsigned)(s + 1);
    const int ng = res_grad_items(X), n = ng + X.nedge;
    for (int i = tid; i < n; i += nthr) {
        if (i < ng) res_grad_item<EXACT>(A, G, X, cur, i);
        else {
            int lr, lc;
            res_locate_edge(X, i - ng, lr, lc);
            res_group<EXACT, 1, true, MG>(A, G, X, cur, nxt, lr, lc, mask, last, box, tag);
        }
    }
}

// item i of a phase -> its cell(s) and what to do with them: 1 = a group of NC cells whose formulas are all central,
// 2 = ONE cell through the general formulas, 0 = nothing.  Phase 0 = the ring (what the neighbours wait for), published
// as it is computed; phase 1 = the interior, computed while those records travel.  With NC > 1 the groups that touch a
// physical edge are skipped and their cells done ONE PER THREAD as extra items of phase 0: the one-sided formulas cost
// several times the central ones and the step's critical path is the slowest thread of the slowest tile.
template <int NC>
FK_HD int res_item(const TileArgs& A, const ResGeom& G, const ResCta& X, int phase, int i, int& lr, int& lc) {
    const int n = phase ? X.ninner : X.nring;
    if (G.tp && i >= n) return 0;   // two-pass form: the cells at a physical edge belong to the gradient phase
    if (NC == 1) {
        if (i >= n) return 0;
        res_locate(G, X, phase, i, lr, lc);
        return 2;
    }
    if (i < n) {
        res_locate(G, X, phase, i, lr, lc);
        if (FK_RES_R2 && phase && i < X.npair) return 3;   // a 2-row block of the interior: central by construction
        const int row = X.r0 + lr, c = X.c0 + lc;
        return (row >= 4 && row + 5 <= A.H && c >= 4 && c + NC + 4 <= A.W) ? 1 : 0;
    }
    if (phase || i >= n + X.nedge) return 0;
    res_locate_edge(X, i - n, lr, lc);
    return 2;
}

template <int NC>
FK_HD void res_thread_setup(const TileArgs& A, const ResGeom& G, const ResCta& X, int tid, int nthr, ResThread& T) {
    T.lr0 = T.lc0 = T.lr1 = T.lc1 = 0;
    T.ty0 = res_item<NC>(A, G, X, 0, tid, T.lr0, T.lc0);
    T.ty1 = res_item<NC>(A, G, X, 1, tid, T.lr1, T.lc1);
    const int U = X.nhalo[0] + X.nhalo[1] + X.nhalo[2] + X.nhalo[3];
    T.hs0 = T.hs1 = 0; T.hd0 = T.hd1 = 0;
    if (tid < U) res_halo_unit(G, X, tid, T.hs0, T.hd0);
    if (tid + nthr < U) res_halo_unit(G, X, tid + nthr, T.hs1, T.hd1);
}

// one phase of step s, its items strided over the CTA's threads
// CL: the cluster transport (ring cells stored into the neighbours' halos) -- its own instantiation, so that the mailbox
// kernels carry none of it (as a run-time branch it cost them 7-14 %: profiles/probe_cluster_r02.md)
template <bool EXACT, int NC, bool MG, bool CL = false, bool TP = false>
FK_HD void res_phase(const TileArgs& A, const ResGeom& G, const ResCta& X, const ResThread& T, int s, int phase,
                     unsigned mask, int tid, int nthr) {   // (called from ONE site, in a phase loop: a single copy)
    const float* cur = X.U0 + (s & 1) * X.nu;
    float* nxt = X.U0 + ((s + 1) & 1) * X.nu;
    const bool last = s == G.nsteps - 1;
    u64* box = (last || phase) ? nullptr : (CL ? reinterpret_cast<u64*>(X.U0)   // (non-null: "publish")
                                               : G.xchg + ((s + 1) & 1) * X.pstride + X.mbox);
    const unsigned tag = (unsigned)(s + 1);
    const int n = phase ? X.ninner : ((NC == 1 || TP) ? X.nring : X.nring + X.nedge);
    for (int i = tid; i < n; i += nthr) {
        int lr = phase ? T.lr1 : T.lr0, lc = phase ? T.lc1 : T.lc0, ty = phase ? T.ty1 : T.ty0;
        if (i != tid) ty = res_item<NC>(A, G, X, phase, i, lr, lc);
        if (ty == 2) res_group<EXACT, 1, true, MG, CL>(A, G, X, cur, nxt, lr, lc, mask, last, box, tag);
        else if (TP && NC == 4 && ty == 1) res_group2p<EXACT, MG>(A, G, X, cur, nxt, lr, lc, mask, last, box, tag);
        else if (NC > 1 && ty == 1) res_group<EXACT, NC, false, MG, CL>(A, G, X, cur, nxt, lr, lc, mask, last, box, tag);
        else if (FK_RES_R2 && NC == 4 && ty == 3) res_block2<EXACT, MG>(A, G, X, cur, nxt, lr, lc, mask, last);
    }
}

// ------------------------------------------------------------------ launch geometry (host)
struct ResPlan {
    ResGeom G;
    int threads;
    long long smem_bytes;
    long long xchg_bytes;   // mailboxes of the launch (both parities)
};

#ifndef FK_RES_THREADS_CAP
#define FK_RES_THREADS_CAP 512   // development: 1024 for an A/B build (64 registers per thread)
#endif
enum { FK_RES_MAX_THREADS = FK_RES_THREADS_CAP };


// thread slots the busiest phase of a tile walks through
inline int res_tile_items(int H, int W, const ResGeom& G, int tile) {
    ResCta X;
    res_tile_geom(H, W, G, tile, X);
    const int p0 = G.nc == 1 ? X.nring : X.nring + X.nedge;
    return p0 > X.ninner ? p0 : X.ninner;
}

// work of a tile in central-cell units: a cell at a physical edge (one-sided formulas, one cell per thread) costs ~6
inline double res_tile_work(int H, int W, const ResGeom& G, int tile) {
    ResCta X;
    res_tile_geom(H, W, G, tile, X);
    return (double)X.th * X.tw + 5.0 * X.nedge;
}

// Tiles for a (batch, H, W) problem on `capacity` co-resident CTAs with `smem_limit` bytes each and `xchg_limit` bytes
// of mailboxes.  A tile next to another tile is at least 8 cells wide/tall (a halo comes from ONE tile, and the edge
// formulas reach 7 cells into the tissue).  Distilled from sweeps on B200 (profiles/probe_resident_r01.log):
//  * cells per thread: 2 up to 1024-cell tiles (such tiles are latency bound: every lane busy), 4 above (fewest
//    instructions); 1 only for tiles of a few dozen cells;
//  * small tiles (nc < 4): the step time follows th * tw + 6 (th + tw) of the largest tile -- its cells plus its
//    perimeter (ring work, halo records) --, grids that divide the tissue evenly are ~15 % faster than ragged ones,
//    and if everything fits one round of <= 512 threads there is no interior phase at all;
//  * large tiles (nc = 4) are issue bound and the step lasts as long as the slowest tile: the tiles at the tissue's
//    edges, whose one-sided cells cost ~6 central ones, get f = 1, .85, .7, .55 or .45 of the others' rows / columns,
//    whichever evens out the work of corner, edge and interior tiles best within shared memory.
// force_eh / force_ewq: > 0 that many rows / column groups in the edge tiles, < 0 even split, 0 planner's choice.
// maps_global: 1 = keep the diffusivity maps in global memory (L2), 0 = in shared memory if any plan fits, else global.
inline bool plan_resident(int H, int W, int batch, int capacity, long long smem_limit, long long xchg_limit, int force_ntr,
                          int force_ntc, int force_threads, int force_nc, int force_eh, int force_ewq, int maps_global,
                          ResPlan& P, int one_pass_smem_maps_only = 0) {
    if (W % 4 != 0 || H < 3 || W < 4 || batch < 1) return false;
    // the last plan is kept: a run calls this once per segment with the same problem
    struct Memo { int key[8]; bool ok; ResPlan plan; };
    static Memo memo = {{-1, 0, 0, 0, 0, 0, 0, 0}, false, ResPlan()};
    const int key[8] = {H, W, batch, capacity, force_ntr * 4096 + force_ntc, force_eh * 4096 + force_ewq,
                        force_threads * 32 + force_nc * 4 + (maps_global ? 2 : 0) + (one_pass_smem_maps_only ? 1 : 0),
                        (int)(smem_limit / 64) + (int)((xchg_limit / 4096) % 1000003)};
    bool same = true;
    for (int i = 0; i < 8; ++i) same = same && memo.key[i] == key[i];
    if (same) { if (memo.ok) P = memo.plan; return memo.ok; }
    const int Q = W >> 2;
    static const double fs[5] = {1.0, 0.85, 0.7, 0.55, 0.45};
    double best = 1e300;
    bool found = false;
    // Plans with the maps in shared memory and plans with the maps in L2 compete on cost (the latter 2 % dearer); nc = 4
    // tiles take the two-pass step when its first derivatives fit shared memory as well, and pay for it when they do
    // not (the one-pass body spends ~2.5 x the instructions per cell) -- so 1024^2 moves its maps to L2 to get it.
    // FK_RES_TWO_PASS=0 (development) switches the two-pass form off.
    for (int mg = maps_global ? 1 : 0; mg < (one_pass_smem_maps_only ? 1 : 2); ++mg)
    for (int ntr = 1; ntr <= H; ++ntr) {
        if (ntr > 1 && H / ntr < 8) break;
        if (force_ntr > 0 && ntr != force_ntr) continue;
        for (int ntc = 1; ntc <= Q; ++ntc) {
            if (ntc > 1 && Q / ntc < 2) break;
            if ((long long)ntr * ntc * batch > capacity) break;
            if (force_ntc > 0 && ntc != force_ntc) continue;
            const int th_even = (H + ntr - 1) / ntr, tw_even = 4 * ((Q + ntc - 1) / ntc);
            int nc = force_nc;
            if (nc != 1 && nc != 2 && nc != 4) nc = th_even * tw_even <= 64 ? 1 : (th_even * tw_even <= 1024 ? 2 : 4);
            const bool forced_edge = force_eh != 0 || force_ewq != 0;
            for (int k = 0; k < ((nc == 4 && !forced_edge) ? 5 : 1); ++k) {
                ResGeom G = ResGeom();
                G.ntr = ntr; G.ntc = ntc; G.nc = nc; G.mg = mg; G.r2 = (nc == 4 && FK_RES_R2) ? 1 : 0;
                if (forced_edge) {
                    if (force_eh > 0 && ntr >= 3) G.eh = force_eh;
                    if (force_ewq > 0 && ntc >= 3) G.ewq = force_ewq;
                } else if (k) {
                    if (ntr >= 3) G.eh = (int)(fs[k] * H / (ntr - 2 + 2 * fs[k]) + 0.5);
                    if (ntc >= 3) G.ewq = (int)(fs[k] * Q / (ntc - 2 + 2 * fs[k]) + 0.5);
                    if (!G.eh && !G.ewq) break;
                }
                if (G.eh && (G.eh < 8 || (H - 2 * G.eh) / (ntr - 2) < 8)) continue;
                if (G.ewq && (G.ewq < 2 || (Q - 2 * G.ewq) / (ntc - 2) < 2)) continue;
                int th = 0, tw = 0;
                for (int t = 0; t < ntr; ++t) { const int h = res_split(H, ntr, G.eh, t + 1) - res_split(H, ntr, G.eh, t); if (h > th) th = h; }
                for (int t = 0; t < ntc; ++t) { const int w = 4 * (res_split(Q, ntc, G.ewq, t + 1) - res_split(Q, ntc, G.ewq, t)); if (w > tw) tw = w; }
                const int tp = (FK_RES_TWO_PASS && !one_pass_smem_maps_only && nc == 4 && res_smem_floats(th, tw, mg, 1) * 4 <= smem_limit) ? 1 : 0;
                const long long smem = res_smem_floats(th, tw, mg, tp) * 4;
                if (smem > smem_limit) continue;
                G.tp = tp;
                G.th_max = th; G.tw_max = tw; G.pitch = tw + 8; G.slots = 8 * tw + 8 * th;
                const long long xbytes = 2LL * batch * ntr * ntc * G.slots * (long long)sizeof(u64);
                if (xbytes > xchg_limit) continue;
                // corner, top-edge, left-edge and interior tile (those that exist)
                const int rows[2] = {0, ntr > 2 ? 1 : ntr - 1}, cols[2] = {0, ntc > 2 ? 1 : ntc - 1};
                double cost;
                if (nc == 4) {
                    cost = 0;
                    for (int a = 0; a < 2; ++a)
                        for (int b = 0; b < 2; ++b) {
                            const double w = res_tile_work(H, W, G, rows[a] * ntc + cols[b]);
                            if (w > cost) cost = w;
                        }
                    cost += 6.0 * (th + tw);
                } else {
                    cost = (double)th * tw + 6.0 * (th + tw);
                    if (H % ntr != 0 || Q % ntc != 0) cost *= 1.15;
                }
                if (nc == 4 && !G.tp && FK_RES_TWO_PASS && !one_pass_smem_maps_only) cost *= 1.8;
                if (mg) cost *= 1.02;
                cost += 0.01 * th + 1e-3 * ntr * ntc;
                if (cost >= best) continue;
                best = cost;
                found = true;
                // one round where <= 512 threads allow it; if even ring + interior + edge items together fit: one phase
                G.single = 1;
                int all = 0, two = 0;
                for (int a = 0; a < 2; ++a)
                    for (int b = 0; b < 2; ++b) {
                        G.single = 1;
                        const int i1 = res_tile_items(H, W, G, rows[a] * ntc + cols[b]);
                        G.single = 0;
                        const int i2 = res_tile_items(H, W, G, rows[a] * ntc + cols[b]);
                        if (i1 > all) all = i1;
                        if (i2 > two) two = i2;
                    }
                {   // the largest tile (tiles of a class differ by a row / a column group)
                    const int q = tw / nc, qe = 4 / nc, g = th * q;
                    int inner = (th > 8 && tw > 8) ? (th - 8) * (q - 2 * qe) : 0;
                    const int ring = g - inner;
                    if (G.r2) inner = ((th - 8) / 2 + ((th - 8) & 1)) * (q - 2 * qe);
                    if (g > all) all = g;
                    if (ring > two) two = ring;
                    if (inner > two) two = inner;
                }
                G.single = all <= FK_RES_MAX_THREADS;
                int threads = force_threads;
                if (threads <= 0) {
                    threads = ((G.single ? all : two) + 31) / 32 * 32;
                    if (threads < 64) threads = 64;
                    if (threads > FK_RES_MAX_THREADS) threads = FK_RES_MAX_THREADS;
                } else if (all > threads) {
                    G.single = 0;
                }
                P.G = G;
                P.threads = threads;
                P.smem_bytes = smem;
                P.xchg_bytes = xbytes;
            }
        }
    }
    for (int i = 0; i < 8; ++i) memo.key[i] = key[i];
    memo.ok = found;
    if (found) memo.plan = P;
    return found;
}

// Tiles of ONE tissue for the cluster transport: at most FK_CLUSTER_MAX CTAs (the non-portable cluster size of sm_100),
// maps in shared memory, no mailboxes.  The same planner, told that the machine has 16 SMs.
enum { FK_CLUSTER_MAX = 16 };
inline bool plan_cluster(int H, int W, long long smem_limit, int force_ntr, int force_ntc, int force_threads, int force_nc,
                         ResPlan& P) {
    if (force_ntr > 0 && force_ntc > 0 && force_ntr * force_ntc > FK_CLUSTER_MAX) return false;
    if (!plan_resident(H, W, 1, FK_CLUSTER_MAX, smem_limit, 1LL << 40, force_ntr, force_ntc, force_threads, force_nc, -1, -1, 0, P,
                       /*one_pass_smem_maps_only=*/1))   // (the cluster kernel has the one-pass body, maps in shared memory)
        return false;
    P.G.cluster = 1;
    P.xchg_bytes = 0;
    return true;
}

// ------------------------------------------------------------------ CPU emulation of one launch (tests/emu)
#if !defined(__CUDACC__)
}  // namespace fk
#include <vector>
namespace fk {
template <bool EXACT>
inline void emu_res_phase(const TileArgs& A, const ResGeom& G, const ResCta& X, const ResThread& T, int s, int phase,
                          unsigned mask) {
    if (G.tp) {
        if (G.mg) res_phase<EXACT, 4, true, false, true>(A, G, X, T, s, phase, mask, 0, 1);
        else res_phase<EXACT, 4, false, false, true>(A, G, X, T, s, phase, mask, 0, 1);
    } else if (G.cluster) {
        if (G.nc == 1) res_phase<EXACT, 1, false, true>(A, G, X, T, s, phase, mask, 0, 1);
        else if (G.nc == 2) res_phase<EXACT, 2, false, true>(A, G, X, T, s, phase, mask, 0, 1);
        else res_phase<EXACT, 4, false, true>(A, G, X, T, s, phase, mask, 0, 1);
    } else if (G.mg) {
        if (G.nc == 1) res_phase<EXACT, 1, true>(A, G, X, T, s, phase, mask, 0, 1);
        else if (G.nc == 2) res_phase<EXACT, 2, true>(A, G, X, T, s, phase, mask, 0, 1);
        else res_phase<EXACT, 4, true>(A, G, X, T, s, phase, mask, 0, 1);
    } else {
        if (G.nc == 1) res_phase<EXACT, 1, false>(A, G, X, T, s, phase, mask, 0, 1);
        else if (G.nc == 2) res_phase<EXACT, 2, false>(A, G, X, T, s, phase, mask, 0, 1);
        else res_phase<EXACT, 4, false>(A, G, X, T, s, phase, mask, 0, 1);
    }
}
// CTAs advance in lock step, phase by phase, which is one legal interleaving of the mailbox protocol; shared memory is
// poisoned with NaN so that a read of a halo nobody filled shows up in the result, and a record that has not arrived
// when its consumer looks for it is an error (-7).
inline int emu_resident_launch(const ResPlan& P, const TileArgs& A, int batch, int exact) {
    ResGeom G = P.G;
    const int ntiles = G.ntr * G.ntc;
    const long long floats = res_smem_floats(G.th_max, G.tw_max, G.mg, G.tp);
    std::vector<std::vector<float>> smem((size_t)ntiles * batch, std::vector<float>((size_t)floats, __builtin_nanf("")));
    std::vector<u64> xchg((size_t)(P.xchg_bytes / sizeof(u64)), 0ull);
    G.xchg = xchg.data();
    std::vector<ResCta> X((size_t)ntiles * batch);
    std::vector<ResThread> T((size_t)ntiles * batch);
    for (int sim = 0; sim < batch; ++sim)
        for (int t = 0; t < ntiles; ++t) {
            ResCta& x = X[(size_t)sim * ntiles + t];
            res_setup(A, G, t, sim, batch, smem[(size_t)sim * ntiles + t].data(), x);
            if (G.cluster) {   // the neighbours' shared memory, as cluster.map_shared_rank gives it on the device
                if (x.has_n) x.peer[0] = smem[(size_t)sim * ntiles + t - G.ntc].data();
                if (x.has_s) x.peer[1] = smem[(size_t)sim * ntiles + t + G.ntc].data();
                if (x.has_w) x.peer[2] = smem[(size_t)sim * ntiles + t - 1].data();
                if (x.has_e) x.peer[3] = smem[(size_t)sim * ntiles + t + 1].data();
            }
            res_load(A, G, x, 0, 1);
            ResThread& th = T[(size_t)sim * ntiles + t];
            if (G.nc == 1) res_thread_setup<1>(A, G, x, 0, 1, th);
            else if (G.nc == 2) res_thread_setup<2>(A, G, x, 0, 1, th);
            else res_thread_setup<4>(A, G, x, 0, 1, th);
        }
    for (int s = 0; s < G.nsteps; ++s) {
        for (size_t i = 0; i < X.size(); ++i) {
            const unsigned mask = res_mask(A, X[i], s);
            if (G.tp) {
                if (exact) { if (G.mg) res_phase_a<true, true>(A, G, X[i], s, mask, 0, 1); else res_phase_a<true, false>(A, G, X[i], s, mask, 0, 1); }
                else { if (G.mg) res_phase_a<false, true>(A, G, X[i], s, mask, 0, 1); else res_phase_a<false, false>(A, G, X[i], s, mask, 0, 1); }
            }
            for (int phase = 0; phase < 2; ++phase) {
                if (exact) emu_res_phase<true>(A, G, X[i], T[i], s, phase, mask);
                else emu_res_phase<false>(A, G, X[i], T[i], s, phase, mask);
            }
        }
        if (s == G.nsteps - 1 || G.cluster) continue;   // (cluster transport: the halos were stored by their producers)
        for (size_t i = 0; i < X.size(); ++i)
            if (!res_halo(G, X[i], T[i], s, 0, 1)) return -7;
    }
    return 0;
}
#endif

}  // namespace fk
