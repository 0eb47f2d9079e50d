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
// The kernel owns all four physical edges (one launch per T steps, nothing else on the hot path): the edge thread of a
// strip applies the reference's one-sided formulas on the padded columns, the first / last row chunk on the padded rows.
// Strip edges: columns within 4s of a strip edge are garbage at level s and are never stored.
//
// Written as FK_HD code over an explicit per-thread state so tests/emu can run it on the CPU
// (threads of a block executed one after the other inside each iteration).
#pragma once
#include <cstdlib>
#include "fk_core.h"
#include "fk_tile.h"

namespace fk {

struct alignas(8) F2 { float x, y; };
struct alignas(16) F4 { float x, y, z, w; };
FK_HD F2 ld2(const float* p) { return *reinterpret_cast<const F2*>(p); }
FK_HD F4 ld4(const float* p) { return *reinterpret_cast<const F4*>(p); }
// read-only global data (diffusivity maps, stimulus fields): non-coherent path on the device
#ifndef FK_MAP_PF
#define FK_MAP_PF 3
#endif
#ifndef FK_MAP_LATE
#define FK_MAP_LATE 1
#endif
// which stages load their maps at the point of use: 0 none (all up front), 1 all, 2 all but the first, 3 the first only
#define FK_MAP_IS_LATE(s) (FK_MAP_LATE == 1 || (FK_MAP_LATE == 2 && (s) > 0) || (FK_MAP_LATE == 3 && (s) == 0))
#ifndef FK_MAP_PF1
#define FK_MAP_PF1 0
#endif
FK_HD void pf_l1(const float* p) {
#if defined(__CUDA_ARCH__)
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#else
    (void)p;
#endif
}
FK_HD void pf_l2(const float* p) {
#if defined(__CUDA_ARCH__)
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
    (void)p;
#endif
}

FK_HD F4 ldg4(const float* p) {
#if defined(__CUDA_ARCH__)
    const float4 t = __ldg(reinterpret_cast<const float4*>(p));
    F4 r; r.x = t.x; r.y = t.y; r.z = t.z; r.w = t.w;
    return r;
#else
    return *reinterpret_cast<const F4*>(p);
#endif
}
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
// keeps a loop-carried offset in its register (the compiler would otherwise rebuild it from the loop counter with a
// 64-bit multiply at every use)
FK_HD long long opaque(long long x) {
#if defined(__CUDA_ARCH__)
    asm volatile("" : "+l"(x));
#endif
    return x;
}
FK_HD int opaque_i(int x) {   // the same for a 32-bit value (stops the compiler from rebuilding it from threadIdx inside the loop)
#if defined(__CUDA_ARCH__)
    asm volatile("" : "+r"(x));
#endif
    return x;
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

// Split block barrier on an mbarrier in shared memory (steady-state loop): a thread ARRIVES when its shared-memory
// writes of an iteration are done and WAITS, in the next iteration, only just before it reads what the others wrote --
// its own columns are loaded and the halo-free arithmetic issued in between.  (The CPU emulation runs the threads of an
// iteration one after the other, so these are no-ops there.)
FK_HD void sb_init(float* bar, int count) {
#if defined(__CUDA_ARCH__)
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
#endif
}
FK_HD void sb_arrive(float* bar) {
#if defined(__CUDA_ARCH__)
    asm volatile("{\n\t.reg .b64 t;\n\tmbarrier.arrive.shared::cta.b64 t, [%0];\n\t}" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
#endif
}
FK_HD void sb_wait(float* bar, int parity) {
#if defined(__CUDA_ARCH__)
    asm volatile(
        "{\n\t.reg .pred p;\n"
        "FK_SB_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra FK_SB_DONE_%=;\n\t"
        "bra FK_SB_WAIT_%=;\n"
        "FK_SB_DONE_%=:\n\t}" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
#endif
}

// non-blocking look at the barrier, issued at the top of an iteration so that its latency overlaps the loads of the
// thread's own columns; sb_wait is only entered when the phase was not complete yet
FK_HD int sb_test(float* bar, int parity) {
#if defined(__CUDA_ARCH__)
    int done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.s32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
    return done;
#else
    return 1;
#endif
}

enum { FK_PF = 3 };      // level-0 rows are fetched this many iterations ahead

// Shared memory has two regions.
//
// LINE-major blocks for what cp.async brings in from global memory (level 0): every group of 8 consecutive threads owns a
// block of 16 rows x 128 bytes -- rows 0..7 the U(0) ring, 8..11 V(0), 12..15 W(0) -- in which lane l of the group holds
// its 4 columns at byte 16 l of every row.  A warp's LDGSTS therefore lands in four whole 128-byte lines and asks L2 for
// every sector once.  (With thread-major destinations, 528 bytes apart, every thread's 16 bytes became a request of their
// own: 2x the sectors from L2 and ~28 shared-memory wavefronts per instruction -- ncu source page of round 2.)  Rows are a
// compile-time 128 bytes apart whatever the CTA's size.  The neighbours' halo values of a level-0 row sit one granule to
// the left / right, in the adjacent block for the first / last lane of a group (StreamMem::hl, hr); the CTA's first and
// last thread read their own granule instead (garbage that only reaches columns which are never stored).
//
// THREAD-major chunks for what the threads write themselves (levels >= 1, u_y): every thread owns one chunk of
// StreamLay<T>::CHUNK floats that holds its 4 columns (one 16-byte granule) of every ring row, so that each access is
// `chunk base + compile-time offset` -- the neighbours' halo values sit at +-CHUNK.  CHUNK is an odd number of granules,
// which keeps the 16-byte accesses of consecutive threads on distinct banks.  One pad chunk on each side of the CTA
// absorbs the halo reads of its first and last thread (garbage that only reaches columns which are never stored).
//
// Ring slots are functions of the CTA-local iteration index i alone, so that the steady-state loop, unrolled by 4
// with the phase i & 3 as a template parameter, addresses every slot with a compile-time offset:
//   U(0)     8 slots: level-0 row rin0 + i (the newest input of stage 0 at iteration i) sits in slot i & 7; it is
//            fetched at iteration i - FK_PF.
//   U(s>=1)  4 slots: the level-s row stage s receives at iteration i goes to slot i & 3 -- the slot of the row it read
//            as its window's oldest row (rho) a moment before; only row rho + 1 is ever read by the neighbours.
//   V/W(0)   4 slots: level-0 v, w of the row stage 0 emits at iteration i: slot i & 3, fetched at iteration i - FK_PF.
//   V/W(s)   4 slots: level-s v, w emitted by stage s-1 at iteration i: slot i & 3, consumed by stage s at iteration
//            i + 4 from the same slot BEFORE stage s-1 overwrites it (the loads of every stage are hoisted).
//   GY(s)    2 slots: u_y of row rho + 1, written at iteration i into slot i & 1 and read at iteration i + 1.
template <int T>
struct StreamLay {
    enum {
        ROWS = 16, ROW_U = 0, ROW_V = 8, ROW_W = 12, PITCH = 32, BLOCK = 16 * 32,   // the line-major region (floats)
        U1 = 0,                          // U(s) = U1 + 16 (s - 1) for s >= 1
        GY = 16 * (T - 1),               // GY(s) = GY + 8 s
        V = GY + 8 * T - 16,             // V(s) = V + 16 s for s >= 1
        W = V + 16 * (T - 1),            // W(s) = W + 16 s for s >= 1
        CHUNK = W + 16 * T + 4           // + one pad granule: (14 T - 11) granules, odd
    };
    static FK_HD int U(int s) { return (int)U1 + 16 * (s - 1); }   // s >= 1
};

// a thread's view of the CTA's shared memory
struct StreamMem {
    float* tb;    // its thread-major chunk
    float* rb;    // its granule of row 0 of its group's line-major block
    int hl, hr;   // where the left / right neighbour's granule of a level-0 row sits, relative to its own (floats)
};
FK_HD float* row_at(const StreamMem& M, int row) { return M.rb + row * 32; }
FK_HD long long stream_line_floats(int NT) { return (long long)((NT + 7) / 8) * 512; }

// Largest CTA of the streaming kernel and the resident CTAs per SM its register budget is compiled for: 2 x 192
// threads at T = 2 (168 registers; shared memory allows no more than 13 warps anyway), 2 x 256 at T = 1.
FK_HD int stream_max_threads(int T) { return T == 2 ? 192 : 256; }
FK_HD int stream_min_ctas(int T) { return T <= 2 ? 2 : 1; }

struct StreamGeom {  // per launch
    int NT;          // threads per CTA
    int CW;          // 4 * NT, columns a CTA reads
    int RH;          // output rows per CTA
    int ntall;       // the first ntall chunks after the first one have RH rows, the following ones RH - 4 (plan_stream
                     // balances the launch: the chunks at a physical top / bottom edge get fewer rows than the others)
    int RH0;         // ... of the first row chunk (shorter when it starts at the physical top edge); the last chunk
                     // takes what is left
    int nstrips, nchunks;
    int cstride;     // output columns per strip (<= CW - 8T)
    int uniformD;    // diffusivity (and so D_x, D_y) is one constant over the interior
    int row0, row1;  // output rows of this launch (every one needs 4T rows of input above and below)
};

// first output row of row chunk c; chunk c covers [stream_chunk_row(c), stream_chunk_row(c + 1))
FK_HD int stream_chunk_row(const StreamGeom& G, int c) {
    if (c <= 0) return G.row0;
    if (c >= G.nchunks) return G.row1;
    const int k = c - 1;   // whole chunks between the first one and chunk c
    return G.row0 + G.RH0 + k * G.RH - 4 * (k > G.ntall ? k - G.ntall : 0);
}

FK_HD long long stream_smem_floats(int T, int NT) { return stream_line_floats(NT) + (long long)(NT + 2) * (56 * T - 44); }

template <int T>
struct StreamState {     // registers of one thread
    float GX[T][4][4];   // u_x rows rho-2 .. rho+1 of each stage (a rotating window inside the unrolled steady loop)
    float gy[T][4];      // u_y of row rho (made one iteration ahead)
    float gypad[T];      // edge threads only: u_y of the PAD column next to the tissue edge (padded index 0 / W+1)
    float prev[T][4];    // the row each stage received one iteration ago (row rho + 3 of its window)
    float sv[T][4][4];   // general body only, CTAs at the physical top / bottom edge: values a stage carries from one
                         // special iteration to the next (see "physical top / bottom edge" below)
    long long g0;        // offset of (tissue `sim`, row rin0 + i, this thread's first column) in the state arrays
    long long gd;        // the same in the diffusivity maps (one shared map, or one per tissue)
};

struct StreamCta {       // uniform per CTA
    int cs;              // first column of the strip
    int r0, r1;          // output rows
    int rin0, rin_end;   // level-0 rows read
    int top, bot;        // this chunk starts at the physical top edge / ends at the physical bottom edge
    int vw_lo, vw_hi;    // level-0 v, w rows read
    float DXcT, DXcB;    // uniform diffusivity: D_x in the tissue's first / last row (one-sided formula)
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
    C.r0 = stream_chunk_row(G, chunk);
    C.r1 = stream_chunk_row(G, chunk + 1);   // (the last chunk takes the remainder, >= 4T rows)
    // A chunk at a physical edge has no rows beyond it: its stages start at row 0 / run dry after row H-1 and use the
    // reference's one-sided formulas there.  Every other chunk reads 4T apron rows on that side.
    C.top = A.phys_top && C.r0 == 0;
    C.bot = A.phys_bot && C.r1 == A.H;
    C.rin0 = C.top ? 0 : C.r0 - 4 * T;
    C.rin_end = C.bot ? A.H : C.r1 + 4 * T;
    C.vw_lo = C.top ? 0 : C.r0 - 4 * (T - 1);
    C.vw_hi = C.bot ? A.H : C.r1 + 4 * (T - 1);
    C.out_c0 = strip * G.cstride;
    C.out_c1 = C.out_c0 + G.cstride < A.W ? C.out_c0 + G.cstride : A.W;
    C.c_end = C.out_c1 + 4 * T < A.W ? C.out_c1 + 4 * T : A.W;  // columns past this are nobody's input
    C.edgeL = strip == 0 ? 0 : -1;
    C.edgeR = C.out_c1 == A.W ? (A.W - 4 - C.cs) / 4 : -1;
    C.boff = (long long)sim * A.plane;
    C.boffD = (long long)sim * A.plane_D;
    C.stims = A.stims ? A.stims + (long long)sim * A.n_stim : nullptr;
    C.niter = C.r1 + 4 * T - C.rin0;   // the last stage emits row r1 - 1 when (virtual) row r1 - 1 + 4T is the newest
    for (int s = 0; s < 8; ++s) C.mask[s] = 0;
    for (int s = 0; s < T; ++s) {
        const double t = A.t0 + (double)s;
        unsigned m = 0;
        for (int i = 0; i < A.n_stim; ++i) {
            const StimDev sd = C.stims[i];
            if (sd.field && stim_on(sd, t, A.t_is_int)) m |= 1u << i;
        }
        C.mask[s] = m;
    }
    C.Dc = C.DXc = C.DYc = C.DYcL = C.DYcR = C.DXcT = C.DXcB = 0.0f;
    if (G.uniformD) {
        const int rmid = (G.row0 + G.row1) / 2;   // a row 2+ away from the top and bottom edges
        const long long g = C.boffD + (long long)rmid * A.W + 4 * T;  // ... and a column 2+ away from left and right
        C.Dc = A.D[g]; C.DXc = A.DX[g]; C.DYc = A.DY[g];
        C.DYcL = A.DY[g - 4 * T];
        C.DYcR = A.DY[g - 4 * T + A.W - 1];
        C.DXcT = C.top ? A.DX[C.boffD + 4 * T] : C.DXc;
        C.DXcB = C.bot ? A.DX[C.boffD + (long long)(A.H - 1) * A.W + 4 * T] : C.DXc;
    }
}

// this thread's chunk of the CTA's shared memory, and its registers at the start of a CTA
template <int T>
FK_HD StreamMem stream_mem(float* smem, int tid, int NT) {
    typedef StreamLay<T> L;
    StreamMem M;
    M.rb = smem + opaque_i((tid >> 3) * (int)L::BLOCK + (tid & 7) * 4);
    M.hl = tid == 0 ? 0 : ((tid & 7) == 0 ? -((int)L::BLOCK - 28) : -4);
    M.hr = tid == NT - 1 ? 0 : ((tid & 7) == 7 ? (int)L::BLOCK - 28 : 4);
    M.tb = smem + stream_line_floats(NT) + (long long)(tid + 1) * L::CHUNK;
    return M;
}
// the split barrier of the steady-state loop lives in the pad granule of the CTA's first (pad) chunk
template <int T>
FK_HD float* stream_bar(float* smem, int NT) { return smem + stream_line_floats(NT) + StreamLay<T>::CHUNK - 4; }

template <int T>
FK_HD void stream_state_init(const TileArgs& A, const StreamCta& C, int tid, StreamState<T>& R) {
    for (int s = 0; s < T; ++s) {
        for (int k = 0; k < 4; ++k) {
            R.GX[s][0][k] = R.GX[s][1][k] = R.GX[s][2][k] = R.GX[s][3][k] = 0.0f;
            R.gy[s][k] = R.prev[s][k] = 0.0f;
            R.sv[s][0][k] = R.sv[s][1][k] = R.sv[s][2][k] = R.sv[s][3][k] = 0.0f;
        }
        R.gypad[s] = 0.0f;
    }
    R.g0 = opaque(C.boff + (long long)C.rin0 * A.W + C.cs + 4 * tid);
    R.gd = opaque(C.boffD + (long long)C.rin0 * A.W + C.cs + 4 * tid);
}

// start the asynchronous fetch of the level-0 rows iteration `i` will consume (gi: offset of row rin0 + i of this
// tissue at this thread's columns) into the given granules; always commits a group
template <int T>
FK_HD void stream_prefetch_to(const TileArgs& A, const StreamCta& C, float* udst, float* vdst, float* wdst, int i,
                              long long gi, bool act) {
    const int n = C.rin0 + i;  // u row pushed at iteration i
    if (act && n < C.rin_end) async_copy16(udst, A.u_in + gi);
    const int rho = n - 4;     // v, w row stage 0 emits at iteration i
    if (act && rho >= C.vw_lo && rho < C.vw_hi) {
        const long long g = gi - 4LL * A.W;
        async_copy16(vdst, A.v_in + g);
        async_copy16(wdst, A.w_in + g);
    }
    async_commit();
}

// the same with the slots computed from i (pipeline prologue)
template <int T>
FK_HD void stream_prefetch(const TileArgs& A, const StreamCta& C, const StreamMem& M, int i, int tid, bool act) {
    typedef StreamLay<T> L;
    stream_prefetch_to<T>(A, C, row_at(M, L::ROW_U + (i & 7)), row_at(M, L::ROW_V + (i & 3)), row_at(M, L::ROW_W + (i & 3)), i,
                          C.boff + (long long)(C.rin0 + i) * A.W + C.cs + 4 * tid, act);
}

// one-sided first derivative / dx (solve.py:232-235, 246-249) on four consecutive values
template <bool EXACT>
FK_HD float edge_deriv(const Consts& K, int kind, float a0, float a1, float a2, float a3) {
    float k0, k1, k2, k3;
    int o0, o1, o2, o3;
    kind_coeffs(kind, k0, k1, k2, k3, o0, o1, o2, o3);
    return deriv<EXACT>(K, kind, k0, k1, k2, k3, a0, a1, a2, a3);
}

enum { FK_WARM = 8 };   // iterations the warm start stands for

// steady-state unroll factor: 4 (a phase knows i & 3, only the halves of the 8-slot ring alternate) while the loop body
// fits the instruction cache, 2 from T = 2 on (32 KB of code per loop otherwise: every instruction line would miss)
#ifndef FK_UNROLL_T2
#define FK_UNROLL_T2 2   // development: 4 = the 4-way loop at T = 2 as well (25 KB of code with the packed arithmetic)
#endif
FK_HD constexpr int stream_unroll(int T) { return T == 2 ? FK_UNROLL_T2 : (T >= 2 ? 2 : 4); }

// may iterations i >= 8T use the condition-free body?  (no stimulus active in any level of this launch)
template <int T>
FK_HD bool stream_steady_ok(const StreamCta& C) {
    unsigned m = 0;
    for (int s = 0; s < T; ++s) m |= C.mask[s];
    return m == 0;
}

// iterations [stream_nfill, stream_nfill + U * stream_nbody) run the unrolled steady-state body, phase (i - nfill) % U
// (a chunk at the bottom edge leaves the steady state before its stage 0 receives the tissue's last row)
template <int T>
FK_HD int stream_steady_end(const StreamCta& C) { return C.bot ? C.niter - 4 * T - 1 : C.niter; }
// First iteration of the unrolled body.  Right after the warm start (iteration 8) in an ordinary chunk: stage 0 is in
// its steady state from there, and the later stages, which are still filling, simply run on whatever their rings hold
// -- nothing of it reaches a valid row (a stage's first valid emission only uses rows the stage before emitted validly)
// and the last stage's stores are masked until its first row (iteration 8T).  In a chunk that starts at the physical
// top edge, where stage s receives row m at iteration m + 4s and needs the general body up to m = 5: 4T + 2, rounded
// up to whole bodies.
template <int T>
FK_HD int stream_fill_len(const StreamCta& C) {
    return C.top ? (4 * T + 2 + stream_unroll(T) - 1) / stream_unroll(T) * stream_unroll(T) : (int)FK_WARM;
}
template <int T>
FK_HD int stream_nfill(const StreamCta& C) {
    const int f = stream_fill_len<T>(C);
    return (f + 4 <= stream_steady_end<T>(C) && stream_steady_ok<T>(C)) ? f : C.niter;
}
template <int T>
FK_HD int stream_nbody(const StreamCta& C) {
    const int n = stream_steady_end<T>(C) - stream_nfill<T>(C);
    return n > 0 ? n / stream_unroll(T) : 0;
}

// u_y (solve.py:50) of one row at the thread's 4 columns: u1 = the row's own values, r1p = its granule (the neighbours'
// granules sit at r1p + hl / r1p + hr: one chunk away in the thread-major rings, StreamMem::hl, hr in a level-0 row); gypad:
// u_y of the pad column at a tissue edge
template <bool EXACT, int T>
FK_HD void stream_make_gy(const Consts& K, const float* r1p, int hl, int hr, const float* u1, bool edgeL, bool edgeR, float* gy,
                          float& gypad) {
    const F2 Lh = ld2(r1p + hl + 2), Rh = ld2(r1p + hr);
    // edge-pad (solve.py:31): the column outside the tissue repeats the edge column
    const float e[8] = {Lh.x, edgeL ? u1[0] : Lh.y, u1[0], u1[1], u1[2], u1[3], edgeR ? u1[3] : Rh.x, Rh.y};
    dcen_span4<EXACT>(K, e, gy);
    if (edgeL) {  // padded columns 0 (the pad) and 1 (tissue column 0) use the forward formula
        gypad = edge_deriv<EXACT>(K, FWD, e[1], e[2], e[3], e[4]);
        gy[0] = edge_deriv<EXACT>(K, FWD, e[2], e[3], e[4], e[5]);
    }
    if (edgeR) {  // padded columns W (tissue column W-1) and W+1 (the pad) use the backward formula
        gy[3] = edge_deriv<EXACT>(K, BWD, e[2], e[3], e[4], e[5]);
        gypad = edge_deriv<EXACT>(K, BWD, e[3], e[4], e[5], e[6]);
    }
}

// ---- warm start: iterations 0..7 of a CTA only stream level-0 rows in (stage 0 first emits at iteration 8), so the
// kernel fetches those 8 rows at once (stream_warm_load), and after one block barrier builds stage 0's registers from
// them directly (stream_warm_start): the u_x window, the previous row, u_y of the first row it will emit (published in
// slot 1 of GY(0), as iteration 7 would have) -- and starts the fetches of iterations 8 .. 8 + FK_PF - 1.  A second
// block barrier later the general body continues at iteration 8.
// not for a chunk at the physical top edge (it emits from iteration 4 on), nor for a bottom chunk so short that its
// stage 0 receives the tissue's last row -- a special iteration -- before iteration 8
template <int T>
FK_HD bool stream_use_warm(const StreamCta& C) { return !C.top && !(C.bot && C.niter - 4 * T - 1 < FK_WARM); }

template <int T>
FK_HD void stream_warm_load(const TileArgs& A, const StreamCta& C, const StreamMem& M, int tid) {
    const int c = C.cs + 4 * tid;
    if (c < C.c_end) {
        const long long g = C.boff + (long long)C.rin0 * A.W + c;
#pragma unroll
        for (int m = 0; m < FK_WARM; ++m) async_copy16(row_at(M, StreamLay<T>::ROW_U + m), A.u_in + g + (long long)m * A.W);
    }
    async_commit();
}

template <bool EXACT, int T>
FK_HD void stream_warm_start(const TileArgs& A, const StreamCta& C, StreamState<T>& R, const StreamMem& M, int tid) {
    typedef StreamLay<T> L;
    float* const tb = M.tb;
    const int c = C.cs + 4 * tid;
    const bool act = c < C.c_end;
    if (act) {
        float u[FK_WARM][4];
#pragma unroll
        for (int m = 0; m < FK_WARM; ++m) unpack4(ld4(row_at(M, L::ROW_U + m)), u[m]);
        // u_x of rows rin0 + 2 .. rin0 + 5: the window (rho-2 .. rho+1) of iteration 8, rho = rin0 + 4
#pragma unroll
        for (int m = 0; m < 4; ++m)
#pragma unroll
            for (int k = 0; k < 4; ++k) R.GX[0][m][k] = dcen<EXACT>(A.K, u[m][k], u[m + 1][k], u[m + 3][k], u[m + 4][k]);
#pragma unroll
        for (int k = 0; k < 4; ++k) R.prev[0][k] = u[7][k];
        stream_make_gy<EXACT, T>(A.K, row_at(M, L::ROW_U + 4), M.hl, M.hr, u[4], tid == C.edgeL, tid == C.edgeR, R.gy[0], R.gypad[0]);
        st4(tb + L::GY + 4 * 1, R.gy[0]);
    }
    // rows 0 .. FK_PF-1 of the ring are free again: fetch what iterations 8 .. 8 + FK_PF - 1 consume
    for (int m = 0; m < FK_PF; ++m) stream_prefetch<T>(A, C, M, FK_WARM + m, tid, act);
    R.g0 = opaque(R.g0 + (long long)FK_WARM * A.W);
    R.gd = opaque(R.gd + (long long)FK_WARM * A.W);
}

// second derivatives, reaction, stimulus and Euler update of one row of one stage (4 cells)
template <bool EXACT, bool HAS_STIM, bool EDGE>
FK_HD void stream_emit(const Consts& K, const float* u0, const float* v, const float* w, const float* gxm2,
                       const float* gxm1, const float* gx0, const float* gxp1, const float* gxp2, const float* g,
                       const float* gy0, const float* Dv, const float* DXv, const float* DYv, const float* stim,
                       bool edgeL, bool edgeR, float* un, float* vn, float* wn, int mode = 0,
                       float (*sv)[4] = nullptr) {
    // mode 0: the ordinary row.  mode 2: u_xx supplied in sv[3] (the tissue's last row).  mode 1: the caller finishes u
    // itself one iteration later (the tissue's first row): u_yy and j_ion (fast numerics: the reaction term g of
    // cell_react_fast) go to sv[2], sv[3], un is not meaningful.
    if (!EXACT) {
        // fast numerics: two cells per instruction (fk_core.h, f2): the same operations, lane for lane, as the scalar
        // loop below runs with Num<false>
        float uxx[4], uyy[4];
        dcen_rows4<false>(K, gxm2, gxm1, gxp1, gxp2, uxx);                       // solve.py:51
        if (mode == 2) {
#pragma unroll
            for (int k = 0; k < 4; ++k) uxx[k] = sv[3][k];
        }
        dcen_span4<false>(K, g, uyy);                                            // solve.py:52
        // the tissue's first / last column: forward / backward formula on u_y of padded columns 1..4 / W-3..W
        if (EDGE && edgeL) uyy[0] = edge_deriv<false>(K, FWD, g[2], g[3], g[4], g[5]);
        if (EDGE && edgeR) uyy[3] = edge_deriv<false>(K, BWD, g[2], g[3], g[4], g[5]);
#pragma unroll
        for (int k = 0; k < 4; k += 2) {
            const f2 uu = f2_at(u0 + k), vv = f2_at(v + k), ww = f2_at(w + k);
            f2 gg, v2, w2;
            cell_react_fast2<HAS_STIM>(K, uu, vv, ww, HAS_STIM ? f2_at(stim + k) : f2_all(0.0f), gg, v2, w2);
            if (mode == 1) { f2_to(sv[2] + k, f2_at(uyy + k)); f2_to(sv[3] + k, gg); }
            f2_to(vn + k, v2);
            f2_to(wn + k, w2);
            f2_to(un + k, cell_u_fast2(K, uu, gg, f2_at(Dv + k), f2_at(DXv + k), f2_at(DYv + k), f2_at(gx0 + k),
                                        f2_at(gy0 + k), f2_at(uxx + k), f2_at(uyy + k)));   // solve.py:55, 59, 70
        }
        return;
    }
    // exact numerics: the four u_xx and the four u_yy each share one range test of their divisions (fk_core.h, DivTrack)
    float uxx[4], uyy[4];
    dcen_rows4<EXACT>(K, gxm2, gxm1, gxp1, gxp2, uxx);                           // solve.py:51
    if (mode == 2) {
#pragma unroll
        for (int k = 0; k < 4; ++k) uxx[k] = sv[3][k];
    }
    dcen_span4<EXACT>(K, g, uyy);                                                // solve.py:52
    // the tissue's first / last column: forward / backward formula on u_y of padded columns 1..4 / W-3..W
    if (EDGE && edgeL) uyy[0] = edge_deriv<EXACT>(K, FWD, g[2], g[3], g[4], g[5]);
    if (EDGE && edgeR) uyy[3] = edge_deriv<EXACT>(K, BWD, g[2], g[3], g[4], g[5]);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float del_u = diffusion<EXACT>(Dv[k], DXv[k], DYv[k], gx0[k], gy0[k], uxx[k], uyy[k]);
        float d_v, d_w, j_ion;
        cell_rhs_parts<EXACT, HAS_STIM>(K, u0[k], v[k], w[k], HAS_STIM ? stim[k] : 0.0f, d_v, d_w, j_ion);
        if (mode == 1) { sv[2][k] = uyy[k]; sv[3][k] = j_ion; }
        vn[k] = euler<EXACT>(v[k], d_v, K.dt);
        wn[k] = euler<EXACT>(w[k], d_w, K.dt);
        un[k] = euler<EXACT>(u0[k], Num<EXACT>::add(del_u, j_ion), K.dt);   // solve.py:59, 70
    }
}

// Where one iteration finds its ring slots.  Ring offsets inside the chunk are added by the user (bj + L::V + 16 s is
// slot j of V(s), and so on), so that they fold into the instructions' immediate fields.
struct StreamPtrs {
    float* bj;     // chunk base shifted to slot j = i & 3 of the 4-slot rings
    float* bj1;    // ... to slot (j + 1) & 3
    float *u0new, *u0r1, *u0r0, *u0pf;   // stage-0 ring granules (row-major region) of iterations i, i-3, i-4, i+FK_PF
    float *v0, *w0, *v0pf, *w0pf;        // level-0 v, w: slot j and slot (j + FK_PF) & 3
};

// general body: everything computed from i
template <int T>
FK_HD StreamPtrs stream_ptrs_any(const StreamMem& M, int i) {
    typedef StreamLay<T> L;
    StreamPtrs P;
    P.bj = M.tb + 4 * (i & 3);
    P.bj1 = M.tb + 4 * ((i + 1) & 3);
    P.u0new = row_at(M, L::ROW_U + (i & 7));
    P.u0r1 = row_at(M, L::ROW_U + ((i + 5) & 7));
    P.u0r0 = row_at(M, L::ROW_U + ((i + 4) & 7));
    P.u0pf = row_at(M, L::ROW_U + ((i + FK_PF) & 7));
    P.v0 = row_at(M, L::ROW_V + (i & 3));
    P.w0 = row_at(M, L::ROW_W + (i & 3));
    P.v0pf = row_at(M, L::ROW_V + ((i + FK_PF) & 3));
    P.w0pf = row_at(M, L::ROW_W + ((i + FK_PF) & 3));
    return P;
}

// the per-body bases of the unrolled loop (body = U consecutive iterations starting at a multiple of U).  They are CARRIED
// from body to body (stream_body_next: a rotation of four offsets, two swaps) rather than rebuilt from the iteration index,
// which cost ~40 integer instructions per body.
template <int T>
struct StreamBody {
    float* tb;
    float* rb;
    // line-major region, offsets in BYTES from rb (so that they fold into `register + uniform register + immediate`).  U = 2 (i even): U(0) slots i & 7, (i + 2) & 7, (i + 4) & 7, (i + 6) & 7
    // (the odd slots follow their even ones: + one row); V(0) slots i & 3 and (i + 2) & 3 (W(0): + 4 rows).  U = 4 (i a
    // multiple of 4): ua, ub = the halves of the U(0) ring holding slot i & 7 / the other one, ve = V(0) slot 0.
    int ua, uc, ub, ud, ve, vf;
    float *tH, *tO;         // U = 2: chunk base shifted to the pair of 4-ring slots holding slot i & 3 / the other pair
};

template <int T>
FK_HD StreamBody<T> stream_body_at(const StreamMem& M, int i) {   // i: first iteration of the body
    typedef StreamLay<T> L;
    StreamBody<T> Y;
    Y.tb = M.tb;
    Y.rb = M.rb;
    Y.ua = (L::ROW_U + (i & 7)) * (int)L::PITCH * 4;
    Y.uc = (L::ROW_U + ((i + 2) & 7)) * (int)L::PITCH * 4;
    Y.ub = (L::ROW_U + ((i + 4) & 7)) * (int)L::PITCH * 4;
    Y.ud = (L::ROW_U + ((i + 6) & 7)) * (int)L::PITCH * 4;
    Y.ve = (L::ROW_V + (i & 3)) * (int)L::PITCH * 4;
    Y.vf = (L::ROW_V + ((i + 2) & 3)) * (int)L::PITCH * 4;
    const int qo = (i & 2) * 4;   // floats: 8 per slot pair
    Y.tH = M.tb + qo;
    Y.tO = M.tb + (8 - qo);
    return Y;
}

// ... of the next body (U iterations later)
template <int T, int U>
FK_HD void stream_body_next(StreamBody<T>& Y) {
    if (U == 2) {
        const int a = Y.ua;
        Y.ua = Y.uc; Y.uc = Y.ub; Y.ub = Y.ud; Y.ud = a;
        const int e = Y.ve;
        Y.ve = Y.vf; Y.vf = e;
        float* h = Y.tH;
        Y.tH = Y.tO; Y.tO = h;
    } else {
        const int a = Y.ua;
        Y.ua = Y.ub; Y.ub = a;
    }
}

// slots of phase PH of a body, every offset a compile-time constant
template <int T, int U, int PH>
FK_HD StreamPtrs stream_ptrs_phase(const StreamBody<T>& Y) {
    StreamPtrs P;
    if (U == 4) {
        P.bj = Y.tb + 4 * PH;
        P.bj1 = Y.tb + 4 * ((PH + 1) & 3);
    } else {   // U == 2: slot j = 2 h + PH
        P.bj = Y.tH + 4 * PH;
        P.bj1 = PH == 0 ? Y.tH + 4 : Y.tO;
    }
    // the rows of the row-major region: uniform offsets (functions of the body's first iteration alone)
    // U(0) slot (i + k) & 7 and V(0) slot (i + k) & 3 of the body that starts at iteration i: a carried base + a constant
    constexpr int PT = (int)StreamLay<T>::PITCH * 4, WV = ((int)StreamLay<T>::ROW_W - (int)StreamLay<T>::ROW_V) * PT;
#define FK_OU(k) (U == 4 ? ((k) < 4 ? Y.ua + PT * (k) : Y.ub + PT * ((k) - 4))                                               \
                         : (((k) >> 1) == 0 ? Y.ua : ((k) >> 1) == 1 ? Y.uc : ((k) >> 1) == 2 ? Y.ub : Y.ud) + PT * ((k) & 1))
#define FK_AT(off) reinterpret_cast<float*>(reinterpret_cast<char*>(Y.rb) + (off))
#define FK_OV(k) (U == 4 ? Y.ve + PT * (k) : ((k) < 2 ? Y.ve : Y.vf) + PT * ((k) & 1))
    P.u0new = FK_AT(FK_OU(PH));
    P.u0r0 = FK_AT(FK_OU((PH + 4) & 7));
    P.u0r1 = FK_AT(FK_OU((PH + 5) & 7));
    P.u0pf = FK_AT(FK_OU((PH + FK_PF) & 7));
    P.v0 = FK_AT(FK_OV(PH));
    P.w0 = FK_AT(FK_OV(PH) + WV);
    P.v0pf = FK_AT(FK_OV((PH + FK_PF) & 3));
    P.w0pf = FK_AT(FK_OV((PH + FK_PF) & 3) + WV);
#undef FK_OU
#undef FK_OV
#undef FK_AT
    return P;
}

// one row iteration of one thread.  `tid` in [0, NT), its 4 columns start at C.cs + 4 tid, its chunk is tb.
//   PH < 0   general body: every per-stage condition evaluated, ring slots computed from i, u_x window slid by moves;
//            the caller separates iterations with a block barrier.
//   PH >= 0  steady state, phase PH == (i - nfill) % U of the U-way unrolled loop: the caller guarantees that every
//            stage has input and emits in this iteration and that no stimulus is active in this launch; every slot in P
//            is a base register plus a compile-time offset and the u_x window is renamed, not moved.  Iterations are
//            separated by the split barrier `bar`: wait for the others' previous iteration just before the first halo
//            read, arrive after the last shared-memory write.
//   UNI      one constant diffusivity (C.Dc, C.DXc, C.DYc) instead of three maps
//   EDGE     this CTA's strip may contain the tissue's first / last column
//   MODE     what the last level's store does besides storing: 0 nothing, FK_STORE_HEUN the closing pass of a fast Heun
//            step, FK_STORE_MIRROR the halo mirror of the row-slab decomposition (rows of TileArgs::mir_* go to the
//            neighbouring GPU's memory as well) -- separate kernel instantiations on the device, run-time flags on the CPU
enum { FK_STORE_PLAIN = 0, FK_STORE_HEUN = 1, FK_STORE_MIRROR = 2 };
template <bool EXACT, int T, int PH, bool UNI, bool EDGE, int U = 4, int MODE = 0>
FK_HD void stream_iter(const TileArgs& A, const StreamCta& C, StreamState<T>& R, const StreamMem& M, int i, int tid,
                       const StreamPtrs& P, float* bar) {
    typedef StreamLay<T> L;
    float* const tb = M.tb;
    constexpr bool ST = PH >= 0;
    // fast Heun's closing pass in the last level's store (TileArgs::hy_*): its own kernel instantiation on the device --
    // even a uniform run-time branch here cost the plain Euler kernel 7 % -- a run-time flag in the CPU emulation
#if defined(__CUDA_ARCH__)
    constexpr bool heun = MODE == FK_STORE_HEUN;
    constexpr bool mirror = MODE == FK_STORE_MIRROR;
#else
    const bool heun = A.hy_u != nullptr;
    const bool mirror = A.mir_u[0] != nullptr || A.mir_u[1] != nullptr;
#endif
    // positions of the u_x window rows (rho-2, rho-1, rho, rho+1) in R.GX[s]: U = 4 rotates by one per phase; U = 2
    // keeps the rows of each parity in a pair of register sets and moves one set per phase
    constexpr int W0 = !ST ? 0 : PH;
    constexpr int W1 = !ST ? 1 : (U == 4 ? (PH + 1) & 3 : 1 - PH);
    constexpr int W2 = !ST ? 2 : (U == 4 ? (PH + 2) & 3 : 2 + PH);
    constexpr int W3 = !ST ? 3 : (U == 4 ? (PH + 3) & 3 : 3 - PH);
    constexpr int CH = L::CHUNK;
    const int c = C.cs + 4 * tid;
    const bool act = c < C.c_end;
    const long long g0 = R.g0;         // row n0 = rin0 + i at this thread's columns
    R.g0 = opaque(g0 + A.W);           // carried, not recomputed from i
    const long long gd = R.gd;
    if (!UNI) R.gd = opaque(gd + A.W);
    // rows fetched FK_PF iterations ago have landed (this thread's own columns; the neighbours'
    // columns of a row are only read three barriers later)
    stream_prefetch_to<T>(A, C, P.u0pf, P.v0pf, P.w0pf, i + FK_PF, g0 + (long long)FK_PF * A.W, act);
    async_wait<FK_PF>();
    // Threads beyond the strip's last needed column: the general body skips them.  The steady-state body lets them
    // compute on whatever their chunk holds (nothing of theirs is stored, nobody reads their columns): an early exit
    // would cost every thread a block of register moves at the join, and the planner never leaves a whole warp idle.
    if (!ST && !act) return;
    const int bar_done = ST ? sb_test(bar, PH & 1) : 1;
    const bool edgeL = EDGE && tid == C.edgeL, edgeR = EDGE && tid == C.edgeR;
    const int n0 = C.rin0 + i;
    // ---- this thread's own columns: nothing here was written by another thread, so it is read before the barrier wait
    // v, w of the rows the stages emit now: slot j of every stage, read before stage s-1 refills slot j of stage s
    float vv[T][4], ww[T][4], u0[T][4], u1[T][4];
    float* r0p[T];
    float* r1p[T];
    float in_u[4] = {0.f, 0.f, 0.f, 0.f};
    bool have_in = ST || n0 < C.rin_end;
#pragma unroll
    for (int s = 0; s < T; ++s) {
        unpack4(ld4(s == 0 ? P.v0 : P.bj + L::V + 16 * s), vv[s]);
        unpack4(ld4(s == 0 ? P.w0 : P.bj + L::W + 16 * s), ww[s]);
        // input rows rho, rho+1 of the stage (row rho+3 is still in registers, row rho+4 arrives below)
        if (s == 0) {
            r0p[s] = P.u0r0; r1p[s] = P.u0r1;
            if (have_in) unpack4(ld4(P.u0new), in_u);
        } else {
            r0p[s] = P.bj + L::U(s); r1p[s] = P.bj1 + L::U(s);
        }
        unpack4(ld4(r0p[s]), u0[s]);
        unpack4(ld4(r1p[s]), u1[s]);
    }
    // heterogeneous diffusivity: the three maps at the rows the stages emit, requested before the wait as well
    float Dm[UNI ? 1 : T][4], DXm[UNI ? 1 : T][4], DYm[UNI ? 1 : T][4];
    bool need_map[T];
#pragma unroll
    for (int s = 0; s < T; ++s) need_map[s] = false;
    if (!UNI) {
#pragma unroll
        for (int s = 0; s < T; ++s) {
            const int rho = n0 - 4 * (s + 1);
            // only rows this stage emits (all of them in the steady state), only threads inside the tissue
            // (the unrolled body also runs while the later stages are still filling: their rows may lie above the tissue)
            const bool need = act && (ST ? rho >= 0 : (rho >= (C.top ? 0 : C.r0 - 4 * (T - 1 - s)) &&
                                                       rho < (C.bot ? A.H : C.r1 + 4 * (T - 1 - s))));
#pragma unroll
            for (int k = 0; k < 4; ++k) Dm[s][k] = DXm[s][k] = DYm[s][k] = 0.0f;
            need_map[s] = need;
            if (need) {
                const long long back = 4LL * (s + 1) * A.W;
                if (!FK_MAP_IS_LATE(s)) {
                    unpack4(ldg4(A.D + gd - back), Dm[s]);
                    unpack4(ldg4(A.DX + gd - back), DXm[s]);
                    unpack4(ldg4(A.DY + gd - back), DYm[s]);
                }
#if FK_MAP_PF
                // the maps come from DRAM once (the later stages re-read them from L2): ask L2 for the rows the first
                // stage emits FK_MAP_PF iterations from now -- no register is held, unlike a load issued that early
                if (s == 0 && rho + FK_MAP_PF < A.H) {
                    const long long ahead = (long long)FK_MAP_PF * A.W - back;
                    pf_l2(A.D + gd + ahead); pf_l2(A.DX + gd + ahead); pf_l2(A.DY + gd + ahead);
                }
#endif
#if FK_MAP_PF1
                if (rho + FK_MAP_PF1 < A.H) {
                    const long long ahead1 = (long long)FK_MAP_PF1 * A.W - back;
                    pf_l1(A.D + gd + ahead1); pf_l1(A.DX + gd + ahead1); pf_l1(A.DY + gd + ahead1);
                }
#endif
            }
        }
    }
    // fast Heun: y at the row the last stage stores, requested before the wait as well -- L2 hits (this CTA read those rows
    // 4T iterations ago), but loaded at the store they were 40 % of the warp samples (ncu: long scoreboard)
    float hyu[4], hyv[4], hyw[4];
    bool hy_ok = false;
    if (heun && UNI) {   // (with diffusivity maps the twelve extra live registers spill: -9 % at 1200^2; uniform D: +6 %)
        const int rho_l = n0 - 4 * T;
        hy_ok = act && c >= C.out_c0 && c < C.out_c1 && rho_l >= C.r0 && rho_l < A.H;
        if (hy_ok) {
            const long long gl = g0 - 4LL * T * A.W;
            unpack4(ldg4(A.hy_u + gl), hyu);
            unpack4(ldg4(A.hy_v + gl), hyv);
            unpack4(ldg4(A.hy_w + gl), hyw);
        }
    }
    if (ST && !bar_done) sb_wait(bar, PH & 1);
#pragma unroll
    for (int s = 0; s < T; ++s) {
        const int rho = n0 - 4 * (s + 1);  // row this stage emits (level s+1); its newest input row is rho+4
        const int lo = C.top ? 0 : C.r0 - 4 * (T - 1 - s), hi = C.bot ? A.H : C.r1 + 4 * (T - 1 - s);
        if (s > 0 && have_in) st4(r0p[s], in_u);  // row rho+4 takes the slot of row rho
        // u_x of row rho+2 (solve.py:49)
        float ngx[4];
        dcen_rows4<EXACT>(A.K, u0[s], u1[s], R.prev[s], in_u, ngx);
        // ---- physical top / bottom edge (general body, chunks that touch it).  m = the newest input row of this stage.
        // Along the rows the reference differentiates the edge-PADDED array twice (solve.py:29-32, 49, 51) with one-sided
        // formulas in padded rows 0, 1 and H, H+1.  In terms of tissue rows, with gx[r] = u_x of row r:
        //   top     gxpad = FWD(u0,u0,u1,u2), gx[0] = FWD(u0..u3), gx[1] = CEN(u0,u0,u2,u3): made when row 3 arrives;
        //           u_xx[0] = FWD(gx[0..3]) needs row 5, ONE ROW MORE than the pipeline's lag: row 0's v, w are emitted
        //           on time (m = 4), its u one iteration later (m = 5) and handed to the next stage retroactively;
        //           u_xx[1] = CEN(gxpad, gx[0], gx[2], gx[3]) is the ordinary formula on that window.
        //   bottom  gx[H-2] = CEN(u[H-4],u[H-3],u[H-1],u[H-1]), gx[H-1] = BWD(u[H-4..H-1]), gxpad = BWD(u[H-3],u[H-2],
        //           u[H-1],u[H-1]) are made when the last row arrives (m = H-1) and fed to the window while the stage runs
        //           dry (m = H .. H+2); u_xx[H-1] = BWD(gx[H-4..H-1]) is taken from the window one iteration early.
        const int m = rho + 4;
        bool defer_u = false, give_uxx = false, top_init = false;
        float tgx[3][4];
        if (!ST && (C.top || C.bot)) {
            float rm2[4] = {0.f, 0.f, 0.f, 0.f};   // row m-2: the only window row the ordinary formulas never read
            if ((C.top && m == 3) || (C.bot && m == A.H - 1))
                unpack4(ld4(s == 0 ? row_at(M, L::ROW_U + ((i + 6) & 7)) : tb + L::U(s) + 4 * ((i + 2) & 3)), rm2);
            if (C.top && m == 3) {
                top_init = true;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float a0 = u1[s][k], a1 = rm2[k], a2 = R.prev[s][k], a3 = in_u[k];   // rows 0, 1, 2, 3
                    tgx[0][k] = edge_deriv<EXACT>(A.K, FWD, a0, a0, a1, a2);
                    tgx[1][k] = edge_deriv<EXACT>(A.K, FWD, a0, a1, a2, a3);
                    tgx[2][k] = dcen<EXACT>(A.K, a0, a0, a2, a3);
                }
            }
            if (C.top && m == 4) defer_u = true;
            if (C.bot && m == A.H - 1) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float a0 = u1[s][k], a1 = rm2[k], a2 = R.prev[s][k], a3 = in_u[k];   // rows H-4 .. H-1
                    R.sv[s][0][k] = dcen<EXACT>(A.K, a0, a1, a3, a3);
                    R.sv[s][1][k] = edge_deriv<EXACT>(A.K, BWD, a0, a1, a2, a3);
                    R.sv[s][2][k] = edge_deriv<EXACT>(A.K, BWD, a1, a2, a3, a3);
                }
            }
            if (C.bot && m >= A.H && m <= A.H + 2) {
#pragma unroll
                for (int k = 0; k < 4; ++k) ngx[k] = m == A.H ? R.sv[s][0][k] : (m == A.H + 1 ? R.sv[s][1][k] : R.sv[s][2][k]);
            }
            if (C.bot && m == A.H + 2) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    R.sv[s][3][k] = edge_deriv<EXACT>(A.K, BWD, R.GX[s][0][k], R.GX[s][1][k], R.GX[s][2][k], R.GX[s][3][k]);
            }
            if (C.bot && m == A.H + 3) give_uxx = true;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) R.prev[s][k] = in_u[k];
        const bool emit = ST || (rho >= lo && rho < hi);
        if (emit) {
            const float* gr = tb + L::GY + 8 * s + 4 * (ST ? (PH + 1) & 1 : (i + 1) & 1);
            const F2 Lh = ld2(gr - CH + 2), Rh = ld2(gr + CH);
            // at a tissue edge the neighbour is the pad column, whose u_y this thread made itself
            const float g[8] = {Lh.x, edgeL ? R.gypad[s] : Lh.y, R.gy[s][0], R.gy[s][1], R.gy[s][2], R.gy[s][3],
                                edgeR ? R.gypad[s] : Rh.x, Rh.y};
            float Dv[4], DXv[4], DYv[4];
            const long long back = 4LL * (s + 1) * A.W;
            const long long grow = g0 - back;   // (sim, rho, c)
            if (UNI) {
#pragma unroll
                for (int k = 0; k < 4; ++k) { Dv[k] = C.Dc; DXv[k] = C.DXc; DYv[k] = C.DYc; }
                if (edgeL) DYv[0] = C.DYcL;
                if (edgeR) DYv[3] = C.DYcR;
                if (!ST && C.top && rho == 0) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) DXv[k] = C.DXcT;
                }
                if (!ST && C.bot && rho == A.H - 1) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) DXv[k] = C.DXcB;
                }
            } else {
                // loaded here, stage by stage (L2 hits thanks to the prefetch above), so that only one stage's maps are
                // live at a time
                if (FK_MAP_IS_LATE(s) && need_map[s]) {
                    unpack4(ldg4(A.D + gd - back), Dm[s]);
                    unpack4(ldg4(A.DX + gd - back), DXm[s]);
                    unpack4(ldg4(A.DY + gd - back), DYm[s]);
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) { Dv[k] = Dm[UNI ? 0 : s][k]; DXv[k] = DXm[UNI ? 0 : s][k]; DYv[k] = DYm[UNI ? 0 : s][k]; }
            }
            const int emode = ST ? 0 : (defer_u ? 1 : (give_uxx ? 2 : 0));
            if (!ST && defer_u) {
#pragma unroll
                for (int k = 0; k < 4; ++k) { R.sv[s][0][k] = u0[s][k]; R.sv[s][1][k] = R.gy[s][k]; }
            }
            float un[4], vn[4], wn[4];
            const unsigned mask = ST ? 0u : C.mask[s];
            if (mask) {  // solve.py:260-269: later stimuli win, zero cells never stimulate
                float stim[4] = {0.f, 0.f, 0.f, 0.f};
                for (int q = 0; q < A.n_stim; ++q)
                    if (mask >> q & 1u) {
                        float f[4];
                        unpack4(ldg4(C.stims[q].field + (grow - C.boff)), f);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            if (f[k] != 0.0f) stim[k] = f[k];
                    }
                stream_emit<EXACT, true, EDGE>(A.K, u0[s], vv[s], ww[s], R.GX[s][W0], R.GX[s][W1], R.GX[s][W2],
                                               R.GX[s][W3], ngx, g, R.gy[s], Dv, DXv, DYv,
                                               stim, edgeL, edgeR, un, vn, wn, emode, R.sv[s]);
            } else {
                stream_emit<EXACT, false, EDGE>(A.K, u0[s], vv[s], ww[s], R.GX[s][W0], R.GX[s][W1], R.GX[s][W2],
                                                R.GX[s][W3], ngx, g, R.gy[s], Dv, DXv, DYv,
                                                nullptr, edgeL, edgeR, un, vn, wn, emode, R.sv[s]);
            }
            if (s == T - 1) {
                if (c >= C.out_c0 && c < C.out_c1 && rho >= C.r0) {   // (rho < r0: the unrolled body while it fills)
                    if (heun) {   // fast Heun: this launch computed E(E(y)); store y + (E(E(y)) - y) / 2
                        if (!hy_ok) {
                            unpack4(ldg4(A.hy_u + grow), hyu);
                            unpack4(ldg4(A.hy_v + grow), hyv);
                            unpack4(ldg4(A.hy_w + grow), hyw);
                        }
                        heun_fold4(hyu, un); heun_fold4(hyv, vn); heun_fold4(hyw, wn);
                    }
                    if (ST || !defer_u) st4(A.u_out + grow, un);   // (a deferred first row's u follows below)
                    st4(A.v_out + grow, vn);
                    st4(A.w_out + grow, wn);
                    if (mirror) {   // the rows a neighbouring slab needs as its halo: stored into its memory as well
#pragma unroll
                        for (int nb = 0; nb < 2; ++nb)
                            if (rho >= A.mir_r0[nb] && rho < A.mir_r1[nb]) {   // (never the tissue's first row: no deferral)
                                const long long gm = grow + A.mir_off[nb];
                                st4(A.mir_u[nb] + gm, un);
                                st4(A.mir_v[nb] + gm, vn);
                                st4(A.mir_w[nb] + gm, wn);
                            }
                    }
                }
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) in_u[k] = un[k];
                st4(P.bj + L::V + 16 * (s + 1), vn);
                st4(P.bj + L::W + 16 * (s + 1), wn);
            }
        }
        if (!ST && C.top && m == 5) {
            // the tissue's first row, one iteration late: u_xx[0] = FWD(gx[0], gx[1], gx[2], gx[3]) with gx[3] the new row
            // of the window, u_x[0] = gx[0]; u_y, u_yy, j_ion and u itself were saved when its v, w were emitted
            const long long grow0 = g0 - 4LL * (s + 1) * A.W - A.W;   // (sim, 0, c)
            float Dv[4], DXv[4], DYv[4], un0[4];
            if (UNI) {
#pragma unroll
                for (int k = 0; k < 4; ++k) { Dv[k] = C.Dc; DXv[k] = C.DXcT; DYv[k] = C.DYc; }
                if (edgeL) DYv[0] = C.DYcL;
                if (edgeR) DYv[3] = C.DYcR;
            } else {
                const long long gd0 = gd - 4LL * (s + 1) * A.W - A.W;
                unpack4(ldg4(A.D + gd0), Dv);
                unpack4(ldg4(A.DX + gd0), DXv);
                unpack4(ldg4(A.DY + gd0), DYv);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float u_xx = edge_deriv<EXACT>(A.K, FWD, R.GX[s][1][k], R.GX[s][2][k], R.GX[s][3][k], ngx[k]);
                if (EXACT) {
                    const float del_u = diffusion<EXACT>(Dv[k], DXv[k], DYv[k], R.GX[s][1][k], R.sv[s][1][k], u_xx, R.sv[s][2][k]);
                    un0[k] = euler<EXACT>(R.sv[s][0][k], Num<EXACT>::add(del_u, R.sv[s][3][k]), A.K.dt);
                } else {
                    un0[k] = cell_u_fast(A.K, R.sv[s][0][k], R.sv[s][3][k], Dv[k], DXv[k], DYv[k], R.GX[s][1][k], R.sv[s][1][k],
                                         u_xx, R.sv[s][2][k]);
                }
            }
            if (s == T - 1) {
                if (heun) {
                    float y4[4];
                    unpack4(ldg4(A.hy_u + grow0), y4); heun_fold4(y4, un0);
                }
                if (c >= C.out_c0 && c < C.out_c1) st4(A.u_out + grow0, un0);
            } else {
                // the next stage should have received this row one iteration ago: its ring slot and its "previous row"
                st4(tb + L::U(s + 1 < T ? s + 1 : s) + 4 * ((i + 3) & 3), un0);
#pragma unroll
                for (int k = 0; k < 4; ++k) R.prev[s + 1 < T ? s + 1 : s][k] = un0[k];
            }
        }
        // u_y of row rho+1 (solve.py:50) for the next iteration, published for the neighbours
        if (ST || (rho + 1 >= lo && rho + 1 < hi)) {
            stream_make_gy<EXACT, T>(A.K, r1p[s], s == 0 ? M.hl : -CH, s == 0 ? M.hr : CH, u1[s], edgeL, edgeR, R.gy[s], R.gypad[s]);
            st4(tb + L::GY + 8 * s + 4 * (ST ? PH & 1 : i & 1), R.gy[s]);
        }
        // the u_x window takes the new row
        if (ST && U == 4) {
#pragma unroll
            for (int k = 0; k < 4; ++k) R.GX[s][W0][k] = ngx[k];
        } else if (ST) {
#pragma unroll
            for (int k = 0; k < 4; ++k) { R.GX[s][W0][k] = R.GX[s][W2][k]; R.GX[s][W2][k] = ngx[k]; }
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                R.GX[s][0][k] = R.GX[s][1][k];
                R.GX[s][1][k] = R.GX[s][2][k];
                R.GX[s][2][k] = R.GX[s][3][k];
                R.GX[s][3][k] = ngx[k];
            }
            if (top_init) {   // the window the first emissions need: (-, gxpad, gx[0], gx[1])
#pragma unroll
                for (int k = 0; k < 4; ++k) { R.GX[s][1][k] = tgx[0][k]; R.GX[s][2][k] = tgx[1][k]; R.GX[s][3][k] = tgx[2][k]; }
            }
        }
        have_in = emit;
    }
    if (ST) sb_arrive(bar);
}

// iteration i of a CTA in whichever form the kernel runs it: general body during the pipeline fill, the tail and any
// launch with an active stimulus; phase (i - nfill) & 3 of the unrolled body in between.  (The CUDA kernel spells the
// same schedule out as three loops; the CPU emulation calls this.)
template <bool EXACT, int T, bool UNI, int U, int PH>
FK_HD void stream_iter_phase(const TileArgs& A, const StreamCta& C, StreamState<T>& R, const StreamMem& M, int i, int tid,
                             const StreamBody<T>& Y, bool edge, float* bar) {
    if (edge) stream_iter<EXACT, T, PH, UNI, true, U>(A, C, R, M, i, tid, stream_ptrs_phase<T, U, PH>(Y), bar);
    else stream_iter<EXACT, T, PH, UNI, false, U>(A, C, R, M, i, tid, stream_ptrs_phase<T, U, PH>(Y), bar);
}

template <bool EXACT, int T, bool UNI>
FK_HD void stream_iter_any(const TileArgs& A, const StreamCta& C, StreamState<T>& R, const StreamMem& M, int i, int tid) {
    constexpr int U = stream_unroll(T);
    const int nfill = stream_nfill<T>(C), nbody = stream_nbody<T>(C);
    if (i < nfill || i >= nfill + U * nbody) {
        stream_iter<EXACT, T, -1, UNI, true>(A, C, R, M, i, tid, stream_ptrs_any<T>(M, i), nullptr);
        return;
    }
    const int ph = (i - nfill) % U;   // nfill is a multiple of U
    const StreamBody<T> Y = stream_body_at<T>(M, i - ph);
    const bool edge = C.edgeL >= 0 || C.edgeR >= 0;
    if (ph == 0) stream_iter_phase<EXACT, T, UNI, U, 0>(A, C, R, M, i, tid, Y, edge, nullptr);
    else if (ph == 1) stream_iter_phase<EXACT, T, UNI, U, 1>(A, C, R, M, i, tid, Y, edge, nullptr);
    else if (ph == 2) stream_iter_phase<EXACT, T, UNI, U == 4 ? 4 : 4, U == 4 ? 2 : 0>(A, C, R, M, i, tid, Y, edge, nullptr);
    else stream_iter_phase<EXACT, T, UNI, U == 4 ? 4 : 4, U == 4 ? 3 : 0>(A, C, R, M, i, tid, Y, edge, nullptr);
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
// rows fewer than the others that the chunk at the physical top / bottom edge gets (plan_stream: top_off, bot_off).
// FK_TOP_OFF / FK_BOT_OFF in the environment: development sweeps.
// (measured on 4096^2, T = 2: 350.5 Gcell-steps/s without, 357 .. 359 for (12, 20), (16, 28), (20, 36); the 128 x 256^2
// ensemble, whose SMs hold six small CTAs each, is indifferent up to (12, 20) and loses beyond)
inline int stream_top_off(int T) { const char* e = getenv("FK_TOP_OFF"); return e ? atoi(e) : 6 * T; }
inline int stream_bot_off(int T) { const char* e = getenv("FK_BOT_OFF"); return e ? atoi(e) : 10 * T; }

template <class OccFn>
inline bool plan_stream(int row0, int row1, int W, int batch, int T, int cta_threads, int rows_per_cta, int num_sms,
                        int uniformD, int max_threads, OccFn occ, StreamPlan& P, int top_off = 0, int bot_off = 0) {
    if (T < 1 || T > 4) return false;
    if (W % 4 != 0) return false;                       // float4 rows
    if (row1 - row0 < 8 || W < 8 * T + 32) return false;  // too small: the general tile kernel does it all
    const int Wint = W, Hint = row1 - row0;
    double best = -1.0;
    const int ns_min = Wint <= 4 * max_threads ? 1 : (Wint + (4 * max_threads - 8 * T) - 1) / (4 * max_threads - 8 * T);
    for (int ns = ns_min; ns < ns_min + 24; ++ns) {
        int stride = (Wint + ns - 1) / ns;
        stride = (stride + 3) / 4 * 4;
        // columns a CTA must hold: its strip plus a 4T apron towards every neighbouring strip (none towards a physical
        // left / right edge: a single strip spans the tissue with no apron at all)
        const int need = stride + (ns == 1 ? 0 : (ns == 2 ? 4 * T : 8 * T));
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
                RH = (RH + 3) / 4 * 4;   // whole bodies of the 4-way unrolled steady-state loop
            }
            // every chunk but the first / last must stay clear of the one-sided rows at a physical edge at every level:
            // at least 4T rows per chunk, and a remainder of fewer than 4T rows is left to the last chunk
            if (RH < 8) RH = 8;
            if (RH < 4 * T) RH = 4 * T;
            if (RH > Hint) RH = Hint;
            int RH0 = RH;
            int nchunks = 1 + (Hint - RH0 + RH - 1) / RH;
            if (nchunks > 1 && Hint - RH0 - (nchunks - 2) * RH < 4 * T) --nchunks;
            int ntall = nchunks;
            // top_off / bot_off: how many rows FEWER than the others the first / last chunk should get.  A chunk at the
            // physical top edge runs its first 4T + 2 iterations in the general body, one at the bottom edge its last
            // 4T + 1: per-CTA timelines (tools/probe_stream_timing.py) put them 18 and 30 steady-state iterations behind
            // a chunk of the same height at T = 2, and a launch is one wave -- it ends with its slowest CTA.  The same
            // number of chunks is kept; the rows taken off go to the others, 4 at a time.
            if (rows_per_cta <= 0 && nchunks >= 4 && (top_off > 0 || bot_off > 0)) {
                const int n = nchunks, lo = 4 * T > 8 ? 4 * T : 8;
                int RHi = (Hint + top_off + bot_off + n - 1) / n;
                RHi = (RHi + 3) / 4 * 4;
                const int surplus = n * RHi - top_off - bot_off - Hint;            // rows to take off again
                int nshort = surplus / 4 < n - 2 ? surplus / 4 : n - 2;             // ... 4 from each of that many chunks
                const int last = RHi - bot_off - (surplus - 4 * nshort);            // ... and the rest from the last one
                if (RHi - top_off >= lo && RHi - 4 >= lo && last >= 4 * T) {
                    RH = RHi; RH0 = RHi - top_off; ntall = n - 2 - nshort;
                }
            }
            const long long ncta = units * nchunks;
            const double rounds = (double)((ncta + slots - 1) / slots);
            // CTAs actually resident on an SM (a small tissue does not fill the machine), their active warps, and
            // the time of one row iteration: T stages, each at least a lone warp's dependent chain (~2 units) and
            // growing with the warps that share the four schedulers
            long long res = (ncta + num_sms - 1) / num_sms;
            if (res > o) res = o;
            const double warps = (double)((need + 127) / 128) * (double)res;
            // (CTAs of more than 4 warps pay a little for the coarser split barrier: measured 242 vs 255 Gcell-steps/s
            // for 6- vs 4-warp CTAs at equal residency)
            const int wpc = (need + 127) / 128;
            const double cost = rounds * (RH + 8.0 * T) * T * (warps + 8.0) / 4.0 * (1.0 + 0.02 * (wpc > 4 ? wpc - 4 : 0));
            if (best < 0 || cost < best) {
                best = cost;
                P.G.NT = NT; P.G.CW = 4 * NT; P.G.RH = RH; P.G.RH0 = RH0; P.G.ntall = ntall;
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
    StreamCta C;
    stream_cta_setup<T>(A, G, strip, chunk, sim, C);
    std::vector<StreamState<T>> R((size_t)G.NT);
    for (int tid = 0; tid < G.NT; ++tid) {
        stream_state_init<T>(A, C, tid, R[tid]);
        const StreamMem M = stream_mem<T>(smem.data(), tid, G.NT);
        if (!stream_use_warm<T>(C)) for (int j = 0; j < FK_PF; ++j) stream_prefetch<T>(A, C, M, j, tid, C.cs + 4 * tid < C.c_end);
        else stream_warm_load<T>(A, C, M, tid);
    }
    if (stream_use_warm<T>(C))
        for (int q = 0; q < G.NT; ++q) {   // after the first block barrier
            const int tid = reverse ? G.NT - 1 - q : q;
            stream_warm_start<EXACT, T>(A, C, R[tid], stream_mem<T>(smem.data(), tid, G.NT), tid);
        }
    for (int i = stream_use_warm<T>(C) ? (int)FK_WARM : 0; i < C.niter; ++i)
        for (int q = 0; q < G.NT; ++q) {
            const int tid = reverse ? G.NT - 1 - q : q;
            const StreamMem M = stream_mem<T>(smem.data(), tid, G.NT);
            if (G.uniformD) stream_iter_any<EXACT, T, true>(A, C, R[tid], M, i, tid);
            else stream_iter_any<EXACT, T, false>(A, C, R[tid], M, i, tid);
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
