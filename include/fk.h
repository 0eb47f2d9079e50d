/* fk.h -- C ABI of libfk.so, the B200 (sm_100a) Fenton-Karma tissue stepper.
 *
 * The reference (epignatelli/cardiax) has no FFI: its boundary for this path is the Python
 * module API of cardiax/solve.py.  Each entry point below replaces one reference function
 * (cited as file:line relative to the reference tree) and is what a binding for that
 * function would call -- see INTEGRATION.md for the ctypes stub.
 *
 * Conventions
 *   - every pointer named *_dev / documented "device" is a CUDA device pointer owned by the
 *     caller; the library never frees or retains it beyond the call
 *   - arrays are row-major fp32, a state is three (batch, H, W) arrays in the reference's
 *     State order v, w, u (solve.py:12-15)
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing
 *     synchronises the device
 *   - return value: 0 on success, < 0 argument error, > 0 a cudaError_t; fk_last_error()
 *     gives the message (thread local).  Nothing throws.
 */
#ifndef FK_H_
#define FK_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FK_ABI_VERSION 2

/* cardiax/params.py:4-18 -- same 14 fields, same order, already converted to fp32. */
typedef struct FkParams {
    float tau_v_plus, tau_v1_minus, tau_v2_minus, tau_w_plus, tau_w_minus, tau_d, tau_0, tau_r, tau_si, k, V_csi, V_c,
        V_v, Cm;
} FkParams;

/* cardiax/stimulus.py:13-21 -- Protocol (start, duration, period in STEP units, fp32 like the
 * reference's loop counter) + field.  `field` is a device pointer to an (H, W) fp32 array (NULL =
 * unused slot). */
/* cardiax/stimulus.py:13-21.  The protocol is in STEP units and keeps the caller's typing, because the reference's
 * schedule (solve.py:262-267) is evaluated by jax in the operands' own types: `deepx.generate.random_protocol`
 * (generate.py:24-27) draws start and period as int32 arrays, `solve.forward` callers pass Python ints or 1e9 floats.
 * Values are doubles (exact for int32 and float32 alike); `int_mask` says which of them the caller held as integers. */
#define FK_STIM_INT_START 1
#define FK_STIM_INT_DURATION 2
#define FK_STIM_INT_PERIOD 4
typedef struct FkStimulus {
    const float* field;   /* (H, W) fp32 on the device; NULL = unused slot */
    double start, duration, period;
    int int_mask;         /* FK_STIM_INT_* bits */
    int reserved;
} FkStimulus;

typedef struct FkOptions {
    int exact;            /* 1: reference operation order, bit-identical to the CPU oracle; 0: fast numerics */
    int steps_per_launch; /* temporal blocking depth T (1..8); 0 = library default */
    int kernel;           /* 0 = auto, 1 = general tile kernel everywhere, 2 = require the streaming kernel,
                             3 = low-latency one-step kernel for small tissues, 5 = resident kernel with the tissue as ONE
                             thread-block cluster (<= 16 tiles, halos through distributed shared memory: the default for
                             single tissues up to 256 x 256 cells), 4 = resident kernel (whole call in one
                             cooperative launch, state in shared memory) */
    int phys_top;         /* is buffer row 0 the physical tissue edge? (0 only for slab decomposition) */
    int phys_bottom;      /* is buffer row H-1 the physical tissue edge? */
    int cta_threads;      /* streaming kernel: threads per CTA (0 = auto) */
    int rows_per_cta;     /* streaming kernel: rows per CTA chunk (0 = auto) */
    int uniform_diffusivity; /* caller asserts the diffusivity map is one constant: the streaming kernel then keeps
                                D, D_x, D_y as three scalars instead of reading three maps */
    int safe_division;    /* exact numerics only: 1 = every division through the IEEE sequence (__fdiv_rn) instead of the
                             3-instruction correctly rounded FMA division by constants (see fk_check_exact_division) */
    int tiles_r, tiles_c; /* resident kernel: tile grid rows x columns, one CTA per tile (0 = auto) */
    int cells_per_thread; /* resident kernel: 1, 2 or 4 adjacent cells per thread (0 = auto) */
    int edge_rows, edge_colgroups; /* resident kernel: rows / 4-column groups of the tiles at the tissue's edges
                                      (0 = auto, < 0 = even split) */
    int maps_global;      /* resident kernel: 1 = read the diffusivity maps from global memory (L2) instead of keeping
                             them in shared memory (0 = only when the tissue would not fit otherwise) */
    int counter_is_int;   /* the loop counter t0, t0 + 1, ... is an int32 (`deepx.generate.sequence` passes integer
                             checkpoints) rather than the float32 of `solve.forward`: decides, with FkStimulus::int_mask,
                             in which type each operation of the stimulus schedule is done (exact above 2^24) */
} FkOptions;

/* Fills *opt with the defaults (exact = 0, auto everything, both edges physical). */
void fk_default_options(FkOptions* opt);

int fk_abi_version(void);
const char* fk_last_error(void);

/* Bytes of device scratch fk_forward_euler / fk_rhs need for this problem. */
size_t fk_workspace_bytes(int H, int W, int batch, int n_stim, int diffusivity_batched);

/* solve._forward_euler (cardiax/solve.py:92-100) == lax.fori_loop(t0, t1, step_euler):
 * advances `batch` independent tissues from counter t0 while counter < t1 (counter += 1).
 *   diffusivity_dev      (H, W) shared by all tissues, or (batch, H, W) when diffusivity_batched
 *   stimuli              HOST array of batch * n_stim entries (tissue-major); the schedule
 *                        `t >= start && mod(start - t + 1, period) < duration` (solve.py:262-267)
 *                        is evaluated on the device, per step, in the caller's typing (FkStimulus, FkOptions::counter_is_int)
 * Outputs must not alias inputs. */
int fk_forward_euler(const float* v_in_dev, const float* w_in_dev, const float* u_in_dev, float* v_out_dev,
                     float* w_out_dev, float* u_out_dev, const float* diffusivity_dev, int diffusivity_batched, int H,
                     int W, int batch, const FkParams* params, const FkStimulus* stimuli, int n_stim, double t0,
                     double t1, float dt, float dx, const FkOptions* opt, void* workspace_dev, size_t workspace_bytes,
                     void* stream);

/* solve._forward_heun (cardiax/solve.py:73-85, 103-111): Heun steps for counter in [t0, t1) -- both stages evaluated at
 * the same counter, new = y + (k1 + k2) * (dt * 0.5) -- with the whole loop on the device.
 *   exact numerics: the literal formula, bit-identical to the CPU oracle: ONE tile-kernel launch per step whose two
 *     shared-memory levels are the predictor and the corrector (tissues below 2^20 cells, or opt->steps_per_launch = 2),
 *     else two right-hand-side launches and two stage kernels per step (opt->steps_per_launch = 1); same bits;
 *   fast numerics: y + (E(E(y)) - y) / 2 with E one Euler step of the streaming / wide kernel, both at counter t, plus a
 *     combine pass -- the same quantity at one extra rounding per step, 4x faster on large tissues.
 * Same arguments as fk_forward_euler; workspace: fk_heun_workspace_bytes. */
size_t fk_heun_workspace_bytes(int H, int W, int batch, int n_stim, int diffusivity_batched);
int fk_forward_heun(const float* v_in_dev, const float* w_in_dev, const float* u_in_dev, float* v_out_dev,
                    float* w_out_dev, float* u_out_dev, const float* diffusivity_dev, int diffusivity_batched, int H,
                    int W, int batch, const FkParams* params, const FkStimulus* stimuli, int n_stim, double t0,
                    double t1, float dt, float dx, const FkOptions* opt, void* workspace_dev, size_t workspace_bytes,
                    void* stream);

/* Building block of the row-slab decomposition of one large tissue over several GPUs (no reference counterpart: the
 * reference runs a tissue on one device).  ONE launch of nsteps (1..4) Euler steps that writes output rows
 * [row0, row1) of a local (H, W) buffer and nothing else; it reads input rows [row0 - 4 nsteps, row1 + 4 nsteps)
 * -- the neighbour's halo rows -- except across an edge that opt->phys_top / phys_bottom declares physical.
 * dx_map_dev / dy_map_dev come from fk_diffusivity_gradients (once per diffusivity map).
 * workspace: at least 32 * n_stim bytes (stimulus table). */
int fk_euler_rows(const float* v_in_dev, const float* w_in_dev, const float* u_in_dev, float* v_out_dev, float* w_out_dev,
                  float* u_out_dev, const float* diffusivity_dev, const float* dx_map_dev, const float* dy_map_dev, int H,
                  int W, const FkParams* params, const FkStimulus* stimuli, int n_stim, double t0, int nsteps, float dt,
                  float dx, const FkOptions* opt, int row0, int row1, void* workspace_dev, size_t workspace_bytes,
                  void* stream);

/* ---- halo exchange of the row-slab decomposition over peer memory (NVLink / NVSwitch); no reference counterpart.
 * One process per GPU.  Each rank allocates its exchange buffers with fk_peer_alloc (cudaMalloc + a CUDA IPC handle),
 * hands the 64-byte handle to its neighbours (any host channel: torch.distributed.all_gather_object) and maps theirs with
 * fk_peer_open.  fk_euler_rows_peer is fk_euler_rows whose result rows [mirror->row0[n], row1[n]) are ALSO stored into
 * the mapped arrays of neighbour n (0 = the slab above, 1 = below) starting at its row dst_row0[n]: by the streaming
 * kernel itself, as it produces them (peer stores that overlap the interior's arithmetic row by row; *fused_out = 1), or,
 * when the call is not a streaming launch, by copies the library enqueues after it (*fused_out = 0).  Either way the
 * rows are in the neighbours' memory once the stream reaches the next operation: fk_peer_signal then writes a sequence
 * number into a flag in the neighbour's memory (system-scope release), and the neighbour orders its next launch after
 * fk_peer_wait(own flag, sequence number) -- a stream memory operation (cuStreamWaitValue32, >=), no SM is held while
 * waiting and no host thread is involved. */
typedef struct FkPeerMirror {
    float* v[2];          /* peer-mapped (H', W) arrays of the neighbour above [0] / below [1]; u[n] NULL = no neighbour */
    float* w[2];
    float* u[2];
    int row0[2], row1[2]; /* local rows to mirror */
    int dst_row0[2];      /* first destination row in the neighbour's arrays */
} FkPeerMirror;

int fk_peer_alloc(size_t bytes, void** dev_ptr_out, unsigned char* handle64_out);
int fk_peer_open(const unsigned char* handle64, void** dev_ptr_out);
int fk_peer_close(void* mapped_dev_ptr);
int fk_peer_free(void* dev_ptr);
int fk_peer_signal(unsigned int* flag_peer_dev, unsigned int value, void* stream);
int fk_peer_wait(const unsigned int* flag_local_dev, unsigned int value, void* stream);
/* plain asynchronous copy between any two device pointers this process can address (own or peer-mapped) */
int fk_peer_copy(void* dst_dev, const void* src_dev, size_t bytes, void* stream);
int fk_euler_rows_peer(const float* v_in_dev, const float* w_in_dev, const float* u_in_dev, float* v_out_dev,
                       float* w_out_dev, float* u_out_dev, const float* diffusivity_dev, const float* dx_map_dev,
                       const float* dy_map_dev, int H, int W, const FkParams* params, const FkStimulus* stimuli,
                       int n_stim, double t0, int nsteps, float dt, float dx, const FkOptions* opt, int row0, int row1,
                       void* workspace_dev, size_t workspace_bytes, void* stream, const FkPeerMirror* mirror,
                       int* fused_out);

/* solve._forward_dormandprince / solve.step_rk (cardiax/solve.py:88-89, 114-124) ==
 * jax.experimental.ode.odeint(step, state, ts, params, diffusivity, stimuli, dx): adaptive Dormand-Prince 5(4) with
 * jax's step-size controller and 4th-order dense output (jax/experimental/ode.py, an un-vendored dependency: algorithm
 * restated in csrc/fk_ode.h).  The right-hand side is solve.step evaluated at the CONTINUOUS fp32 time t (the reference
 * passes its step-unit checkpoints as times and never uses dt here).
 *   ts                HOST array of n_ts fp32 output times, increasing
 *   v/w/u_out_dev     (n_ts, batch, H, W) each: the state at every ts[i]; entry 0 is the initial state
 *   rtol, atol, mxstep  odeint's keyword defaults are 1.4e-8, 1.4e-8, +inf
 *   stats3            optional HOST {attempted steps, accepted steps, right-hand-side evaluations}
 * The arrays stay on the device; the step-size controller runs on the host and reads ONE reduced scalar per attempted
 * step, so this entry point synchronises `stream` (the reference's odeint is a device-side while loop).
 * A batch is integrated as ONE system (one common step size), like a raveled pytree. */
size_t fk_dopri5_workspace_bytes(int H, int W, int batch, int n_stim, int diffusivity_batched);
int fk_odeint_dopri5(const float* v0_dev, const float* w0_dev, const float* u0_dev, float* v_out_dev, float* w_out_dev,
                     float* u_out_dev, const float* diffusivity_dev, int diffusivity_batched, int H, int W, int batch,
                     const FkParams* params, const FkStimulus* stimuli, int n_stim, const float* ts, int n_ts, float dx,
                     float rtol, float atol, double mxstep, const FkOptions* opt, void* workspace_dev,
                     size_t workspace_bytes, void* stream, long long* stats3);

/* io.imresize (cardiax/io.py:118-124) == jax.image.resize(a, a.shape[:-2] + (Ho, Wo), "bilinear"): separable
 * anti-aliased triangle filter, half-pixel centres, weights renormalised at the edges.  One launch resizes n_planes
 * (H, W) arrays -- e.g. the v, w, u of a snapshot -- into one packed (n_planes, Ho, Wo) array.
 *   planes            HOST array of n_planes device pointers
 *   workspace_ready   0: the filter tables of this (H, W, Ho, Wo) are built and uploaded into the workspace first;
 *                     1: the caller kept the workspace of an earlier call with the same shapes (nothing is uploaded
 *                        for n_planes <= 8: a snapshot costs one kernel launch) */
size_t fk_resize_workspace_bytes(int H, int W, int Ho, int Wo, int n_planes);
int fk_resize_bilinear(const float* const* planes, int n_planes, int H, int W, float* out_dev, int Ho, int Wo,
                       void* workspace_dev, size_t workspace_bytes, int workspace_ready, void* stream);

/* metrics.electrogram (cardiax/metrics.py:13-22): out[f] = sum_ij x[f][i][j] * sqrt((j - p0)^2 + (i - p1)^2) for every
 * (H, W) frame of x (the reference multiplies by the distance; its ogrid only broadcasts for H == W). */
int fk_electrogram(const float* x_dev, int frames, int H, int W, float p0, float p1, float* out_dev, void* stream);

/* Exact numerics divide by the run's constants (time constants, dx) with q = RN(a * RN(1/b)); RN(q + (a - b q) RN(1/b)).
 * This verifies that sequence against __fdiv_rn on the device for every significand x three exponents x both signs x
 * every divisor of (params, dx) and returns the number of mismatches (0 expected; if not, set
 * FkOptions.safe_division).  Synchronises `stream`. */
int fk_check_exact_division(const FkParams* params, float dx, long long* mismatches, void* stream);

/* solve.step (cardiax/solve.py:26-65): the time derivatives (d_v, d_w, d_u) at counter t. */
int fk_rhs(const float* v_dev, const float* w_dev, const float* u_dev, float* dv_dev, float* dw_dev, float* du_dev,
           const float* diffusivity_dev, int diffusivity_batched, int H, int W, int batch, const FkParams* params,
           const FkStimulus* stimuli, int n_stim, double t, float dx, const FkOptions* opt, void* workspace_dev,
           size_t workspace_bytes, void* stream);

/* solve.gradient (cardiax/solve.py:225-254) along one axis of an N-D array viewed as
 * (outer, n, inner); n >= 5.  NOT divided by dx, exactly like the reference. */
int fk_gradient(const float* a_dev, float* out_dev, long long outer, long long n, long long inner, void* stream);

/* solve.stimulate (cardiax/solve.py:257-271) on one (H, W) array. */
int fk_stimulate(double t, int t_is_int, const float* x_dev, float* out_dev, int H, int W, const FkStimulus* stimuli,
                 int n_stim, void* workspace_dev, size_t workspace_bytes, void* stream);

/* D_x, D_y of solve.py:53-54 (gradient of the edge-padded map / dx, cropped), the two static
 * maps the step kernels read next to D. */
int fk_diffusivity_gradients(const float* diffusivity_dev, float* dx_out_dev, float* dy_out_dev, int H, int W,
                             int batch, float dx, int phys_top, int phys_bottom, void* stream);

/* Instrumentation for bench.py (not part of the reference surface).
 * fk_launch_count: kernels this library has launched since it was loaded.
 * fk_profile_enable(1): record a CUDA event pair, on the launch stream, around every step-kernel launch;
 * fk_profile_collect: wait for them, return summed device milliseconds and launch counts of the streaming kernel
 * and of the general tile kernel since the last collect (plus the cell-steps those streaming launches produced), and
 * reset. */
long long fk_launch_count(void);
/* geometry of the most recent streaming-kernel launch: {T, cta_threads, strips, columns per strip, rows per CTA,
 * row chunks, resident CTAs per SM, dynamic shared memory bytes} */
void fk_last_plan(int* out8);
/* name of the step kernel launched most recently: "fk_stream_kernel", "fk_resident_kernel", "fk_wide_kernel",
 * "fk_tile_kernel" (resident launches report {steps, cta_threads, tile columns, tile width, tile height, tile rows,
 * cells per thread (+ 8 if the diffusivity maps stay in L2), shared memory bytes} through fk_last_plan) */
const char* fk_last_kernel(void);
/* Development aid: with FK_RES_TIMING=1 in the environment, CTA (0, 0) of every resident launch accumulates SM cycles
 * spent in {ring groups, interior groups, waiting for and copying the halo, block barrier} (out8[0..3]) and the step
 * count (out8[6]); this returns the counters of the last launch (synchronises the device). */
int fk_resident_timing(unsigned long long* out8);
void fk_profile_enable(int on);
int fk_profile_collect(double* stream_ms, long long* stream_launches, double* tile_ms, long long* tile_launches,
                       double* stream_cell_steps);
/* launches that were NOT timed since the last call because 2^19 event pairs were already pending (0 in any sane
 * collection interval); resets the counter.  A non-zero value means the collected sums are a sample, not the total. */
long long fk_profile_dropped(void);

#ifdef __cplusplus
}
#endif
#endif /* FK_H_ */
