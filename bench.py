#!/usr/bin/env python
"""bench.py -- Gcell-steps/s of the fp32 Fenton-Karma Euler loop (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

One "step" = one checkpoint segment of SEG Euler steps (solve._forward_euler over [t, t+SEG)), the
unit the reference's `forward` loop launches (cardiax/solve.py:198-217).

Workloads (BASELINE.json configs):
  fk4096   4096 x 4096 homogeneous D = 1e-3, PARAMSET_5, random rectangular excitations  (N = 1 default)
  slab     row-slab decomposition of a (2048*N) x 16384 tissue over N GPUs (N = 8: BASELINE config 5, 16384^2),
           NCCL halo exchange over NVLink overlapped with the interior (N > 1 default)
  ens256   ensemble of independent 256 x 256 tissues (128 per GPU), heterogeneous D, 3 stimuli each, no comms
  fk512    512 x 512 scar map + S1-S2 (config 2)
  fk128    128 x 128 plane wave (config 1, the README benchmark shape)

Rank 0 prints ONE JSON line.  `value` is device-timed (CUDA events, barrier + synchronize on both
sides, max over ranks) with the state resident in HBM; `e2e` goes through the public
cardiax.solve API from pinned HOST buffers with the H2D and D2H copies inside the timed region.
`--impl reference` times the reference's CPU path (the C/OpenMP oracle port: the reference itself
needs a 2021 JAX that is not installable here) on the host cores for the same metric.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

ALG_BYTES = 28.0  # read u, v, w, D + write u, v, w (fp32) per cell-step
SEG = 500         # Euler steps per checkpoint segment (experiments/generate_fd_data_256.py:11-12)


def ncu_traffic(workload):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture (profiles/), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic_r02.json")) as f:
            return float(json.load(f)[workload]["dram_bytes_per_launch"])
    except Exception:
        return None


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


# ------------------------------------------------------------------ clocks
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML in a thread of this process (a handful of driver
    queries every 50 ms).  An `nvidia-smi -lms` child did the same job until its queries were seen to cost the unprofiled
    4096^2 pass 5 % (tools/probe_step_jitter.py) and its start-up to stall launches; it remains the fallback."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index, self.nvml, self.run = [], None, index, None, False

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            pynvml.nvmlDeviceGetClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.nvml, self.run = pynvml, True
            threading.Thread(target=self._poll, daemon=True).start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _poll(self):
        N = self.nvml
        names = (("hw_slowdown", N.nvmlClocksEventReasonHwSlowdown), ("hw_thermal_slowdown", N.nvmlClocksEventReasonHwThermalSlowdown),
                 ("sw_thermal_slowdown", N.nvmlClocksEventReasonSwThermalSlowdown), ("sw_power_cap", N.nvmlClocksEventReasonSwPowerCap))
        while self.run:
            try:
                sm = N.nvmlDeviceGetClockInfo(self.h, N.NVML_CLOCK_SM)
                mx = N.nvmlDeviceGetMaxClockInfo(self.h, N.NVML_CLOCK_SM)
                r = N.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                self.rows.append([str(self.index), str(sm), str(mx), ""] + ["Active" if r & bit else "Not Active" for _, bit in names])
            except Exception:
                pass
            time.sleep(0.05)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.nvml is not None:
            time.sleep(0.06)
            self.run = False
        elif self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        else:
            time.sleep(0.15)
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in list(self.rows):
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "via": "nvml" if self.nvml is not None else "nvidia-smi"}


# ------------------------------------------------------------------ workloads (host side, NumPy)
# Stimulus masks come from `S`: the product's own cardiax_b200.stimulus on the GPU arm, the oracle's builders on the
# reference (CPU) arm -- the same fields (tests/test_reference_pin.py::test_mask_builders).  Scar maps come from the
# product's generator (cardiax_b200.generate.random_diffusivity: NumPy / SciPy, seeded) on both arms.
def _np(x):
    return x.detach().cpu().numpy() if hasattr(x, "detach") else np.asarray(x)


def _stim(S, proto, field):
    return (tuple(float(np.asarray(_np(p)).reshape(-1)[0]) for p in proto), np.ascontiguousarray(_np(field), dtype=np.float32))


def scar_map(shape, seed):
    from cardiax_b200 import generate
    return np.ascontiguousarray(generate.random_diffusivity(np.random.default_rng(seed), shape), dtype=np.float32)


def ens_scar_map(shape, seed):
    """Scar map of the ensemble tissues: the generator's map smoothed by 3 more cells.  The generator's sharpest maps
    (Gaussian taper of one cell: D falls tenfold within three cells) make the reference's own discretisation -- the
    D_x u_x + D_y u_y terms on one-sided boundary stencils -- unstable at a tissue corner within ~1500 steps in a few
    tissues out of 128 (u -> -inf in exact and in fast numerics alike, CPU emulation and GPU); a throughput of NaNs is
    not a measurement, so the bench runs the ensemble on maps that stay finite and counts the tissues that do."""
    from scipy import ndimage
    return np.ascontiguousarray(ndimage.gaussian_filter(scar_map(shape, seed), 3.0, mode="nearest"), dtype=np.float32)


def make_fk4096(S=None, H=4096, W=4096, seed=0):
    rng = np.random.default_rng(seed)
    u = np.zeros((H, W), np.float32)
    for _ in range(max(4, H * W // 400000)):
        r, c = rng.integers(0, H - 64), rng.integers(0, W - 64)
        u[r:r + 48, c:c + 48] = 1.0
    return dict(v=np.ones((H, W), np.float32), w=np.ones((H, W), np.float32), u=u,
                D=np.full((H, W), 1e-3, np.float32), stimuli=[], params="5")


def make_fk128(S):
    """BASELINE config 1 (the README benchmark): 128 x 128, PARAMSET_3, D = 1e-3, NORTH stripe, one stimulus."""
    shape = (128, 128)
    s = S.linear(shape, 0, 0.2, 20.0, S.Protocol(0, 2, 1e9))
    return dict(v=np.ones(shape, np.float32), w=np.ones(shape, np.float32), u=np.zeros(shape, np.float32),
                D=np.full(shape, 1e-3, np.float32), stimuli=[_stim(S, s.protocol, s.field)], params="3")


def make_fk512(S, seed=0):
    """BASELINE config 2: 512 x 512 scar map, S1 at step 0, S2 (rotated by 90 degrees) at step 40 000
    (experiments/generate_fd_data_256.py:22-40)."""
    shape = (512, 512)
    ang = float(np.random.default_rng(seed).uniform(0, 180))
    s1 = S.triangular(shape, 0, ang, 0.2, 20.0, S.Protocol(0, 2, 1e9))
    s2 = S.triangular(shape, 0, ang + 90, 0.5, 20.0, S.Protocol(40000, 2, 1e9))
    return dict(v=np.ones(shape, np.float32), w=np.ones(shape, np.float32), u=np.zeros(shape, np.float32),
                D=scar_map(shape, seed), stimuli=[_stim(S, s.protocol, s.field) for s in (s1, s2)], params="3")


def make_fk1200(S=None):
    """the reference's data-generation tissue: 1200 x 1200, scar-map D (deepx/generate.py:92)"""
    wq = make_fk4096(None, 1200, 1200)
    wq.update(D=scar_map((1200, 1200), 0), params="3")
    return wq


def make_ens256(S, nsims, seed=0):
    """BASELINE config 4, one GPU's share: `nsims` independent 256 x 256 tissues, scar D, three stimuli each drawn like
    deepx/generate.py:59-76 (type in {rectangular, triangular, linear}, amplitude 20, duration 2)."""
    shape = (256, 256)
    rng = np.random.default_rng(seed)
    D = np.empty((nsims,) + shape, np.float32)
    stim = []
    for b in range(nsims):
        D[b] = ens_scar_map(shape, seed * 100003 + b)
        ss = []
        for k in range(3):
            kind = rng.integers(0, 3)
            proto = S.Protocol(int(1 + k * 25000 + rng.integers(0, 2)), 2, int(rng.integers(400, 10 ** 9)))
            if kind == 0:
                size = rng.integers(10, 85, 2)
                s = S.rectangular(shape, rng.integers(int(size.min()), 256, 2), size, 20.0, proto)
            elif kind == 1:
                s = S.triangular(shape, int(rng.integers(0, 3)), float(abs(rng.normal()) * 45),
                                 float(abs(rng.normal()) * 0.2), 20.0, proto)
            else:
                s = S.linear(shape, int(rng.integers(0, 3)), float(abs(rng.normal()) * 0.2), 20.0, proto)
            ss.append(_stim(S, s.protocol, s.field))
        stim.append(ss)
    return dict(v=np.ones((nsims,) + shape, np.float32), w=np.ones((nsims,) + shape, np.float32),
                u=np.zeros((nsims,) + shape, np.float32), D=D, stimuli=stim, params="3")


def build_work(workload, S, rank, world):
    if workload == "fk4096":
        return make_fk4096(), {"workload": "fk4096: 4096x4096 homogeneous D=1e-3, PARAMSET_5, random rectangular excitations",
                               "grid": [4096, 4096]}
    if workload == "slab":
        return make_fk4096(None, 2048, 16384, seed=rank), {
            "workload": "slab: (2048*N)x16384 tissue (16384^2 at N=8), row slabs of 2048 rows per GPU; halo rows stored into "
                        "the neighbouring GPUs' memory by the step kernel itself (peer-mapped buffers over NVLink, flags)",
            "grid": [2048 * world, 16384]}
    if workload == "fk512":
        return make_fk512(S), {"workload": "fk512: 512x512 scar-map D, S1-S2 cross-field stimuli, PARAMSET_3", "grid": [512, 512]}
    if workload == "fk128":
        return make_fk128(S), {"workload": "fk128: 128x128 plane wave, PARAMSET_3, D=1e-3 (README benchmark shape)", "grid": [128, 128]}
    if workload == "ens256":
        return make_ens256(S, 128, seed=rank), {
            "workload": "ens256: 128 independent 256x256 tissues per GPU, scar D, 3 random stimuli each, no communication",
            "grid": [128 * world, 256, 256]}
    raise SystemExit("unknown workload " + workload)


# ------------------------------------------------------------------ reference arm / cpu baseline
def cpu_run(work, euler_steps, repeats=1):
    """C/OpenMP port of the reference loop on all host cores; returns (cell_steps_per_s, cores, seconds)."""
    import oracle as O
    from oracle import c_oracle as C
    cores = C.set_threads(os.cpu_count() or 1)
    st = O.State(work["v"], work["w"], work["u"])
    if st.u.ndim == 3:  # ensemble: one tissue is the sample
        st = O.State(st.v[0], st.w[0], st.u[0])
        D, stim = work["D"][0], work["stimuli"][0]
    else:
        D, stim = work["D"], work["stimuli"]
    stim = [O.Stimulus(O.Protocol(*p), f) for p, f in stim]
    cells = st.u.size
    t0 = time.perf_counter()
    for _ in range(repeats):
        C.forward_euler(st, 0, euler_steps, O.PARAMSETS[work["params"]], D, stim, 0.01, 0.01)
    dt = time.perf_counter() - t0
    return cells * euler_steps * repeats / dt, cores, dt


def reference_arm(args, workload, work, config):
    cells = work["u"].shape[-1] * work["u"].shape[-2]
    n_euler = max(1, int(2.0e8 // cells))  # ~2e8 cell-steps (about a second of CPU work) per bench step
    for _ in range(args.warmup):
        cpu_run(work, 1)
    t0 = time.perf_counter()
    cores = 1
    for _ in range(args.steps):
        _, cores, _ = cpu_run(work, n_euler)
    dt = time.perf_counter() - t0
    val = cells * n_euler * args.steps / dt / 1e9
    sample = "%s: one tissue, %d Euler steps per bench step" % (workload, n_euler)
    print(json.dumps({
        "impl": "reference", "metric": "Gcell-steps/s fp32 FK", "value": val, "unit": "Gcell-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config,
        "cpu_baseline": {"value": val, "unit": "Gcell-steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "Gcell-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference = C/OpenMP restatement of cardiax/solve.py (oracle/fk_oracle.c); the JAX reference is not installable here",
    }))


# ------------------------------------------------------------------ main
def numa_bind(local_rank):
    """Pin this process (and so the first-touch placement of its pinned buffers) to the CPUs of the NUMA node its GPU
    hangs off: at 8 ranks the host copies of the e2e leg otherwise cross the socket interconnect."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
        dom = torch.cuda.get_device_properties(local_rank).pci_domain_id
        dev_id = torch.cuda.get_device_properties(local_rank).pci_device_id
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/numa_node" % (dom, bus, dev_id)
        node = int(open(path).read().strip())
        if node < 0:
            return None
        cpus = []
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus += list(range(int(lo), int(hi or lo) + 1))
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None)
    ap.add_argument("--seg", type=int, default=SEG)
    ap.add_argument("--numerics", default="fast", choices=["fast", "exact"])
    ap.add_argument("--T", type=int, default=0)
    ap.add_argument("--cta-threads", type=int, default=0)
    ap.add_argument("--rows-per-cta", type=int, default=0)
    ap.add_argument("--kernel", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--halo-launches", type=int, default=8)
    ap.add_argument("--comm", default=None, choices=["peer", "dist"], help="slab halo exchange: fused peer stores (default) or NCCL")
    ap.add_argument("--no-extra", action="store_true",
                    help="skip the lines beside the headline (other BASELINE configs, ensemble, single-GPU slab shape, checks)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    workload = args.workload or ("fk4096" if args.gpus == 1 else "slab")
    l2_note = {"euler_steps_per_step": args.seg, "dt": 0.01, "dx": 0.01, "numerics": args.numerics,
               "l2": "state (3 arrays of >= 64 MiB per GPU) larger than the 126 MB L2" if workload in ("fk4096", "slab")
               else "working set is L2 resident by nature of the config; an L2 flush buffer is written between steps"}

    if args.impl == "reference":
        if rank != 0:
            return
        import oracle as O   # the reference arm is the one place bench.py executes oracle/
        work, config = build_work(workload, O, rank, world)
        config.update(l2_note)
        return reference_arm(args, workload, work, config)

    import ctypes
    import torch
    from cardiax_b200 import _lib, options, solve, stimulus
    from cardiax_b200 import params as fkparams

    assert torch.cuda.is_available(), "bench.py needs a CUDA device"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = numa_bind(local_rank) if world > 1 else None
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    if _lib.needs_build():
        _lib.build()
    L = _lib.lib()
    options.verbose = False
    options.numerics = args.numerics
    options.steps_per_launch = args.T
    options.cta_threads, options.rows_per_cta, options.kernel = args.cta_threads, args.rows_per_cta, args.kernel
    work, config = build_work(workload, stimulus, rank, world)
    config.update(l2_note)
    pset = lambda name: getattr(fkparams, "PARAMSET_" + name)  # noqa: E731
    params = pset(work["params"])
    seg = args.seg

    def dev_stim(stimuli):
        if len(stimuli) and isinstance(stimuli[0], list):
            return [dev_stim(s) for s in stimuli]
        return [stimulus.Stimulus(stimulus.Protocol(*p), torch.as_tensor(f).to(dev)) for p, f in stimuli]

    def dev_state(wk):
        return solve.State(*[torch.as_tensor(wk[k]).to(dev) for k in "vwu"])

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return float(x)
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed_segments(step, n_warm, n_timed, flush_buf=None):
        """Device time (ms, CUDA events on the launch stream) of n_timed calls of step(i); L2 flushed outside the timing."""
        for i in range(n_warm):
            step(i)
        torch.cuda.synchronize()
        # (no host synchronisation between the segments: the first segments after one run slower -- 21.9 and 23.6 ms for the
        # first two 500-step segments of the ensemble against 20.2 for the others -- and with one after every segment
        # every segment was a first one)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_timed)]
        for i in range(n_timed):
            if flush_buf is not None:
                flush_buf.fill_(1.0)
            ev[i][0].record()
            step(n_warm + i)
            ev[i][1].record()
        torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b in ev)

    gstim = dev_stim(work["stimuli"])
    D = torch.as_tensor(work["D"]).to(dev)
    state0 = dev_state(work)
    cells = state0.u.numel()
    flush = None
    if workload in ("fk512", "ens256", "fk128"):
        flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)

    runner = None
    if workload == "slab" and world > 1:
        from cardiax_b200 import slab
        runner = slab.SlabRunner(state0, D, params, gstim, 0.01, 0.01, rank, world, steps_per_launch=args.T or 0,
                                 halo_launches=args.halo_launches, comm=args.comm)
        step_fn = lambda st, t: runner.advance(None, t, t + seg, copy=False)  # noqa: E731  (the state stays resident)
    else:
        step_fn = lambda st, t: solve._forward_euler(st, t, t + seg, params, D, gstim, 0.01, 0.01)  # noqa: E731

    # ---- device-resident measurement
    st, t = state0, 0
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()   # (before the warm-up: NVML's / nvidia-smi's start-up takes driver locks that can stall launches)
        time.sleep(0.2)
    for _ in range(args.warmup):
        if flush is not None:
            flush.fill_(1.0)   # (the fill kernel's first use loads its module: ~10 ms that do not belong in the timed region)
        st = step_fn(st, t); t += seg
    barrier()
    if rank == 0:
        sampler.rows.clear()   # only the samples taken during the timed region count
    # The timed region of `value`: exactly K steps as the product runs them (no per-launch events: the pair the library
    # records around every launch when profiling is on costs ~6 % of a 4096^2 step -- 357 vs 378 Gcell-steps/s --
    # so the per-launch figures of `roofline` come from a second pass of the same K steps right after, with its own clock)
    def timed_pass(st, t, profile):
        L.fk_profile_enable(1 if profile else 0)
        L.fk_profile_collect(None, None, None, None, None)
        L.fk_profile_dropped()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
        barrier()
        e0.record()
        for k in range(args.steps):
            if flush is not None:
                flush.fill_(1.0)
            st = step_fn(st, t); t += seg
            marks[k].record()   # (one event per 500-step segment: how evenly the steps of the region ran)
        e1.record()
        barrier()
        per_step = [a.elapsed_time(b) for a, b in zip([e0] + marks[:-1], marks)]
        return st, t, e0.elapsed_time(e1), per_step

    launches0 = L.fk_launch_count()
    st, t, ms, step_ms = timed_pass(st, t, False)
    launches = L.fk_launch_count() - launches0
    # One step far above the others is the HOST falling behind (the GPU idles while a descheduled launch thread catches up:
    # seen in about one run out of ten on the virtualised boxes, tools/probe_step_jitter.py), not the path being measured:
    # such a pass is timed once more, on every rank, and the first attempt is reported beside it (`retimed`).
    retimed = None
    if args.steps >= 4 and max_over_ranks(float(max(step_ms) > 1.5 * float(np.median(step_ms)))) > 0.0:
        retimed = {"first_attempt_ms_per_step": ms / args.steps, "first_attempt_step_ms": [round(x, 3) for x in step_ms]}
        st, t, ms, step_ms = timed_pass(st, t, False)
    clocks = sampler.stop() if rank == 0 else None
    st, t, ms_prof, _ = timed_pass(st, t, not os.environ.get("FK_BENCH_NOPROF"))
    sm_ms, sm_n, tl_ms, tl_n, sm_cs = ctypes.c_double(), ctypes.c_longlong(), ctypes.c_double(), ctypes.c_longlong(), ctypes.c_double()
    L.fk_profile_collect(ctypes.byref(sm_ms), ctypes.byref(sm_n), ctypes.byref(tl_ms), ctypes.byref(tl_n), ctypes.byref(sm_cs))
    prof_dropped = int(L.fk_profile_dropped())
    plan = _lib.last_plan()
    main_kernel_name = _lib.last_kernel()
    L.fk_profile_enable(0)
    if flush is not None:  # take the flush writes out: time them alone
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); f0.record()
        for _ in range(args.steps):
            flush.fill_(1.0)
        f1.record(); torch.cuda.synchronize()
        ms -= f0.elapsed_time(f1)
        ms_prof -= f0.elapsed_time(f1)
    assert all(bool(torch.isfinite(x).all()) for x in st)
    ms = max_over_ranks(ms)
    total_cells = cells * world
    value = total_cells * seg * args.steps / (ms * 1e-3) / 1e9
    t_end_main = t

    # ---- end to end through the public API from pinned host buffers: every step copies its input state host->device
    # and its result device->host inside the timed region.  The copies run on their own streams (double-buffered device
    # and host buffers), so step i+1's upload and step i-1's download overlap step i's kernels -- the way
    # cardiax_b200.io streams snapshots out; nothing is skipped and the last download is inside the timed region.
    NB = 2
    host_in = [torch.as_tensor(work[k]).pin_memory() for k in "vwu"]
    host_out = [[torch.empty_like(x).pin_memory() for x in host_in] for _ in range(NB)]
    dev_in = [[torch.empty(x.shape, dtype=torch.float32, device=dev) for x in host_in] for _ in range(NB)]
    h2d = sum(x.numel() * 4 for x in host_in)
    d2h = sum(x.numel() * 4 for x in host_out[0])
    s_in, s_out, s_run = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev), torch.cuda.current_stream()
    ev_in = [torch.cuda.Event() for _ in range(NB)]
    ev_free = [torch.cuda.Event() for _ in range(NB)]      # device input buffer consumed
    ev_done = [torch.cuda.Event() for _ in range(NB)]
    ev_out = [torch.cuda.Event() for _ in range(NB)]       # host output buffer written
    results = [None] * NB

    E2E_SKIP = os.environ.get("FK_E2E_SKIP", "")   # development: "h2d", "d2h" or both -- which copy the e2e loop leaves out

    def e2e_step(i, t):
        b = i % NB
        with torch.cuda.stream(s_in):
            s_in.wait_event(ev_free[b])
            if "h2d" not in E2E_SKIP:
                for d, h in zip(dev_in[b], host_in):
                    d.copy_(h, non_blocking=True)
            ev_in[b].record(s_in)
        s_run.wait_event(ev_in[b])
        s = solve.State(*dev_in[b])
        if runner is not None:
            s = runner.advance(s, t, t + seg, copy=True)    # loads the uploaded rows, refreshes the neighbours' halos
        else:
            s = step_fn(s, t)
        ev_free[b].record(s_run)
        ev_done[b].record(s_run)
        results[b] = s          # keep the tensors alive until their download has been issued
        with torch.cuda.stream(s_out):
            s_out.wait_event(ev_done[b])
            s_out.wait_event(ev_out[b])   # (host buffer b was last written NB steps ago on this same stream)
            for o, x in zip(host_out[b], s):
                if "d2h" not in E2E_SKIP:
                    o.copy_(x, non_blocking=True)
                x.record_stream(s_out)
            ev_out[b].record(s_out)

    # the host link as this box gives it (diagnostic next to e2e: a box with a busy or degraded PCIe link shows here)
    link = {}
    for name, dst, src in (("h2d_gbs", dev_in[0][0], host_in[0]), ("d2h_gbs", host_out[0][0], dev_in[0][0])):
        dst.copy_(src, non_blocking=True); torch.cuda.synchronize()
        l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0.record(); dst.copy_(src, non_blocking=True); l1.record(); torch.cuda.synchronize()
        link[name] = src.numel() * 4 / (l0.elapsed_time(l1) * 1e-3) / 1e9
    # ... and with every rank copying in BOTH directions at once (what the e2e loop does): on a box whose host memory is
    # the shared resource this, not the kernels, bounds e2e at large N
    barrier()
    l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0.record()
    s_in.wait_event(l0); s_out.wait_event(l0)
    with torch.cuda.stream(s_in):
        for d, h in zip(dev_in[0], host_in):
            d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s_out):
        for o, x in zip(host_out[1], dev_in[1]):
            o.copy_(x, non_blocking=True)
    s_run.wait_stream(s_in); s_run.wait_stream(s_out)
    l1.record(); torch.cuda.synchronize()
    both_ms = max_over_ranks(l0.elapsed_time(l1))
    link["all_ranks_both_directions_ms"] = both_ms
    link["all_ranks_both_directions_gbs_each_way"] = h2d / (both_ms * 1e-3) / 1e9
    for b in range(NB):
        ev_free[b].record(s_run); ev_out[b].record(s_out)
    for i in range(NB):      # every host / device buffer used once before the clock starts: the first DMA into freshly pinned
        e2e_step(i, 0)       # pages took 200-300 ms (first_attempt_step_ms of profiles/bench_r03j_default.json)
    barrier()
    def e2e_pass():
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
        g0.record()
        for i in range(args.steps):
            e2e_step(i, i * seg)
            marks[i].record()          # (on the run stream: when step i's kernels are done)
        s_run.wait_stream(s_out)     # the last result is on the host before the clock stops
        g1.record()
        barrier()
        return g0.elapsed_time(g1), [a_.elapsed_time(b_) for a_, b_ in zip([g0] + marks[:-1], marks)]

    ems, e2e_step_ms = e2e_pass()
    e2e_retimed = None
    if args.steps >= 4 and max_over_ranks(float(max(e2e_step_ms) > 1.5 * float(np.median(e2e_step_ms)))) > 0.0:
        e2e_retimed = {"first_attempt_ms_per_step": ems / args.steps, "first_attempt_step_ms": [round(x, 3) for x in e2e_step_ms]}
        ems, e2e_step_ms = e2e_pass()   # (same rule as the device-resident pass: one step far above the others = the host fell behind)
    assert all(bool(torch.isfinite(x).all()) for x in host_out[(args.steps - 1) % NB])
    ems = max_over_ranks(ems)
    e2e = total_cells * seg * args.steps / (ems * 1e-3) / 1e9
    del host_in, host_out, dev_in, results

    other = {} if not args.no_extra else None
    exit_code = 0

    # ---- slab: the exchange's exposed time, and the result checked against a single-GPU recomputation
    if runner is not None and other is not None:
        n_grp = 6
        grp_steps = runner.M * runner.T
        def slab_ms(comm_on):
            runner.comm_enabled = comm_on
            barrier()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            runner.advance(None, 0, n_grp * grp_steps, copy=False)
            a1.record()
            barrier()
            runner.comm_enabled = True
            return max_over_ranks(a0.elapsed_time(a1))
        slab_ms(True)
        with_comm = min(slab_ms(True) for _ in range(3))
        runner.load(state0)          # (the runs without communication leave wrong halos behind)
        slab_ms(False)
        without = min(slab_ms(False) for _ in range(3))
        other["exposed_us_per_exchange"] = {"value": (with_comm - without) * 1e3 / n_grp, "groups": n_grp,
                                            "steps_per_group": grp_steps, "ms_with_exchange": with_comm, "ms_without": without,
                                            "comm": runner.comm.kind, "fused_mirror_launches": getattr(runner.be, "fused_mirrors", None)}
        # check: a band around every rank boundary recomputed by ONE GPU as a tissue of its own.  With K = 4 n + 8 rows
        # on each side the recomputation's artificial top / bottom edges cannot reach the compared rows in n steps.
        n_chk, band = 40, 64
        K = 4 * n_chk + 8 + band
        runner.load(state0)
        out = runner.advance(None, 0, n_chk, copy=True)
        ok = True
        if rank + 1 < world:      # this rank checks the boundary below it: the neighbour sends its first K rows
            top_in = [torch.empty((K, state0.u.shape[1]), device=dev) for _ in range(3)]
            top_out = [torch.empty((band, state0.u.shape[1]), device=dev) for _ in range(3)]
        ops = []
        if rank > 0:
            ops += [dist.P2POp(dist.isend, x[:K].contiguous(), rank - 1) for x in state0]
            ops += [dist.P2POp(dist.isend, x[:band].contiguous(), rank - 1) for x in out]
        if rank + 1 < world:
            ops += [dist.P2POp(dist.irecv, x, rank + 1) for x in top_in]
            ops += [dist.P2POp(dist.irecv, x, rank + 1) for x in top_out]
        for w_ in dist.batch_isend_irecv(ops):
            w_.wait()
        if rank + 1 < world:
            win = solve.State(*[torch.cat([mine[-K:], theirs]) for mine, theirs in zip(state0, top_in)])
            ref = solve._forward_euler(win, 0, n_chk, params, torch.full(win.u.shape, 1e-3, device=dev), [], 0.01, 0.01)
            ok = all(torch.equal(r_[K - band:K], o[-band:]) and torch.equal(r_[K:K + band], t_)
                     for r_, o, t_ in zip(ref, out, top_out))
        flag = torch.tensor([int(ok)], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        other["slab_check"] = ("bit-equal" if int(flag.item()) else "MISMATCH") + \
            ": %d rows around each of the %d rank boundaries after %d steps vs a single-GPU recomputation" % (2 * band, world - 1, n_chk)
        if not int(flag.item()):
            exit_code = 1
        runner.close()
        runner = None

    # ---- beside the headline, at EVERY N: BASELINE config 4 (ensemble, 128 tissues per GPU, no communication) and the
    # slab shape on one GPU (the denominator of SURVEY 8e's weak-scaling efficiency)
    if other is not None:
        del st, state0, D
        fl = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
        wk = make_ens256(stimulus, 128, seed=rank)
        gs, Dk, pk = dev_stim(wk["stimuli"]), torch.as_tensor(wk["D"]).to(dev), pset(wk["params"])
        box = [dev_state(wk)]
        def ens_step(i):
            box[0] = solve._forward_euler(box[0], i * seg, (i + 1) * seg, pk, Dk, gs, 0.01, 0.01)
        barrier()
        n_seg = 6
        tot = max_over_ranks(timed_segments(ens_step, 3, n_seg, fl))
        other["ens256"] = {"value": world * box[0].u.numel() * seg * n_seg / (tot * 1e-3) / 1e9, "unit": "Gcell-steps/s",
                           "tissues": 128 * world, "tissues_per_gpu": 128, "grid": [256, 256], "segments": n_seg,
                           "euler_steps_per_segment": seg, "kernel": _lib.last_kernel(), "launch_geometry": _lib.last_plan(),
                           "finite_tissues_this_rank": int(sum(bool(torch.isfinite(box[0].u[b]).all()) for b in range(128))),
                           "note": "BASELINE config 4 sharded over the ranks with no communication; time = max over ranks"}
        other["ens256"]["frac_of_28B_roofline_per_gpu"] = ALG_BYTES * other["ens256"]["value"] / world / peaks()[0]
        del box, gs, Dk
        wk = make_fk4096(None, 2048, 16384, seed=rank)
        Dk, pk = torch.as_tensor(wk["D"]).to(dev), pset(wk["params"])
        box = [dev_state(wk)]
        def slab1_step(i):
            box[0] = solve._forward_euler(box[0], i * seg, (i + 1) * seg, pk, Dk, [], 0.01, 0.01)
        barrier()
        n_seg = 4
        tot = max_over_ranks(timed_segments(slab1_step, 2, n_seg))
        slab_n1 = box[0].u.numel() * seg * n_seg / (tot * 1e-3) / 1e9
        other["slab_n1"] = {"value": slab_n1, "unit": "Gcell-steps/s", "grid": [2048, 16384],
                            "note": "ONE GPU running a 2048x16384 tissue standalone (every rank measures it at the same time; slowest rank)"}
        if workload == "slab" and world > 1:
            other["efficiency_vs_slab_n1"] = value / (world * slab_n1)
        del box, Dk, fl

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        sys.exit(exit_code)

    # ---- BASELINE configs 2 and 1 beside the headline (N = 1 default run only): a few segments each, device resident
    if world == 1 and workload == "fk4096" and other is not None:
        fl = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
        for name, mk in (("fk512", make_fk512), ("fk128", make_fk128), ("fk1200", make_fk1200)):
            wk = mk(stimulus)
            gs, Dk = dev_stim(wk["stimuli"]), torch.as_tensor(wk["D"]).to(dev)
            pk = pset(wk["params"])
            box = [dev_state(wk)]
            def seg_step(i):
                box[0] = solve._forward_euler(box[0], i * seg, (i + 1) * seg, pk, Dk, gs, 0.01, 0.01)
            n_seg = 8
            tot = timed_segments(seg_step, 3, n_seg, fl)
            cs = box[0].u.numel() * seg * n_seg
            other[name] = {"value": cs / (tot * 1e-3) / 1e9, "unit": "Gcell-steps/s", "us_per_euler_step": tot * 1e3 / (seg * n_seg),
                           "kernel": _lib.last_kernel(), "launch_geometry": _lib.last_plan(), "segments": n_seg,
                           "euler_steps_per_segment": seg}
        # BASELINE config 2 at its REAL protocol: S2 fires at step 40 000 of 1e5 (generate_fd_data_256.py:22-40) -- the
        # whole run, 200 segments of 500 steps, timed end to end on the device
        wk = make_fk512(stimulus)
        gs, Dk, pk = dev_stim(wk["stimuli"]), torch.as_tensor(wk["D"]).to(dev), pset("3")
        sk = dev_state(wk)
        torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        u_before = None
        for i in range(200):
            if i == 80:
                u_before = sk.u
            sk = solve._forward_euler(sk, i * 500, (i + 1) * 500, pk, Dk, gs, 0.01, 0.01)
            if i == 80:
                s2_jump = float((sk.u - u_before).abs().max().item())    # S2 fired inside this segment (host read: one sync)
        a1.record(); torch.cuda.synchronize()
        other["fk512_full_protocol"] = {"value": sk.u.numel() * 1e5 / (a0.elapsed_time(a1) * 1e-3) / 1e9, "unit": "Gcell-steps/s",
                                        "euler_steps": 100000, "seconds": a0.elapsed_time(a1) * 1e-3, "s2_fired_at_step_40000": s2_jump > 0.05,
                                        "max_u_change_in_the_s2_segment": s2_jump, "finite": bool(torch.isfinite(sk.u).all())}
        # the Heun integrator (solve.py:73-85, 103-111; deepx/DataGeneration.ipynb selects it) on the headline tissue
        wk = work
        sk = dev_state(wk)
        Dk, pk, hs = torch.as_tensor(wk["D"]).to(dev), pset(wk["params"]), 40
        sk = solve._forward_heun(sk, 0, 4, pk, Dk, [], 0.01, 0.01)
        torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fl.fill_(1.0)
        a0.record()
        sk = solve._forward_heun(sk, 4, 4 + hs, pk, Dk, [], 0.01, 0.01)
        a1.record(); torch.cuda.synchronize()
        other["heun_fk4096"] = {"value": sk.u.numel() * hs / (a0.elapsed_time(a1) * 1e-3) / 1e9, "unit": "Gcell-steps/s (Heun steps)",
                                "kernel": _lib.last_kernel(), "heun_steps": hs}
        # exact numerics (bit-identical to the reference's source, tests/test_reference_pin.py) on the headline tissue
        options.numerics = "exact"
        sk = dev_state(wk)
        sk = solve._forward_euler(sk, 0, 20, pk, Dk, [], 0.01, 0.01)
        torch.cuda.synchronize()
        a0.record()
        sk = solve._forward_euler(sk, 20, 120, pk, Dk, [], 0.01, 0.01)
        a1.record(); torch.cuda.synchronize()
        other["exact_fk4096"] = {"value": sk.u.numel() * 100 / (a0.elapsed_time(a1) * 1e-3) / 1e9, "unit": "Gcell-steps/s",
                                 "kernel": _lib.last_kernel()}
        options.numerics = args.numerics
        # the snapshot path of deepx.generate.sequence on the reference's 1200^2 tissue: 500 Euler steps, resize kernel to
        # 256^2, pinned D2H on a side stream, writer thread -- per-segment time with and without snapshots
        from cardiax_b200 import io as fio
        wk = make_fk1200()
        Dk, pk = torch.as_tensor(wk["D"]).to(dev), pset("3")
        sk = dev_state(wk)
        n_seg = 8
        snaps = np.zeros((n_seg + 2, 3, 256, 256), np.float32)
        res = {}
        for mode in ("solver_only", "with_snapshots"):
            writer = fio.AsyncSnapshotWriter(snaps, (3, 256, 256)) if mode == "with_snapshots" else None
            for i in range(2):
                sk = solve._forward_euler(sk, i * seg, (i + 1) * seg, pk, Dk, [], 0.01, 0.01)
                if writer: writer.submit(sk, i)
            torch.cuda.synchronize()
            w0 = time.perf_counter()
            for i in range(2, n_seg + 2):
                sk = solve._forward_euler(sk, i * seg, (i + 1) * seg, pk, Dk, [], 0.01, 0.01)
                if writer: writer.submit(sk, i)
            if writer: writer.close()
            torch.cuda.synchronize()
            res[mode] = (time.perf_counter() - w0) / n_seg * 1e3
        other["snapshots_fk1200"] = {"ms_per_500_step_segment": res, "snapshot": "3 x 1200^2 -> 3 x 256^2 fp32, async"}
        del fl
        # the reference's README table (README.md:25-31): wall seconds of forward() over 1e3 steps, field size varied,
        # two stimuli (cardiax's fenton_karma notebooks); host wall clock, call + synchronize, best of 3
        table = {}
        for n in (64, 128, 256, 512, 1024):
            shp = (n, n)
            s1 = stimulus.linear(shp, stimulus.Direction.NORTH, 0.2, 20.0, stimulus.Protocol(0, 2, 1e9))
            s2 = stimulus.linear(shp, stimulus.Direction.EAST, 0.2, 20.0, stimulus.Protocol(200, 2, 1e9))
            Dn = torch.full(shp, 1e-3, device=dev)
            best = 1e30
            for _ in range(4):
                s0 = solve.init(shp)
                torch.cuda.synchronize()
                w0 = time.perf_counter()
                out_states = solve.forward(s0, [0, 1000], pset("3"), Dn, [s1, s2], 0.01, 0.01)
                torch.cuda.synchronize()
                best = min(best, time.perf_counter() - w0)
            assert bool(torch.isfinite(out_states[-1].u).all())
            table[str(n)] = best
        other["readme_table_seconds_1e3_steps"] = {
            "ours_b200": table,
            "reference_published": {"jax_cpu_2vcpu": {"64": 0.336, "128": 0.904, "256": 2.94, "512": 11.1, "1024": 45.0},
                                    "jax_gpu_t4": {"64": 0.193, "128": 0.189, "256": 0.199, "512": 0.237, "1024": 0.613},
                                    "jax_tpu": {"64": 0.059, "128": 0.074, "256": 0.119, "512": 0.272, "1024": 0.842},
                                    "source": "reference README.md:28-31 (older API, other hardware: context only)"}}

    peak, peak_kind = peaks()
    roof = None
    main_kernel = main_kernel_name
    traffic = ncu_traffic(workload)
    if sm_n.value > 0:
        # streaming kernel: T steps over the rows/columns it owns per launch (counted by the library per launch)
        cs_per_launch = sm_cs.value / sm_n.value
        avg_ms = sm_ms.value / sm_n.value
        achieved = ALG_BYTES * cs_per_launch / (avg_ms * 1e-3) / 1e9
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_kind": peak_kind, "kernel": main_kernel,
                "dram_frac": (traffic / (avg_ms * 1e-3) / 1e9 / peak) if traffic else None,
                "alg_bytes_per_launch": ALG_BYTES * cs_per_launch, "avg_launch_ms": avg_ms,
                "kernel_share_of_step": sm_ms.value / ms_prof, "tile_kernel_share_of_step": tl_ms.value / ms_prof,
                "launches_timed": int(sm_n.value), "launches_not_timed": prof_dropped,
                "timed_in": "a second pass of the same %d steps with a CUDA event pair around every launch (%.3f ms per step "
                            "against %.3f in the pass `value` is timed in, which records none)" % (
                                args.steps, ms_prof / args.steps, ms / args.steps),
                "note": ("the streaming kernel covers the whole tissue, physical edges included; the rest of the step is "
                         "fk_dgrad_kernel (once per call) and launch gaps; dram_frac = the ncu-measured DRAM bytes of one "
                         "launch / its event-timed duration / peak") if main_kernel == "fk_stream_kernel" else
                        ("one resident launch per segment: the state lives in shared memory for all of its Euler steps and "
                         "halos travel through L2, so HBM sees only the segment's first load and last store; the fraction "
                         "compares the ALGORITHMIC 28 B per cell-step with the HBM peak like every other line"),
                "launch_geometry": plan}
    elif tl_n.value > 0:
        achieved = ALG_BYTES * cells * seg * args.steps / (tl_ms.value * 1e-3) / 1e9
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "peak_kind": peak_kind, "kernel": "fk_tile_kernel",
                "kernel_share_of_step": tl_ms.value / ms_prof}
    cpu = None
    if not args.no_cpu and world == 1:
        n_e = max(1, int(1.5e9 // cells)) if workload in ("fk4096",) else max(1, int(1.5e9 // (work["u"].shape[-1] * work["u"].shape[-2])))
        v, cores, secs = cpu_run(work, n_e)
        cpu = {"value": v / 1e9, "unit": "Gcell-steps/s", "cores": cores, "kind": "port",
               "sample": "%s: one tissue, %d Euler steps, %.1f s" % (workload, n_e, secs)}
    out = {
        "metric": "Gcell-steps/s fp32 FK", "value": value, "unit": "Gcell-steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config, "clocks": clocks,
        "e2e": {"value": e2e, "unit": "Gcell-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "host_link_gbs": link, "numa_node": numa, "step_ms": [round(x, 3) for x in e2e_step_ms],
                "retimed": e2e_retimed},
        "step_ms": [round(x, 3) for x in step_ms], "retimed": retimed,
        "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu, "other_configs": other,
        "hbm_roofline_frac_whole_step": ALG_BYTES * value / world / peak,
    }
    if other:   # the multi-GPU evidence also at the top level of the line
        for k in ("slab_check", "efficiency_vs_slab_n1"):
            if k in other:
                out[k] = other[k]
        if "exposed_us_per_exchange" in other:
            out["exposed_us_per_exchange"] = other["exposed_us_per_exchange"]["value"]
        if "ens256" in other:
            out["ensemble"] = {k: other["ens256"][k] for k in ("value", "unit", "tissues", "frac_of_28B_roofline_per_gpu")}
        if "slab_n1" in other:
            out["slab_n1"] = other["slab_n1"]["value"]
    print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()
    sys.exit(exit_code)


if __name__ == "__main__":
    main()
