"""Development probe: per-segment times of fk4096 (500 Euler steps per call), with and without the bench's nvidia-smi sampler."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from cardiax_b200 import _lib, options, solve, params as P

wk = bench.make_fk4096()
dev = torch.device("cuda:0")
st = solve.State(*[torch.as_tensor(wk[k]).to(dev) for k in "vwu"])
D = torch.as_tensor(wk["D"]).to(dev)
options.verbose = False
for _ in range(3):
    st = solve._forward_euler(st, 0, 500, P.PARAMSET_5, D, [], 0.01, 0.01)
torch.cuda.synchronize()
for label in ("no sampler", "sampler", "no sampler", "sampler"):
    sampler = None
    if label == "sampler":
        sampler = bench.ClockSampler(0)
        sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(13)]
    t0 = time.time()
    ev[0].record()
    for i in range(12):
        st = solve._forward_euler(st, 0, 500, P.PARAMSET_5, D, [], 0.01, 0.01)
        ev[i + 1].record()
    t_enq = time.time() - t0
    torch.cuda.synchronize()
    if sampler:
        sampler.stop()
    ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(12)]
    print("%-10s enqueue %.3f s; per-segment ms: %s" % (label, t_enq, " ".join("%.1f" % m for m in ms)))
