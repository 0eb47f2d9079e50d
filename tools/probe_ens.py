"""Ensemble of 256^2 tissues (BASELINE config 4): the streaming kernel on the whole batch vs the resident kernel on
sub-batches that fit the SMs' shared memory; heterogeneous-D streaming at 4096^2 for reference."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from cardiax_b200 import _lib, options, params, solve, stimulus
import oracle as O

options.verbose = False


def timed(fn, n=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e-3


nsims, steps = 128, 500
work = bench.make_ens256(nsims)
P = O.PARAMSETS["3"]
st = solve.State(*[torch.as_tensor(work[k]).cuda() for k in "vwu"])
D = torch.as_tensor(work["D"]).cuda()
stim = [[stimulus.Stimulus(stimulus.Protocol(*s.protocol), torch.as_tensor(s.field).cuda()) for s in ss] for ss in work["stimuli"]]
s = timed(lambda: solve._forward_euler(st, 0, steps, P, D, stim, 0.01, 0.01))
print("whole batch of %d: %s  %.1f Gcs/s" % (nsims, _lib.last_kernel(), nsims * 65536 * steps / s / 1e9))
ref = solve._forward_euler(st, 0, steps, P, D, stim, 0.01, 0.01)
for chunk in (8, 12, 16, 20, 24, 28, 32):
    def run():
        outs = []
        for c0 in range(0, nsims, chunk):
            sl = slice(c0, min(nsims, c0 + chunk))
            outs.append(solve._forward_euler(solve.State(st.v[sl], st.w[sl], st.u[sl]), 0, steps, P, D[sl], stim[sl], 0.01, 0.01))
        return outs
    options.kernel = 4
    try:
        s = timed(run)
        outs = run()
        same = all(torch.equal(torch.cat([o.u for o in outs]), ref.u) for _ in (0,))
        print("chunks of %d: %s %s  %.1f Gcs/s  identical=%s" % (chunk, _lib.last_kernel(), _lib.last_plan(), nsims * 65536 * steps / s / 1e9, same))
    except Exception as e:  # noqa: BLE001
        print("chunks of %d: %s" % (chunk, str(e)[:100]))
    options.kernel = 0

H = 4096
u = torch.zeros((H, H), device="cuda"); u[100:200, 100:300] = 1.0
s0 = solve.State(torch.ones((H, H), device="cuda"), torch.ones((H, H), device="cuda"), u)
for name, Dm in (("uniform", torch.full((H, H), 1e-3, device="cuda")), ("hetero", torch.rand((H, H), device="cuda") * 9e-4 + 1e-4)):
    for T in (0, 1, 2, 3):
        options.steps_per_launch = T
        try:
            s = timed(lambda: solve._forward_euler(s0, 0, 60, O.PARAMSETS["5"], Dm, [], 0.01, 0.01))
            print("4096^2 %s D, T=%d: %s %s %.1f Gcs/s" % (name, T, _lib.last_kernel(), _lib.last_plan(), H * H * 60 / s / 1e9))
        except Exception as e:  # noqa: BLE001
            print(name, T, str(e)[:100])
    options.steps_per_launch = 0
