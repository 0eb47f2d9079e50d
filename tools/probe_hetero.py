"""Streaming kernel with a heterogeneous diffusivity map: 4096^2, 2400^2 and the 128 x 256^2 ensemble."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from cardiax_b200 import _lib, options, params, solve, stimulus
options.verbose = False


def timed(fn, n=5, warm=2):
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


print("library:", _lib.SO_PATH)
for H, steps in ((4096, 60), (2400, 100)):
    yy, xx = torch.meshgrid(torch.arange(H, device="cuda", dtype=torch.float32), torch.arange(H, device="cuda", dtype=torch.float32), indexing="ij")
    D = 1e-4 + 9e-4 * (0.5 + 0.5 * torch.sin(xx / 7.0) * torch.cos(yy / 9.0))
    u = torch.zeros((H, H), device="cuda"); u[100:200, 100:300] = 1.0
    s0 = solve.State(torch.ones((H, H), device="cuda"), torch.ones((H, H), device="cuda"), u)
    s = timed(lambda: solve._forward_euler(s0, 0, steps, params.PARAMSET_5, D, [], 0.01, 0.01))
    print("%d^2 scar-map D: %s %.1f Gcell-steps/s" % (H, _lib.last_kernel(), H * H * steps / s / 1e9))
work = bench.make_ens256(stimulus, 128)
st = solve.State(*[torch.as_tensor(work[k]).cuda() for k in "vwu"])
D = torch.as_tensor(work["D"]).cuda()
stim = [[stimulus.Stimulus(stimulus.Protocol(*p), torch.as_tensor(f).cuda()) for p, f in ss] for ss in work["stimuli"]]
s = timed(lambda: solve._forward_euler(st, 100, 600, params.PARAMSET_3, D, stim, 0.01, 0.01), n=3, warm=1)
print("ens256 (128 x 256^2): %s %.1f Gcell-steps/s" % (_lib.last_kernel(), 128 * 65536 * 500 / s / 1e9))
