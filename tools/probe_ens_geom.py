"""ens256 (128 x 256^2): streaming-kernel geometry sweep (rows per CTA chunk), and Heun on small tissues through the
resident kernel (options.kernel = 4) vs two wide launches."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import oracle as O
from cardiax_b200 import _lib, options, params, solve, stimulus
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


work = bench.make_ens256(128)
st = solve.State(*[torch.as_tensor(work[k]).cuda() for k in "vwu"])
D = torch.as_tensor(work["D"]).cuda()
stim = [[stimulus.Stimulus(stimulus.Protocol(*s.protocol), torch.as_tensor(s.field).cuda()) for s in ss] for ss in work["stimuli"]]
for T in (0, 1, 2):
    for rows in (0, 48, 64, 88, 96, 128, 136, 256):
        options.steps_per_launch, options.rows_per_cta, options.kernel = T, rows, (2 if rows or T else 0)
        try:
            s = timed(lambda: solve._forward_euler(st, 100, 600, O.PARAMSETS["3"], D, stim, 0.01, 0.01))
            print("ens256 T=%d rows_per_cta=%d: %s %s %.1f Gcs/s" % (T, rows, _lib.last_kernel(), _lib.last_plan(), 128 * 65536 * 500 / s / 1e9), flush=True)
        except Exception as e:  # noqa: BLE001
            print("ens256 T=%d rows=%d: %s" % (T, rows, str(e)[:80]))
options.steps_per_launch, options.rows_per_cta, options.kernel = 0, 0, 0
for H in (128, 256, 512):
    u = torch.zeros((H, H), device="cuda"); u[10:60, 20:90] = 1.0
    s0 = solve.State(torch.ones((H, H), device="cuda"), torch.ones((H, H), device="cuda"), u)
    Dm = torch.full((H, H), 1e-3, device="cuda")
    for k in (0, 4):
        options.kernel = k
        s = timed(lambda: solve._forward_heun(s0, 0, 200, params.PARAMSET_3, Dm, [], 0.01, 0.01))
        print("heun %d^2 kernel=%d: %s %.1f us per step" % (H, k, _lib.last_kernel(), s / 200 * 1e6), flush=True)
    options.kernel = 0
