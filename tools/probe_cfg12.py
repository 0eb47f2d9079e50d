"""BASELINE configs 1 and 2 (bench.make_fk128 / make_fk512) on the default path: us per Euler step (development probe)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from cardiax_b200 import _lib, options, params, solve, stimulus
options.verbose = False
print("library:", _lib.SO_PATH)
for name, mk in (("fk128", bench.make_fk128), ("fk512", bench.make_fk512), ("fk1024", None)):
    if mk is None:
        wk = bench.make_fk4096(None, 1024, 1024); wk["params"] = "3"
    else:
        wk = mk(stimulus)
    gs = [stimulus.Stimulus(stimulus.Protocol(*p), torch.as_tensor(f).cuda()) for p, f in wk["stimuli"]]
    D = torch.as_tensor(wk["D"]).cuda()
    P = getattr(params, "PARAMSET_" + wk["params"])
    box = [solve.State(*[torch.as_tensor(wk[k]).cuda() for k in "vwu"])]
    def seg(i):
        box[0] = solve._forward_euler(box[0], i * 500, (i + 1) * 500, P, D, gs, 0.01, 0.01)
    for i in range(3):
        seg(i)
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(3, 11):
            seg(i)
        b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) * 1e3 / (8 * 500))
    print("%-7s %s %.3f us per Euler step" % (name, _lib.last_kernel(), best), _lib.last_plan())
