"""Resident kernel: does a 1024-thread CTA (FK_RES_THREADS_CAP=1024 build) hide the latency the 512-thread one cannot?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from cardiax_b200 import _lib, options, params, solve, stimulus
options.verbose = False
print("library:", _lib.SO_PATH)
for name, mk in (("fk128", bench.make_fk128), ("fk512", bench.make_fk512)):
    wk = mk(stimulus)
    gs = [stimulus.Stimulus(stimulus.Protocol(*p), torch.as_tensor(f).cuda()) for p, f in wk["stimuli"]]
    D = torch.as_tensor(wk["D"]).cuda()
    P = getattr(params, "PARAMSET_" + wk["params"])
    for nc, thr in ((0, 0), (2, 0), (4, 0), (2, 1024), (4, 1024), (1, 1024), (2, 768), (4, 768)):
        options.kernel, options.cells_per_thread, options.cta_threads = 4, nc, thr
        try:
            box = [solve.State(*[torch.as_tensor(wk[k]).cuda() for k in "vwu"])]
            def seg(i):
                box[0] = solve._forward_euler(box[0], i * 500, (i + 1) * 500, P, D, gs, 0.01, 0.01)
            for i in range(2):
                seg(i)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for i in range(2, 8):
                seg(i)
            b.record(); torch.cuda.synchronize()
            p = _lib.last_plan()
            print("%-6s nc %d thr %4d: %.3f us per step  (%dx%d tiles of %dx%d, %d thr, nc %d)" % (
                name, nc, thr, a.elapsed_time(b) * 1e3 / 3000, p["tile_rows"], p["tile_cols"], p["tile_h"], p["tile_w"], p["cta_threads"], p["cells_per_thread"]), flush=True)
        except Exception as e:
            print("%-6s nc %d thr %4d: %s" % (name, nc, thr, str(e)[:80]))
