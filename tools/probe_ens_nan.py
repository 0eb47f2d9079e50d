"""Development probe: the ens256 bench workload, streaming kernel against the general tile kernel, segment by segment."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from cardiax_b200 import _lib, options, solve, stimulus, params as P

nsims = int(sys.argv[1]) if len(sys.argv) > 1 else 128
seg = int(sys.argv[2]) if len(sys.argv) > 2 else 20
nseg = int(sys.argv[3]) if len(sys.argv) > 3 else 6
work = bench.make_ens256(stimulus, nsims)
dev = torch.device("cuda:0")
gstim = [[stimulus.Stimulus(stimulus.Protocol(*p), torch.as_tensor(f).to(dev)) for p, f in ss] for ss in work["stimuli"]]
D = torch.as_tensor(work["D"]).to(dev)
st0 = solve.State(*[torch.as_tensor(work[k]).to(dev) for k in "vwu"])
prm = getattr(P, "PARAMSET_" + work["params"])
options.verbose = False
sa, sb, t = st0, st0, 0
for k in range(nseg):
    options.kernel = 0
    sa = solve._forward_euler(sa, t, t + seg, prm, D, gstim, 0.01, 0.01)
    torch.cuda.synchronize()
    ka, pa = _lib.last_kernel(), _lib.last_plan()
    options.kernel = 1
    sb = solve._forward_euler(sb, t, t + seg, prm, D, gstim, 0.01, 0.01)
    torch.cuda.synchronize()
    t += seg
    for name, a, b in zip("vwu", sa, sb):
        bad = ~torch.isfinite(a)
        d = (a - b).abs()
        d[bad] = 0
        msg = "seg %d %s %s: nonfinite %d (tile kernel: %d), max|diff| %.3g" % (k, ka, name, int(bad.sum()), int((~torch.isfinite(b)).sum()), float(d.max()))
        if bad.any():
            idx = bad.nonzero()
            msg += " first bad (sim,row,col) %s last %s; sims %s rows %s..%s cols %s..%s" % (
                idx[0].tolist(), idx[-1].tolist(), sorted(set(idx[:, 0].tolist()))[:8], int(idx[:, 1].min()), int(idx[:, 1].max()),
                int(idx[:, 2].min()), int(idx[:, 2].max()))
        print(msg)
    if k == 0:
        print(pa)
