"""Small end-to-end cases for compute-sanitizer (development tool): streaming kernel with physical edges, hetero and
uniform D, stimuli on and off, a batch, the wide and tile kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oracle as O
from oracle import c_oracle as C
from cardiax_b200 import options, solve, stimulus
from tests import common
options.verbose = False
P3 = O.PARAMSETS["3"]
def run(shape, n, kernel, T, n_stim, uniform, numerics="exact", batch=1, **kw):
    st, D, stim = common.random_case(shape, seed=3, n_stim=n_stim)
    if uniform: D = np.full(shape, 1e-3, np.float32)
    options.numerics, options.kernel, options.steps_per_launch = numerics, kernel, T
    options.cta_threads, options.rows_per_cta = kw.get("nt", 0), kw.get("rh", 0)
    gs = [stimulus.Stimulus(stimulus.Protocol(*s.protocol), torch.as_tensor(s.field).cuda()) for s in stim]
    if batch > 1:
        sts = solve.State(*[torch.as_tensor(np.stack([x] * batch)).cuda() for x in st])
    else:
        sts = solve.State(*[torch.as_tensor(x).cuda() for x in st])
    out = solve._forward_euler(sts, 0, n, P3, torch.as_tensor(D).cuda(), gs, 0.01, 0.01)
    torch.cuda.synchronize()
    ref = C.forward_euler(st, 0, n, P3, D, stim, 0.01, 0.01)
    for a, b in zip(out, ref):
        a = a.cpu().numpy()
        a = a[0] if batch > 1 else a
        assert np.array_equal(a, b) if numerics == "exact" else np.abs(a - b).max() < 1e-4
    print("ok", shape, n, kernel, T, n_stim, uniform, numerics, batch, kw, flush=True)
run((72, 160), 4, 2, 2, 0, 0, nt=32, rh=16)
run((72, 160), 4, 2, 2, 0, 1, nt=32, rh=16, numerics="fast")
run((96, 288), 6, 2, 3, 3, 0, nt=32, rh=20)
run((40, 64), 3, 2, 1, 0, 0)
run((130, 516), 4, 2, 4, 0, 0, nt=64, rh=17)
run((64, 256), 4, 2, 2, 2, 0, batch=3, numerics="fast")
run((40, 64), 3, 3, 0, 2, 0)
run((37, 53), 3, 1, 2, 2, 0)
run((640, 1664), 4, 2, 2, 0, 1, numerics="fast")   # the planner's own geometry: >= 3 strips, balanced row chunks, block order
run((640, 1664), 4, 2, 2, 2, 0)
run((512, 256), 4, 2, 2, 0, 0, batch=4, numerics="fast")   # one strip per tissue (the ensemble shape)
print("all ok")
