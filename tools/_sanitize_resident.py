"""Small end-to-end cases of the resident kernel for compute-sanitizer (development tool): 1/2/4 cells per thread, uneven
and edge-shrunk tiles, several rounds per phase, stimuli, a batch, single-phase and two-phase plans."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oracle as O
from oracle import c_oracle as C
from cardiax_b200 import _lib, options, solve, stimulus
from tests import common
options.verbose = False
P3 = O.PARAMSETS["3"]
def run(shape, n, tiles, nc, threads=0, edge=(0, 0), n_stim=2, numerics="exact", batch=1):
    st, D, stim = common.random_case(shape, seed=3, n_stim=n_stim)
    options.numerics, options.kernel, options.steps_per_launch = numerics, 4, 0
    options.cta_threads, options.tiles, options.cells_per_thread, options.edge_tile = threads, tiles, nc, edge
    gs = [stimulus.Stimulus(stimulus.Protocol(*s.protocol), torch.as_tensor(s.field).cuda()) for s in stim]
    if batch > 1:
        sts = solve.State(*[torch.as_tensor(np.stack([x] * batch)).cuda() for x in st])
    else:
        sts = solve.State(*[torch.as_tensor(x).cuda() for x in st])
    out = solve._forward_euler(sts, 0, n, P3, torch.as_tensor(D).cuda(), gs, 0.01, 0.01)
    torch.cuda.synchronize()
    assert _lib.last_kernel() == "fk_resident_kernel"
    ref = C.forward_euler(st, 0, n, P3, D, stim, 0.01, 0.01)
    for a, b in zip(out, ref):
        a = a.cpu().numpy()
        a = a[0] if batch > 1 else a
        assert np.array_equal(a, b) if numerics == "exact" else np.abs(a - b).max() < 1e-4
    print("ok", shape, n, tiles, nc, threads, edge, n_stim, numerics, batch, _lib.last_plan(), flush=True)
run((64, 96), 6, (0, 0), 0)
run((64, 96), 6, (2, 3), 4)
run((64, 96), 5, (4, 1), 2, threads=64)
run((96, 160), 5, (5, 6), 4, edge=(12, 4))
run((96, 160), 5, (5, 6), 1, threads=96)
run((128, 128), 6, (0, 0), 0, numerics="fast")
run((64, 64), 5, (2, 2), 2, batch=3, numerics="fast")
run((256, 256), 4, (0, 0), 0, n_stim=3)
print("all ok")
