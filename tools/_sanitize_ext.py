"""Small cases of the kernels around the Euler loop for compute-sanitizer (development tool): snapshot resize (odd shapes,
up- and down-scaling, > 8 planes), Dormand-Prince (stage / finish / reduce / interp kernels), electrogram, Heun (fused tile
kernel; fast path with the closing pass folded into the wide and the streaming kernel)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oracle as O
from oracle import fk_oracle_ext as X
from cardiax_b200 import _lib, io, metrics, options, solve, stimulus
from tests import common
options.verbose = False
P3 = O.PARAMSETS["3"]
rng = np.random.default_rng(0)
for shape, size in (((3, 24, 36), (8, 9)), ((20, 20), (7, 13)), ((12, 12), (24, 30)), ((11, 75, 61), (16, 13)), ((5, 5), (1, 1)), ((1, 7), (3, 2))):
    a = rng.random(shape).astype(np.float32)
    got = io.imresize(torch.as_tensor(a).cuda(), size).cpu().numpy()
    assert np.abs(got - X.resize_bilinear(a, size)).max() < 1e-5
    print("ok resize", shape, size, flush=True)
x = rng.random((3, 33, 33)).astype(np.float32)
assert np.allclose(metrics.electrogram(torch.as_tensor(x).cuda(), (3, 7)).cpu().numpy(), X.electrogram(x, (3, 7)), rtol=2e-7)
print("ok electrogram", flush=True)
shape = (24, 28)
st, D = common.smooth_case(shape, 1)
stim = [O.linear(shape, 0, 0.3, 20.0, O.Protocol(0, 2, 50))]
gst = [stimulus.Stimulus(stimulus.Protocol(*s.protocol), torch.as_tensor(s.field).cuda()) for s in stim]
state = solve.State(*[torch.as_tensor(v).cuda() for v in st])
options.numerics, options.ode_rtol, options.ode_atol = "exact", 1e-5, 1e-5
ts = np.array([0, 0.5, 1.0], np.float32)
out = solve._forward_dormandprince(state, ts, P3, torch.as_tensor(D).cuda(), gst, 0.01, 0.01)
ref = X.odeint_dopri5(st, ts, P3, D, stim, 0.01, rtol=1e-5, atol=1e-5)
assert all(np.array_equal(a.cpu().numpy(), b) for a, b in zip(out, ref))
print("ok dopri5", solve.last_ode_stats, flush=True)
for shp, numerics, kernel in (((40, 72), "exact", "fk_tile_kernel"), ((64, 96), "fast", "fk_wide_kernel"), ((1024, 1056), "fast", "fk_stream_kernel")):
    st, D = common.smooth_case(shp, 4)
    stim = [O.linear(shp, 0, 0.3, 20.0, O.Protocol(2, 2, 50))]
    gst = [stimulus.Stimulus(stimulus.Protocol(*s.protocol), torch.as_tensor(s.field).cuda()) for s in stim]
    state = solve.State(*[torch.as_tensor(v).cuda() for v in st])
    options.numerics = numerics
    n = 5 if shp[0] < 1000 else 4
    out = solve._forward_heun(state, 0, n, P3, torch.as_tensor(D).cuda(), gst, 0.01, 0.01)
    torch.cuda.synchronize()
    assert _lib.last_kernel() == kernel, _lib.last_kernel()
    if shp[0] < 1000:
        ref = O.forward_heun(st, 0, n, P3, D, stim, 0.01, 0.01)
        for a, b in zip(out, ref):
            d = np.abs(a.cpu().numpy() - b)
            # fast numerics: 2e-5 everywhere except a cell that sits on a gate threshold (u == V_c to 1 ulp flips the
            # gate one step earlier or later: one step of dv = v dt / tau_v_plus = 3e-3)
            assert np.array_equal(a.cpu().numpy(), b) if numerics == "exact" else \
                ((d > 2e-5 * max(1.0, float(np.abs(b).max()))).sum() <= 2 and d.max() < 5e-3)
    else:
        assert bool(torch.isfinite(out.u).all())
    print("ok heun", shp, numerics, kernel, flush=True)
print("all ok")
