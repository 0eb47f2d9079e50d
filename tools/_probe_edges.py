"""Time one streaming launch (T = 2) of a 4096^2 tissue with and without physical top/bottom edges (development tool)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oracle as O
from cardiax_b200 import _lib, options, solve
H = W = 4096
L = _lib.lib()
dev = torch.device("cuda")
u = torch.zeros((H, W), device=dev); u[H // 4:H // 4 + 64, :] = 1.0
v = torch.ones((H, W), device=dev); w = torch.ones((H, W), device=dev)
D = torch.full((H, W), 1e-3, device=dev)
P = solve._params_struct(O.PARAMSETS["5"])
outs = [torch.empty_like(u) for _ in range(3)]
nbytes = L.fk_workspace_bytes(H, W, 1, 0, 0)
ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
arr = (_lib.FkStimulus * 1)()
for phys in (1, 0, 1, 0):
    o = solve._options(D, P, 0.01, phys_top=phys, phys_bottom=phys, steps_per_launch=2, kernel=2)
    def call():
        _lib.check(L.fk_forward_euler(v.data_ptr(), w.data_ptr(), u.data_ptr(), outs[0].data_ptr(), outs[1].data_ptr(), outs[2].data_ptr(),
                                      D.data_ptr(), 0, H, W, 1, ctypes.byref(P), arr, 0, 0.0, 2.0, np.float32(0.01), np.float32(0.01),
                                      ctypes.byref(o), ws.data_ptr(), nbytes, solve._stream()))
    for _ in range(5): call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): call()
    e1.record(); torch.cuda.synchronize()
    print("phys edges", phys, "%.1f us per call (dgrad + one T=2 launch)" % (e0.elapsed_time(e1) * 1e3 / 50), _lib.last_plan())
