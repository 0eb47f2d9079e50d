"""A few fast-numerics Heun steps on the 4096^2 headline tissue (for ncu: the HEUN instantiation of the streaming kernel)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import oracle as O
from cardiax_b200 import _lib, options, solve
options.verbose = False
H = 4096
u = torch.zeros((H, H), device="cuda"); u[100:200, 100:300] = 1.0
s0 = solve.State(torch.ones((H, H), device="cuda"), torch.ones((H, H), device="cuda"), u)
D = torch.full((H, H), 1e-3, device="cuda")
out = solve._forward_heun(s0, 0, 12, O.PARAMSETS["5"], D, [], 0.01, 0.01)
torch.cuda.synchronize()
print(_lib.last_kernel(), _lib.last_plan())
