"""One call of the streaming kernel on a 4096^2 tissue with a heterogeneous diffusivity map (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import oracle as O
from cardiax_b200 import _lib, options, solve
options.verbose = False
H = 4096
yy, xx = torch.meshgrid(torch.arange(H, device="cuda", dtype=torch.float32), torch.arange(H, device="cuda", dtype=torch.float32), indexing="ij")
D = 1e-4 + 9e-4 * (0.5 + 0.5 * torch.sin(xx / 7.0) * torch.cos(yy / 9.0))
u = torch.zeros((H, H), device="cuda"); u[100:200, 100:300] = 1.0
s0 = solve.State(torch.ones((H, H), device="cuda"), torch.ones((H, H), device="cuda"), u)
out = solve._forward_euler(s0, 0, 40, O.PARAMSETS["5"], D, [], 0.01, 0.01)
torch.cuda.synchronize()
print(_lib.last_kernel(), _lib.last_plan())
