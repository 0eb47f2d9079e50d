import sys, os, time
sys.path.insert(0, os.getcwd())
import torch
import oracle as O
from cardiax_b200 import _lib, options, solve
options.verbose=False
H=W=4096
P=O.PARAMSETS["5"]
D=torch.full((H,W),1e-3,device="cuda")
def ic(kind):
    u=torch.zeros((H,W),device="cuda")
    if kind=="bands": u[H//4:H//4+64,:]=1.0; u[:,W//3:W//3+32]=1.0
    if kind=="hband": u[H//4:H//4+64,:]=1.0
    if kind=="vband": u[:,W//3:W//3+32]=1.0
    if kind=="hband_in": u[H//4:H//4+64,64:-64]=1.0
    if kind=="vband_in": u[64:-64,W//3:W//3+32]=1.0
    if kind=="block": u[1000:1048,1000:1048]=1.0
    return solve.State(torch.ones((H,W),device="cuda"),torch.ones((H,W),device="cuda"),u)
def run(kind,n,numerics="fast",kernel=0):
    options.numerics=numerics; options.kernel=kernel
    st=ic(kind)
    solve._forward_euler(st,0,n,P,D,[],0.01,0.01); torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    t0=time.perf_counter(); e0.record()
    s=solve._forward_euler(st,0,n,P,D,[],0.01,0.01)
    e1.record(); t1=time.perf_counter(); torch.cuda.synchronize()
    print(kind,n,numerics,kernel,"gpu ms/step %.4f"%(e0.elapsed_time(e1)/n),"host ms %.2f"%((t1-t0)*1e3), _lib.last_plan()["cta_threads"], flush=True)
for kind in ("zero","block","hband_in","vband_in","hband","vband","bands"):
    run(kind,48)
run("bands",480)
run("bands",48,"fast",1)
run("block",48,"exact"); run("bands",48,"exact"); run("zero",48,"exact")
