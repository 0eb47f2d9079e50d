"""Development probe: per-CTA timeline of the streaming kernel on fk4096 (library built with -DFK_STREAM_TIMING).

    FK_SO=.../libfk_timing.so python tools/probe_stream_timing.py [T]
"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from cardiax_b200 import _lib, options, solve, params as P

T = int(sys.argv[1]) if len(sys.argv) > 1 else 2
ens = len(sys.argv) > 2 and sys.argv[2] == "ens256"
dev = torch.device("cuda:0")
options.verbose = False
options.steps_per_launch = T
if ens:   # (only tissue 0's CTAs are recorded: blockIdx.y == 0)
    from cardiax_b200 import stimulus
    wk = bench.make_ens256(stimulus, 128)
    st = solve.State(*[torch.as_tensor(wk[k]).to(dev) for k in "vwu"])
    D = torch.as_tensor(wk["D"]).to(dev)
    for _ in range(3):
        st = solve._forward_euler(st, 100, 140, P.PARAMSET_3, D, [], 0.01, 0.01)
else:
    wk = bench.make_fk4096()
    st = solve.State(*[torch.as_tensor(wk[k]).to(dev) for k in "vwu"])
    D = torch.as_tensor(wk["D"]).to(dev)
    for _ in range(3):
        st = solve._forward_euler(st, 0, 40, P.PARAMSET_5, D, [], 0.01, 0.01)
torch.cuda.synchronize()
plan = _lib.last_plan()
print(_lib.last_kernel(), plan)
nb = 128 if ens else 1
n = plan["strips"] * plan["row_chunks"] * nb
buf = (ctypes.c_ulonglong * (3 * n))()
fn = getattr(_lib.lib(), "fk_stream_timing_T%d_E0" % T)
fn.argtypes = [ctypes.c_void_p, ctypes.c_int]
assert fn(buf, n) == 0
a = np.array(buf, dtype=np.int64).reshape(n, 3)
t0 = a[:, 0].min()
start, end, sm = (a[:, 0] - t0) / 1e3, (a[:, 1] - t0) / 1e3, a[:, 2]
dur = end - start
ns = plan["strips"]
strip, chunk = np.arange(n) % ns, (np.arange(n) // ns) % plan["row_chunks"]
print("launch span %.1f us; CTA duration min/median/max %.1f / %.1f / %.1f us; start max %.1f us" % (end.max(), dur.min(), np.median(dur), dur.max(), start.max()))
print("by strip: " + "  ".join("%d: %.1f" % (s, dur[strip == s].mean()) for s in range(ns)))
for c in (0, 1, plan["row_chunks"] // 2, plan["row_chunks"] - 2, plan["row_chunks"] - 1):
    print("chunk %d: mean duration %.1f us, end %.1f" % (c, dur[chunk == c].mean(), end[chunk == c].mean()))
sm_end = {}
for s_, e_ in zip(sm, end):
    sm_end[s_] = max(sm_end.get(s_, 0.0), e_)
ends = np.array(sorted(sm_end.values()))
print("SMs used %d; SM finish time min/median/mean/max %.1f / %.1f / %.1f / %.1f us" % (len(ends), ends.min(), np.median(ends), ends.mean(), ends.max()))
cnt = np.bincount(sm.astype(int), minlength=148)
print("CTAs per SM: " + str(np.bincount(cnt)))
edge = (strip == 0) | (strip == ns - 1)
per_sm_edges = np.zeros(int(sm.max()) + 1)
for s_, e_ in zip(sm, edge):
    per_sm_edges[int(s_)] += e_
for k in range(4):
    sel = [sm_end[s_] for s_ in sm_end if per_sm_edges[int(s_)] == k]
    if sel:
        print("SMs with %d edge CTAs: %d, mean finish %.1f us" % (k, len(sel), np.mean(sel)))
late = np.argsort(-end)[:12]
print("last CTAs (strip, chunk, sm, start, dur): " + "; ".join("(%d,%d,%d,%.1f,%.1f)" % (strip[i], chunk[i], sm[i], start[i], dur[i]) for i in late))
