"""Pinned H2D / D2H bandwidth as bench.py's e2e leg sees it, with and without binding the process to the GPU's NUMA node."""
import os, sys, time
import torch


def bw(n_mb=192, reps=5):
    h = torch.empty(n_mb * 1024 * 1024 // 4, dtype=torch.float32).pin_memory()
    d = torch.empty_like(h, device="cuda")
    out = {}
    for name, fn in (("h2d", lambda: d.copy_(h, non_blocking=True)), ("d2h", lambda: h.copy_(d, non_blocking=True))):
        fn(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record(); torch.cuda.synchronize()
        out[name] = n_mb * reps / 1024 / (a.elapsed_time(b) * 1e-3)
    # both directions at once on two streams
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    h2 = torch.empty_like(h).pin_memory(); d2 = torch.empty_like(d)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        with torch.cuda.stream(s1):
            d.copy_(h, non_blocking=True)
        with torch.cuda.stream(s2):
            h2.copy_(d2, non_blocking=True)
    torch.cuda.synchronize()
    out["both"] = 2 * n_mb * reps / 1024 / (time.perf_counter() - t0)
    return out


print("cpus allowed:", len(os.sched_getaffinity(0)), "of", os.cpu_count())
print("default:", {k: "%.1f GB/s" % v for k, v in bw().items()})
try:
    import pynvml
    pynvml.nvmlInit()
    hnd = pynvml.nvmlDeviceGetHandleByIndex(0)
    words = pynvml.nvmlDeviceGetCpuAffinity(hnd, (os.cpu_count() + 63) // 64)
    cpus = [i for i in range(os.cpu_count()) if words[i // 64] >> (i % 64) & 1]
    print("gpu-local cpus:", len(cpus), cpus[:4], "...", cpus[-4:])
    print("pcie gen/width:", pynvml.nvmlDeviceGetCurrPcieLinkGeneration(hnd), pynvml.nvmlDeviceGetCurrPcieLinkWidth(hnd),
          "max", pynvml.nvmlDeviceGetMaxPcieLinkGeneration(hnd), pynvml.nvmlDeviceGetMaxPcieLinkWidth(hnd))
    ok = sorted(set(cpus) & os.sched_getaffinity(0))
    if ok:
        os.sched_setaffinity(0, ok)
        print("bound to gpu-local cpus:", {k: "%.1f GB/s" % v for k, v in bw().items()})
except Exception as e:  # noqa: BLE001
    print("nvml:", e)
os.system("nvidia-smi topo -m 2>/dev/null | head -12; numactl -H 2>/dev/null | head -6; cat /sys/devices/system/node/online 2>/dev/null")
