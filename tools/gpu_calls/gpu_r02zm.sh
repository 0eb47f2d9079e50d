#!/bin/bash
# 2 GPUs: slab parity on NVLink after the streaming kernel's layout / block order / chunk changes, bench at N = 2
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/r02zm_multi_n2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02zm_multi_n2.log
tail -5 gpurun_out/r02zm_multi_n2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02zm_bench_n2.json 2> gpurun_out/r02zm_bench_n2.err; echo "bench rc=$?"
tail -3 gpurun_out/r02zm_bench_n2.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02zm_bench_n2.json").read().strip().splitlines()[-1])
print("value %.1f e2e %.1f" % (d["value"], d["e2e"]["value"]), {k: d.get(k) for k in ("slab_check", "exposed_us_per_exchange", "efficiency_vs_slab_n1", "ensemble")})
oc = d.get("other_configs") or {}
for k in ("ens256", "slab_n1"):
    if k in oc: print("  ", k, oc[k].get("value"))
PY
