#!/bin/bash
# 8 GPUs: slab parity on NVLink + bench at N = 8 (16384^2 slab, 1024-tissue ensemble beside it)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/r03k_multi_n8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r03k_multi_n8.log
tail -3 gpurun_out/r03k_multi_n8.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r03k_bench_n8.json 2> gpurun_out/r03k_bench_n8.err; echo "bench rc=$?"
tail -2 gpurun_out/r03k_bench_n8.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r03k_bench_n8.json").read().strip().splitlines()[-1])
print("value %.1f e2e %.1f" % (d["value"], d["e2e"]["value"]), {k: d.get(k) for k in ("slab_check", "exposed_us_per_exchange", "efficiency_vs_slab_n1", "ensemble")}, d["step_ms"], d["clocks"])
oc = d.get("other_configs") or {}
for k in ("ens256", "slab_n1"):
    if k in oc: print("  ", k, oc[k].get("value"))
PY
