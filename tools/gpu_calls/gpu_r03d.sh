#!/bin/bash
# 4 GPUs: bench at N = 4 (8192 x 16384 slab, 512-tissue ensemble beside it)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r03d_bench_n4.json 2> gpurun_out/r03d_bench_n4.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r03d_bench_n4.json").read().strip().splitlines()[-1])
print("value %.1f e2e %.1f" % (d["value"], d["e2e"]["value"]), {k: d.get(k) for k in ("slab_check", "exposed_us_per_exchange", "efficiency_vs_slab_n1", "ensemble", "retimed")}, d["step_ms"])
PY
