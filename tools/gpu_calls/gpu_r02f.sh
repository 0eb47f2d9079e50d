#!/bin/bash
# round 2, GPU call F (N GPUs): slab parity (world = min(N, 4)), bench at N with the fused peer exchange
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02f_smi_n$N.txt
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/r02f_multi_n$N.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02f_multi_n$N.log
tail -3 gpurun_out/r02f_multi_n$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02f_bench_n$N.json 2> gpurun_out/r02f_bench_n$N.err; echo "bench rc=$?"
tail -3 gpurun_out/r02f_bench_n$N.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02f_bench_n$N.json").read().strip().splitlines()[-1])
    print("value %.1f e2e %.1f" % (d["value"], d["e2e"]["value"]), {k: d.get(k) for k in ("slab_check", "exposed_us_per_exchange", "efficiency_vs_slab_n1", "ensemble", "slab_n1")})
except Exception as e:
    print("FAILED", e)
PY
