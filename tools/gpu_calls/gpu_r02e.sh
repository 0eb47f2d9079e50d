#!/bin/bash
# round 2, GPU call E (2 GPUs): slab parity on NVLink (peer + NCCL exchange), bench at N = 2 with both exchanges
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02e_smi.txt
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/r02e_multi_n2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02e_multi_n2.log
tail -5 gpurun_out/r02e_multi_n2.log
for comm in peer dist; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --comm $comm > gpurun_out/r02e_bench_n2_$comm.json 2> gpurun_out/r02e_bench_n2_$comm.err; echo "bench $comm rc=$?"
tail -3 gpurun_out/r02e_bench_n2_$comm.err
done
python - <<'PY'
import json
for c in ("peer", "dist"):
    try:
        d = json.loads(open("gpurun_out/r02e_bench_n2_%s.json" % c).read().strip().splitlines()[-1])
        print(c, "value %.1f e2e %.1f" % (d["value"], d["e2e"]["value"]), {k: d.get(k) for k in ("slab_check", "exposed_us_per_exchange", "efficiency_vs_slab_n1")})
        oc = d.get("other_configs") or {}
        for k in ("ens256", "slab_n1"):
            if k in oc: print("  ", k, oc[k].get("value"))
    except Exception as e:
        print(c, "FAILED", e)
PY
