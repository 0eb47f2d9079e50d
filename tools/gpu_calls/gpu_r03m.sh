#!/bin/bash
# the library found stale by every rank of a torchrun job at once: one rank builds, the others wait for it
mkdir -p gpurun_out
echo "// touched" >> cardiax_b200/csrc/fk_aux.h
python -c "from cardiax_b200 import _lib; print('stale:', _lib.needs_build())"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 6 --warmup 2 --no-extra --no-cpu > gpurun_out/r03m_bench_n2.json 2> gpurun_out/r03m_bench_n2.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r03m_bench_n2.json").read().strip().splitlines()[-1])
print("value %.1f e2e %.1f" % (d["value"], d["e2e"]["value"]), d["step_ms"], d["retimed"])
PY
python -c "from cardiax_b200 import _lib; print('stale afterwards:', _lib.needs_build())"
