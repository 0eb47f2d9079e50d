#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r02zi_bench.json 2> gpurun_out/r02zi_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02zi_bench.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], "launch ms", d["roofline"]["avg_launch_ms"], "share", d["roofline"]["kernel_share_of_step"])
for k, v in d.get("other_configs", {}).items():
    print(k, json.dumps(v)[:300])
print(d.get("cpu_baseline"), d.get("clocks"))
PY
