#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r03j_bench.json 2> gpurun_out/r03j_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r03j_bench.json"))
print("value %.1f e2e %.1f frac %.3f" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"]), d["step_ms"], d["retimed"], d["clocks"])
print({k: (v.get("value") if isinstance(v, dict) else v) for k, v in d["other_configs"].items() if k != "readme_table_seconds_1e3_steps"})
print(d["cpu_baseline"])
PY
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 | cut -c1-300
