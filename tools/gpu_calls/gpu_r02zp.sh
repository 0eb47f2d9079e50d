#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 6 --warmup 3 --no-extra --no-cpu"
for v in "--cta-threads 96" "--cta-threads 160" "--cta-threads 192" "--cta-threads 64" "--T 1" "--T 3"; do
  n=$(echo "x$v" | tr -d ' -')
  timeout 200 $B $v > gpurun_out/r02zp_$n.json 2> gpurun_out/r02zp_$n.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02zp_*.json")):
    try:
        d = json.load(open(f))
        print(f.split("r02zp_")[1], "value %.1f launch_ms %.4f" % (d["value"], d["roofline"]["avg_launch_ms"]), d["roofline"]["launch_geometry"])
    except Exception as e:
        print(f, "FAILED", e)
PY
