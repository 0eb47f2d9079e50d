#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02v_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02v_smoke.log
B="python bench.py --steps 3 --warmup 3 --no-extra --no-cpu --numerics exact --seg 100"
for v in "--T 1 --cta-threads 32" "--T 1 --cta-threads 96" "--T 1 --cta-threads 128" "--T 1 --cta-threads 192"; do
  n=$(echo "x$v" | tr -d ' -')
  timeout 200 $B $v > gpurun_out/r02v_exact_$n.json 2> gpurun_out/r02v_exact_$n.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02v_exact_*.json")):
    try:
        d = json.load(open(f))
        print(f.split("r02v_exact_")[1], "value %.1f launch_ms %.4f" % (d["value"], d["roofline"]["avg_launch_ms"]), d["roofline"]["launch_geometry"])
    except Exception as e:
        print(f, "FAILED", e)
PY
