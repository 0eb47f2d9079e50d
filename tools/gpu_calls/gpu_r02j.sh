#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 3 --no-extra --no-cpu --numerics exact --seg 100"
for v in "" "--T 1" "--T 3" "--cta-threads 64" "--cta-threads 96" "--cta-threads 192" "--cta-threads 256"; do
  n=$(echo "x$v" | tr -d ' -')
  timeout 200 $B $v > gpurun_out/r02j_exact_$n.json 2> gpurun_out/r02j_exact_$n.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02j_exact_*.json")):
    try:
        d = json.load(open(f))
        print(f.split("r02j_exact_")[1], "value %.1f launch_ms %.4f" % (d["value"], d["roofline"]["avg_launch_ms"]), d["roofline"]["launch_geometry"])
    except Exception as e:
        print(f, "FAILED", e)
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fk_stream_kernel -s 4 -c 1 -o gpurun_out/prof_stream_exact_r02j python bench.py --steps 1 --warmup 2 --seg 8 --no-cpu --no-extra --numerics exact > gpurun_out/r02j_ncu.log 2>&1
tail -2 gpurun_out/r02j_ncu.log
