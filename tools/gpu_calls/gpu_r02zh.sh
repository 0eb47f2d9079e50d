#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 6 --warmup 3 --no-extra --no-cpu"
for d in 0 8 12 16 20; do
  FK_FIRST_DISCOUNT=$d timeout 200 $B > gpurun_out/r02zh_d$d.json 2> gpurun_out/r02zh_d$d.err
  FK_FIRST_DISCOUNT=$d timeout 200 $B --workload ens256 > gpurun_out/r02zh_ens_d$d.json 2> gpurun_out/r02zh_ens_d$d.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02zh_*.json")):
    try:
        d = json.load(open(f))
        print(f.split("r02zh_")[1], "value %.1f launch_ms %.4f" % (d["value"], d["roofline"]["avg_launch_ms"]), d["roofline"]["launch_geometry"])
    except Exception as e:
        print(f, "FAILED", e)
PY
