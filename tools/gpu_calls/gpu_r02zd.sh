#!/bin/bash
# row-major level-0 rings (coalesced LDGSTS destinations): parity + bench A/B against r02u (292.2)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r02zd_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02zd_pytest.log
tail -3 gpurun_out/r02zd_pytest.log
B="python bench.py --steps 6 --warmup 3 --no-extra --no-cpu"
for v in "--T 2" "--T 1" "--T 3"; do
  n=$(echo "x$v" | tr -d ' -')
  timeout 200 $B $v > gpurun_out/r02zd_fast_$n.json 2> gpurun_out/r02zd_fast_$n.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02zd_fast_*.json")):
    try:
        d = json.load(open(f))
        print(f.split("r02zd_fast_")[1], "value %.1f launch_ms %.4f" % (d["value"], d["roofline"]["avg_launch_ms"]), d["roofline"]["launch_geometry"])
    except Exception as e:
        print(f, "FAILED", e)
PY
