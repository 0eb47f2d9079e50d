#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r02zj_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02zj_pytest.log
tail -2 gpurun_out/r02zj_pytest.log
B="python bench.py --steps 6 --warmup 3 --no-extra --no-cpu"
for v in "--T 2" "--workload ens256" "--T 3"; do
  n=$(echo "x$v" | tr -d ' -')
  timeout 200 $B $v > gpurun_out/r02zj_fast_$n.json 2> gpurun_out/r02zj_fast_$n.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02zj_fast_*.json")):
    try:
        d = json.load(open(f))
        print(f.split("r02zj_fast_")[1], "value %.1f launch_ms %.4f" % (d["value"], d["roofline"]["avg_launch_ms"]), d["roofline"]["launch_geometry"])
    except Exception as e:
        print(f, "FAILED", e)
PY
