#!/bin/bash
# line-major level-0 rings with carried slot offsets: parity + bench (fk4096 T = 1/2/3, ens256) + ncu --set full of the headline kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r02ze_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02ze_pytest.log
tail -3 gpurun_out/r02ze_pytest.log
B="python bench.py --steps 6 --warmup 3 --no-extra --no-cpu"
for v in "--T 2" "--T 1" "--T 3" "--workload ens256"; do
  n=$(echo "x$v" | tr -d ' -')
  timeout 200 $B $v > gpurun_out/r02ze_fast_$n.json 2> gpurun_out/r02ze_fast_$n.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02ze_fast_*.json")):
    try:
        d = json.load(open(f))
        print(f.split("r02ze_fast_")[1], "value %.1f launch_ms %.4f" % (d["value"], d["roofline"]["avg_launch_ms"]), d["roofline"]["launch_geometry"])
    except Exception as e:
        print(f, "FAILED", e)
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fk_stream_kernel -s 4 -c 1 -o gpurun_out/prof_stream_r02ze -f python bench.py --steps 1 --warmup 3 --seg 8 --no-cpu --no-extra > gpurun_out/r02ze_ncu.log 2>&1
ls -la gpurun_out/prof_stream_r02ze.ncu-rep
