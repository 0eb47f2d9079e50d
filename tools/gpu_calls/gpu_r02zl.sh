#!/bin/bash
# balanced row chunks: sweep of the rows taken off the top / bottom chunk
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r02zl_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02zl_pytest.log
tail -2 gpurun_out/r02zl_pytest.log
B="python bench.py --steps 6 --warmup 3 --no-extra --no-cpu"
for tb in "0 0" "16 28" "12 20" "20 36" "16 40" "8 28"; do
  set -- $tb
  FK_TOP_OFF=$1 FK_BOT_OFF=$2 timeout 200 $B > gpurun_out/r02zl_t$1_b$2.json 2> gpurun_out/r02zl_t$1_b$2.err
  FK_TOP_OFF=$1 FK_BOT_OFF=$2 timeout 200 $B --workload ens256 > gpurun_out/r02zl_ens_t$1_b$2.json 2> gpurun_out/r02zl_ens_t$1_b$2.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02zl_*.json")):
    try:
        d = json.load(open(f))
        print(f.split("r02zl_")[1], "value %.1f launch_ms %.4f" % (d["value"], d["roofline"]["avg_launch_ms"]), d["roofline"]["launch_geometry"]["rows_per_cta"], d["roofline"]["launch_geometry"]["row_chunks"])
    except Exception as e:
        print(f, "FAILED", e)
PY
