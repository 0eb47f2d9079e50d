#!/bin/bash
# round 2, GPU call C: full parity suite (typed schedule, config-2 fixture, all-paramset fast), A/B of kernel variants
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -s > gpurun_out/r02c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02c_pytest.log
grep -E "config 2 step|passed|failed|rc=" gpurun_out/r02c_pytest.log | tail -20
B="python bench.py --steps 8 --warmup 3 --no-extra --no-cpu"
run() { # name, env..., args
  name=$1; shift
  env "$@" > /dev/null 2>&1
}
for v in base u4; do
  so=""; [ $v != base ] && so="FK_SO=$PWD/cardiax_b200/csrc/build/alt_$v/libfk_$v.so"
  env $so timeout 200 $B > gpurun_out/r02c_fk4096_$v.json 2> gpurun_out/r02c_fk4096_$v.err
done
for nt in 64 96 160 192; do
  timeout 200 $B --cta-threads $nt > gpurun_out/r02c_fk4096_nt$nt.json 2> gpurun_out/r02c_fk4096_nt$nt.err
done
for v in base late2 late0; do
  so=""; [ $v != base ] && so="FK_SO=$PWD/cardiax_b200/csrc/build/alt_$v/libfk_$v.so"
  env $so timeout 300 python tools/probe_hetero.py > gpurun_out/r02c_hetero_$v.log 2>&1
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02c_fk4096_*.json")):
    try:
        d = json.load(open(f))
        print(f.split("r02c_")[1], "value %.1f launch_ms %.4f" % (d["value"], d["roofline"]["avg_launch_ms"]), d["roofline"]["launch_geometry"])
    except Exception as e:
        print(f, "FAILED", e)
PY
tail -n 4 gpurun_out/r02c_hetero_*.log
