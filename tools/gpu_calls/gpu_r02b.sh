#!/bin/bash
# round 2, GPU call B: parity suite (incl. the two-process peer-exchange tests on one GPU), full default bench, ncu capture
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r02b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_pytest.log
tail -15 gpurun_out/r02b_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r02b_bench.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r02b_bench.json"))
    print("value %.1f e2e %.1f frac %.3f launch_ms %.4f share %.3f" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["avg_launch_ms"], d["roofline"]["kernel_share_of_step"]), d["clocks"])
    for k, v in (d["other_configs"] or {}).items():
        print(k, {a: b for a, b in v.items() if a in ("value", "us_per_euler_step", "kernel", "seconds", "s2_fired_at_step_40000", "ms_per_500_step_segment", "ours_b200")} if isinstance(v, dict) else v)
except Exception as e:
    print("bench FAILED", e)
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fk_stream_kernel -s 4 -c 1 -o gpurun_out/prof_stream_r02b python bench.py --steps 1 --warmup 3 --seg 8 --no-cpu --no-extra > gpurun_out/r02b_ncu.log 2>&1
tail -2 gpurun_out/r02b_ncu.log
