#!/bin/bash
# round 2, GPU call ZU (after the shared-memory layout, block order, balanced chunks, throttle): full parity suite, full default bench, ncu launch list of the bench command, ncu --set full of the headline kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02zu_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02zu_pytest.log
tail -3 gpurun_out/r02zu_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02zu_bench.json 2> gpurun_out/r02zu_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r02zu_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02zu_bench_ref.json 2> gpurun_out/r02zu_bench_ref.err; echo "ref rc=$?"
cat gpurun_out/r02zu_bench_ref.json | cut -c1-400
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r02zu_bench.json"))
    print("value %.1f e2e %.1f frac %.3f launch_ms %.4f share %.3f" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["avg_launch_ms"], d["roofline"]["kernel_share_of_step"]), d["clocks"])
    for k, v in (d["other_configs"] or {}).items():
        print(k, {a: b for a, b in v.items() if a in ("value", "us_per_euler_step", "kernel", "seconds", "s2_fired_at_step_40000", "ms_per_500_step_segment", "ours_b200")} if isinstance(v, dict) else v)
except Exception as e:
    print("bench FAILED", e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02zu_bench.csv python bench.py --steps 2 --warmup 1 --seg 40 --no-cpu --no-extra > gpurun_out/r02zu_ncu_list.log 2>&1
tail -2 gpurun_out/r02zu_ncu_list.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fk_stream_kernel -s 4 -c 1 -o gpurun_out/prof_stream_r02zu python bench.py --steps 1 --warmup 3 --seg 8 --no-cpu --no-extra > gpurun_out/r02zu_ncu.log 2>&1
tail -2 gpurun_out/r02zu_ncu.log
