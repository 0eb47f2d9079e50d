#!/bin/bash
# round 2, GPU call A: parity suite, bench with the packed-fp32 streaming kernel, A/B against horizontal packing, ncu capture
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02a_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02a_pytest.log
tail -5 gpurun_out/r02a_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-extra > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
FK_SO=$PWD/cardiax_b200/csrc/build/alt_h1/libfk_h1.so timeout 300 python bench.py --steps 10 --warmup 3 --no-extra --no-cpu > gpurun_out/r02a_bench_h1.json 2> gpurun_out/r02a_bench_h1.err
timeout 300 python bench.py --steps 6 --warmup 3 --no-extra --no-cpu --workload ens256 > gpurun_out/r02a_bench_ens.json 2> gpurun_out/r02a_bench_ens.err
python - <<'PY'
import json
for f in ("r02a_bench", "r02a_bench_h1", "r02a_bench_ens"):
    try:
        d = json.load(open("gpurun_out/%s.json" % f))
        print(f, "value %.1f e2e %.1f frac %.3f launch_ms %.4f" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["avg_launch_ms"]), d["clocks"])
    except Exception as e:
        print(f, "FAILED", e)
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fk_stream_kernel -s 4 -c 1 -o gpurun_out/prof_stream_r02a python bench.py --steps 1 --warmup 3 --seg 8 --no-cpu --no-extra > gpurun_out/r02a_ncu.log 2>&1
tail -3 gpurun_out/r02a_ncu.log
