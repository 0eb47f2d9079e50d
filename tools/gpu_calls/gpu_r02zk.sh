#!/bin/bash
mkdir -p gpurun_out
FK_SO=$PWD/cardiax_b200/csrc/build/alt_timing/libfk_timing.so timeout 300 python tools/probe_stream_timing.py 2 > gpurun_out/r02zk_timing.log 2>&1
cat gpurun_out/r02zk_timing.log
B="python bench.py --steps 6 --warmup 3 --no-extra --no-cpu"
for r in 1 2; do timeout 200 $B | python -c "import json,sys; d=json.load(sys.stdin); print(d['value'], d['roofline']['avg_launch_ms'])"; done
