#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/probe_exact.py > gpurun_out/r02h_probe_exact.log 2>&1
cat gpurun_out/r02h_probe_exact.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fk_stream_kernel -s 2 -c 1 -o gpurun_out/prof_stream_exact_r02h python bench.py --steps 1 --warmup 1 --seg 8 --no-cpu --no-extra --numerics exact > gpurun_out/r02h_ncu.log 2>&1
tail -2 gpurun_out/r02h_ncu.log
