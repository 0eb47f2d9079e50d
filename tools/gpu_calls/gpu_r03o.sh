#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fk_stream_kernel -s 6 -c 1 -o gpurun_out/prof_stream_ens256_r03o -f python bench.py --workload ens256 --steps 1 --warmup 2 --seg 40 --no-cpu --no-extra > gpurun_out/r03o_ncu.log 2>&1
tail -2 gpurun_out/r03o_ncu.log
ls -la gpurun_out/prof_stream_ens256_r03o.ncu-rep
