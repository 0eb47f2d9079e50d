#!/bin/bash
mkdir -p gpurun_out
(for c in 0 4 68; do FK_RES_TIMING=$c timeout 300 python tools/probe_res_timing.py 128 2>&1 | grep -A1 fast; done
for c in 0 6 78; do FK_RES_TIMING=$c timeout 300 python tools/probe_res_timing.py 512 2>&1 | grep -A1 fast; done) > gpurun_out/r02n_res_timing.log 2>&1
cat gpurun_out/r02n_res_timing.log
