#!/bin/bash
mkdir -p gpurun_out
for dbg in 0 1 2 3 4 5; do
echo "== FK_CL_DEBUG=$dbg (1: no remote stores, 2: relaxed cluster barrier, 4: block barrier only)"
FK_CL_DEBUG=$dbg PROBE_KERNEL=5 FK_RES_TIMING=1 timeout 300 python tools/probe_res_timing.py 64 128 2>&1 | grep -A1 fast
done > gpurun_out/r02l_cluster_timing.log 2>&1
cat gpurun_out/r02l_cluster_timing.log
