#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -s > gpurun_out/r02d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02d_pytest.log
grep -E "config 2 step|passed|failed|rc=|Error" gpurun_out/r02d_pytest.log | tail -24
FK_RES_TIMING=1 timeout 300 python tools/probe_res_timing.py > gpurun_out/r02d_res_timing.log 2>&1
cat gpurun_out/r02d_res_timing.log
