#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "cluster or resident or config2" > gpurun_out/r02o_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02o_pytest.log
tail -3 gpurun_out/r02o_pytest.log
timeout 600 python tools/probe_small_res.py > gpurun_out/r02o_small_res.log 2>&1
cat gpurun_out/r02o_small_res.log
