#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_long_horizon.py -m gpu -q -x -k "cluster or resident or config2 or bitwise or long or breakup" > gpurun_out/r02zc_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02zc_pytest.log
tail -3 gpurun_out/r02zc_pytest.log
timeout 300 python tools/probe_cfg12.py > gpurun_out/r02zc_cfg12.log 2>&1
cat gpurun_out/r02zc_cfg12.log
