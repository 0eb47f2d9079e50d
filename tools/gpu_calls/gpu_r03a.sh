#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "four_corners or large_field" > gpurun_out/r03a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r03a_pytest.log
tail -15 gpurun_out/r03a_pytest.log
