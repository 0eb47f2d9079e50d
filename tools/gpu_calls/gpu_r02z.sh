#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02z_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02z_pytest.log
tail -3 gpurun_out/r02z_pytest.log
timeout 300 python tools/probe_cfg12.py > gpurun_out/r02z_cfg12.log 2>&1
cat gpurun_out/r02z_cfg12.log
