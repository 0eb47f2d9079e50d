#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02i_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02i_pytest.log
tail -5 gpurun_out/r02i_pytest.log
timeout 600 python tools/probe_exact.py > gpurun_out/r02i_probe_exact.log 2>&1
cat gpurun_out/r02i_probe_exact.log
