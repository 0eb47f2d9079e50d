#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -m gpu -q -x > gpurun_out/r02zo_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02zo_pytest_gpu.log
tail -4 gpurun_out/r02zo_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02zo_smoke.log 2>&1; tail -2 gpurun_out/r02zo_smoke.log
timeout 300 python tools/probe_hetero.py 2>&1 | tail -3
