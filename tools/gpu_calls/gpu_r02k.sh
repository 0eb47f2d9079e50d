#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "cluster or resident" > gpurun_out/r02k_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02k_pytest.log
tail -12 gpurun_out/r02k_pytest.log
timeout 900 python tools/probe_cluster.py > gpurun_out/r02k_probe_cluster.log 2>&1
cat gpurun_out/r02k_probe_cluster.log
