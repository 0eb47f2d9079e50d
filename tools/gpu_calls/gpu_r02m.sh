#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "cluster or resident" > gpurun_out/r02m_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02m_pytest.log
tail -3 gpurun_out/r02m_pytest.log
(for k in 4 5; do echo "== kernel $k"; PROBE_KERNEL=$k FK_RES_TIMING=1 timeout 300 python tools/probe_res_timing.py 64 128 256 512 1024 2>&1 | grep -A1 fast; done) > gpurun_out/r02m_res_timing.log 2>&1
cat gpurun_out/r02m_res_timing.log
