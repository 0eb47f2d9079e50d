#!/bin/bash
# compute-sanitizer memcheck + racecheck on the streaming kernel after the shared-memory layout change
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/_sanitize_case.py > gpurun_out/r03g_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r03g_memcheck.log
tail -4 gpurun_out/r03g_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 9 python tools/_sanitize_case.py > gpurun_out/r03g_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/r03g_racecheck.log
tail -6 gpurun_out/r03g_racecheck.log
