#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/probe_crossover.py > gpurun_out/r03c_crossover.log 2>&1
cut -c1-150 gpurun_out/r03c_crossover.log
