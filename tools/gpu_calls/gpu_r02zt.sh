#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/probe_step_jitter.py 2>&1 | tail -6
echo NO_PDL
FK_NO_PDL=1 timeout 300 python tools/probe_step_jitter.py 2>&1 | tail -6
