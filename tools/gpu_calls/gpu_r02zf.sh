#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/probe_ens_nan.py 128 500 13 > gpurun_out/r02zf_ens_smooth.log 2>&1
grep " u:" gpurun_out/r02zf_ens_smooth.log | cut -c1-200
