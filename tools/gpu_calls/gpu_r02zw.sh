#!/bin/bash
mkdir -p gpurun_out
FK_SO=$PWD/cardiax_b200/csrc/build/alt_timing/libfk_timing.so timeout 300 python tools/probe_stream_timing.py 2 ens256 > gpurun_out/r02zw_timing_ens.log 2>&1
cat gpurun_out/r02zw_timing_ens.log
FK_SO=$PWD/cardiax_b200/csrc/build/alt_timing/libfk_timing.so timeout 300 python tools/probe_stream_timing.py 2 > gpurun_out/r02zw_timing_fk4096.log 2>&1
cat gpurun_out/r02zw_timing_fk4096.log
