#!/bin/bash
mkdir -p gpurun_out
for v in base A; do
  so=""; [ $v != base ] && so="FK_SO=$PWD/cardiax_b200/csrc/build/alt_$v/libfk_$v.so"
  env $so timeout 300 python tools/probe_hetero.py > gpurun_out/r02y_hetero_$v.log 2>&1
  cat gpurun_out/r02y_hetero_$v.log
done
