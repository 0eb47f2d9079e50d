#!/bin/bash
mkdir -p gpurun_out
for v in base L3 P1 P2 PF2; do
  so=""; [ $v != base ] && so="FK_SO=$PWD/cardiax_b200/csrc/build/alt_$v/libfk_$v.so"
  env $so timeout 300 python tools/probe_hetero.py > gpurun_out/r03b_hetero_$v.log 2>&1
  echo $v; tail -3 gpurun_out/r03b_hetero_$v.log
done
