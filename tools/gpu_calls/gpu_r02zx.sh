#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 2 --no-extra --no-cpu --numerics exact"
for v in "" "--T 2" "--T 1 --cta-threads 128" "--T 1 --cta-threads 96" "--T 2 --cta-threads 64"; do
  timeout 300 $B $v | python -c "import json,sys; d=json.load(sys.stdin); print('exact [$v]', round(d['value'],1), d['roofline']['launch_geometry'])"
done
