#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 8 --warmup 3 --no-extra --no-cpu"
for r in 1 2; do
timeout 200 $B | python -c "import json,sys; d=json.load(sys.stdin); print('prof   ', d['value'], d['ms_per_step'], d['roofline']['avg_launch_ms'], d['roofline']['kernel_share_of_step'])"
FK_BENCH_NOPROF=1 timeout 200 $B | python -c "import json,sys; d=json.load(sys.stdin); print('noprof ', d['value'], d['ms_per_step'])"
done
