#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 10 --warmup 3 --no-extra --no-cpu"
for r in 1 2 3 4 5 6 7 8; do
timeout 200 $B | python -c "import json,sys; d=json.load(sys.stdin); print('run %.1f' % d['value'], d['step_ms'], d['clocks']['sm_mhz'], d['clocks']['samples'])"
done
