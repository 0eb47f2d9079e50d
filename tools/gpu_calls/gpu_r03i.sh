#!/bin/bash
mkdir -p gpurun_out
for r in 1 2 3 4 5; do
timeout 300 python bench.py --steps 10 --warmup 3 --no-extra --no-cpu 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('value %.1f e2e %.1f' % (d['value'], d['e2e']['value']), d['e2e']['step_ms'], d['e2e']['retimed'], d['retimed'])"
done
