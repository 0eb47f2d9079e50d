#!/bin/bash
mkdir -p gpurun_out
for w in fk512 fk128 ens256; do
timeout 300 python bench.py --workload $w --steps 6 --warmup 3 --no-extra --no-cpu 2> gpurun_out/r03h_$w.err | python -c "import json,sys; d=json.load(sys.stdin); print('$w value %.1f e2e %.1f frac %.3f' % (d['value'], d['e2e']['value'], d['roofline']['frac']), d['step_ms'], d['retimed'], d['roofline']['kernel'])" || tail -3 gpurun_out/r03h_$w.err
done
