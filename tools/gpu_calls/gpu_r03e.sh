#!/bin/bash
mkdir -p gpurun_out
for skip in none d2h h2d h2d,d2h; do
FK_E2E_SKIP=$skip timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 6 --warmup 2 --no-extra --no-cpu 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('skip=$skip value %.1f e2e %.1f  (%.1f ms per e2e step, %.1f compute)' % (d['value'], d['e2e']['value'], 4*2048*16384*500/d['e2e']['value']/1e6, d['ms_per_step']), d['e2e']['host_link_gbs']['all_ranks_both_directions_ms'])"
done
