#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r02zy_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02zy_pytest.log
tail -2 gpurun_out/r02zy_pytest.log
B="python bench.py --steps 10 --warmup 3 --no-extra --no-cpu"
for r in 1 2; do
timeout 200 $B | python -c "import json,sys; d=json.load(sys.stdin); print('fk4096 %.1f' % d['value'], d['roofline']['avg_launch_ms'])"
timeout 200 $B --workload ens256 | python -c "import json,sys; d=json.load(sys.stdin); print('ens256 %.1f' % d['value'], d['roofline']['avg_launch_ms'])"
done
