#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -m gpu -q -x > gpurun_out/r02zr_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02zr_pytest_gpu.log
tail -3 gpurun_out/r02zr_pytest_gpu.log
B="python bench.py --steps 8 --warmup 3 --no-extra --no-cpu"
for r in 1 2; do
timeout 200 $B | python -c "import json,sys; d=json.load(sys.stdin); print('pdl    ', d['value'], d['ms_per_step'], d['roofline']['avg_launch_ms'], d['roofline']['kernel_share_of_step'])"
FK_NO_PDL=1 timeout 200 $B | python -c "import json,sys; d=json.load(sys.stdin); print('plain  ', d['value'], d['ms_per_step'], d['roofline']['avg_launch_ms'], d['roofline']['kernel_share_of_step'])"
done
timeout 200 $B --workload ens256 | python -c "import json,sys; d=json.load(sys.stdin); print('ens pdl    ', d['value'], d['ms_per_step'])"
FK_NO_PDL=1 timeout 200 $B --workload ens256 | python -c "import json,sys; d=json.load(sys.stdin); print('ens plain  ', d['value'], d['ms_per_step'])"
