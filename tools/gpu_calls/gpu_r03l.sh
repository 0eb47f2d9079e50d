#!/bin/bash
mkdir -p gpurun_out
python -c "from cardiax_b200 import _lib; print('needs_build on the box:', _lib.needs_build())"
timeout 1500 python -m pytest tests/ -m gpu -q -x > gpurun_out/r03l_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r03l_pytest_gpu.log
tail -3 gpurun_out/r03l_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python bench.py --steps 10 --warmup 3 --no-extra --no-cpu 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('value %.1f e2e %.1f' % (d['value'], d['e2e']['value']), d['e2e']['step_ms'][:3], d['e2e']['retimed'], d['retimed'])"
