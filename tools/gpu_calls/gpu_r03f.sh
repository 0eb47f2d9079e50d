#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 8 --warmup 3 --no-extra --no-cpu"
for v in base PH U4 base PH U4; do
  so=""; [ $v != base ] && so="FK_SO=$PWD/cardiax_b200/csrc/build/alt_$v/libfk_$v.so"
  env $so timeout 200 $B | python -c "import json,sys; d=json.load(sys.stdin); print('$v fk4096 %.1f' % d['value'], d['roofline']['avg_launch_ms'], d['retimed'] is not None)"
done
for v in base PH U4; do
  so=""; [ $v != base ] && so="FK_SO=$PWD/cardiax_b200/csrc/build/alt_$v/libfk_$v.so"
  env $so timeout 200 $B --workload ens256 | python -c "import json,sys; d=json.load(sys.stdin); print('$v ens256 %.1f' % d['value'], d['roofline']['avg_launch_ms'])"
done
