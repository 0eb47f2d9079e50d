#!/bin/bash
# round 2, GPU call G: exact numerics after the branch-free division (parity suite + exact throughput), host topology probe
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02g_pytest.log
tail -5 gpurun_out/r02g_pytest.log
timeout 300 python bench.py --steps 4 --warmup 3 --no-extra --no-cpu --numerics exact --seg 100 > gpurun_out/r02g_exact.json 2> gpurun_out/r02g_exact.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r02g_exact.json"))
    print("exact fk4096: value %.1f launch_ms %.4f" % (d["value"], d["roofline"]["avg_launch_ms"]), d["roofline"]["launch_geometry"])
except Exception as e:
    print("FAILED", e)
PY
tail -3 gpurun_out/r02g_exact.err
(nvidia-smi topo -m; lscpu | head -25; cat /sys/devices/system/node/online; ls /sys/bus/pci/devices | head -50; for d in /sys/bus/pci/devices/*; do echo $d $(cat $d/numa_node 2>/dev/null) $(cat $d/class 2>/dev/null); done | grep 0x0302) > gpurun_out/r02g_topo.txt 2>&1
