#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02w_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02w_smoke.log
