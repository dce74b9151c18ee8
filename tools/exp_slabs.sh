#!/bin/bash
# occupancy-aware slab count of the decision kernel: parity (SVC suite) and timing of configs 2 / 4, then ncu captures
set -u
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/test_gpu_svc.py tests/test_gpu_scoring.py -q 2>&1 | tail -5 ) | tee gpurun_out/slabs_pytest.log
rm -f gpurun_out/slabs.log
for wl in c2 c4; do
  timeout 300 python tools/run_workload.py $wl 4 2>&1 | grep -E "rep[2-3]|parity" | cut -c1-330 | tee -a gpurun_out/slabs.log
done
bash tools/exp_ncu_new.sh
echo done
