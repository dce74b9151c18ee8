#!/bin/bash
# device-side choice of the solver instance (no mid-search host sync) + Lasso/ElasticNet parity; timing of configs 2 and 4
set -u
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/test_gpu_enet.py tests/test_gpu_ridge.py tests/test_gpu_svc.py tests/test_gpu_splitters.py -q 2>&1 | tail -30 ) | tee gpurun_out/guard_pytest.log
rm -f gpurun_out/guard.log
for wl in c2 c4; do
  echo "=== $wl" | tee -a gpurun_out/guard.log
  timeout 300 python tools/run_workload.py $wl 5 2>&1 | grep -E "rep[1-4]|parity" | cut -c1-330 | tee -a gpurun_out/guard.log
done
timeout 300 python tools/run_workload.py lasso_1024 3 2>&1 | grep -E "rep2|parity" | cut -c1-400 | tee -a gpurun_out/guard.log
echo done
