#!/bin/bash
# device-side choice of the solver instance (no mid-search host sync), cached memory plan, Lasso/ElasticNet and multinomial
# LogisticRegression parity; timing of configs 2, 4, 3
set -u
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/test_gpu_logreg.py tests/test_gpu_enet.py tests/test_gpu_scoring.py tests/test_gpu_svc.py -q 2>&1 | tail -40 ) | tee gpurun_out/guard_pytest.log
rm -f gpurun_out/guard.log
for wl in c2 c4 c3; do
  echo "=== $wl" | tee -a gpurun_out/guard.log
  timeout 300 python tools/run_workload.py $wl 6 2>&1 | grep -E "rep[1-5]|parity" | cut -c1-330 | tee -a gpurun_out/guard.log
done
echo done
