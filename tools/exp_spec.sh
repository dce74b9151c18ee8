#!/bin/bash
set -u
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_gpu_svc.py -x -q 2>&1 | tail -5 ) | tee gpurun_out/spec_pytest.log
export B200GS_PRINT_US=1
rm -f gpurun_out/spec_c2.log
for sp in 1 0; do
  echo "=== c2 default schedule SPEC=$sp" | tee -a gpurun_out/spec_c2.log
  B200GS_SMO_SPEC=$sp timeout 300 python tools/run_workload.py c2 3 2>&1 | grep -E "rep2|us/iter|parity" | cut -c1-200 | tee -a gpurun_out/spec_c2.log
done
echo "=== prof SPEC=1" | tee -a gpurun_out/spec_c2.log
B200GS_SMO_PROF=1 timeout 300 python tools/run_workload.py c2 1 2>&1 | grep -E "prof\]" | tee -a gpurun_out/spec_c2.log
echo done
