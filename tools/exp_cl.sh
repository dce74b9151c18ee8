#!/bin/bash
set -u
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_gemm.py -x -q 2>&1 | tail -4 ) | tee gpurun_out/cl_pytest.log
export B200GS_PRINT_US=1
echo "=== prof cluster kernel (default schedule)" | tee gpurun_out/cl_prof.log
B200GS_SMO_PROF=1 timeout 300 python tools/run_workload.py c2 1 2>&1 | grep -E "prof|us/iter" | tee -a gpurun_out/cl_prof.log
echo "=== prof cluster kernel alone (10 problems only)" | tee -a gpurun_out/cl_prof.log
B200GS_SMO_PROF=1 B200GS_SMO_CLUSTER=4 B200GS_SMO_CLUSTER_N=100000 timeout 300 python tools/exp_one.py 2>&1 | tail -3 | tee -a gpurun_out/cl_prof.log
echo done
