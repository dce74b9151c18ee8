#!/bin/bash
set -u
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_logreg.py tests/test_gpu_ridge.py -q 2>&1 | tail -12 ) | tee gpurun_out/t1_pytest.log
timeout 300 python tools/run_workload.py c5 3 2>&1 | grep -E "rep2|parity" | cut -c1-330
echo done
