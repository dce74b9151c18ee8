#!/bin/bash
# Ridge general splitters + quadratic-form scoring kernel + where the end-to-end time of configs 3/5 goes (host side)
set -u
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_splitters.py tests/test_gpu_ridge.py tests/test_gpu_scoring.py -m gpu -q 2>&1 | tail -25 ) | tee gpurun_out/r2_pytest.log
timeout 300 python tools/run_workload.py c5 3 2>&1 | grep -E "rep2|parity" | cut -c1-330 | tee gpurun_out/r2_c5.log
for k in ${E2E_KEYS:-c3 c5}; do echo "=== $k"; timeout 600 python tools/exp_e2e.py $k 2>&1 | head -32 | cut -c1-200; done | tee gpurun_out/r2_e2e.log
echo done
