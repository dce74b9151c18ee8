#!/bin/bash
# persistent tcgen05 GEMM: tests (own timeout: a protocol bug traps, it must not hang the box), then configs 3 / 5 and an ncu capture
set -u
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_ridge.py tests/test_gpu_logreg.py tests/test_gpu_splitters.py tests/test_gpu_scoring.py -x -q 2>&1 | tail -12 ) | tee gpurun_out/gemm_pytest.log
for k in c3 c5; do timeout 300 python tools/run_workload.py $k 3 2>&1 | grep -E "rep2|parity" | cut -c1-400 | tee -a gpurun_out/gemm_c35.log; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_nt -s 2 -c 3 -o gpurun_out/gemm_r02_c3 -f python tools/run_workload.py c3 1 > gpurun_out/gemm_ncu_c3.log 2>&1; tail -2 gpurun_out/gemm_ncu_c3.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_nt -c 2 -o gpurun_out/gemm_r02_c5 -f python tools/run_workload.py c5 1 > gpurun_out/gemm_ncu_c5.log 2>&1; tail -2 gpurun_out/gemm_ncu_c5.log
timeout 600 ncu --set full --clock-control none -k regex:gemm_nt -c 1 -o gpurun_out/gemm_r02_gram10k -f python tools/exp_gram_tc.py > gpurun_out/gemm_ncu_gram.log 2>&1; tail -2 gpurun_out/gemm_ncu_gram.log
echo done
