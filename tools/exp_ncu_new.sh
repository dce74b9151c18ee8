#!/bin/bash
# ncu captures of the kernels added this round besides the SMO ones: decision values, Lasso coordinate descent, Ridge quadratic forms
set -u
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decision_kernel -s 1 -c 1 -o gpurun_out/decision_r02 -f python tools/run_workload.py c2 1 > gpurun_out/ncu_decision.log 2>&1; tail -2 gpurun_out/ncu_decision.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decision_kernel -s 1 -c 1 -o gpurun_out/decision_r02_c4 -f python tools/run_workload.py c4 1 > gpurun_out/ncu_decision4.log 2>&1; tail -2 gpurun_out/ncu_decision4.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:enet_cd -c 1 -o gpurun_out/enet_r02 -f python tools/run_workload.py lasso_1024 1 > gpurun_out/ncu_enet.log 2>&1; tail -2 gpurun_out/ncu_enet.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ridge_quad -c 1 -o gpurun_out/quad_r02 -f python tools/run_workload.py c5 1 > gpurun_out/ncu_quad.log 2>&1; tail -2 gpurun_out/ncu_quad.log
ls -la gpurun_out/*.ncu-rep | tail -5
echo done
