#!/bin/bash
# Round 2: A/B of the single-CTA SMO kernels on one B200 (run under gpurun).  Output: gpurun_out/lean_*.log
set -u
mkdir -p gpurun_out
export B200GS_PRINT_US=1
( timeout 1200 python -m pytest tests/test_gpu_svc.py -x -q 2>&1 | tail -15 ) > gpurun_out/lean_pytest.log; cat gpurun_out/lean_pytest.log
for cfg in "LEAN=0" "LEAN=1 G=4" "LEAN=1 G=8" "LEAN=1 G=4 CLN=0" "LEAN=1 G=8 CLN=0" "LEAN=1 G=4 CL=8 CLN=10" "LEAN=1 G=4 CL=2 CLN=10"; do
  unset B200GS_SMO_LEAN B200GS_LEAN_G B200GS_SMO_CLUSTER_N B200GS_SMO_CLUSTER
  for kv in $cfg; do
    case $kv in
      LEAN=*) export B200GS_SMO_LEAN=${kv#LEAN=};;
      G=*) export B200GS_LEAN_G=${kv#G=};;
      CLN=*) export B200GS_SMO_CLUSTER_N=${kv#CLN=};;
      CL=*) export B200GS_SMO_CLUSTER=${kv#CL=};;
    esac
  done
  echo "=== c2 $cfg" | tee -a gpurun_out/lean_c2.log
  timeout 300 python tools/run_workload.py c2 2 2>&1 | grep -v "^$" | tee -a gpurun_out/lean_c2.log
done
unset B200GS_SMO_LEAN B200GS_LEAN_G B200GS_SMO_CLUSTER_N B200GS_SMO_CLUSTER
for g in 4 8; do
  echo "=== c4 LEAN G=$g" | tee -a gpurun_out/lean_c4.log
  B200GS_LEAN_G=$g timeout 400 python tools/run_workload.py c4 2 2>&1 | tee -a gpurun_out/lean_c4.log
done
echo "=== prof G=4 (all single)" | tee -a gpurun_out/lean_prof.log
B200GS_SMO_PROF=1 B200GS_SMO_CLUSTER_N=0 B200GS_LEAN_G=4 timeout 300 python tools/run_workload.py c2 1 2>&1 | tee -a gpurun_out/lean_prof.log
echo "=== prof G=8 (all single)" | tee -a gpurun_out/lean_prof.log
B200GS_SMO_PROF=1 B200GS_SMO_CLUSTER_N=0 B200GS_LEAN_G=8 timeout 300 python tools/run_workload.py c2 1 2>&1 | tee -a gpurun_out/lean_prof.log
echo done
