#!/bin/bash
set -u
mkdir -p gpurun_out
export B200GS_PRINT_US=1 B200GS_SMO_LEAN=1
( timeout 900 python -m pytest tests/test_gpu_scoring.py tests/test_gpu_svc.py -x -q 2>&1 | tail -12 ) | tee gpurun_out/sched_pytest.log
rm -f gpurun_out/sched_c2.log
for cfg in "default" "CLN=10 EXN=20" "CLN=10 EXN=10" "CLN=10 EXN=30" "CLN=14 EXN=20" "CLN=10 EXN=0" "CL=8 CLN=10 EXN=20"; do
  unset B200GS_SMO_CLUSTER_N B200GS_SMO_CLUSTER B200GS_SMO_EXCLUSIVE_N
  for kv in $cfg; do
    case $kv in
      CLN=*) export B200GS_SMO_CLUSTER_N=${kv#CLN=};;
      CL=*) export B200GS_SMO_CLUSTER=${kv#CL=};;
      EXN=*) export B200GS_SMO_EXCLUSIVE_N=${kv#EXN=};;
    esac
  done
  echo "=== c2 $cfg" | tee -a gpurun_out/sched_c2.log
  timeout 300 python tools/run_workload.py c2 3 2>&1 | grep -E "rep2|us/iter|parity" | cut -c1-260 | tee -a gpurun_out/sched_c2.log
done
unset B200GS_SMO_CLUSTER_N B200GS_SMO_CLUSTER B200GS_SMO_EXCLUSIVE_N
echo "=== e2e" | tee gpurun_out/sched_e2e.log
timeout 300 python tools/exp_e2e.py 2>&1 | tee -a gpurun_out/sched_e2e.log
echo "=== e2e LEAN=0" | tee -a gpurun_out/sched_e2e.log
B200GS_SMO_LEAN=0 timeout 300 python tools/exp_e2e.py 2>&1 | head -6 | tee -a gpurun_out/sched_e2e.log
echo done
