#!/bin/bash
# Round 2, second pass on the slot-layout SMO kernel: full GPU test-suite, then old vs lean on config 2 / config 4.
set -u
mkdir -p gpurun_out
export B200GS_PRINT_US=1
( B200GS_SMO_LEAN=1 timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/lean2_pytest.log; cat gpurun_out/lean2_pytest.log
rm -f gpurun_out/lean2_c2.log gpurun_out/lean2_c4.log gpurun_out/lean2_prof.log
for cfg in "LEAN=0" "LEAN=1" "LEAN=1 CLN=0" "LEAN=1 CLN=20" "LEAN=1 CLN=30"; do
  unset B200GS_SMO_LEAN B200GS_SMO_CLUSTER_N B200GS_SMO_CLUSTER
  for kv in $cfg; do
    case $kv in
      LEAN=*) export B200GS_SMO_LEAN=${kv#LEAN=};;
      CLN=*) export B200GS_SMO_CLUSTER_N=${kv#CLN=};;
      CL=*) export B200GS_SMO_CLUSTER=${kv#CL=};;
    esac
  done
  echo "=== c2 $cfg" | tee -a gpurun_out/lean2_c2.log
  timeout 300 python tools/run_workload.py c2 2 2>&1 | grep -v "^$" | tee -a gpurun_out/lean2_c2.log
done
unset B200GS_SMO_LEAN B200GS_SMO_CLUSTER_N B200GS_SMO_CLUSTER
for l in 0 1; do
  echo "=== c4 LEAN=$l" | tee -a gpurun_out/lean2_c4.log
  B200GS_SMO_LEAN=$l timeout 400 python tools/run_workload.py c4 2 2>&1 | tee -a gpurun_out/lean2_c4.log
done
echo "=== prof lean (all single)" | tee -a gpurun_out/lean2_prof.log
B200GS_SMO_LEAN=1 B200GS_SMO_PROF=1 B200GS_SMO_CLUSTER_N=0 timeout 300 python tools/run_workload.py c2 1 2>&1 | tee -a gpurun_out/lean2_prof.log
echo "=== ncu lean (10 problems)" 
B200GS_SMO_LEAN=1 B200GS_SMO_CLUSTER_N=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:smo_lean -c 1 -o gpurun_out/lean_r02 -f python tools/exp_one.py > gpurun_out/lean2_ncu.log 2>&1
tail -3 gpurun_out/lean2_ncu.log
echo done
