#!/bin/bash
# simulation-based three-tier schedule: parity (SVC suite), config 2 / 4 timing with the tier timeline, against the previous split
set -u
mkdir -p gpurun_out
rm -f gpurun_out/sched2.log
for mode in default "10 35"; do
  unset B200GS_SMO_CLUSTER_N B200GS_SMO_EXCLUSIVE_N
  if [ "$mode" != "default" ]; then set -- $mode; export B200GS_SMO_CLUSTER_N=$1 B200GS_SMO_EXCLUSIVE_N=$2; fi
  echo "=== c2 split=$mode" | tee -a gpurun_out/sched2.log
  B200GS_SMO_TIMELINE=1 timeout 300 python tools/run_workload.py c2 4 2>&1 | grep -E "rep[2-3]|parity|timeline\] (cluster|exclusive|shared)" | tail -6 | cut -c1-210 | sed 's/profile.*ms_solve/ms_solve/' | tee -a gpurun_out/sched2.log
done
unset B200GS_SMO_CLUSTER_N B200GS_SMO_EXCLUSIVE_N
echo "=== c4" | tee -a gpurun_out/sched2.log
timeout 300 python tools/run_workload.py c4 3 2>&1 | grep -E "rep2|parity" | cut -c1-210 | sed 's/profile.*ms_solve/ms_solve/' | tee -a gpurun_out/sched2.log
( timeout 900 python -m pytest tests/test_gpu_svc.py -q 2>&1 | tail -3 ) | tee -a gpurun_out/sched2.log
echo done
