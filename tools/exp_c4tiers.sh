#!/bin/bash
# config 4 (throughput-bound): does giving the longest problems an SM of their own (now as SM-time-efficient as sharing) help?
set -u
mkdir -p gpurun_out
rm -f gpurun_out/c4tiers.log
for exn in default 20 40 70; do
  unset B200GS_SMO_EXCLUSIVE_N
  [ "$exn" != "default" ] && export B200GS_SMO_EXCLUSIVE_N=$exn
  echo "=== c4 EXCLUSIVE_N=$exn" | tee -a gpurun_out/c4tiers.log
  B200GS_SMO_TIMELINE=1 timeout 300 python tools/run_workload.py c4 3 2>&1 | grep -E "rep2|parity|timeline\] (cluster|exclusive|shared)|#0:|#1:" | tail -6 | cut -c1-210 | sed 's/profile.*ms_solve/ms_solve/' | tee -a gpurun_out/c4tiers.log
done
echo done
