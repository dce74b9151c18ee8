#!/bin/bash
# exclusive-SM tier: 512 threads x 16 slots vs 1024 threads x 8 slots (B200GS_LEAN_EXCL_WIDE), same box, with the tier timeline
set -u
mkdir -p gpurun_out
rm -f gpurun_out/wide.log
for wide in 0 1 0 1; do
  export B200GS_LEAN_EXCL_WIDE=$wide
  echo "=== c2 EXCL_WIDE=$wide" | tee -a gpurun_out/wide.log
  B200GS_SMO_TIMELINE=1 timeout 300 python tools/run_workload.py c2 3 2>&1 | grep -E "rep2|parity|timeline\] (cluster|exclusive|shared)|#1[0-1]:" | tail -7 | cut -c1-210 | sed 's/profile.*ms_solve/ms_solve/' | tee -a gpurun_out/wide.log
done
B200GS_LEAN_EXCL_WIDE=1 timeout 900 python -m pytest tests/test_gpu_svc.py -q 2>&1 | tail -3 | tee -a gpurun_out/wide.log
echo done
