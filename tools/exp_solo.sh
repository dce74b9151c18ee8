#!/bin/bash
set -u
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_gpu_svc.py -x -q 2>&1 | tail -4 ) | tee gpurun_out/solo_pytest.log
export B200GS_PRINT_US=1
rm -f gpurun_out/solo.log
for wl in c4 c2; do
  for solo in default 1 0; do
    unset B200GS_LEAN_SOLO
    [ "$solo" != "default" ] && export B200GS_LEAN_SOLO=$solo
    echo "=== $wl LEAN_SOLO=$solo" | tee -a gpurun_out/solo.log
    timeout 300 python tools/run_workload.py $wl 3 2>&1 | grep -E "rep2|us/iter|parity" | cut -c1-200 | tee -a gpurun_out/solo.log
  done
done
echo done
