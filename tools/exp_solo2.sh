#!/bin/bash
# same-box A/B of the two slot-kernel instances (shared scalar update behind a barrier vs every-warp update) per tier, with the
# tier timeline (B200GS_SMO_TIMELINE): which one should the exclusive-SM tier use?
set -u
mkdir -p gpurun_out
rm -f gpurun_out/solo2.log
for solo in default 0 1 default 0 1; do
  unset B200GS_LEAN_SOLO
  [ "$solo" != "default" ] && export B200GS_LEAN_SOLO=$solo
  echo "=== c2 LEAN_SOLO=$solo" | tee -a gpurun_out/solo2.log
  B200GS_SMO_TIMELINE=1 timeout 300 python tools/run_workload.py c2 3 2>&1 | grep -E "rep2|timeline\] (cluster|exclusive|shared)|#1[0-2]:" | tail -7 | cut -c1-210 | sed 's/profile.*ms_solve/ms_solve/' | tee -a gpurun_out/solo2.log
done
echo done
