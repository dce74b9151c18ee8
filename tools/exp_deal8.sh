#!/bin/bash
# N-GPU weak-scaling bench line under the three dealing strategies (short: no CPU arm, no secondary configs)
set -u
N=${1:-8}
mkdir -p gpurun_out
for deal in ${DEALS:-cost affinity groups}; do
  B200GS_DEAL=$deal timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/deal_n${N}_${deal}.json 2> gpurun_out/deal_n${N}_${deal}.err
  python - <<PY
import json
try:
    j = json.loads(open("gpurun_out/deal_n${N}_${deal}.json").read().strip().splitlines()[-1])
    print("${deal}", "N=${N}", "value", round(j["value"], 1), "ms", round(j["ms_per_step"], 1), "e2e", round(j["e2e"]["value"], 1), "launches", j["gpu_launches"])
except Exception as e:
    print("${deal}", "failed", e)
PY
  tail -2 gpurun_out/deal_n${N}_${deal}.err
done
echo done
