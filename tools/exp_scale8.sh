#!/bin/bash
# same-box weak-scaling check: N=1 and N=8 bench lines (short: no CPU arm, no secondary configs), cost dealing
set -u
mkdir -p gpurun_out
timeout 300 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/scale_n8.json 2> gpurun_out/scale_n8.err
python - <<'PY'
import json
for n in (1, 8):
    try:
        j = json.loads(open("gpurun_out/scale_n%d.json" % n).read().strip().splitlines()[-1])
        print("N=%d value %.1f ms %.1f e2e %.1f launches %d" % (n, j["value"], j["ms_per_step"], j["e2e"]["value"], j["gpu_launches"]))
        print("   phases", {k: (round(v["max"], 1), round(v["mean"], 1)) for k, v in j["phases_ms"].items()})
    except Exception as e:
        print(n, "failed", e)
PY
tail -2 gpurun_out/scale_n8.err
echo done
