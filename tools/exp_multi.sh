#!/bin/bash
# multi-GPU check on an N-GPU box (gpurun --gpus N): torchrun bench line (cost dealing vs whole-gamma-group dealing),
# the in-process scheduler test and the in_process block of the N=1 bench line
set -u
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | tee gpurun_out/multi_gpus.log
( timeout 900 python -m pytest tests/test_gpu_svc.py -q -k "in_process" 2>&1 | tail -4 ) | tee gpurun_out/multi_pytest.log
for deal in ${DEALS:-cost groups}; do
  B200GS_DEAL=$deal timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $N --steps 3 --warmup 2 > gpurun_out/multi_n${N}_${deal}.json 2> gpurun_out/multi_n${N}_${deal}.err
  tail -c 400 gpurun_out/multi_n${N}_${deal}.json; tail -3 gpurun_out/multi_n${N}_${deal}.err
done
timeout 900 python bench.py --gpus 1 --steps 2 --warmup 2 --no-cpu-baseline --no-secondary > gpurun_out/multi_n1_inproc.json 2> gpurun_out/multi_n1_inproc.err; tail -c 700 gpurun_out/multi_n1_inproc.json
echo done
