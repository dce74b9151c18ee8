#!/bin/bash
# Round-end measurement on one B200 (run under gpurun): tests, bench line, ncu launch list, SMO DRAM traffic, kernel captures.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 2500 gpurun_out/bench_n1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:smo_ -c 2 --csv \
    --log-file gpurun_out/smo_dram_r01.csv python tools/run_workload.py c2 1 > gpurun_out/smo_dram_run.log 2>&1
B200GS_SMO_CLUSTER=4 B200GS_SMO_CLUSTER_N=100000 timeout 600 ncu --set full --clock-control none --import-source on -k regex:smo_colown -c 1 \
    -o gpurun_out/colown_v3_r01 -f python tools/exp_one.py > gpurun_out/colown_v3_ncu.log 2>&1
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 400 gpurun_out/bench_ref.json
echo done
