#!/bin/bash
# Round-2 closing measurement on one B200 (run under gpurun): full GPU suite, bench line (+ reference arm), smoke, the ncu launch
# list of the bench command and captures of the kernels added this round.  Numbers printed under ncu are never bench values.
set -u
mkdir -p gpurun_out
( timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -15 ) | tee gpurun_out/final_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err; tail -c 300 gpurun_out/final_bench_n1.json; tail -2 gpurun_out/final_bench_n1.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err; tail -c 300 gpurun_out/final_bench_ref.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/final_smoke.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-secondary > gpurun_out/bench_under_ncu.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_r02.csv | head -30
[ "${NCU_NEW:-0}" = "1" ] && bash tools/exp_ncu_new.sh
echo done
