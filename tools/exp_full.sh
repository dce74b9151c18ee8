#!/bin/bash
# full GPU test-suite + bench line (new contract) + reference arm, one B200
set -u
mkdir -p gpurun_out
( timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -15 ) | tee gpurun_out/full_pytest.log
for k in c5 c3; do timeout 300 python tools/run_workload.py $k 3 2>&1 | grep -E "rep2|parity" | cut -c1-330 | tee -a gpurun_out/full_c35.log; done
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench5_n1.json 2> gpurun_out/bench5_n1.err; tail -c 600 gpurun_out/bench5_n1.json; tail -3 gpurun_out/bench5_n1.err
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/full_smoke.log
echo done
