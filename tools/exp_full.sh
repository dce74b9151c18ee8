#!/bin/bash
# full GPU test-suite + bench line (new contract) + reference arm, one B200
set -u
mkdir -p gpurun_out
( timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -15 ) | tee gpurun_out/full_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench4_n1.json 2> gpurun_out/bench4_n1.err; tail -c 600 gpurun_out/bench4_n1.json; tail -3 gpurun_out/bench4_n1.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench4_ref.json 2> gpurun_out/bench4_ref.err; tail -c 900 gpurun_out/bench4_ref.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/full_smoke.log
echo done
