#!/bin/bash
# First GPU job: smoke, parity tests, short bench. Everything is logged under gpurun_out/.
set +e
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
echo "== smoke" | tee gpurun_out/smoke.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" >> gpurun_out/smoke.log 2>&1
echo "smoke exit $?" | tee -a gpurun_out/smoke.log
echo "== pytest gpu"
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
echo "== bench"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?"
tail -5 gpurun_out/bench.err
cat gpurun_out/bench.json
