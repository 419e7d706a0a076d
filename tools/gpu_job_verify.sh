#!/bin/bash
# Quick verification on a fresh box: smoke, GPU parity tests, one bench line.  usage: tools/gpu_job_verify.sh <tag>
set +e
TAG=${1:-verify}
mkdir -p gpurun_out
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_${TAG}.log 2>&1
echo "smoke exit $?"; tail -2 gpurun_out/smoke_${TAG}.log
echo "== pytest gpu"
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_${TAG}.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu_${TAG}.log
echo "== bench"
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
echo "bench exit $?"; tail -3 gpurun_out/bench_${TAG}.err; python -c "
import json;d=json.load(open('gpurun_out/bench_${TAG}.json'));r=d['roofline'];print('value',d['value'],'kernel_ms',r['kernel_ms'],'frac',r['frac'],'e2e',d['e2e']['value'],'cpu',d['cpu_baseline']['value'],d['clocks'])"
