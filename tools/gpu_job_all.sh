#!/bin/bash
# Full GPU job: smoke, parity tests, bench, ncu launch list + one full capture of the recon kernel.
# usage: tools/gpu_job_all.sh <tag>   -- everything is logged under gpurun_out/
set +e
TAG=${1:-run}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu_${TAG}.txt 2>&1
nproc >> gpurun_out/gpu_${TAG}.txt
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_${TAG}.log 2>&1
echo "smoke exit $?"; tail -2 gpurun_out/smoke_${TAG}.log
echo "== pytest gpu"
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_${TAG}.log 2>&1
echo "pytest exit $?"; tail -5 gpurun_out/pytest_gpu_${TAG}.log
echo "== bench"
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
echo "bench exit $?"; tail -3 gpurun_out/bench_${TAG}.err; cat gpurun_out/bench_${TAG}.json
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 4 --warmup 3 --skip-extras > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
echo "launch list exit $?"
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:recon_ -s 5 -c 2 -f -o gpurun_out/recon_${TAG} \
    python bench.py --steps 4 --warmup 3 --skip-extras > gpurun_out/bench_under_ncu_full_${TAG}.log 2>&1
echo "full capture exit $?"
ls -la gpurun_out/
