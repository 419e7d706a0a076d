#!/bin/bash
# ncu launch list of the bench command + one full capture of the recon kernel.
set +e
mkdir -p gpurun_out
TAG=${1:-v0}
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 4 --warmup 3 --skip-extras > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
echo "launch list exit $?"
ncu --set full --clock-control none --import-source on -k regex:recon_kernel -s 5 -c 2 -f -o gpurun_out/recon_${TAG} \
    python bench.py --steps 4 --warmup 3 --skip-extras > gpurun_out/bench_under_ncu_full_${TAG}.log 2>&1
echo "full capture exit $?"
ls -la gpurun_out/
