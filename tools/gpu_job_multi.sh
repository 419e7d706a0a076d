#!/bin/bash
# Multi-GPU bench exactly as the driver launches it, plus the reference arm.
# usage: tools/gpu_job_multi.sh N tag
set +e
N=${1:-2}; TAG=${2:-multi}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus_${TAG}.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_${TAG}_n${N}.json 2> gpurun_out/bench_${TAG}_n${N}.err
echo "bench N=$N exit $?"; tail -3 gpurun_out/bench_${TAG}_n${N}.err; cat gpurun_out/bench_${TAG}_n${N}.json | cut -c1-600
timeout 600 python bench.py --impl reference --gpus 1 --steps 4 --warmup 3 > gpurun_out/bench_${TAG}_ref.json 2> gpurun_out/bench_${TAG}_ref.err
echo "reference exit $?"; cat gpurun_out/bench_${TAG}_ref.json | cut -c1-400
