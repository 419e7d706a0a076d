#!/bin/bash
# round 2: native single-stream driver, then the bench on 2 GPUs (torchrun, one process per GPU)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
true
true
true
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > $O/r2f_bench_n2.json 2> $O/r2f_bench_n2.err
tail -c 300 $O/r2f_bench_n2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2f_bench_n2.json'))
print({k:d[k] for k in ('value','n_gpus','ms_per_step')}, d['e2e']['value'], d['e2e']['ms_per_step'], d['host_parse'], d['parity'])
PY
