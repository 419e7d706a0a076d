#!/bin/bash
# round 2, second session: the bench on 8 GPUs (torchrun, one process per GPU) with the v15 kernel + the reference arm on that box
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 3 > $O/r02_v15_bench_n8.json 2> $O/r2k_bench_n8.err
tail -c 300 $O/r2k_bench_n8.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_v15_bench_n8.json'))
print({k:d[k] for k in ('value','n_gpus','ms_per_step')}, 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'e2e_side', d['e2e_from_side_info']['value'], d['host_parse']['ms_per_step'], d['host_parse']['threads'], d['parity']['bit_exact_vs_oracle_per_rank'])
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 8 --steps 4 --warmup 3 > $O/r02_v15_bench_n8_reference_arm.json 2> $O/r2k_ref.err
cut -c1-300 $O/r02_v15_bench_n8_reference_arm.json
