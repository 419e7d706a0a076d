#!/bin/bash
# round 2: the 8-GPU box: device-to-host bandwidth probe, then the bench on 8 GPUs (torchrun, one process per GPU)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
nvidia-smi topo -m > $O/r2g_topo.txt 2>&1
lscpu | head -25 >> $O/r2g_topo.txt
timeout 150 tools/d2h_bw.bin > $O/r2g_d2h_bw_8gpu.txt 2>&1
tail -14 $O/r2g_d2h_bw_8gpu.txt
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 3 > $O/r2g_bench_n8.json 2> $O/r2g_bench_n8.err
tail -c 300 $O/r2g_bench_n8.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2g_bench_n8.json'))
print({k:d[k] for k in ('value','n_gpus','ms_per_step')}, 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'e2e_side', d['e2e_from_side_info']['value'], d['host_parse']['ms_per_step'], d['host_parse']['threads'], d['parity']['bit_exact_vs_oracle_per_rank'])
PY
