#!/bin/bash
# round 2, first GPU job: parity of the v14 kernel (three RGBA store variants) + A/B bench against v13
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/r2a_smi.txt
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40 > $O/r2a_pytest_default.log
for v in tma_dense direct; do
  H263CU_LIB=$PWD/h263_rs_b200/libh263cu_$v.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -15 > $O/r2a_pytest_$v.log
done
for v in default v13 tma_dense direct default v13; do
  L=$PWD/h263_rs_b200/libh263cu_$v.so
  [ $v = default ] && L=$PWD/h263_rs_b200/libh263cu.so
  H263CU_LIB=$L timeout 300 python bench.py --steps 20 --warmup 3 --skip-extras > $O/r2a_bench_$v.$RANDOM.json 2> $O/r2a_bench_$v.err
done
grep -h -o '"kernel_ms_per_launch_events": [0-9.]*' $O/r2a_bench_*.json
