#!/bin/bash
# round 2, third GPU job: full GPU test-suite, full default bench, A/B variants and ablations of the v14 kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > $O/r2c_pytest.log
tail -3 $O/r2c_pytest.log
grep -q "passed" $O/r2c_pytest.log && ! grep -q "failed" $O/r2c_pytest.log || { echo "tests not green: stopping"; exit 1; }
timeout 300 python bench.py > $O/r2c_bench_full.json 2> $O/r2c_bench_full.err
for v in default v13 direct abl1 abl2 abl4 abl8 abl15 default v13; do
  L=$PWD/h263_rs_b200/libh263cu_$v.so
  [ $v = default ] && L=$PWD/h263_rs_b200/libh263cu.so
  H263CU_LIB=$L timeout 100 python bench.py --steps 20 --warmup 3 --skip-extras > $O/r2c_bench_$v.$RANDOM.json 2> $O/r2c_bench_$v.err
done
for f in $O/r2c_bench_*.json; do echo $f $(grep -h -o '"kernel_ms_per_launch_events": [0-9.]*' $f); done
