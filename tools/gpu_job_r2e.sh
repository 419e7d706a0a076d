#!/bin/bash
# round 2: GPU test-suite + full bench (single GPU)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > $O/r2e_pytest.log
tail -3 $O/r2e_pytest.log
grep -q "passed" $O/r2e_pytest.log && ! grep -q "failed" $O/r2e_pytest.log || { echo "tests not green: stopping"; exit 1; }
timeout 300 python bench.py > $O/r2e_bench_full.json 2> $O/r2e_bench_full.err
tail -c 600 $O/r2e_bench_full.err
