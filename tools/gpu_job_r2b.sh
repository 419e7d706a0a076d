#!/bin/bash
# round 2, second GPU job: TMA swizzle probe, full GPU test-suite on the dense-tile build, ncu capture of the v14 kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
tools/tma_swizzle_probe.bin > $O/r2b_swizzle_probe.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40 > $O/r2b_pytest.log
ncu --set full --clock-control none --import-source on -k regex:recon_ -s 5 -c 2 -f -o $O/recon_v14 \
    python bench.py --steps 4 --warmup 3 --skip-extras > $O/r2b_bench_under_ncu.log 2>&1
echo "ncu exit $?"
tail -3 $O/r2b_pytest.log
