#!/bin/bash
# compute-sanitizer (memcheck, racecheck) on the small parity cases; logs under gpurun_out/.
set +e
mkdir -p gpurun_out
K='config1_qcif or cif_borders or 4mv or unaligned or many_events or all_intra or tiny'
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -q -x -k "$K" > gpurun_out/sanitizer_memcheck.log 2>&1
echo "memcheck exit $?"; tail -4 gpurun_out/sanitizer_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -q -x -k "config1_qcif or many_events" > gpurun_out/sanitizer_racecheck.log 2>&1
echo "racecheck exit $?"; tail -4 gpurun_out/sanitizer_racecheck.log
timeout 300 python bench.py --steps 10 --warmup 3 --skip-extras > gpurun_out/bench_quick.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_quick.json')); print('kernel_ms', d['roofline']['kernel_ms'], 'frac', d['roofline']['frac'])"
