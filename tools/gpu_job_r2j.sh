#!/bin/bash
# round 2, second session: evidence for profiles/ of the v16 recon kernel: smoke, full bench line, ncu launch list + full capture,
# compute-sanitizer memcheck / racecheck
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2j_smoke.log 2>&1; echo "smoke exit $?"; tail -1 $O/r2j_smoke.log
timeout 900 python bench.py > $O/r02_v16_bench.json 2> $O/r2j_bench.err; echo "bench exit $?"; cut -c1-400 $O/r02_v16_bench.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $O/r02_v16_launches.csv \
    python bench.py --steps 4 --warmup 3 --skip-extras > $O/r2j_bench_under_ncu.log 2>&1; echo "launch list exit $?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:recon_ -s 5 -c 2 -f -o $O/recon_v16 \
    python bench.py --steps 4 --warmup 3 --skip-extras > $O/r2j_bench_under_ncu_full.log 2>&1; echo "full capture exit $?"
K='config1_qcif or cif_borders or 4mv or unaligned or many_events or all_intra or tiny or decode_step or readback or interleave or pipelined or group or disposable'
timeout 700 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py tests/test_gpu_handbuilt.py -q -x -k "$K or 64" > $O/r02_v16_sanitizer_memcheck.log 2>&1
echo "memcheck exit $?"; tail -3 $O/r02_v16_sanitizer_memcheck.log
timeout 700 compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -q -x -k "config1_qcif or many_events or unaligned_47 or pipelined" > $O/r02_v16_sanitizer_racecheck.log 2>&1
echo "racecheck exit $?"; tail -3 $O/r02_v16_sanitizer_racecheck.log
