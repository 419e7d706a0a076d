#!/bin/bash
# round 2: evidence for profiles/: smoke, ncu launch list + full capture of the final recon kernel, a full capture of the
# config-4 kernels, compute-sanitizer memcheck / racecheck of the v14 kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2h_smoke.log 2>&1; echo "smoke exit $?"; tail -1 $O/r2h_smoke.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $O/r02_v14_launches.csv \
    python bench.py --steps 4 --warmup 3 --skip-extras > $O/r2h_bench_under_ncu.log 2>&1; echo "launch list exit $?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:recon_ -s 5 -c 2 -f -o $O/recon_v14b \
    python bench.py --steps 4 --warmup 3 --skip-extras > $O/r2h_bench_under_ncu_full.log 2>&1; echo "full capture exit $?"
timeout 400 ncu --set full --clock-control none -k regex:"deblock_rgba_tile|recon_tile" -s 6 -c 2 -f -o $O/config4_v14 \
    python -m pytest tests/test_gpu_full_size.py -q -x -k "config4" > $O/r2h_config4_under_ncu.log 2>&1; echo "config4 capture exit $?"
K='config1_qcif or cif_borders or 4mv or unaligned or many_events or all_intra or tiny or decode_step or readback or interleave or pipelined or group'
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py tests/test_gpu_handbuilt.py -q -x -k "$K or 64" > $O/r02_sanitizer_memcheck.log 2>&1
echo "memcheck exit $?"; tail -3 $O/r02_sanitizer_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -q -x -k "config1_qcif or many_events or unaligned_47 or pipelined" > $O/r02_sanitizer_racecheck.log 2>&1
echo "racecheck exit $?"; tail -3 $O/r02_sanitizer_racecheck.log
ls -la $O | grep -E "r02_|recon_v14b|config4_v14"
