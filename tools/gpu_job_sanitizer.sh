#!/bin/bash
# compute-sanitizer (memcheck, racecheck) on the small parity cases and the batched decode_step path; logs under gpurun_out/.
set +e
mkdir -p gpurun_out
K='config1_qcif or cif_borders or 4mv or unaligned or many_events or all_intra or tiny or decode_step or readback or interleave or beyond_the_range'
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -q -x -k "$K" > gpurun_out/sanitizer_memcheck.log 2>&1
echo "memcheck exit $?"; tail -4 gpurun_out/sanitizer_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -q -x -k "config1_qcif or many_events or decode_step or beyond_the_range" > gpurun_out/sanitizer_racecheck.log 2>&1
echo "racecheck exit $?"; tail -4 gpurun_out/sanitizer_racecheck.log
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_full_size.py -q -x -k "config4" > gpurun_out/sanitizer_memcheck_fullsize.log 2>&1
echo "memcheck full-size (256 x 4CIF, deblock, border vectors) exit $?"; tail -4 gpurun_out/sanitizer_memcheck_fullsize.log
