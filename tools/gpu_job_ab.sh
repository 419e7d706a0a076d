#!/bin/bash
# A/B of experiment builds and run-time hints: bench only (kernel time per step from CUDA events).
# usage: tools/gpu_job_ab.sh [--test] spec...    spec = <lib variant|default>[,ENV=VAL...]
# e.g.   tools/gpu_job_ab.sh default default,H263CU_HINTS=0 wide cta4
set +e
mkdir -p gpurun_out
if [ "$1" = "--test" ]; then
  shift
  timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/ab_pytest.log 2>&1
  echo "pytest exit $?"; tail -3 gpurun_out/ab_pytest.log
fi
n=0
for spec in "$@"; do
  IFS=',' read -ra parts <<< "$spec"
  v=${parts[0]}
  envs=("${parts[@]:1}")
  n=$((n+1))
  tag=$(echo "$spec" | tr ',=' '__')_$n
  (
    if [ "$v" != "default" ]; then export H263CU_LIB=$PWD/h263_rs_b200/libh263cu_$v.so; fi
    for e in "${envs[@]}"; do export "$e"; done
    timeout 600 python bench.py --steps 20 --warmup 3 --skip-extras > gpurun_out/ab_$tag.json 2> gpurun_out/ab_$tag.err
  )
  python - "$tag" <<'PY'
import json,sys
v=sys.argv[1]
try:
    d=json.load(open('gpurun_out/ab_%s.json'%v))
    print("%-40s kernel_ms %.4f  frac %.3f  step_ms %.4f  parity %s"%(v,d['roofline']['kernel_ms'],d['roofline']['frac'],d['ms_per_step'],(d.get('parity') or {}).get('bit_exact_vs_oracle')))
except Exception as e:
    print(v,'FAILED',e); print(open('gpurun_out/ab_%s.err'%v).read()[-800:])
PY
done
